"""``build_fcos_vovnet_fpn_backbone`` -- VoVNetV2-eSE + FPN + P6/P7 on libcm2.

Replaces ``centermask/modeling/backbone/vovnet.py:527-555`` (builder), ``:380-489`` (VoVNet),
``:263-376`` (OSA), ``fpn.py:17-35`` (LastLevelP6P7) and detectron2's ``FPN`` [d2].
"""
from .. import runtime
from ..config import validate_cfg
from ..arch import backbone_param_spec, vovnet_blocks
from .compat import BACKBONE_REGISTRY, ShapeSpec
from .params import PackedModule, attach_params


class VoVNetFPN(PackedModule):
    """Drop-in for ``FPN(bottom_up=VoVNet, ...)``: ``forward(x[N,3,H,W]) -> {"p3".."p7": [N,C,h,w]}``.

    Returned tensors are zero-copy channels_last-strided views of the engine's NHWC buffers; they are
    overwritten by the next ``forward`` call of the same model (clone to keep)."""

    def __init__(self, cfg, input_shape=None):
        super().__init__()
        self.cfg = cfg
        validate_cfg(cfg, "backbone")
        attach_params(self, backbone_param_spec(cfg))
        fc = cfg.MODEL.FPN.OUT_CHANNELS
        self._size_divisibility = 32
        levels = [int(f[-1]) for f in cfg.MODEL.FPN.IN_FEATURES]
        levels += [max(levels) + 1 + i for i in range(cfg.MODEL.FCOS.TOP_LEVELS)]
        self._out_features = ["p{}".format(l) for l in levels]
        self._out_feature_channels = {k: fc for k in self._out_features}
        self._out_feature_strides = {"p{}".format(l): 2 ** l for l in levels}
        vovnet_blocks(cfg.MODEL.VOVNET.CONV_BODY)            # validates CONV_BODY early

    @property
    def size_divisibility(self):
        return self._size_divisibility

    def output_shape(self):
        return {k: ShapeSpec(channels=self._out_feature_channels[k], stride=self._out_feature_strides[k])
                for k in self._out_features}

    def _pack(self):
        eng = runtime.engine_for(self.cfg)
        if self._packed is None or self._engine is not eng:
            self._packed = eng.pack_backbone(self.state_dict())
            self._engine = eng
        return eng, self._packed

    def forward_fmap(self, x):
        """x: engine FMap of the normalised, padded batch."""
        eng, P = self._pack()
        return eng.run_backbone(x, P)

    def forward(self, x):
        from ..engine import as_fmap
        eng, P = self._pack()
        feats = eng.run_backbone(as_fmap(x, eng.dtype, eng.device), P)
        return {k: v.nchw() for k, v in feats.items()}


@BACKBONE_REGISTRY.register()
def build_fcos_vovnet_fpn_backbone(cfg, input_shape=None):
    """Same name / signature as ``vovnet.py:527-555``."""
    return VoVNetFPN(cfg, input_shape)
