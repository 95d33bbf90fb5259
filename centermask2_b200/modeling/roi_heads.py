"""``CenterROIHeads`` with the SAG-Mask head, the MaskIoU head and the keypoint head on libcm2.

Replaces ``centermask/modeling/centermask/center_heads.py:295-553`` (inference branches),
``pooler.py:192-366`` (ROIPooler eager path), ``sam.py:12-97`` (SpatialAttention(MaskHead)),
``mask_head.py:174-216`` (mask_rcnn_inference), ``maskiou_head.py:50-120`` and, when ``MODEL.KEYPOINT_ON``,
``keypoint_head.py:95-222`` (KRCNNConvDeconvUpsampleHead + keypoint_rcnn_inference -> heatmaps_to_keypoints [d2]).
"""
import torch

from .. import runtime
from ..config import validate_cfg
from ..arch import keypoint_head_param_spec, mask_head_param_spec, maskiou_head_param_spec
from ..engine import as_fmap
from .compat import ROI_HEADS_REGISTRY, Registry, ShapeSpec
from .params import PackedModule, attach_params

# The reference registers its heads in two more registries (mask_head.py:17, maskiou_head.py:10);
# the names are kept so that cfg.MODEL.ROI_MASK_HEAD.NAME / ROI_MASKIOU_HEAD.NAME resolve.
ROI_MASK_HEAD_REGISTRY = Registry("ROI_MASK_HEAD")
ROI_MASKIOU_HEAD_REGISTRY = Registry("ROI_MASKIOU_HEAD")
ROI_KEYPOINT_HEAD_REGISTRY = Registry("ROI_KEYPOINT_HEAD")                       # keypoint_head.py:13


@ROI_MASK_HEAD_REGISTRY.register()
class SpatialAttentionMaskHead(PackedModule):
    """``sam.py:31-97``: ``SpatialAttentionMaskHead(cfg, input_shape).forward(x[R, C, res, res]) -> logits[R, num_mask_classes,
    2res, 2res]`` (all classes, no sigmoid -- what ``mask_rcnn_inference`` consumes, mask_head.py:174-216).  Parameters under
    the reference's names (``mask_fcn{k}``, ``spatialAtt.conv``, ``deconv``, ``predictor``).  Inside ``CenterROIHeads`` the
    same parameters feed the fused ROI plan, which evaluates only the class of each ROI."""

    def __init__(self, cfg, input_shape):
        super().__init__()
        self.cfg = cfg
        validate_cfg(cfg, "roi_heads")
        self.in_channels = input_shape.channels
        attach_params(self, mask_head_param_spec(cfg, input_shape.channels))

    def _pack(self):
        from .. import packing
        eng = runtime.engine_for(self.cfg)
        if self._packed is None or self._engine is not eng:
            sd = self.state_dict()
            P = eng.pack_mask_head(sd)
            c = sd["predictor.weight"].shape[1]
            P["pred_conv"] = packing.conv_bias(sd, "predictor", [c], 1, 0, False, eng.dtype, eng.device, eng.tc)
            self._packed, self._engine = P, eng
        return eng, self._packed

    def forward(self, x):
        eng, P = self._pack()
        out = eng.run_mask_head_logits(as_fmap(x, eng.dtype, eng.device), P, P["pred_conv"])
        return out.view.permute(0, 3, 1, 2)


@ROI_MASKIOU_HEAD_REGISTRY.register()
class MaskIoUHead(PackedModule):
    """``maskiou_head.py:63-120``: ``MaskIoUHead(cfg, input_shape).forward(x[R, C, res, res], mask[R, 1, 2res, 2res]) ->
    [R, num_classes]`` (max-pool the mask, concatenate, 4 convs, 3 linear layers); ``input_shape.channels`` / ``.width`` as
    ``CenterROIHeads`` passes them (center_heads.py:353-356)."""

    def __init__(self, cfg, input_shape):
        super().__init__()
        self.cfg = cfg
        self.in_channels = input_shape.channels
        self.resolution = input_shape.width
        attach_params(self, maskiou_head_param_spec(cfg, input_shape.channels, input_shape.width))

    def _pack(self):
        eng = runtime.engine_for(self.cfg)
        if self._packed is None or self._engine is not eng:
            self._packed = eng.pack_maskiou_head(self.state_dict(), "", self.in_channels, self.resolution)
            self._engine = eng
        return eng, self._packed

    def forward(self, x, mask):
        eng, P = self._pack()
        probs = mask.to(device=eng.device, dtype=torch.float32).contiguous()
        out = eng.run_maskiou_head(as_fmap(x, eng.dtype, eng.device), probs, P)
        return out.buf.reshape(out.n, -1)


def build_mask_head(cfg, input_shape):
    """``mask_head.py:284-289``: the head named by ``cfg.MODEL.ROI_MASK_HEAD.NAME``."""
    return ROI_MASK_HEAD_REGISTRY.get(cfg.MODEL.ROI_MASK_HEAD.NAME)(cfg, input_shape)


def build_maskiou_head(cfg, input_shape):
    """``maskiou_head.py:123-128``: the head named by ``cfg.MODEL.ROI_MASKIOU_HEAD.NAME``."""
    return ROI_MASKIOU_HEAD_REGISTRY.get(cfg.MODEL.ROI_MASKIOU_HEAD.NAME)(cfg, input_shape)


@ROI_KEYPOINT_HEAD_REGISTRY.register()
class KRCNNConvDeconvUpsampleHead(object):
    """Marker for ``cfg.MODEL.ROI_KEYPOINT_HEAD.NAME`` (keypoint_head.py:168); parameters live under
    ``CenterROIHeads.keypoint_head.*``."""


@ROI_HEADS_REGISTRY.register()
class CenterROIHeads(PackedModule):
    def __init__(self, cfg, input_shape):
        super().__init__()
        self.cfg = cfg
        validate_cfg(cfg, "roi_heads")
        self.in_features = list(cfg.MODEL.ROI_HEADS.IN_FEATURES)                 # center_heads.py:121
        self.mask_on = bool(cfg.MODEL.MASK_ON)
        self.maskiou_on = bool(cfg.MODEL.MASKIOU_ON)
        self.keypoint_on = bool(cfg.MODEL.KEYPOINT_ON)                           # center_heads.py:360
        chans = {input_shape[f].channels for f in self.in_features}
        assert len(chans) == 1, chans                                            # center_heads.py:327
        self.strides = [input_shape[f].stride for f in self.in_features]
        in_ch = next(iter(chans))
        res = cfg.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION
        # the heads are real sub-modules built through their registries (center_heads.py:337-356); they own the parameters
        # (``mask_head.*`` / ``maskiou_head.*`` keys) and can be called on their own; the fused ROI plan packs the same tensors
        if self.mask_on:
            self.mask_head = build_mask_head(cfg, ShapeSpec(channels=in_ch, width=res, height=res))
        if self.maskiou_on:
            self.maskiou_head = build_maskiou_head(cfg, ShapeSpec(channels=in_ch, width=res, height=res))
        if self.keypoint_on:
            ROI_KEYPOINT_HEAD_REGISTRY.get(cfg.MODEL.ROI_KEYPOINT_HEAD.NAME)
            self.kp_in_features = list(cfg.MODEL.ROI_KEYPOINT_HEAD.IN_FEATURES)  # center_heads.py:363
            self.kp_strides = [input_shape[f].stride for f in self.kp_in_features]
            assert input_shape[self.kp_in_features[0]].channels == in_ch
            attach_params(self, {"keypoint_head." + k: v for k, v in keypoint_head_param_spec(cfg, in_ch).items()})

    def _head_gens(self):
        """Weight generations of the head sub-modules: a head whose weights were loaded on its own re-packs the plan too."""
        return tuple(getattr(self, n)._pack_gen for n in ("mask_head", "maskiou_head") if hasattr(self, n))

    def graph_token(self):
        return ("cm2w", id(self), (self._pack_gen,) + self._head_gens())

    def _pack(self):
        eng = runtime.engine_for(self.cfg)
        gens = self._head_gens()
        if self._packed is None or self._engine is not eng or self._packed_gens != gens:
            if self._packed is not None and self._engine is not None:
                self._engine.drop_graphs(id(self))
            self._packed = eng.pack_roi_heads(self.state_dict())
            self._engine, self._packed_gens = eng, gens
        return eng, self._packed

    def run(self, feats, det, image_sizes):
        """Device-only: feats list of FMaps, det fixed-size buffers -> (probs [N*R,1,m,m], mask_scores [N*R])."""
        eng, P = self._pack()
        return eng.run_roi_heads(feats, self.strides, det, image_sizes, P)

    def run_keypoints(self, feats, det, image_sizes):
        """Device-only: feats list of FMaps (kp_in_features), det fixed-size buffers -> f32 [N, R, K, 4] (x, y, logit, score)."""
        eng, P = self._pack()
        return eng.run_keypoint_head(feats, self.kp_strides, det, image_sizes, P)

    def forward(self, images, features, proposals, targets=None):
        """center_heads.py:384-411 (inference): returns ``(list[Instances], {})``."""
        assert targets is None and not self.training, "inference only"
        return self.forward_with_given_boxes(features, proposals), {}

    def forward_with_given_boxes(self, features, instances):
        """center_heads.py:413-444: adds ``pred_masks`` [R,1,28,28] and, when the batch has at least one
        detection, ``mask_scores`` [R] to the *same* Instances objects and returns them."""
        assert instances[0].has("pred_boxes") and instances[0].has("pred_classes")
        if not self.mask_on and not self.keypoint_on:
            return instances
        eng, _ = self._pack()
        det = _det_from_instances(instances, eng)
        sizes = [tuple(i.image_size) for i in instances]
        r_cap = det["boxes"].shape[1]
        total = sum(len(i) for i in instances)
        if self.mask_on:
            feats = [as_fmap(features[f], eng.dtype, eng.device) for f in self.in_features]
            probs, mask_scores = self.run(feats, det, sizes)
            for k, inst in enumerate(instances):
                m = len(inst)
                inst.pred_masks = probs[k * r_cap:k * r_cap + m].clone()              # mask_head.py:215-216
                if self.maskiou_on and total > 0:                                     # center_heads.py:511-517
                    inst.mask_scores = mask_scores[k * r_cap:k * r_cap + m].clone()   # maskiou_head.py:59-60
        if self.keypoint_on:                                                          # center_heads.py:442
            feats = [as_fmap(features[f], eng.dtype, eng.device) for f in self.kp_in_features]
            kp = self.run_keypoints(feats, det, sizes)
            for k, inst in enumerate(instances):
                m = kp[k, :len(inst)]
                inst.pred_keypoints = torch.cat([m[..., :2], m[..., 3:4]], dim=-1)    # keypoint_head.py:116 (x, y, score)
        return instances


def _det_from_instances(instances, eng):
    """Fixed-size ROI slot buffers for a list of Instances: reuse FCOS's device buffers when the
    Instances came from our own proposal generator, else pack the given boxes."""
    tags = [getattr(i, "_cm2_det", None) for i in instances]
    if all(t is not None for t in tags) and all(t[0] is tags[0][0] and t[1] == k for k, t in enumerate(tags)) \
            and tags[0][0]["boxes"].shape[0] == len(instances):
        det = tags[0][0]
        # The tag only says where the Instances came from.  The engine's detection buffers are shared: another FCOS forward
        # since then (generation stamp) or a caller that refined pred_boxes / pred_classes in place (content check) means
        # the buffers no longer describe THESE instances -- then the given boxes are packed instead.
        if det.get("gen") == eng._det_gen and all(
                len(i) <= det["boxes"].shape[1] and torch.equal(det["boxes"][k, :len(i)], i.pred_boxes.tensor)
                and torch.equal(det["classes"][k, :len(i)], i.pred_classes) for k, i in enumerate(instances)):
            return det
    n = len(instances)
    r_cap = max(1, max(len(i) for i in instances))
    dev = eng.device
    det = dict(boxes=torch.zeros((n, r_cap, 4), dtype=torch.float32, device=dev),
               scores=torch.zeros((n, r_cap), dtype=torch.float32, device=dev),
               classes=torch.zeros((n, r_cap), dtype=torch.int64, device=dev),
               count=torch.tensor([len(i) for i in instances], dtype=torch.int32, device=dev))
    for k, inst in enumerate(instances):
        m = len(inst)
        if m:
            det["boxes"][k, :m] = inst.pred_boxes.tensor.to(dev)
            det["classes"][k, :m] = inst.pred_classes.to(dev)
            if inst.has("scores"):
                det["scores"][k, :m] = inst.scores.to(dev)
    return det
