"""``FCOS`` proposal generator on libcm2.

Replaces ``centermask/modeling/fcos/fcos.py:28-240`` (FCOS, FCOSHead, Scale) and the inference half of
``fcos/fcos_outputs.py:372-495`` (decode, threshold, per-level top-k, class-aware NMS, post top-k).
Semantics notes: the per-level pre-NMS top-k is applied (upstream
semantics, ``PRE_NMS_TOPK_TEST``); the fork commented it out (``fcos_outputs.py:444-449``), both agree
whenever a level has <= PRE_NMS_TOPK_TEST candidates above the threshold.
"""
import torch

from .. import runtime
from ..config import validate_cfg
from ..arch import fcos_param_spec
from ..engine import as_fmap
from .compat import PROPOSAL_GENERATOR_REGISTRY, Boxes, Instances
from .params import PackedModule, attach_params


@PROPOSAL_GENERATOR_REGISTRY.register()
class FCOS(PackedModule):
    def __init__(self, cfg, input_shape):
        super().__init__()
        self.cfg = cfg
        validate_cfg(cfg, "fcos")
        self.in_features = list(cfg.MODEL.FCOS.IN_FEATURES)                 # fcos.py:35
        self.fpn_strides = list(cfg.MODEL.FCOS.FPN_STRIDES)
        chans = {input_shape[f].channels for f in self.in_features}
        assert len(chans) == 1, "Each level must have the same channel!"    # fcos.py:166
        attach_params(self, fcos_param_spec(cfg, chans.pop()))
        if cfg.MODEL.FCOS.USE_DEFORMABLE:
            raise NotImplementedError("MODEL.FCOS.USE_DEFORMABLE is out of scope (broken in the reference, SURVEY 2 #15)")

    def _pack(self):
        eng = runtime.engine_for(self.cfg)
        if self._packed is None or self._engine is not eng:
            self._packed = eng.pack_fcos(self.state_dict())
            self._engine = eng
        return eng, self._packed

    def compute_locations(self, features):
        """fcos.py:120-144 (kept for API parity; the kernels recompute locations on the fly)."""
        out = []
        for f, s in zip(features, self.fpn_strides):
            h, w = f.shape[-2:]
            xs = torch.arange(0, w * s, step=s, dtype=torch.float32, device=f.device)
            ys = torch.arange(0, h * s, step=s, dtype=torch.float32, device=f.device)
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            out.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), dim=1) + s // 2)
        return out

    def detect(self, feats):
        """feats: list of engine FMaps.  Device-only: returns the fixed-size detection buffers."""
        eng, P = self._pack()
        head = eng.run_fcos_head(feats, P)
        return eng.run_fcos_post(head, reg_scale=P["reg_scale"])

    def forward(self, images, features, gt_instances=None):
        """``fcos.py:61-118``: returns ``(list[Instances], {})`` with fields pred_boxes, scores,
        pred_classes, locations (``fcos_outputs.py:458-462``), sorted by descending score."""
        assert gt_instances is None and not self.training, "inference only"
        eng, _ = self._pack()
        feats = [as_fmap(features[f], eng.dtype, eng.device) for f in self.in_features]
        det = self.detect(feats)
        return instances_from_det(det, images.image_sizes), {}


def instances_from_det(det, image_sizes):
    """Fixed-size device buffers -> per-image ``Instances`` (one host sync for the counts)."""
    counts = det["count"].tolist()
    over = (det["cand_count"] > det["cand_cap"]).any().item() if "cand_count" in det else False
    if over:
        raise RuntimeError("FCOS candidate buffer overflow (> {} candidates above threshold in one level); "
                           "results would be order-dependent".format(det["cand_cap"]))
    out = []
    for i, (k, size) in enumerate(zip(counts, image_sizes)):
        inst = Instances(tuple(size))
        inst.pred_boxes = Boxes(det["boxes"][i, :k].clone())
        inst.scores = det["scores"][i, :k].clone()
        inst.pred_classes = det["classes"][i, :k].clone()
        inst.locations = det["locations"][i, :k].clone()
        inst._cm2_det = (det, i)
        out.append(inst)
    return out
