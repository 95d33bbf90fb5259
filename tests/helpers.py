"""Shared helpers for the parity tests (oracle side + comparison metrics)."""
import os

import numpy as np
import torch

from oracle.cases import CASES, case_cfg
from centermask2_b200.synth import synthetic_state_dict, synthetic_images

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances stated by BASELINE.json north_star for the fp32 variant.
BOX_TOL_PX = 1e-2
SCORE_TOL = 1e-3
MASK_IOU_MIN = 0.99


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def build_case(name, gold=None):
    """cfg, state_dict (with the golden's calibrated cls bias) and inputs of a named case."""
    overrides, sizes, wseed, iseed, target = CASES[name][:5]
    cfg = case_cfg(name)
    sd = synthetic_state_dict(cfg, seed=wseed)
    if gold is None:
        gold = load_golden(name)
    key = "proposal_generator.fcos_head.cls_logits.bias"
    sd[key] = torch.full_like(sd[key], gold["cls_bias"])
    inputs = []
    for i, (h, w) in enumerate(sizes):
        inputs.extend(synthetic_images(1, h, w, seed=iseed + i))
    return cfg, sd, inputs


def weights_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


def unpack_masks(post):
    shape = post["pred_masks_shape"]
    n = int(np.prod(shape))
    bits = np.unpackbits(post["pred_masks_packed"].numpy())[:n]
    return torch.from_numpy(bits.reshape(shape).astype(bool))


def mask_iou(a, b):
    a = a.reshape(a.shape[0], -1).bool()
    b = b.reshape(b.shape[0], -1).bool()
    inter = (a & b).sum(1).double()
    union = (a | b).sum(1).double()
    return torch.where(union > 0, inter / union.clamp(min=1), torch.ones_like(inter))


def assert_detections_match(got, ref, box_tol=BOX_TOL_PX, score_tol=SCORE_TOL, what=""):
    """got / ref: dicts with pred_boxes, scores, pred_classes, locations (sorted by score desc)."""
    assert len(got["scores"]) == len(ref["scores"]), "{}: {} vs {} detections".format(
        what, len(got["scores"]), len(ref["scores"]))
    if len(ref["scores"]) == 0:
        return
    # identical kept set *and* order: class + originating location identify a candidate
    assert torch.equal(got["pred_classes"].cpu().long(), ref["pred_classes"].long()), what + ": classes/order differ"
    assert torch.equal(got["locations"].cpu().float(), ref["locations"].float()), what + ": locations/order differ"
    db = (got["pred_boxes"].cpu().float() - ref["pred_boxes"].float()).abs().max().item()
    ds = (got["scores"].cpu().float() - ref["scores"].float()).abs().max().item()
    assert db <= box_tol, "{}: box diff {} px".format(what, db)
    assert ds <= score_tol, "{}: score diff {}".format(what, ds)
    if "mask_scores" in ref:
        # mask_scores = score * maskiou; with random-init weights the MaskIoU output is not confined to [0, 1]
        # (hundreds for the deep V-99 body), so the absolute tolerance is scaled by the reference magnitude.
        dm = (got["mask_scores"].cpu().float() - ref["mask_scores"].float()).abs().max().item()
        bound = score_tol * max(1.0, ref["mask_scores"].abs().max().item())
        assert dm <= bound, "{}: mask_score diff {} > {}".format(what, dm, bound)


def assert_masks_match(got, ref, what=""):
    """Pasted bool masks: IoU >= 0.99 per instance (north star).  A mask of fewer than 100 pixels cannot lose a
    single pixel without dropping below 0.99, and a pixel whose interpolated probability sits within float
    rounding of the 0.5 threshold may legitimately flip; so on such tiny masks one differing pixel is tolerated."""
    iou = mask_iou(got, ref)
    a = got.reshape(got.shape[0], -1).bool()
    b = ref.reshape(ref.shape[0], -1).bool()
    diff = (a ^ b).sum(1)
    area = (a | b).sum(1)
    ok = (iou >= MASK_IOU_MIN) | ((area < 100) & (diff <= 1))
    assert bool(ok.all()), "{}: mask IoU {} (diff px {}, area {})".format(what, iou[~ok].tolist(), diff[~ok].tolist(), area[~ok].tolist())


def assert_keypoints_match(got, ref, what="", xy_tol=1e-2, score_rtol=2e-3, max_moved_frac=0.0):
    """pred_keypoints [R, K, 3] = (x, y, score).  The location is an arg-max over a bicubic-resized map: where two
    pixels of the resized map are within float rounding of each other the arg-max may legitimately move, so at most
    ``max_moved_frac`` of the keypoints may differ in location -- and those must still carry the same score (the score
    is 1 / sum(exp(map - max)), a function of the maximum VALUE only)."""
    got, ref = got.cpu().float(), ref.float()
    assert got.shape == ref.shape, "{}: {} vs {}".format(what, tuple(got.shape), tuple(ref.shape))
    if ref.numel() == 0:
        return
    moved = ((got[..., :2] - ref[..., :2]).abs() > xy_tol).any(dim=-1)
    frac = moved.float().mean().item()
    assert frac <= max_moved_frac, "{}: {:.2%} of the keypoints moved (max diff {} px)".format(
        what, frac, (got[..., :2] - ref[..., :2]).abs().max().item())
    ds = ((got[..., 2] - ref[..., 2]).abs() / ref[..., 2].abs().clamp(min=1e-6)).max().item()
    assert ds <= score_rtol, "{}: keypoint score relative diff {}".format(what, ds)


def pack_lowres(low):
    """[R, K, 2res, 2res] -> the conv engine's phase layout [R, res, res, 4, K] (kp_lowres_offset)."""
    r, k, s, _ = low.shape
    res = s // 2
    return low.reshape(r, k, res, 2, res, 2).permute(0, 2, 4, 3, 5, 1).reshape(r, res, res, 4, k).contiguous()


def keypoint_boxes(g, r, extent=200.0):
    """Boxes of assorted sizes, including sub-pixel ones (clamped to 1 px by heatmaps_to_keypoints) and off-image ones."""
    xy = torch.rand(r, 2, generator=g) * extent - 20.0
    wh = torch.exp(torch.rand(r, 2, generator=g) * 6.0 - 1.0)          # 0.37 .. 150 px
    return torch.cat([xy, xy + wh], dim=1).float()
