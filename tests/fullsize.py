"""Full-size (BASELINE configs[0]/[2] shape: V-39-eSE-FPN, 800x1333) end-to-end parity reports.

``deviation_report(precision)`` runs the registered B200 model on two seeded 800x1333 images and compares every
output field with the CPU oracle (``oracle/restate.py``): the pure fp32 restatement for the fp32 engines, the
bf16-rounding restatement (``restate.bf16_sim``) for the bf16 tensor-core engine.  The report is a plain dict of
measured deviations; ``tests/test_gpu_fullsize.py`` asserts the gates on it and ``tools/parity_report.py`` writes it
to ``profiles/`` so that the measured numbers travel with the repository.

Detections are matched by (class, originating location): that pair identifies a candidate independently of its rank.
"""
import functools
import os

import torch

H, W = 800, 1333
N_IMAGES = 2
# (weight seed, image seed) of the draw.  CM2_FULLSIZE_SEEDS="weights,images" changes the default for tools/parity_report.py;
# every function below also takes ``seeds`` explicitly (tests/test_gpu_fullsize.py checks a second draw: the measured deviations
# must not be a property of one set of random-init weights).
DEFAULT_SEEDS = tuple(int(v) for v in os.environ.get("CM2_FULLSIZE_SEEDS", "101,202").split(","))
WEIGHT_SEED, IMAGE_SEED = DEFAULT_SEEDS
CAND_TARGET = 800
BIAS_KEY = "proposal_generator.fcos_head.cls_logits.bias"


def _fields(inst):
    return {k: (v.tensor if hasattr(v, "tensor") else v).detach().cpu() for k, v in inst.get_fields().items()}


@functools.lru_cache(maxsize=None)
def workload(seeds=DEFAULT_SEEDS):
    """cfg-independent pieces: state_dict with the calibrated cls bias (picked from the fp32 oracle's own logits, so the
    workload does not depend on any device result) and the uint8 inputs."""
    from centermask2_b200.config import get_cfg
    from centermask2_b200.synth import synthetic_images, synthetic_state_dict, calibrate_cls_bias
    from oracle import restate
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml")
    sd = synthetic_state_dict(cfg, seed=seeds[0])
    imgs = synthetic_images(N_IMAGES, H, W, seed=seeds[1])
    for b in imgs:
        b["image"] = b["image"].to(torch.uint8)           # what a data loader hands over
    sd[BIAS_KEY] = torch.zeros_like(sd[BIAS_KEY])
    tr = {}
    restate.run_model(_float_inputs(imgs), sd, cfg, postprocess=False, trace=tr)
    sd[BIAS_KEY] = torch.full_like(sd[BIAS_KEY], calibrate_cls_bias(tr["logits"], CAND_TARGET))
    return sd, imgs


def _float_inputs(imgs):
    return [dict(b, image=b["image"].float()) for b in imgs]


@functools.lru_cache(maxsize=None)
def oracle_outputs(bf16, seeds=DEFAULT_SEEDS):
    """(raw, post, trace) of the oracle; ``bf16`` selects the bf16-rounding restatement."""
    from centermask2_b200.config import get_cfg
    from oracle import restate
    sd, imgs = workload(seeds)
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml")
    tr = {}
    with restate.bf16_sim(bool(bf16)):
        raw = restate.run_model(_float_inputs(imgs), sd, cfg, postprocess=False, trace=tr)
        post = [restate.detector_postprocess(d, H, W) for d in raw]
    return raw, post, tr


def _keys(d):
    return [(int(c), float(l[0]), float(l[1])) for c, l in zip(d["pred_classes"], d["locations"])]


def margins(raw, tr, thresh=0.05):
    """How well-posed set equality is on this draw (SURVEY 8d margin check): distance of the closest class probability
    to the candidate threshold, and the smallest score gap between consecutive kept detections."""
    gap = min((torch.sigmoid(lg) - thresh).abs().min().item() for lg in tr["logits"])
    sgap = min(((d["scores"][:-1] - d["scores"][1:]).abs().min().item() if len(d["scores"]) > 1 else 1.0) for d in raw)
    return {"min_prob_distance_to_threshold": gap, "min_score_gap_between_kept": sgap}


def roialign_margin_px(boxes, image_size, feat_hw, res=14):
    """Distance, in image pixels, of every ROI's closest ROIAlign sample to a DISCONTINUITY of the reference operator.

    torchvision's roi_align (the op behind detectron2's ROIAlign, pooler.py:249-255) drops a sample whose coordinate is
    below -1 or above the feature extent and clamps it to the border otherwise: a sample that crosses x = -1 jumps from
    f[.., 0] to 0.  A box that differs by less than the box tolerance can therefore change the pooled feature by O(1)
    when a sample sits that close to the jump -- downstream fields of such a ROI (mask, mask score) are ill-posed at the
    stated box tolerance, exactly like a score within rounding of the threshold (SURVEY 8d margin check).
    ``feat_hw``: {level index: (h, w)} of the pooled levels (strides 8, 16, 32)."""
    from oracle import restate
    if boxes.numel() == 0:
        return boxes.new_zeros((0,))
    lv = restate.assign_levels_by_ratio(boxes, image_size[0] * image_size[1], 3, 5)
    out = []
    for b, l in zip(boxes.double(), lv.tolist()):
        stride = 8 << l
        fh, fw = feat_hw[l]
        best = float("inf")
        for lo, hi, extent in ((b[0], b[2], fw), (b[1], b[3], fh)):
            start = lo.item() / stride - 0.5
            size = (hi.item() - lo.item()) / stride
            grid = max(1, int(-(-size // res)))
            bin_ = size / res
            idx = torch.arange(res * grid, dtype=torch.float64)
            pos = start + (idx // grid) * bin_ + ((idx % grid) + 0.5) * bin_ / grid
            d = torch.minimum((pos + 1.0).abs(), (pos - extent).abs()).min().item() * stride
            best = min(best, d)
        out.append(best)
    return torch.tensor(out)


def compare(got_raw, got_post, ref_raw, ref_post, feat_hw=None, box_tol=1e-2):
    """Per-field deviations over the detections both sides kept, plus the set / order agreement.  ROIs whose ROIAlign
    sampling sits within ``box_tol`` of a discontinuity of the reference operator (``roialign_margin_px``) are counted
    in ``roialign_ill_posed`` and left out of the mask / mask-score statistics."""
    from tests.helpers import mask_iou
    rep = {"images": len(ref_raw), "kept_ref": [], "kept_got": [], "common": [], "same_order": True, "box_px": 0.0,
           "score": 0.0, "mask_score_rel": 0.0, "mask_prob": 0.0, "post_box_px": 0.0, "mask_iou_min": 1.0,
           "mask_iou_below_0p99": 0, "masks_compared": 0, "roialign_ill_posed": 0, "roialign_margin_px_min": None}
    ious = []
    for g, gp, r, rp in zip(got_raw, got_post, ref_raw, ref_post):
        kg, kr = _keys(g), _keys(r)
        rep["kept_ref"].append(len(kr))
        rep["kept_got"].append(len(kg))
        pos_g = {k: i for i, k in enumerate(kg)}
        common = [k for k in kr if k in pos_g]
        rep["common"].append(len(common))
        rep["same_order"] = rep["same_order"] and kg == kr
        if not common:
            continue
        ig = torch.tensor([pos_g[k] for k in common])
        ir = torch.tensor([i for i, k in enumerate(kr) if k in pos_g])
        rep["box_px"] = max(rep["box_px"], (g["pred_boxes"][ig] - r["pred_boxes"][ir]).abs().max().item())
        rep["score"] = max(rep["score"], (g["scores"][ig] - r["scores"][ir]).abs().max().item())
        well = torch.ones(len(kr), dtype=torch.bool)
        if feat_hw is not None:
            margin = roialign_margin_px(r["pred_boxes"], r["image_size"], feat_hw)
            well = margin >= box_tol
            rep["roialign_ill_posed"] += int((~well).sum())
            m = margin.min().item()
            rep["roialign_margin_px_min"] = m if rep["roialign_margin_px_min"] is None else min(m, rep["roialign_margin_px_min"])
            sel = well[ir]
            ig, ir = ig[sel], ir[sel]
            if len(ir) == 0:
                continue
        if "mask_scores" in r:
            ms_r = r["mask_scores"][ir]
            rel = ((g["mask_scores"][ig] - ms_r).abs() / ms_r.abs().clamp(min=1.0)).max().item()
            rep["mask_score_rel"] = max(rep["mask_score_rel"], rel)
        rep["mask_prob"] = max(rep["mask_prob"], (g["pred_masks"][ig] - r["pred_masks"][ir]).abs().max().item())
        # post-processed: boxes dropped by detector_postprocess vanish on both sides alike only when the boxes agree, so
        # match again by key
        kgp, krp = _keys(gp), _keys(rp)
        pos_gp = {k: i for i, k in enumerate(kgp)}
        bad = {k for k, ok in zip(kr, well.tolist()) if not ok}
        cp = [k for k in krp if k in pos_gp and k not in bad]
        if cp:
            jg = torch.tensor([pos_gp[k] for k in cp])
            jr = torch.tensor([i for i, k in enumerate(krp) if k in pos_gp and k not in bad])
            rep["post_box_px"] = max(rep["post_box_px"], (gp["pred_boxes"][jg] - rp["pred_boxes"][jr]).abs().max().item())
            a, b = gp["pred_masks"][jg], rp["pred_masks"][jr]
            iou = mask_iou(a, b)
            ious.append(iou)
            # same exemption as helpers.assert_masks_match: a mask of < 100 px cannot lose one pixel and stay >= 0.99
            a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
            tiny_ok = ((a2 | b2).sum(1) < 100) & ((a2 ^ b2).sum(1) <= 1)
            rep["mask_iou_failures"] = rep.get("mask_iou_failures", 0) + int(((iou < 0.99) & ~tiny_ok).sum())
    if ious:
        iou = torch.cat(ious)
        rep["mask_iou_min"] = iou.min().item()
        rep["mask_iou_median"] = iou.median().item()
        rep["mask_iou_below_0p99"] = int((iou < 0.99).sum())
        rep["masks_compared"] = int(iou.numel())
    rep["overlap"] = sum(rep["common"]) / max(1, sum(rep["kept_ref"]))
    return rep


def device_outputs(precision, seeds=DEFAULT_SEEDS):
    """(raw, post) of the registered B200 model at ``precision`` (fields on the CPU)."""
    import centermask2_b200 as cm
    from centermask2_b200 import runtime
    from centermask2_b200.config import get_cfg
    sd, imgs = workload(seeds)
    runtime.reset()
    try:
        cfg = get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.B200.PRECISION", precision])
        model = cm.build_model(cfg)
        model.load_state_dict(sd)
        raw = [_fields(i) for i in model.inference(imgs, do_postprocess=False)]
        post = [_fields(o["instances"]) for o in model(imgs)]
        torch.cuda.synchronize()
    finally:
        runtime.reset()
    return raw, post


def deviation_report(precision, seeds=DEFAULT_SEEDS):
    bf16 = precision == "bf16"
    ref_raw, ref_post, tr = oracle_outputs(bf16, seeds)
    got_raw, got_post = device_outputs(precision, seeds)
    feat_hw = {l: tuple(tr["features"]["p{}".format(3 + l)].shape[-2:]) for l in range(3)}
    rep = compare(got_raw, got_post, ref_raw, ref_post, feat_hw)
    rep["precision"] = precision
    rep["seeds"] = list(seeds)
    rep["oracle"] = "restate.bf16_sim()" if bf16 else "restate (fp32)"
    rep["margins"] = margins(ref_raw, tr)
    if bf16:
        # for the record: the bf16 engine against the PURE fp32 oracle (what a user switching precision sees)
        f_raw, f_post, _ = oracle_outputs(False, seeds)
        raw32 = compare(got_raw, got_post, f_raw, f_post, feat_hw)
        rep["vs_fp32_oracle"] = {k: raw32[k] for k in ("overlap", "box_px", "score", "mask_score_rel", "mask_iou_min",
                                                        "mask_iou_median", "mask_iou_below_0p99", "masks_compared")
                                 if k in raw32}
    return rep
