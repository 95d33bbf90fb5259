#!/usr/bin/env python
"""BASELINE config 5: FCOS post-process + SAG-Mask microbenchmark (no convolutions).

    python tools/micro_post.py [--batch 32] [--cand 1000] [--rois 100] [--out profiles/r1_micro_post_b32.json]

5 FPN levels of the 800x1344 pyramid, head outputs drawn directly: logits N(mu_l, 1) with mu_l chosen
so that ~`cand` entries per level and image exceed the 0.05 threshold, regression |N(0,1)|*4 (stride units),
centerness N(0,1); ROI stage fed `rois` boxes per image with log-uniform areas so that all three levels are used;
random bf16 P3-P5 features and 28x28 mask probabilities.  Every kernel is FIRST CHECKED against the oracle at this
size (``check_*``: restate.fcos_postprocess for decode / top-k / NMS over all images; torchvision roi_align with the
reference's level rule, the SAM restatement and restate.paste_masks on the first images), then timed with CUDA events
(3 warm-ups, 10 launches; working sets exceed the 126 MB L2 except for top-k / NMS) and reported as achieved GB/s over its
ALGORITHMIC bytes against the measured HBM copy peak.  One JSON object per kernel.  ``bench.py --config post``
calls ``run()``.
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from centermask2_b200 import lib                                    # noqa: E402
from centermask2_b200.config import get_cfg                        # noqa: E402
from centermask2_b200.engine import Engine, FMap                   # noqa: E402

H, W = 800, 1344
LEVELS = [(100, 168, 8), (50, 84, 16), (25, 42, 32), (13, 21, 64), (7, 11, 128)]


def norm_isf(p):
    """Inverse survival function of N(0,1) (Acklam-free: bisection on erfc)."""
    lo, hi = -10.0, 10.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if 0.5 * math.erfc(mid / math.sqrt(2.0)) > p:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def timed(fn, reps=10, warm=3):
    """Device time per call: `reps` calls captured into one CUDA graph (what the engine replays in production), the
    graph replayed 3 times between two events.  Host-side launch overhead (ctypes argument marshalling costs more than
    some of these kernels run) is thereby kept out of the number."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if os.environ.get("CM2_MICRO_EAGER") == "1":
        graph, replays = None, 1
    else:
        graph, replays = torch.cuda.CUDAGraph(), 3
        with torch.cuda.graph(graph):
            for _ in range(reps):
                fn()
        graph.replay()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        if graph is None:
            for _ in range(reps):
                fn()
        else:
            graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays)


# ---------------------------------------------------------------------------------------------------
# correctness at the benchmark size (the oracle is the checker, never the thing timed)
# ---------------------------------------------------------------------------------------------------
def check_fcos_post(det, head, cfg, n):
    """decode + per-level top-k + class-aware NMS + post top-k of ALL images against restate.fcos_postprocess
    (fcos_outputs.py:372-495): identical kept set and order, boxes <= 1e-2 px, scores <= 1e-3."""
    from oracle import restate
    from tests.helpers import assert_detections_match
    logits = [lg.view.permute(0, 3, 1, 2).float().cpu() for lg, _ in head]
    regs = [torch.relu(rc.view[..., :4]).permute(0, 3, 1, 2).float().cpu() for _, rc in head]
    ctrs = [rc.view[..., 4:5].permute(0, 3, 1, 2).float().cpu() for _, rc in head]
    ref = restate.fcos_postprocess(logits, regs, ctrs, [(H, 1333)] * n, cfg)
    counts = det["count"].tolist()
    for i, r in enumerate(ref):
        k = counts[i]
        got = {"pred_boxes": det["boxes"][i, :k].cpu(), "scores": det["scores"][i, :k].cpu(),
               "pred_classes": det["classes"][i, :k].cpu(), "locations": det["locations"][i, :k].cpu()}
        assert_detections_match(got, r, what="micro_post fcos_post image {}".format(i))
    return sum(counts)


def check_roialign(roi, lvl, feats, boxes, n_check, R):
    """ROIAlign + ratio level rule (pooler.py:80-118, 320-366) of the first images against torchvision on the same
    bf16 features; tolerance = one bf16 rounding of the output."""
    import torchvision
    from oracle import restate
    b = boxes[:n_check].reshape(-1, 4).cpu()
    lv_ref = restate.assign_levels_by_ratio(b, float(H * W), 3, 5)
    assert torch.equal(lvl[:n_check * R].cpu().long(), lv_ref), "level assignment differs from pooler.py:80-118"
    rois = torch.cat([torch.arange(n_check).repeat_interleave(R).float()[:, None], b], dim=1)
    ref = torch.zeros((n_check * R, feats[0].c, 14, 14))
    for li, f in enumerate(feats):
        sel = (lv_ref == li).nonzero().squeeze(1)
        fm = f.view[:n_check].permute(0, 3, 1, 2).float().cpu()
        ref[sel] = torchvision.ops.roi_align(fm, rois[sel], 14, 1.0 / (8 << li), 0, True)
    got = roi.view[:n_check * R].permute(0, 3, 1, 2).float().cpu()
    err = (got - ref).abs()
    bound = ref.abs() / 128 + ref.abs().mean() / 128 + 1e-3
    assert (err <= bound).all(), "roialign: max err {} beyond bf16 rounding".format(err.max().item())


def check_sam(roi, att, w18, n_rois):
    """SpatialAttention (sam.py:23-28) of the first ROIs against its torch restatement on the same bf16 input."""
    import torch.nn.functional as F
    x = roi.view[:n_rois].permute(0, 3, 1, 2).float().cpu()
    pooled = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    ref = x * torch.sigmoid(F.conv2d(pooled, w18.cpu().reshape(1, 2, 3, 3), None, 1, 1))
    got = att.view[:n_rois].permute(0, 3, 1, 2).float().cpu()
    err = (got - ref).abs()
    assert (err <= ref.abs() / 64 + 1e-2).all(), "spatial attention: max err {}".format(err.max().item())


def check_paste(masks, probs, bx, valid, n_masks):
    """paste_masks_in_image [d2] (restate.paste_masks; deploy_utils.py:129-158) of the first masks: IoU >= 0.99 each and
    >= 99.999 % of the pixels identical (a pixel whose interpolated probability is within rounding of 0.5 may flip)."""
    from oracle import restate
    from tests.helpers import mask_iou
    ok = valid.reshape(-1)[:n_masks].bool().cpu()
    ref = restate.paste_masks(probs[:n_masks, 0].cpu(), bx.reshape(-1, 4)[:n_masks].cpu(), H, 1333)
    got = masks[:n_masks].bool().cpu()
    assert not got[~ok].any()
    iou = mask_iou(got[ok], ref[ok])
    same = (got[ok] == ref[ok]).float().mean().item()
    assert iou.min().item() >= 0.99 and same >= 0.99999, "paste: min IoU {}, identical pixels {}".format(iou.min().item(), same)


def run(batch=32, cand=1000, rois=100, check=True, seed=5, old=False, out=None):
    """Check (``check=True``) and time every kernel of BASELINE configs[4]; returns the list of per-kernel rows."""
    class A(object):
        pass
    args = A()
    args.batch, args.cand, args.rois, args.old, args.out = batch, cand, rois, old, out
    n, R = args.batch, args.rois
    dev = "cuda"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.FCOS.POST_NMS_TOPK_TEST", R, "MODEL.B200.PRECISION", "bf16"])
    eng = Engine(cfg, "bf16", dev)
    g = torch.Generator(device=dev).manual_seed(seed)
    ncls = 80
    thr_logit = math.log(0.05 / 0.95)
    head = []
    for (h, w, s) in LEVELS:
        frac = min(0.5, args.cand / float(h * w * ncls))
        mu = thr_logit - norm_isf(frac)
        lg = torch.randn((n, h, w, ncls), device=dev, generator=g) + mu
        rc = torch.zeros((n, h, w, 16), device=dev)
        rc[..., :4] = torch.randn((n, h, w, 4), device=dev, generator=g).abs() * 4.0
        rc[..., 4] = torch.randn((n, h, w), device=dev, generator=g)
        head.append((FMap(lg, 0), FMap(rc, 0)))
    results = []

    def report(name, ms, alg_bytes, note=""):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        row = {"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(alg_bytes / 1e6, 1), "GBps": round(gbs, 1),
               "frac_of_hbm_peak": round(gbs / hbm, 3), "batch": n, "checked_against_oracle": bool(check), "note": note}
        results.append(row)
        print(json.dumps(row), file=sys.stderr)

    # ---- decode + threshold (A11): 22 400 x 85 fp32 per image
    det = eng.run_fcos_post(head)
    torch.cuda.synchronize()
    if check:
        kept = check_fcos_post(det, head, cfg, n)
        print("checked: decode / top-k / NMS of {} images == restate.fcos_postprocess ({} detections)".format(n, kept), file=sys.stderr)
    cand = det["cand_count"].float().mean().item()
    B = eng.buffer
    L = len(LEVELS)
    cap = det["cand_cap"]
    cb = dict(boxes=B("cand_boxes", (n, L, cap, 4), torch.float32, False), score=B("cand_score", (n, L, cap), torch.float32, False),
              cls=B("cand_cls", (n, L, cap), torch.int32, False), flat=B("cand_flat", (n, L, cap), torch.int32, False),
              count=B("cand_count", (n, L), torch.int32))
    cbuf = lib.cand_buffers(cb["boxes"], cb["score"], cb["cls"], cb["flat"], cb["count"])

    def decode():
        cb["count"].zero_()
        lib.fcos_decode_levels([lg.view for lg, _ in head], [rc.view for _, rc in head], [s for _, _, s in LEVELS], [1.0] * L, 0.05, False,
                               cap, cbuf)
    ms = timed(decode)
    report("fcos_decode (5 levels)", ms, n * 22400 * 85 * 4, "{:.0f} candidates / level / image".format(cand))
    if args.old:
        os.environ["CM2_DECODE_VARIANT"] = "0"
        report("fcos_decode (5 levels) [v0: CTA per image row]", timed(decode), n * 22400 * 85 * 4)
        os.environ["CM2_DECODE_VARIANT"] = "1"

    # ---- per-level top-k + class-aware NMS + post top-k (A12-A14): latency-bound
    ms_all = timed(lambda: eng.run_fcos_post(head))
    report("fcos_select + nms (per image CTA)", max(ms_all - ms, 1e-6), n * L * args.cand * 24,
           "latency-bound: {} candidates in, {} kept / image".format(int(cand * L), int(det["count"].float().mean().item())))

    # ---- ROIAlign with level assignment (A15 + A16)
    c = 256
    feats = [eng.fmap("mf{}".format(i), n, h, w, c) for i, (h, w, _) in enumerate(LEVELS[:3])]
    for f in feats:
        f.view.copy_(torch.randn(f.view.shape, device=dev, generator=g).to(torch.bfloat16))
    area = torch.exp(torch.rand((n, R), device=dev, generator=g) * (math.log(0.9 * H * W) - math.log(32.0 * 32.0)) + math.log(32.0 * 32.0))
    ar = torch.exp((torch.rand((n, R), device=dev, generator=g) - 0.5) * 1.4)
    bw, bh = torch.sqrt(area * ar).clamp(max=W - 1.0), torch.sqrt(area / ar).clamp(max=H - 1.0)
    x0 = torch.rand((n, R), device=dev, generator=g) * (W - bw)
    y0 = torch.rand((n, R), device=dev, generator=g) * (H - bh)
    boxes = torch.stack([x0, y0, x0 + bw, y0 + bh], dim=2).contiguous()
    counts = torch.full((n,), R, dtype=torch.int32, device=dev)
    img_area = torch.full((n,), float(H * W), device=dev)
    roi = eng.fmap("mroi", n * R, 14, 14, c)
    lvl = torch.zeros((n * R,), dtype=torch.int32, device=dev)
    def roialign():
        lib.roialign_fpn([f.view for f in feats], [8, 16, 32], boxes, counts, n, R, img_area, 0, 0, roi.view, lvl)
    roialign()
    torch.cuda.synchronize()
    if check:
        check_roialign(roi, lvl, feats, boxes, min(n, 4), R)
        print("checked: ROIAlign + level rule of {} ROIs == torchvision roi_align on the same features".format(min(n, 4) * R), file=sys.stderr)
    ms = timed(roialign)
    hist = torch.bincount(lvl.long(), minlength=3).tolist()
    report("roialign_fpn (+ level assignment)", ms, n * (R * c * 196 * 2 + c * 22050 * 2), "ROIs per level {}".format(hist))
    if args.old:
        got2 = roi.view.float().clone()
        os.environ["CM2_ROIALIGN_VARIANT"] = "1"
        roialign()
        torch.cuda.synchronize()
        dev_ = (roi.view.float() - got2).abs().max().item()
        print("roialign column walk vs merged taps: max |diff| = {:.4g} (bf16 outputs, max |value| {:.3g})".format(
            dev_, got2.abs().max().item()), file=sys.stderr)
        report("roialign_fpn [v1: CTA per ROI, merged taps]", timed(roialign), n * (R * c * 196 * 2 + c * 22050 * 2))
        os.environ["CM2_ROIALIGN_VARIANT"] = "0"
        report("roialign_fpn [v0: thread per (bin, 8 channels), sample loop]", timed(roialign), n * (R * c * 196 * 2 + c * 22050 * 2))
        del os.environ["CM2_ROIALIGN_VARIANT"]

    # ---- spatial attention (A18)
    att = eng.fmap("matt", n * R, 14, 14, c)
    w18 = torch.randn((18,), device=dev, generator=g)
    lib.spatial_attention(roi.view, att.view, w18)
    torch.cuda.synchronize()
    if check:
        check_sam(roi, att, w18, min(n, 4) * R)
        print("checked: spatial attention of {} ROIs == sam.py:23-28 restated".format(min(n, 4) * R), file=sys.stderr)
    ms = timed(lambda: lib.spatial_attention(roi.view, att.view, w18))
    report("spatial_attention", ms, n * R * c * 196 * 2 * 2)

    # ---- paste-back (A23)
    probs = torch.rand((n * R, 1, 28, 28), device=dev, generator=g)
    sizes = [(H, 1333)] * n
    bx, valid = eng.rescale_boxes(boxes, sizes, sizes)
    chunk = max(1, min(n, 600 // R))                       # <= 65535 ROIs per launch and a bounded output buffer
    masks = torch.empty((chunk * R, H, 1333), dtype=torch.uint8, device=dev)

    def paste():
        for i0 in range(0, n, chunk):
            k = min(chunk, n - i0)
            lib.paste_masks(probs[i0 * R:(i0 + k) * R], bx[i0:i0 + k], valid[i0:i0 + k], masks, k * R, 28, H, 1333, 0.5)
    if check:
        lib.paste_masks(probs[:chunk * R], bx[:chunk], valid[:chunk], masks, chunk * R, 28, H, 1333, 0.5)
        torch.cuda.synchronize()
        check_paste(masks, probs, bx, valid, min(chunk * R, 2 * R))
        print("checked: paste-back of {} masks == paste_masks_in_image restated".format(min(chunk * R, 2 * R)), file=sys.stderr)
    ms = timed(paste, reps=3, warm=1)
    inbox = float(((bx[..., 2] - bx[..., 0]) * (bx[..., 3] - bx[..., 1])).mean().item()) / (H * 1333)
    report("paste_masks", ms, n * R * H * 1333 + n * R * 784 * 4, "mean box area {:.1%} of the image".format(inbox))
    if args.old:
        os.environ["CM2_PASTE_VARIANT"] = "0"
        report("paste_masks [v0: memset + window kernel]", timed(paste, reps=3, warm=1), n * R * H * 1333 + n * R * 784 * 4)
        os.environ["CM2_PASTE_VARIANT"] = "1"
    if args.out:
        with open(args.out, "w") as f:
            for r in results:
                f.write(json.dumps(r) + "\n")
    eng.release()
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--cand", type=int, default=1000)
    ap.add_argument("--rois", type=int, default=100)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-check", action="store_true", help="skip the oracle comparison in front of every timing")
    ap.add_argument("--old", action="store_true", help="also time the previous implementation of each kernel (CM2_*_VARIANT=0)")
    args = ap.parse_args()
    run(batch=args.batch, cand=args.cand, rois=args.rois, check=not args.no_check, old=args.old, out=args.out)


if __name__ == "__main__":
    main()
