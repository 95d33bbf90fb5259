#!/usr/bin/env python
"""ROIAlign-only slice of tools/micro_post.py (same boxes, same features) for tuning sweeps:

    python tools/micro_roi.py [--variants 3,2,1] [--sort]

Prints ms / GB/s over the algorithmic bytes per CM2_ROIALIGN_VARIANT (every variant twice, ABAB, after a clock warm-up) and
the deviation of every variant from variant 0.
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from micro_post import H, W, LEVELS, timed                          # noqa: E402
from centermask2_b200 import lib                                    # noqa: E402
from centermask2_b200.config import get_cfg                        # noqa: E402
from centermask2_b200.engine import Engine                         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--rois", type=int, default=100)
    ap.add_argument("--variants", default="2,1")
    ap.add_argument("--sort", action="store_true", help="boxes of the whole batch in descending area order (tail experiment)")
    ap.add_argument("--sides", default="", help="lo,hi: box side range in image pixels (sqrt of the area; default 32 .. "
                    "sqrt(0.9 H W)) -- which box sizes a variant is good at")
    ap.add_argument("--split", default="", help="experiment: comma list of k -- per image the k largest boxes go to one launch "
                    "and the other R - k to a second launch on a side stream, both inside the timed graph; every pair of "
                    "variants (large-box launch | small-box launch) out of 2 and 4 is timed (do the two kernels use different units?)")
    ap.add_argument("--check-only", action="store_true", help="one launch per variant, deviation only (for compute-sanitizer)")
    ap.add_argument("--mma-configs", default="", help="variant 4 only: comma list of stages:split settings "
                    "(CM2_ROIALIGN_STAGES / _SPLIT), e.g. 4:1,3:1,4:0")
    args = ap.parse_args()
    n, R, dev, c = args.batch, args.rois, "cuda", 256
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.FCOS.POST_NMS_TOPK_TEST", R, "MODEL.B200.PRECISION", "bf16"])
    eng = Engine(cfg, "bf16", dev)
    g = torch.Generator(device=dev).manual_seed(5)
    feats = [eng.fmap("mf{}".format(i), n, h, w, c) for i, (h, w, _) in enumerate(LEVELS[:3])]
    for f in feats:
        f.view.copy_(torch.randn(f.view.shape, device=dev, generator=g).to(torch.bfloat16))
    a_lo, a_hi = 32.0 * 32.0, 0.9 * H * W
    if args.sides:
        a_lo, a_hi = [float(t) ** 2 for t in args.sides.split(",")]
    area = torch.exp(torch.rand((n, R), device=dev, generator=g) * (math.log(a_hi) - math.log(a_lo)) + math.log(a_lo))
    ar = torch.exp((torch.rand((n, R), device=dev, generator=g) - 0.5) * 1.4)
    bw, bh = torch.sqrt(area * ar).clamp(max=W - 1.0), torch.sqrt(area / ar).clamp(max=H - 1.0)
    x0 = torch.rand((n, R), device=dev, generator=g) * (W - bw)
    y0 = torch.rand((n, R), device=dev, generator=g) * (H - bh)
    boxes = torch.stack([x0, y0, x0 + bw, y0 + bh], dim=2).contiguous()
    if args.sort:       # experiment only: slots of all images in descending area order (the image index of a slot stays)
        order = torch.argsort((bw * bh).flatten(), descending=True)
        boxes = boxes.view(-1, 4)[order].view(n, R, 4).contiguous()
    counts = torch.full((n,), R, dtype=torch.int32, device=dev)
    img_area = torch.full((n,), float(H * W), device=dev)
    roi = eng.fmap("mroi", n * R, 14, 14, c)
    lvl = torch.zeros((n * R,), dtype=torch.int32, device=dev)

    def roialign():
        lib.roialign_fpn([f.view for f in feats], [8, 16, 32], boxes, counts, n, R, img_area, 0, 0, roi.view, lvl)
    alg = n * (R * c * 196 * 2 + c * 22050 * 2)
    os.environ["CM2_ROIALIGN_VARIANT"] = "0"
    roialign()
    torch.cuda.synchronize()
    ref = roi.view.float().clone()
    if args.check_only:
        for v in [int(t) for t in args.variants.split(",")]:
            os.environ["CM2_ROIALIGN_VARIANT"] = str(v)
            roi.view.zero_()
            roialign()
            torch.cuda.synchronize()
            print("variant {}: max|diff vs v0| {:.4g}".format(v, (roi.view.float() - ref).abs().max().item()), flush=True)
        return
    for _ in range(300):                  # bring the clocks up before anything is timed
        roialign()
    torch.cuda.synchronize()
    if args.split:
        order = torch.argsort((bw * bh), dim=1, descending=True)
        sorted_boxes = torch.gather(boxes, 1, order[:, :, None].expand(-1, -1, 4)).contiguous()
        side = torch.cuda.Stream()
        ev0, ev1 = torch.cuda.Event(), torch.cuda.Event()
        for k in [int(t) for t in args.split.split(",")]:
            parts = []
            for lo, hi, tag in ((0, k, "a"), (k, R, "b")):
                m = hi - lo
                parts.append((sorted_boxes[:, lo:hi].contiguous(), torch.full((n,), m, dtype=torch.int32, device=dev), m,
                              eng.fmap("mroi_{}{}".format(tag, k), n * m, 14, 14, c), torch.zeros((n * m,), dtype=torch.int32, device=dev),
                              torch.empty((n * m + 1024,), dtype=torch.int32, device=dev)))

            def launch(part, variant):
                bx, cnt, m, out, lv, ws = part
                os.environ["CM2_ROIALIGN_VARIANT"] = str(variant)
                lib.roialign_fpn([f.view for f in feats], [8, 16, 32], bx, cnt, n, m, img_area, 0, 0, out.view, lv, ws)

            for va, vb in ((2, 2), (4, 2), (4, 4), (2, 4)):
                def both():
                    main = torch.cuda.current_stream()
                    ev0.record(main)
                    side.wait_event(ev0)
                    launch(parts[0], va)
                    with torch.cuda.stream(side):
                        launch(parts[1], vb)
                        ev1.record(side)
                    main.wait_event(ev1)

                def serial():
                    launch(parts[0], va)
                    launch(parts[1], vb)
                ms_c, ms_s = timed(both), timed(serial)
                print("split {:3d} largest | {:3d} others: variants {} | {}: concurrent {:.4f} ms   back to back {:.4f} ms".format(
                    k, R - k, va, vb, ms_c, ms_s), flush=True)
        return

    def diagnose(got):
        """Where a variant leaves variant 0: worst ROI slot, its box / level, the bins and channels that differ."""
        diff = (got - ref).abs()
        per_slot = diff.flatten(1).max(1)[0]
        bad = (per_slot > 0.1).nonzero().flatten()
        print("  slots off by > 0.1: {} of {}   (first: {})".format(bad.numel(), per_slot.numel(), bad[:8].tolist()))
        if bad.numel() == 0:
            return
        sl = int(per_slot.argmax())
        b = boxes.view(-1, 4)[sl].tolist()
        print("  worst slot {} box {} level {} size at level {:.1f} x {:.1f}".format(
            sl, [round(t, 1) for t in b], int(lvl[sl]), (b[2] - b[0]) / (8 << int(lvl[sl])), (b[3] - b[1]) / (8 << int(lvl[sl]))))
        d = diff[sl]                                       # [14, 14, c] view of the slot
        print("  max diff per bin row   :", [round(t, 2) for t in d.flatten(1).max(1)[0].tolist()])
        print("  max diff per bin column:", [round(t, 2) for t in d.transpose(0, 1).flatten(1).max(1)[0].tolist()])
        ch = d.flatten(0, 1).max(0)[0]
        print("  channels off by > 0.1: {} of {} (first: {})".format(int((ch > 0.1).sum()), ch.numel(), (ch > 0.1).nonzero().flatten()[:16].tolist()))
        print("  got[0, :, 0:4]", got[sl][0, :4, 0:4].flatten().tolist())
        print("  ref[0, :, 0:4]", ref[sl][0, :4, 0:4].flatten().tolist())

    mma_cfgs = [t.split(":") for t in args.mma_configs.split(",") if t] or [None]
    for rep in range(2):                  # every configuration twice (ABAB) so that drift shows
        for v in [int(t) for t in args.variants.split(",")]:
            os.environ["CM2_ROIALIGN_VARIANT"] = str(v)
            for mc in (mma_cfgs if v == 4 else [None]):
                k = ""
                if mc is not None:
                    os.environ["CM2_ROIALIGN_STAGES"], os.environ["CM2_ROIALIGN_SPLIT"] = mc
                    k = " (stages {} split {})".format(*mc)
                roi.view.zero_()
                ms = timed(roialign)
                got = roi.view.float()
                d = (got - ref).abs().max().item()
                print("variant {}{}: {:.4f} ms  {:7.1f} GB/s  {:.3f} of HBM peak   max|diff vs v0| {:.4g}  mean|diff| {:.3g}".format(
                    v, k, ms, alg / ms / 1e6, alg / ms / 1e6 / hbm, d, (got - ref).abs().mean().item()), flush=True)
                if rep == 0 and d > 0.1:
                    diagnose(got)


if __name__ == "__main__":
    main()
