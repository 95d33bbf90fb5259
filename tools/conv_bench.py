"""Microbenchmark of the tensor-core convolution engine on the layer shapes of V-39 @ 800x1344.

    CM2_TC_VARIANT={0,1,2,3} python tools/conv_bench.py [--batch 8] [--only name]

Prints device time (CUDA events, mean of 5 launches after 2 warm-ups, inputs far larger than nothing
special: the same buffers are re-read, so small layers are L2-resident -- this is a kernel tuning tool,
not the bench) and effective TFLOP/s per layer.  Also the target of the `ncu --set full` captures.
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centermask2_b200 import lib, packing  # noqa: E402

BF = torch.bfloat16


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.batch
    r = 50 * n
    shapes = [
        # name, images, h, w, sources, cout, k, out_mode
        ("stem1_1x1_32_64", n, 400, 672, [32], 64, 1, 0),
        ("stem2_3x3_64", n, 400, 672, [64], 64, 3, 0),
        ("stem2_phase_out", n, 400, 672, [64], 64, 3, 2),
        ("stem3_s2_64_128", n, 400, 672, [64], 128, 3, 3),
        ("osa2_3x3_128", n, 200, 336, [128], 128, 3, 0),
        ("osa2_cat_768_256", n, 200, 336, [128] * 6, 256, 1, 0),
        ("osa3_3x3_160", n, 100, 168, [160], 160, 3, 0),
        ("osa3_cat_1056_512", n, 100, 168, [256] + [160] * 5, 512, 1, 0),
        ("osa4_3x3_192", n, 50, 84, [192], 192, 3, 0),
        ("osa5_3x3_224", n, 25, 42, [224], 224, 3, 0),
        ("fpn_inner3_512_256", n, 100, 168, [512], 256, 1, 0),
        ("fcos_tower_p3", n, 100, 168, [256], 256, 3, 0),
        ("fcos_tower_p4", n, 50, 84, [256], 256, 3, 0),
        ("fcos_logits_p3", n, 100, 168, [256], 80, 3, 0),
        ("fcos_regctr_p3", n, 100, 168, [256], 16, 3, 0),
        ("iou_fc1_12544_1024", 1, 1, r, [12544], 1024, 1, 0),
        ("mask_fcn", r, 14, 14, [256], 256, 3, 0),
        ("mask_deconv", r, 14, 14, [256], 1024, 1, 1),
    ]
    dev = "cuda"
    print("# variant {} batch {}".format(os.environ.get("CM2_TC_VARIANT", "0"), n))
    for name, nn, h, w, srcs, cout, k, out_mode in shapes:
        if args.only and args.only not in name:
            continue
        g = torch.Generator().manual_seed(1)
        cin = sum(srcs)
        stride, src_phase = 1, False
        if out_mode == 3:                                   # 3x3 / stride 2 reading phase planes (stem_3, P6, P7)
            stride, src_phase = 2, True
            bufs = [torch.zeros((4, nn, h // 2 + 2, w // 2 + 2, c), dtype=BF, device=dev) for c in srcs]
            for b in bufs:
                b[:, :, 1:-1, 1:-1].normal_()
        else:
            bufs = [torch.zeros((nn, h + 2, w + 2, c), dtype=BF, device=dev) for c in srcs]
            for b in bufs:
                b[:, 1:-1, 1:-1].normal_()
        wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
        cw = packing.ConvW(wt, srcs, stride, k // 2, torch.ones(cout), torch.zeros(cout), True, BF, dev, True)
        om = out_mode
        if out_mode == 0:
            out = torch.zeros((nn, h + 2, w + 2, cout), dtype=BF, device=dev)[:, 1:-1, 1:-1]
            views = [b[:, 1:-1, 1:-1] for b in bufs]
        elif out_mode == 1:
            out = torch.zeros((nn, 2 * h, 2 * w, cout // 4), dtype=BF, device=dev)
            views = [b[:, 1:-1, 1:-1] for b in bufs]
        elif out_mode == 2:
            out = torch.zeros((4, nn, h // 2 + 2, w // 2 + 2, cout), dtype=BF, device=dev)[0, :, 1:-1, 1:-1]
            views = [b[:, 1:-1, 1:-1] for b in bufs]
        else:
            om = 0
            out = torch.zeros((nn, h // 2 + 2, w // 2 + 2, cout), dtype=BF, device=dev)[:, 1:-1, 1:-1]
            views = [b[0, :, 1:-1, 1:-1] for b in bufs]

        def run():
            ok = lib.conv2d(views, cw.w_tc, out, cout, k, stride, k // 2, scale=cw.scale, shift=cw.shift, relu=True,
                            out_mode=om, engine=lib.ENGINE_TC, probe=True, src_phase=src_phase)
            assert ok, lib.last_error()
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        gflop = 2.0 * nn * (h // stride) * (w // stride) * cin * k * k * cout / 1e9
        print("{:22s} {:9.1f} GFLOP {:8.4f} ms {:8.1f} TFLOP/s".format(name, gflop, ms, gflop / ms))


if __name__ == "__main__":
    main()
