#!/usr/bin/env python
"""Microbenchmark of cm2_keypoints_decode: 16 images x 50 ROI slots x 17 keypoints, boxes 20..400 px (log-uniform).
``python tools/micro_kp.py [--iters N]``; CM2_KP_VARIANT=0 selects the flat per-pixel variant."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from centermask2_b200 import lib
from tests.helpers import pack_lowres

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
g = torch.Generator().manual_seed(1)
n, r_cap, k, res = 16, 50, 17, 14
low = pack_lowres(torch.randn(n * r_cap, k, 2 * res, 2 * res, generator=g) * 2.5).cuda()
xy = torch.rand(n * r_cap, 2, generator=g) * 600
wh = torch.exp(torch.rand(n * r_cap, 2, generator=g) * 3.0 + 3.0)            # 20 .. 400 px
boxes = torch.cat([xy, xy + wh], 1).view(n, r_cap, 4).cuda()
cnt = torch.full((n,), r_cap, dtype=torch.int32).cuda()
out = torch.zeros(n, r_cap, k, 4).cuda()
for _ in range(3):
    lib.keypoints_decode(low, boxes, cnt, n, r_cap, res, k, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    lib.keypoints_decode(low, boxes, cnt, n, r_cap, res, k, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
pix = float((wh.clamp(min=1).ceil().prod(1)).sum()) * k
print(json.dumps({"kernel": "keypoints_decode", "variant": os.environ.get("CM2_KP_VARIANT", "1"), "ms": ms, "roi_slots": n * r_cap,
                  "keypoints": k, "resized_pixels": pix, "Gpix_per_s": pix / ms / 1e6,
                  "note": "16 images x 50 ROIs, boxes 20..400 px; fp32-issue bound"}))
