#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/micro_keypoints.txt
timeout 500 python -m pytest tests/test_gpu_keypoints.py tests/test_gpu_model.py -m gpu -q --tb=short -k "keypoint" -s 2>&1 | tail -40 | cut -c1-300 | tee gpurun_out/pytest_keypoints.log
for V in 1 0; do CM2_KP_VARIANT=$V timeout 200 python - <<'PY' 2>&1 | tail -2 | tee -a gpurun_out/micro_keypoints.txt
import torch, json
from centermask2_b200 import lib
from tests.helpers import pack_lowres
g = torch.Generator().manual_seed(1)
n, r_cap, k, res = 16, 50, 17, 14
low = pack_lowres(torch.randn(n * r_cap, k, 28, 28, generator=g) * 2.5).cuda()
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
for _ in range(20):
    lib.keypoints_decode(low, boxes, cnt, n, r_cap, res, k, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
pix = float((wh.clamp(min=1).ceil().prod(1)).sum()) * k
print(json.dumps({"kernel": "keypoints_decode", "ms": ms, "roi_slots": n * r_cap, "keypoints": k, "resized_pixels": pix,
                  "Gpix_per_s": pix / ms / 1e6, "variant": __import__("os").environ.get("CM2_KP_VARIANT"), "note": "16 images x 50 ROIs, boxes 20..400 px; fp32-issue bound"}))
PY
done
