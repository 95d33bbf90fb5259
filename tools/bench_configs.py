#!/usr/bin/env python
"""Device-resident throughput of the OTHER BASELINE configs (not the headline; bench.py measures configs[2]):

    python tools/bench_configs.py [--out profiles/rNN_bench_configs.jsonl] [--only lite,v99,v19_slim_dw]

  lite        configs[1]: CenterMask2-Lite V-19-eSE-FPN (upstream Lite recipe as cfg overrides), short side 512
              (512x853 -> 512x864), batch 8, bf16
  v99         configs[3]: V-99-eSE-FPN, batch 8, 800x1333, bf16 (dense-conv roofline stress)
  v19_slim_dw the depthwise body V-19-slim-dw-eSE, batch 16, 800x1333, bf16

Same step as bench.py (inputs resident in HBM, one CUDA-graph replay per step, CUDA events, 3 warm-ups); one JSON line
per config with img/s, ms/step and the achieved dense TFLOP/s over the algorithmic conv FLOPs of the whole step time
(NOT the conv-only roofline of bench.py: every kernel of the step is in the denominator)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                        # noqa: E402

CONFIGS = {
    "lite": ("lite", 8, 512, 853),
    "v99": (["MODEL.VOVNET.CONV_BODY", "V-99-eSE"], 8, 800, 1333),
    "v19_slim_dw": (["MODEL.VOVNET.CONV_BODY", "V-19-slim-dw-eSE"], 16, 800, 1333),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=",".join(CONFIGS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import centermask2_b200 as cm
    from centermask2_b200.arch import conv_gflop_per_image
    from centermask2_b200.config import get_cfg, lite_overrides
    from centermask2_b200.synth import synthetic_state_dict
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    lines = []
    for name in args.only.split(","):
        over, batch, h, w = CONFIGS[name]
        cfg = get_cfg(bench.CFG_FILE, lite_overrides() if over == "lite" else over)
        cfg.merge_from_list(["MODEL.B200.PRECISION", "bf16"])
        bench.H, bench.W = h, w                                    # bench.make_images reads the module globals
        model = cm.build_model(cfg)
        model.load_state_dict(synthetic_state_dict(cfg, seed=bench.WEIGHT_SEED))
        host = bench.make_images(batch, 0, pinned=False)
        bench.calibrate_on_gpu(model, cfg, host)
        dev = [b["image"].cuda() for b in host]
        step = bench.make_device_step(model, cfg, dev, (h, w), graph=True)
        for _ in range(3):
            det, _ms = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
        gflop = conv_gflop_per_image(cfg, hp, wp, cfg.MODEL.FCOS.POST_NMS_TOPK_TEST)
        line = {"config": name, "body": cfg.MODEL.VOVNET.CONV_BODY, "batch": batch, "image": [h, w], "dtype": "bf16",
                "img_per_s": round(batch / ms * 1e3, 1), "ms_per_step": round(ms, 3), "gflop_per_image": round(gflop, 1),
                "tflops_whole_step": round(gflop * batch / ms, 1),
                "frac_of_sustained_bf16": round(gflop * batch / ms / peaks["bf16_tflops_sustained"], 3),
                "detections_per_image": det["count"].float().mean().item(), "data": "synthetic", "cuda_graph": True}
        print(json.dumps(line))
        lines.append(line)
        del model, step, dev
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
