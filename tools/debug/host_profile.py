#!/usr/bin/env python
"""Where the host time of the end-to-end loop goes: cProfile over GeneralizedRCNN.inference_records (bench.py's e2e loop)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                             # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    import centermask2_b200 as cm
    from centermask2_b200.synth import synthetic_state_dict
    cfg = bench.make_cfg("bf16")
    model = cm.build_model(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=bench.WEIGHT_SEED))
    host_inputs = bench.make_images(16, 0, pinned=True)
    bench.calibrate_on_gpu(model, cfg, host_inputs)

    def run(k):
        for res in model.inference_records((host_inputs for _ in range(k))):
            pass
    run(8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    print("inference_records: {:.2f} ms / step (wall)".format((time.perf_counter() - t0) / steps * 1e3))
    pr = cProfile.Profile()
    pr.enable()
    run(steps)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(40)
    st.sort_stats("tottime").print_stats(22)
    # device-side split of the same loop (CUPTI): which kernels the result encoding adds to the step
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(6)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))


if __name__ == "__main__":
    main()
