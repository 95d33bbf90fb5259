#!/usr/bin/env python
"""Does alternating between several captured graphs of the same step cost launch latency that replaying ONE graph does not?
(inference_records replays one graph per pipeline slot.)  Times 30 replays of one executable graph against 30 replays cycling
over three, optionally with cudaGraphUpload of the next executable on a side stream."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                             # noqa: E402


def main():
    import centermask2_b200 as cm
    from centermask2_b200 import runtime
    from centermask2_b200.synth import synthetic_state_dict
    cfg = bench.make_cfg("bf16")
    model = cm.build_model(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=bench.WEIGHT_SEED))
    host_inputs = bench.make_images(16, 0, pinned=True)
    bench.calibrate_on_gpu(model, cfg, host_inputs)
    dev_images = [b["image"].cuda() for b in host_inputs]
    eng = runtime.engine_for(cfg)
    plan = bench.make_device_step(model, cfg, dev_images, (bench.H, bench.W), graph=False)
    keys = [("switch", i) for i in range(3)]
    for k in keys:
        for _ in range(3):
            eng.graphed(k, plan, keep=model.packed_refs())
    torch.cuda.synchronize()
    execs = [eng._graphs[k][0] for k in keys]
    rt = ctypes.CDLL("libcudart.so.12")
    side = torch.cuda.Stream()

    def timed(order, upload):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in order[:3]:
            execs[i].replay()
        torch.cuda.synchronize()
        e0.record()
        for j, i in enumerate(order):
            if upload and j + 1 < len(order):
                h = execs[order[j + 1]].raw_cuda_graph_exec()
                rc = rt.cudaGraphUpload(ctypes.c_void_p(h), ctypes.c_void_p(side.cuda_stream))
                assert rc == 0, rc
            execs[i].replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / len(order)

    back = torch.cuda.Stream()
    pin = torch.empty(1 << 14, dtype=torch.float32, pin_memory=True)
    dev = torch.zeros(1 << 14, device="cuda")

    def timed_events(mode):
        """one graph, 30 replays; mode 0: nothing between the replays, 1: an event record, 2: event record + a D2H copy on another
        stream that waits for it, 3: a tiny eager kernel between the replays"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            execs[0].replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(30):
            execs[0].replay()
            if mode in (1, 2):
                ev = torch.cuda.Event()
                ev.record()
                if mode == 2:
                    back.wait_event(ev)
                    with torch.cuda.stream(back):
                        pin.copy_(dev, non_blocking=True)
            elif mode == 3:
                dev.add_(1.0)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 30

    for rep in range(3):
        print("between replays: nothing {:.3f}  event {:.3f}  event + D2H on a side stream {:.3f}  eager kernel {:.3f}  ms / step".format(
            timed_events(0), timed_events(1), timed_events(2), timed_events(3)))
    same = [0] * 30
    cyc = [i % 3 for i in range(30)]
    for rep in range(2):
        print("one graph      : {:.3f} ms / step".format(timed(same, False)))
        print("three graphs   : {:.3f} ms / step".format(timed(cyc, False)))
        try:
            print("three + upload : {:.3f} ms / step".format(timed(cyc, True)))
        except Exception as e:                           # noqa: BLE001
            print("upload failed:", e)


if __name__ == "__main__":
    main()
