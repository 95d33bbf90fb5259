#!/usr/bin/env python
"""Kernel timeline of the device-resident step (CUPTI via torch.profiler; the image has no nsys).

    python tools/trace_step.py --batch 16 --steps 3 [--out gpurun_out/trace_b16.txt]

Prints, per kernel name, launches / step, total device time / step and share, plus the GPU idle time inside
the step (wall of the step on the device minus the union of kernel intervals) -- i.e. how launch-bound the
host side is.  Numbers under the profiler are diagnostic only, never bench values.
"""
import argparse
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--e2e", action="store_true", help="trace GeneralizedRCNN.inference_stream (host images in, records out) instead")
    ap.add_argument("--records", action="store_true", help="with --e2e: trace GeneralizedRCNN.inference_records (bench.py's e2e loop)")
    ap.add_argument("--gaps", type=float, default=0.0, help="also list device idle gaps longer than this many microseconds")
    args = ap.parse_args()
    import centermask2_b200 as cm
    from centermask2_b200.synth import synthetic_state_dict
    cfg = bench.make_cfg(args.precision)
    model = cm.build_model(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=bench.WEIGHT_SEED))
    host_inputs = bench.make_images(args.batch, 0, pinned=True)
    bench.calibrate_on_gpu(model, cfg, host_inputs)
    dev_images = [b["image"].cuda() for b in host_inputs]
    step = bench.make_device_step(model, cfg, dev_images, (bench.H, bench.W), graph=bool(args.graph))
    r_cap = cfg.MODEL.FCOS.POST_NMS_TOPK_TEST

    side = torch.cuda.Stream()

    def stream(k):
        if args.records:
            for res in model.inference_records((host_inputs for _ in range(k))):
                pass
            return
        for out, done in model.inference_stream((host_inputs for _ in range(k)), with_event=True):
            side.wait_event(done)
            with torch.cuda.stream(side):
                rec = bench.compact_results(out, r_cap).to("cpu", non_blocking=True)
            side.synchronize()
    for _ in range(3):
        step()
    if args.e2e:
        stream(8)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        if args.e2e:
            stream(args.steps)
        else:
            for _ in range(args.steps):
                step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = collections.defaultdict(lambda: [0, 0.0])
    ivals = []
    for e in evs:
        nme = re.sub(r"\(.*", "", e.name)
        nme = re.sub(r"<.*", "", nme)
        dur = e.time_range.end - e.time_range.start
        agg[nme][0] += 1
        agg[nme][1] += dur
        ivals.append((e.time_range.start, e.time_range.end, nme))
    ivals.sort()
    busy, cur_s, cur_e, last_name, gaps = 0.0, None, None, None, []
    for s, e, nme in ivals:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
                if args.gaps and s - cur_e > args.gaps:
                    gaps.append((s - cur_e, last_name, nme))
            cur_s, cur_e, last_name = s, e, nme
        else:
            if e > cur_e:
                cur_e, last_name = e, nme
    if cur_e is not None:
        busy += cur_e - cur_s
    wall = ivals[-1][1] - ivals[0][0] if ivals else 0.0
    lines = ["# batch {} steps {} graph {}".format(args.batch, args.steps, args.graph),
             "# device wall {:.3f} ms/step  busy {:.3f} ms/step  idle {:.1f} %".format(
                 wall / 1e3 / args.steps, busy / 1e3 / args.steps, 100.0 * (1 - busy / wall) if wall else 0.0),
             "# kernel  launches/step  ms/step  share_of_busy"]
    tot = sum(v for _, v in agg.values())
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("{:58s} {:7.1f} {:9.4f} {:6.1f}%".format(k[:58], n / args.steps, v / 1e3 / args.steps, 100.0 * v / tot))
    if args.gaps:
        lines.append("# idle gaps > {} us: {} gaps, {:.3f} ms/step".format(args.gaps, len(gaps), sum(g for g, _, _ in gaps) / 1e3 / args.steps))
        agg_g = collections.defaultdict(lambda: [0, 0.0])
        for g, a, b in gaps:
            agg_g[(a[:40], b[:40])][0] += 1
            agg_g[(a[:40], b[:40])][1] += g
        for (a, b), (n, v) in sorted(agg_g.items(), key=lambda kv: -kv[1][1])[:15]:
            lines.append("#   {:6.1f} us x {:3d}  after {}  before {}".format(v / n, n, a, b))
    txt = "\n".join(lines)
    print(txt)
    if args.out:
        with open(args.out, "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
