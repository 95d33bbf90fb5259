#!/usr/bin/env python
"""bench.py -- img/s of the CenterMask2 V-39-eSE-FPN inference path at 800x1333 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16|fp32] [--impl reference]

One process per GPU (torchrun for N > 1); every rank runs the same per-GPU batch of synthetic images
(weak scaling, no data-path collective; one final all_gather of the compact result records).
A *step* is one pass of the hot path over one batch: normalise+pad -> VoVNet-eSE + FPN -> FCOS head ->
decode / top-k / NMS -> ROIAlign -> SAG-Mask + MaskIoU -> box rescale + mask paste-back.

JSON line keys: see the task contract; `value` = device-resident inputs, `e2e` = through
GeneralizedRCNN.forward with pinned host images (H2D inside) and a D2H read of the compact results.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 800, 1333
CFG_FILE = "centermask_V_39_eSE_FPN.yaml"
WEIGHT_SEED, IMAGE_SEED = 101, 202
CAND_TARGET = 800            # candidates / level / image above the 0.05 threshold


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """`nvidia-smi -lms` running for the whole process (its start-up takes longer than a short timed region);
    only samples whose arrival time falls inside a `window()` -- the timed regions -- are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        self.index, self.rows, self.proc, self.windows = index, [], None, []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(int(period_ms))],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    @contextlib.contextmanager
    def window(self):
        t0 = time.perf_counter()
        yield
        self.windows.append((t0, time.perf_counter()))

    def close(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not any(a <= t <= b + 0.02 for a, b in self.windows):
                continue
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def make_cfg(precision):
    from centermask2_b200.config import get_cfg
    cfg = get_cfg(CFG_FILE)
    cfg.merge_from_list(["MODEL.B200.PRECISION", precision])
    return cfg


def make_images(batch, rank, pinned):
    from centermask2_b200.synth import synthetic_images
    base = synthetic_images(min(batch, 4), H, W, seed=IMAGE_SEED + 1000 * rank)
    out = []
    for i in range(batch):
        img = base[i % len(base)]["image"].to(torch.uint8)          # what a data loader hands over: uint8 BGR CHW
        if i >= len(base):
            img = torch.roll(img, shifts=17 * i, dims=2)
        out.append({"image": img.pin_memory() if pinned else img, "height": H, "width": W})
    return out


def calibrate_on_gpu(model, cfg, inputs):
    """Pick cls_logits.bias so ~CAND_TARGET candidates/level survive the threshold (stock init gives none)."""
    from centermask2_b200 import runtime
    from centermask2_b200.synth import calibrate_cls_bias
    key = "proposal_generator.fcos_head.cls_logits.bias"
    sd = model.state_dict()
    sd[key] = torch.zeros_like(sd[key])
    model.load_state_dict(sd)
    eng = runtime.engine_for(cfg)
    x, _ = eng.preprocess([b["image"].to(eng.device) for b in inputs[:2]])
    feats = model.backbone.forward_fmap(x)
    fcos = model.proposal_generator
    e, P = fcos._pack()
    head = e.run_fcos_head([feats[f] for f in fcos.in_features], P)
    logits = [lg.view.float().permute(0, 3, 1, 2) for lg, _ in head]
    b = calibrate_cls_bias(logits, CAND_TARGET)
    sd[key] = torch.full_like(sd[key], b)
    model.load_state_dict(sd)
    return b


def compact_results(results, r_cap):
    """Fixed-size result record per image (centermask2_b200/parallel.py; SURVEY 5 'Distributed communication backend')."""
    from centermask2_b200 import parallel
    return parallel.pack_records([r["instances"] for r in results], r_cap)


def make_device_step(model, cfg, dev_images, sizes_out, graph=True):
    """Hot path with inputs resident in HBM; results stay on the device in fixed-size buffers (no host sync).
    Returns a callable running one step: the launch plan replayed as a CUDA graph (``graph=True``, the product
    path of ``GeneralizedRCNN.inference``) or launched eagerly through the C ABI."""
    from centermask2_b200 import runtime
    eng = runtime.engine_for(cfg)
    fcos, roi = model.proposal_generator, model.roi_heads
    n = len(dev_images)
    out_sizes = [tuple(sizes_out)] * n

    def plan():
        x, sizes = eng.preprocess(dev_images, 32)
        feats = model.backbone.forward_fmap(x)
        det = fcos.detect([feats[f] for f in fcos.in_features])
        probs, mask_scores = roi.run([feats[f] for f in roi.in_features], det, sizes)
        boxes, valid = eng.rescale_boxes(det["boxes"], sizes, out_sizes)
        r_cap = det["boxes"].shape[1]
        masks = eng.buffer("bench_masks", (n, r_cap, sizes_out[0], sizes_out[1]), torch.uint8, zero=False)
        eng.paste_batch(probs, boxes, valid, out_sizes, masks=masks)
        return det, mask_scores

    if graph and eng.use_graphs:
        return lambda: eng.graphed(("bench_step", model.graph_tokens(), n, tuple(sizes_out)), plan, keep=model.packed_refs())
    return plan


def device_step(model, cfg, dev_images, sizes_out):
    """One eager step (used by the instrumented per-launch timing pass)."""
    return make_device_step(model, cfg, dev_images, sizes_out, graph=False)()


def conv_time_per_step(model, cfg, dev_images, steps, layers=None):
    """Sum of the device time of every convolution launch in one step (CUDA events around each launch, the launches of a
    step enqueued behind a device-side delay so that the events measure kernel time, not host launch latency)."""
    from centermask2_b200 import runtime
    eng = runtime.engine_for(cfg)
    events = []
    orig = eng.conv

    def timed(name, srcs, w, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(name, srcs, w, *a, **k)
        e1.record()
        events.append((e0, e1))
        if layers is not None:
            x = srcs[0]
            ho = (x.h + 2 * w.pad - w.k) // w.stride + 1
            wo = (x.w + 2 * w.pad - w.k) // w.stride + 1
            layers.append((name, 2.0 * x.n * ho * wo * sum(w.src_c) * w.k * w.k * w.cout / 1e9, (e0, e1)))
        return r
    orig_seg = eng.conv_seg

    def timed_seg(name, x, w, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_seg(name, x, w, *a, **k)
        e1.record()
        events.append((e0, e1))
        if layers is not None:
            px = sum(n * h * ww for _, n, h, ww in x.segs)
            layers.append((name, 2.0 * px * sum(w.src_c) * w.k * w.k * w.cout / 1e9, (e0, e1)))
        return r
    eng.conv = timed
    eng.conv_seg = timed_seg
    try:
        for _ in range(steps):
            # keep the GPU behind the host while the step is enqueued: with an empty queue every (event, launch) pair
            # would also time the few microseconds of host-side argument marshalling between the two calls
            torch.cuda._sleep(40_000_000)                # ~20 ms of device-side spinning
            device_step(model, cfg, dev_images, (H, W))
        torch.cuda.synchronize()
    finally:
        eng.conv = orig
        eng.conv_seg = orig_seg
    total = sum(a.elapsed_time(b) for a, b in events)
    return total / steps, len(events) // steps


def cpu_baseline(steps, warmup, images_per_step=1):
    """The oracle (fp32 restatement executing the reference's ATen/torchvision CPU ops) on the host cores."""
    from centermask2_b200.config import get_cfg
    from centermask2_b200.synth import synthetic_state_dict, synthetic_images, calibrate_cls_bias
    from oracle import restate
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = get_cfg(CFG_FILE)
    sd = synthetic_state_dict(cfg, seed=WEIGHT_SEED)
    inputs = synthetic_images(images_per_step, H, W, seed=IMAGE_SEED)
    key = "proposal_generator.fcos_head.cls_logits.bias"
    sd[key] = torch.zeros_like(sd[key])
    tr = {}
    with contextlib.redirect_stdout(io.StringIO()):
        restate.run_model(inputs, sd, cfg, postprocess=False, trace=tr)
        sd[key] = torch.full_like(sd[key], calibrate_cls_bias(tr["logits"], CAND_TARGET))
        for _ in range(warmup):
            restate.run_model(inputs, sd, cfg, postprocess=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            restate.run_model(inputs, sd, cfg, postprocess=True)
        dt = time.perf_counter() - t0
    return {"value": images_per_step * steps / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": "{} step(s) x {} image(s) 800x1333 V-39-eSE fp32, oracle/restate.py (torch {} CPU ops)".format(
                steps, images_per_step, torch.__version__), "ms_per_step": dt / steps * 1e3}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("CM2_PRECISION", "bf16"), choices=["bf16", "fp32", "fp32_simt"],
                    help="bf16: bf16 activations, tcgen05 convolutions; fp32: fp32 activations, split-precision (f16 hi/lo) "
                         "tcgen05 convolutions with fp32-grade accuracy; fp32_simt: fp32 on CUDA cores")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-ms", type=int, default=int(os.environ.get("CM2_CLOCK_MS", "20")),
                    help="nvidia-smi sampling period for the clocks line (the sampler runs during the timed regions)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--profile-step", action="store_true",
                    help="warm up, then run exactly one eager step between cudaProfilerStart/Stop and exit (for "
                         "`ncu --profile-from-start off`); prints nothing that counts as a bench value")
    ap.add_argument("--layers", default=None, help="write a per-conv-layer timing table (one step) to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return 0
        base = cpu_baseline(max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": "img/s at 800x1333 V-39-eSE", "value": base["value"], "unit": "img/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "CenterMask2 V-39-eSE-FPN 800x1333 (BASELINE configs[2]); each step = 1 image on the host cores"},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import centermask2_b200 as cm
    from centermask2_b200 import lib, runtime
    from centermask2_b200.arch import conv_gflop_per_image
    from centermask2_b200.synth import synthetic_state_dict

    clk = ClockSampler(local, args.clock_ms)
    cfg = make_cfg(args.precision)
    model = cm.build_model(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=WEIGHT_SEED))
    host_inputs = make_images(args.batch, rank, pinned=True)
    bias = calibrate_on_gpu(model, cfg, host_inputs)
    dev_images = [b["image"].cuda() for b in host_inputs]
    eng = runtime.engine_for(cfg)
    if args.no_graph:
        eng.use_graphs = False
    r_cap = cfg.MODEL.FCOS.POST_NMS_TOPK_TEST

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.profile_step:
        eager = make_device_step(model, cfg, dev_images, (H, W), graph=False)
        for _ in range(warmup):
            eager()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        eager()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        clk.close()
        return 0

    # ---- device-resident throughput ("value")
    step = make_device_step(model, cfg, dev_images, (H, W), graph=not args.no_graph)
    for _ in range(warmup):
        det, _ms = step()
    barrier()
    dets_per_image = det["count"].float().mean().item()
    cand = det["cand_count"].float().mean().item()
    l0 = lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clk.window():
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count - l0
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = args.batch * world / (ms_step / 1e3)

    # ---- end to end through the public API (pinned host images in, compact results out).
    # (a) GeneralizedRCNN.forward(batched_inputs) per step, synchronous: H2D, compute and read-back in sequence;
    # (b) GeneralizedRCNN.inference_stream(batches): the same work per step, the next step's H2D copy issued on a copy
    #     stream while the current step computes.  (b) is the reported e2e value; (a) is kept beside it.
    def e2e_step():
        out = model(host_inputs)
        rec = compact_results(out, r_cap)
        return rec.cpu()

    side = torch.cuda.Stream()
    h_rec = torch.empty((args.batch, r_cap, 8), dtype=torch.float32, pin_memory=True)

    def e2e_stream(k):
        # the record of every step is packed and copied to the host on a side stream that waits for that step only
        # (on the compute stream the copy would queue behind the steps already enqueued ahead)
        for out, done in model.inference_stream((host_inputs for _ in range(k)), with_event=True):
            side.wait_event(done)
            with torch.cuda.stream(side):
                rec = compact_results(out, r_cap)
                h_rec.copy_(rec, non_blocking=True)
            side.synchronize()
        return h_rec.clone()
    for _ in range(2):
        e2e_step()
    e2e_stream(3)
    barrier()
    with clk.window():
        e0.record()
        for _ in range(args.steps):
            rec = e2e_step()
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync_ms = t.item() / args.steps
    with clk.window():
        e0.record()
        rec = e2e_stream(args.steps)
        e1.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item() / args.steps
    h2d = sum(b["image"].numel() * b["image"].element_size() for b in host_inputs)
    d2h = rec.numel() * rec.element_size()

    # ---- final result gather (the only collective; not on the hot path)
    from centermask2_b200 import parallel
    allrec = parallel.gather_records(rec.cuda(), args.batch * world)       # [batch*world, r_cap, 8] on every rank
    assert allrec.shape[0] == args.batch * world

    # ---- roofline of the dominant kernel family (convolutions)
    conv_ms, conv_launches = conv_time_per_step(model, cfg, dev_images, max(2, min(args.steps, 5)))
    if args.layers and rank == 0:
        rows = []
        conv_time_per_step(model, cfg, dev_images, 1, layers=rows)
        with open(args.layers, "w") as f:
            f.write("# batch {} precision {}\n# name gflop ms tflops\n".format(args.batch, args.precision))
            for nme, gf, (a, b) in rows:
                ms_l = a.elapsed_time(b)
                f.write("{:24s} {:10.3f} {:9.4f} {:9.1f}\n".format(nme, gf, ms_l, gf / ms_l if ms_l > 0 else 0.0))
    gflop_img = conv_gflop_per_image(cfg, 800, 1344, r_cap)          # algorithmic FLOPs, R = slots computed
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")      # written by tools/ncu_step_summary.py from an ncu capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("images_per_gpu") == args.batch and tj.get("precision") == args.precision:
            traffic = tj.get("conv_dram_bytes_per_step")
    clk.close()
    hbm, tf_burst, tf_sus, src = peaks()
    achieved = gflop_img * args.batch / conv_ms                       # GFLOP / ms = TFLOP/s
    peak = tf_sus
    line = {
        "metric": "img/s at 800x1333 V-39-eSE", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "CenterMask2 V-39-eSE-FPN, {} synthetic 800x1333 images per GPU per step (BASELINE configs[2], "
                               "image data-parallel), random-init weights, POST_NMS_TOPK 50".format(args.batch),
                   "images_per_gpu": args.batch, "precision": args.precision, "cls_bias": bias,
                   "detections_per_image": dets_per_image, "candidates_per_level": cand,
                   "l2": "activations per step ({} images) far exceed the 126 MB L2; no explicit flush".format(args.batch),
                   "cuda_graph": bool(eng.use_graphs and not args.no_graph), "parallelism": "dp{}".format(world)},
        "clocks": clk.summary(),
        "e2e": {"value": args.batch * world / (e2e_ms / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "forward_per_call": {"value": args.batch * world / (e2e_sync_ms / 1e3), "ms_per_step": e2e_sync_ms},
                "note": "GeneralizedRCNN.inference_stream(batches): every step copies its pinned uint8 host images to the "
                        "device (on a copy stream, overlapping the previous step), runs the step, returns Instances with "
                        "pasted bool masks (on device) and copies the compact result record to the host; "
                        "forward_per_call = the same through GeneralizedRCNN.forward(batched_inputs), nothing overlapped"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "conv (all launches of one step)", "launches_per_step": conv_launches,
                     "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / ms_step,
                     "gflop_per_image": gflop_img, "peak_source": src + " (bf16 sustained cuBLAS)",
                     "frac_of_burst": achieved / tf_burst},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(2, 1)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
