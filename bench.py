#!/usr/bin/env python
"""bench.py -- img/s of the CenterMask2 inference path (BASELINE.json metric: img/s at 800x1333 V-39-eSE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config v39|lite|v99|post] [--scaling weak|strong]
                    [--batch B] [--precision bf16|fp32|fp32_simt] [--impl reference]

One process per GPU (torchrun for N > 1), image data-parallel, no data-path collective; one final all_gather of the
result records.  A *step* is one pass of the hot path over one batch: normalise+pad -> VoVNet-eSE + FPN -> FCOS head ->
decode / top-k / NMS -> ROIAlign -> SAG-Mask + MaskIoU -> box rescale + mask paste-back.

  --config   v39  (default) BASELINE configs[2]: V-39-eSE-FPN, 800x1333, 16 images per step
             lite BASELINE configs[1]: CenterMask2-Lite V-19-eSE-FPN, 512x853, 8 images per step
             v99  BASELINE configs[3]: V-99-eSE-FPN, 800x1333, 8 images per step
             post BASELINE configs[4]: FCOS post-process + ROI-stage kernels alone, batch 32 (tools/micro_post.py)
  --scaling  weak: --batch images per GPU per step (default).  strong: --batch images per step in TOTAL, sharded over the
             ranks (configs[2]: 16 -> 16 / 8 / 4 / 2 images per GPU at 1 / 2 / 4 / 8 GPUs).

JSON line: `value` = device-resident inputs; `e2e` = GeneralizedRCNN.inference_records(batches): pinned host images in
(H2D inside the timed region), the whole result on the host -- detection records plus the masks as COCO run lengths.
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
SOAK_SECONDS = 3.0           # graph replays in front of every timed region: the region then runs at sustained clocks

# name -> (cfg overrides | "lite", images per step, image height, width, BASELINE.json configs[] index)
CONFIGS = {
    "v39": ([], 16, 800, 1333, 2),
    "lite": ("lite", 8, 512, 853, 1),
    "v99": (["MODEL.VOVNET.CONV_BODY", "V-99-eSE"], 8, 800, 1333, 3),
}
METRIC = {"v39": "img/s at 800x1333 V-39-eSE", "lite": "img/s at 512x853 V-19-eSE Lite", "v99": "img/s at 800x1333 V-99-eSE",
          "post": "img/s through FCOS post-process + ROI-stage kernels (BASELINE configs[4])"}


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

    def __init__(self, index, period_ms=20, enabled=True):
        self.index, self.rows, self.proc, self.windows = index, [], None, []
        if not enabled:
            return
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
                "samples": len(sm), "sampled_gpu": self.index}


def pin_rank_to_cores(local, world):
    """One process per GPU on one host: give every rank its own slice of the host cores (8 ranks that all float over the
    same cores -- each with framework thread pools as wide as the machine -- cost the N = 8 end-to-end run 30 %)."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return None
    per = max(1, len(cpus) // max(1, world))
    mine = cpus[local * per:(local + 1) * per] or cpus
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return None
    torch.set_num_threads(max(1, min(len(mine), 4)))
    return len(mine)


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def make_cfg(precision, config="v39"):
    from centermask2_b200.config import get_cfg, lite_overrides
    over = CONFIGS[config][0]
    cfg = get_cfg(CFG_FILE, lite_overrides() if over == "lite" else list(over))
    cfg.merge_from_list(["MODEL.B200.PRECISION", precision])
    return cfg


def make_images(batch, rank, pinned, h=None, w=None):
    from centermask2_b200.synth import synthetic_images
    h, w = h or H, w or W
    base = synthetic_images(min(batch, 4), h, w, seed=IMAGE_SEED + 1000 * rank)
    out = []
    for i in range(batch):
        img = base[i % len(base)]["image"].to(torch.uint8)          # what a data loader hands over: uint8 BGR CHW
        if i >= len(base):
            img = torch.roll(img, shifts=17 * i, dims=2)
        out.append({"image": img.pin_memory() if pinned else img, "height": h, "width": w})
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
    """Fixed-size result record per image from ``Instances`` (centermask2_b200/parallel.py)."""
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
        boxes, valid = eng.rescale_boxes(det["boxes"], sizes, out_sizes, det["count"])
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


def _numel_bytes(t):
    return t.numel() * t.element_size()


def instrumented_step(model, cfg, dev_images, sizes_out, steps, layers=None):
    """One eager pass per step with CUDA events around every convolution launch and every bandwidth-bound kernel family
    (the launches of a step are enqueued behind a device-side delay so that the events measure kernel time, not host
    launch latency).  Returns (conv ms / step, conv launches / step, {family: (ms / step, algorithmic bytes / step, launches)})."""
    from centermask2_b200 import lib, runtime
    from centermask2_b200.engine import PhaseMap
    eng = runtime.engine_for(cfg)
    conv_events, fam = [], {}

    def ev_pair():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    orig_conv, orig_seg, orig_stem = eng.conv, eng.conv_seg, eng.stem1_fused

    def timed_stem(name, raw, P):
        # the fused stem (normalise + pad + stem_1, csrc/stem.cu) is a convolution launch: its time counts as conv time
        if "stem1_fused" not in P:
            return orig_stem(name, raw, P)
        e0, e1 = ev_pair()
        e0.record()
        r = orig_stem(name, raw, P)
        e1.record()
        conv_events.append((e0, e1))
        if layers is not None:
            layers.append((name, 2.0 * raw.n * (raw.hp // 2) * (raw.wp // 2) * 27 * 64 / 1e9, (e0, e1)))
        return r

    def timed_conv(name, srcs, w, *a, **k):
        e0, e1 = ev_pair()
        e0.record()
        r = orig_conv(name, srcs, w, *a, **k)
        e1.record()
        conv_events.append((e0, e1))
        if layers is not None:
            x = srcs[0]
            if isinstance(x, PhaseMap):
                ho, wo = x.h, x.w
            else:
                ho, wo = (x.h + 2 * w.pad - w.k) // w.stride + 1, (x.w + 2 * w.pad - w.k) // w.stride + 1
            layers.append((name, 2.0 * x.n * ho * wo * sum(w.src_c) * w.k * w.k * w.cout / 1e9, (e0, e1)))
        return r

    def timed_seg(name, x, w, *a, **k):
        e0, e1 = ev_pair()
        e0.record()
        r = orig_seg(name, x, w, *a, **k)
        e1.record()
        conv_events.append((e0, e1))
        if layers is not None:
            px = sum(n * h * ww for _, n, h, ww in x.segs)
            cout = 5 if "regctr" in name else w.cout                 # 4 box + 1 centerness columns; the other 11 are padding
            layers.append((name, 2.0 * px * sum(w.src_c) * w.k * w.k * cout / 1e9, (e0, e1)))
        return r

    # bandwidth-bound kernel families: lib function -> (label, algorithmic bytes of one call from its arguments)
    def seg_bytes(flat, segs, *a, **k):
        return 2 * sum(n * h * w for _, n, h, w in segs) * flat.shape[1] * flat.element_size()          # read + write, interior pixels

    def ese_bytes(x, gate, identity, full, pool):
        b = _numel_bytes(x) + (_numel_bytes(identity) if identity is not None else 0)
        return b + (_numel_bytes(full) if full is not None else 0) + (_numel_bytes(pool) if pool is not None else 0)

    def roi_bytes(feats, strides, boxes, count, n, r_cap, area, crit, ratio, out, *a, **k):
        return _numel_bytes(out) + sum(_numel_bytes(f) for f in feats)                                   # every level map once + the output

    def paste_bytes(probs, boxes, valid, out, r, m, oh, ow, thr):
        return r * oh * ow + r * m * m * 4

    def decode_bytes(logits, regctrs, *a, **k):
        return sum(lg.shape[0] * lg.shape[1] * lg.shape[2] * (lg.shape[3] + 5) * 4 for lg in logits)

    def im2col_bytes(imgs, mean, std, hp, wp, out, *a, **k):
        return sum(_numel_bytes(im) for im in imgs) + len(imgs) * (hp // 2) * (wp // 2) * 32 * out.element_size()

    def seg_split_bytes(flat, out_split, segs, *a, **k):
        px = sum(n * h * w for _, n, h, w in segs)
        return px * (flat.shape[1] * flat.element_size() + out_split.shape[1] * out_split.element_size())     # fp32 in, [hi | lo] f16 out

    families = {"groupnorm_apply_seg": ("groupnorm apply (+ReLU)", seg_bytes), "groupnorm_apply_seg_split": ("groupnorm apply (+ReLU)", seg_split_bytes),
                "ese_apply_pool": ("eSE apply (+ max-pool)", ese_bytes),
                "roialign_fpn": ("ROIAlign (+ level assignment)", roi_bytes), "paste_masks": ("mask paste-back", paste_bytes),
                "spatial_attention": ("spatial attention", lambda x, out, w: 2 * _numel_bytes(x)),
                "fcos_decode_levels": ("FCOS decode + threshold", decode_bytes),
                "preprocess_im2col_batch": ("normalise + pad + stem im2col", im2col_bytes),
                "split_f16x2": ("split-precision operand split (inside conv_ms)", lambda x, out: _numel_bytes(x) + _numel_bytes(out))}
    saved = {}
    for fn_name, (label, nbytes) in families.items():
        orig = getattr(lib, fn_name)
        saved[fn_name] = orig

        def make(orig=orig, label=label, nbytes=nbytes):
            def wrapped(*a, **k):
                e0, e1 = ev_pair()
                e0.record()
                r = orig(*a, **k)
                e1.record()
                fam.setdefault(label, []).append((e0, e1, nbytes(*a, **k)))
                return r
            return wrapped
        setattr(lib, fn_name, make())
    eng.conv, eng.conv_seg, eng.stem1_fused = timed_conv, timed_seg, timed_stem
    branches, eng.branch_streams = eng.branch_streams, False      # one kernel at a time: every event pair times its own launch only
    try:
        for _ in range(steps):
            # keep the GPU behind the host while the step is enqueued: with an empty queue every (event, launch) pair
            # would also time the few microseconds of host-side argument marshalling between the two calls
            torch.cuda._sleep(40_000_000)                # ~20 ms of device-side spinning
            device_step(model, cfg, dev_images, sizes_out)
        torch.cuda.synchronize()
    finally:
        eng.conv, eng.conv_seg, eng.stem1_fused = orig_conv, orig_seg, orig_stem
        eng.branch_streams = branches
        for fn_name, orig in saved.items():
            setattr(lib, fn_name, orig)
    conv_total = sum(a.elapsed_time(b) for a, b in conv_events)
    fams = {label: (sum(a.elapsed_time(b) for a, b, _ in v) / steps, sum(nb for _, _, nb in v) / steps, len(v) // steps)
            for label, v in fam.items()}
    return conv_total / steps, len(conv_events) // steps, fams


def cpu_baseline(steps, warmup, images_per_step=1, config="v39"):
    """The oracle (fp32 restatement executing the reference's ATen/torchvision CPU ops) on the host cores."""
    from centermask2_b200.synth import synthetic_state_dict, synthetic_images, calibrate_cls_bias
    from oracle import restate
    _, _, h, w, _ = CONFIGS[config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = make_cfg("fp32", config)
    sd = synthetic_state_dict(cfg, seed=WEIGHT_SEED)
    inputs = synthetic_images(images_per_step, h, w, seed=IMAGE_SEED)
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
            "sample": "{} step(s) x {} image(s) {}x{} {} fp32, oracle/restate.py (torch {} CPU ops)".format(
                steps, images_per_step, h, w, cfg.MODEL.VOVNET.CONV_BODY, torch.__version__), "ms_per_step": dt / steps * 1e3}


def workload_name(config, batch):
    _, _, h, w, idx = CONFIGS[config]
    body = {"v39": "V-39-eSE-FPN", "lite": "Lite V-19-eSE-FPN (upstream Lite recipe as cfg overrides)", "v99": "V-99-eSE-FPN"}[config]
    return ("CenterMask2 {}, {} synthetic {}x{} images per GPU per step (BASELINE configs[{}], image data-parallel), "
            "random-init weights, POST_NMS_TOPK 50".format(body, batch, h, w, idx))


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="v39", choices=["v39", "lite", "v99", "post"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (weak) / per step in total (strong); "
                                                            "default: the config's batch")
    ap.add_argument("--precision", default=os.environ.get("CM2_PRECISION", "bf16"), choices=["bf16", "fp32", "fp32_simt"],
                    help="bf16: bf16 activations, tcgen05 convolutions; fp32: fp32 activations, split-precision (f16 hi/lo) "
                         "tcgen05 convolutions with fp32-grade accuracy; fp32_simt: fp32 on CUDA cores")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-soak", action="store_true", help="skip the clock-settling replays in front of the timed regions")
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
        config = "v39" if args.config == "post" else args.config
        base = cpu_baseline(max(1, args.steps), max(0, args.warmup), config=config)
        line = {"impl": "reference", "metric": METRIC[config], "value": base["value"], "unit": "img/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(config, 1) + "; each step = 1 image on the host cores"},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    torch.cuda.set_device(local)
    cores = pin_rank_to_cores(local, world) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    if args.config == "post":
        return main_post(args, rank, world, local)

    import centermask2_b200 as cm
    from centermask2_b200 import lib, parallel, runtime
    from centermask2_b200.arch import conv_gflop_per_image
    from centermask2_b200.synth import synthetic_state_dict

    _, cfg_batch, h, w, _ = CONFIGS[args.config]
    total_batch = args.batch or cfg_batch
    if args.scaling == "strong":
        batch = len(parallel.shard_range(total_batch, rank, world))        # this rank's contiguous shard of the step's images
        if batch == 0:
            raise SystemExit("strong scaling: {} images cannot be sharded over {} ranks".format(total_batch, world))
    else:
        batch = total_batch
    images_per_step = total_batch if args.scaling == "strong" else batch * world

    clk = ClockSampler(local, args.clock_ms, enabled=(rank == 0))          # one sampler per job: rank 0's GPU
    cfg = make_cfg(args.precision, args.config)
    model = cm.build_model(cfg)
    model.load_state_dict(synthetic_state_dict(cfg, seed=WEIGHT_SEED))
    host_inputs = make_images(batch, rank, pinned=True, h=h, w=w)
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

    def max_over_ranks(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    if args.profile_step:
        eager = make_device_step(model, cfg, dev_images, (h, w), graph=False)
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
    step = make_device_step(model, cfg, dev_images, (h, w), graph=not args.no_graph)
    for _ in range(warmup):
        det, _ms = step()
    torch.cuda.synchronize()

    def soak(fn):
        """Run ``fn`` back to back for SOAK_SECONDS: the power-capped clocks settle before the timed region starts."""
        if args.no_soak:
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < SOAK_SECONDS:
            for _ in range(4):
                fn()
            torch.cuda.synchronize()

    soak(step)
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
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = lib.launch_count - l0
    value = images_per_step / (ms_step / 1e3)

    # ---- end to end through the public API.
    # (a) GeneralizedRCNN.inference_records(batches): pinned host images in, the whole result on the host (detection
    #     records + masks as COCO run lengths), software-pipelined -- the reported e2e value;
    # (b) GeneralizedRCNN.forward(batched_inputs) per step, synchronous, Instances with bool masks left on the device and
    #     the compact records copied to the host -- kept beside it (what a drop-in caller of the reference API gets).
    d2h = {"bytes": 0, "runs": 0, "n": 0}

    def e2e_records(k):
        last = None
        for res in model.inference_records((host_inputs for _ in range(k))):
            d2h["bytes"] += res.nbytes
            d2h["runs"] += res.rle_runs.numel()
            d2h["n"] += 1
            last = res
        return last

    def e2e_forward():
        out = model(host_inputs)
        return compact_results(out, r_cap).cpu()

    for _ in range(2):
        e2e_forward()
    e2e_records(8)                                # every slot of the pipeline: one eager pass + its graph capture
    soak(step)
    barrier()
    with clk.window():
        e0.record()
        for _ in range(args.steps):
            rec = e2e_forward()
        e1.record()
        barrier()
    e2e_sync_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    soak(step)
    barrier()
    d2h.update(bytes=0, runs=0, n=0)
    with clk.window():
        e0.record()
        last = e2e_records(args.steps)
        e1.record()
        barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    h2d = sum(_numel_bytes(b["image"]) for b in host_inputs)
    d2h_step = d2h["bytes"] / max(1, d2h["n"])
    runs_step = d2h["runs"] / max(1, d2h["n"])

    # ---- final result gather (the only collective; not on the hot path): records + run lengths of every rank's last batch
    local_res = last.clone()
    if args.scaling == "strong":
        allres = parallel.gather_results(local_res, total_batch, device=torch.device("cuda", local))
        assert allres.records.shape[0] == total_batch
    else:
        allrec = parallel.gather_records(local_res.records.cuda(), batch * world)
        assert allrec.shape[0] == batch * world

    # ---- roofline of the dominant kernel family (convolutions) + the bandwidth-bound families, timed in the step
    conv_ms, conv_launches, fams = instrumented_step(model, cfg, dev_images, (h, w), max(2, min(args.steps, 5)))
    if args.layers and rank == 0:
        rows = []
        instrumented_step(model, cfg, dev_images, (h, w), 1, layers=rows)
        with open(args.layers, "w") as f:
            f.write("# config {} batch {} precision {}\n# name gflop ms tflops\n".format(args.config, batch, args.precision))
            for nme, gf, (a, b) in rows:
                ms_l = a.elapsed_time(b)
                f.write("{:24s} {:10.3f} {:9.4f} {:9.1f}\n".format(nme, gf, ms_l, gf / ms_l if ms_l > 0 else 0.0))
    hp, wp = (h + 31) // 32 * 32, (w + 31) // 32 * 32
    gflop_img = conv_gflop_per_image(cfg, hp, wp, r_cap)             # algorithmic FLOPs of what is executed, R = slots computed
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")      # written by tools/ncu_step_summary.py from an ncu capture
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("images_per_gpu") == batch and tj.get("precision") == args.precision and tj.get("config", "v39") == args.config:
            traffic = tj.get("conv_dram_bytes_per_step")
    clk.close()
    hbm, tf_burst, tf_sus, src = peaks()
    achieved = gflop_img * batch / conv_ms                            # GFLOP / ms = TFLOP/s
    split = args.precision == "fp32"
    # the fp32 engine issues three f16 MMAs per algorithmic product (x_hi W_hi + x_lo W_hi + x_hi W_lo): its matching
    # peak is a third of the dense 16-bit tensor peak
    peak = tf_sus / 3.0 if split else tf_sus
    hbm_rows = [{"kernel": label, "launches_per_step": cnt, "ms_per_step": ms_f, "algorithmic_bytes_per_step": nb,
                 "achieved": nb / (ms_f * 1e-3) / 1e9 if ms_f > 0 else None, "peak": hbm, "unit": "GB/s",
                 "frac": nb / (ms_f * 1e-3) / 1e9 / hbm if ms_f > 0 else None}
                for label, (ms_f, nb, cnt) in sorted(fams.items(), key=lambda kv: -kv[1][0])]
    line = {
        "metric": METRIC[args.config], "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp32": "f32 (split f16 tensor-core operands)", "fp32_simt": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": workload_name(args.config, batch),
                   "images_per_gpu": batch, "images_per_step": images_per_step, "precision": args.precision, "cls_bias": bias,
                   "detections_per_image": dets_per_image, "candidates_per_level": cand,
                   "l2": "activations per step far exceed the 126 MB L2; no explicit flush",
                   "soak_s": 0.0 if args.no_soak else SOAK_SECONDS, "host_cores_per_rank": cores,
                   "cuda_graph": bool(eng.use_graphs and not args.no_graph), "parallelism": "dp{}".format(world)},
        "clocks": clk.summary(),
        "e2e": {"value": images_per_step / (e2e_ms / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_step, "ms_per_step": e2e_ms, "rle_runs_per_step": runs_step,
                "forward_per_call": {"value": images_per_step / (e2e_sync_ms / 1e3), "ms_per_step": e2e_sync_ms,
                                     "d2h_bytes_per_step": _numel_bytes(rec)},
                "note": "GeneralizedRCNN.inference_records(batches): every step copies its pinned uint8 host images to the "
                        "device (copy stream, overlapping the previous step), runs the step incl. mask paste-back, encodes "
                        "the pasted masks as COCO run lengths on the device and copies the detection records (box, score, "
                        "class, mask score, location) plus the run lengths of all masks to pinned host memory; "
                        "forward_per_call = GeneralizedRCNN.forward(batched_inputs), synchronous, Instances with bool masks "
                        "left on the device + compact records to the host"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "conv (all launches of one step)", "launches_per_step": conv_launches,
                     "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / ms_step,
                     "gflop_per_image": gflop_img,
                     "peak_source": src + (" (f16 sustained cuBLAS / 3: three MMAs per product)" if split else
                                           " (bf16 sustained cuBLAS; {} s of replays precede every timed region)".format(
                                               0 if args.no_soak else SOAK_SECONDS)),
                     "frac_of_burst": achieved / (tf_burst / 3.0 if split else tf_burst),
                     "whole_step_tflops": gflop_img * batch / ms_step,
                     "hbm": hbm_rows},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(2, 1, config=args.config)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main_post(args, rank, world, local):
    """BASELINE configs[4]: the FCOS post-process + ROI-stage kernels alone (tools/micro_post.py): every kernel is first
    checked against the oracle at the benchmark size, then timed; one JSON line with the per-kernel HBM rooflines."""
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import micro_post
    batch = args.batch or 32
    rows = micro_post.run(batch=batch, cand=1000, rois=100, check=(rank == 0), seed=5 + rank)
    total_ms = sum(r["ms"] for r in rows)
    t = torch.tensor([total_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    hbm, _, _, src = peaks()
    dom = max(rows, key=lambda r: r["ms"])
    line = {"metric": METRIC["post"], "value": batch * world / (t.item() / 1e3), "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t.item(), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 features, f32 head outputs", "data": "synthetic",
            "config": {"workload": "FCOS post-process + SAG-Mask ROI-stage kernels: 5 FPN levels of 800x1344, ~1000 candidates / level / "
                                   "image, 100 ROIs / image, batch {} (BASELINE configs[4])".format(batch), "images_per_gpu": batch,
                       "l2": "every working set except top-k / NMS exceeds the 126 MB L2", "checked_against_oracle": rank == 0},
            "gpu_launches": len(rows), "e2e": None,
            "roofline": {"bound": "hbm", "achieved": dom["GBps"], "peak": hbm, "unit": "GB/s", "frac": dom["GBps"] / hbm, "traffic": None,
                         "kernel": dom["kernel"], "peak_source": src, "kernels": rows}}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
