"""Launch plans of the CenterMask2 inference path on ``libcm2.so``.

This module is the host side of the hot path: it owns the activation buffers (NHWC with a one-pixel
zero halo, see ``include/cm2.h``), turns the static layer tables of ``arch.py`` plus packed weights
(``packing.py``) into sequences of C-ABI calls, and never touches torch compute ops -- torch is used
for device memory, streams and (at the API boundary only) indexing of the fixed-size result buffers.

Reference call sequence being replaced: ``GeneralizedRCNN.inference`` [d2] as mirrored in-tree at
``/root/reference/tester.py:24-75``: backbone (``vovnet.py:471-481`` + FPN [d2] + ``fpn.py:32-35``)
-> ``FCOS.forward`` (``fcos/fcos.py:61-118``) -> ``CenterROIHeads.forward`` (``center_heads.py:384-444``)
-> ``detector_postprocess`` [d2].
"""
import collections
import math
import os

import torch

from . import lib, packing
from .arch import vovnet_is_depthwise, vovnet_blocks


class FMap(object):
    """A feature map stored NHWC inside a buffer with ``halo`` zero pixels on every side."""

    __slots__ = ("buf", "halo")

    def __init__(self, buf, halo):
        self.buf, self.halo = buf, halo

    @property
    def view(self):
        h = self.halo
        return self.buf[:, h:self.buf.shape[1] - h, h:self.buf.shape[2] - h, :] if h else self.buf

    @property
    def n(self):
        return self.buf.shape[0]

    @property
    def h(self):
        return self.buf.shape[1] - 2 * self.halo

    @property
    def w(self):
        return self.buf.shape[2] - 2 * self.halo

    @property
    def c(self):
        return self.buf.shape[3]

    def nchw(self):
        """Zero-copy [N, C, H, W]-shaped (channels_last-strided) tensor for the registry-level API."""
        t = self.view.permute(0, 3, 1, 2)
        t._cm2_fmap = self
        return t


class PhaseMap(object):
    """A feature map stored as four stride-2 phase planes, ``buf[q, n, ceil(H/2)+2, ceil(W/2)+2, c]`` with
    q = (y&1)*2 + (x&1): the layout the tensor-core engine reads for 3x3 / stride-2 convolutions
    (``include/cm2.h``: ``src_phase`` / ``out_mode`` 2).  ``view`` is the interior of plane 0."""

    __slots__ = ("buf",)
    halo = 1

    def __init__(self, buf):
        self.buf = buf

    @property
    def view(self):
        return self.buf[0, :, 1:-1, 1:-1, :]

    @property
    def n(self):
        return self.buf.shape[1]

    @property
    def h(self):
        return self.buf.shape[2] - 2

    @property
    def w(self):
        return self.buf.shape[3] - 2

    @property
    def c(self):
        return self.buf.shape[4]


class SharedHaloFMap(FMap):
    """A feature map whose images SHARE their zero frame (csrc/conv_tc.cu ``halo_kind`` 2): flat ``[rows, c]`` buffer, line pitch
    ``w + 1`` and image pitch ``(h + 1)(w + 1)`` pixels -- the zero pixel right of a line is the zero pixel left of the next, the
    zero line under an image the zero line above the next.  For the 14x14 ROI maps a 3x3 convolution then runs over 225 GEMM rows
    per ROI instead of 256.  Only convolutions write it (interior + zeros on the shared frame) and view-based kernels read it."""

    __slots__ = ("_shape",)

    def __init__(self, flat, n, h, w):
        assert flat.dim() == 2 and flat.is_contiguous() and flat.shape[0] >= SharedHaloFMap.rows(n, h, w)
        self.buf, self.halo, self._shape = flat, 1, (n, h, w)

    @staticmethod
    def rows(n, h, w):
        return n * (h + 1) * (w + 1) + (w + 1) + 1

    @property
    def view(self):
        n, h, w = self._shape
        c = self.buf.shape[1]
        return torch.as_strided(self.buf, (n, h, w, c), ((h + 1) * (w + 1) * c, (w + 1) * c, c, 1), self.buf.storage_offset() + (w + 2) * c)

    n = property(lambda self: self._shape[0])
    h = property(lambda self: self._shape[1])
    w = property(lambda self: self._shape[2])
    c = property(lambda self: self.buf.shape[1])


class RawInput(object):
    """The un-normalised input batch on its way to the fused stem (``cm2_stem1_fused_batch``: normalise + pad + stem_1 in
    one pass): planar ``[3, h, w]`` device images of one dtype and the padded extent.  Never materialised as a tensor."""

    __slots__ = ("images", "hp", "wp")

    def __init__(self, images, hp, wp):
        self.images, self.hp, self.wp = images, hp, wp

    n = property(lambda self: len(self.images))
    h = property(lambda self: self.hp)
    w = property(lambda self: self.wp)
    c = 3


class SplitFMap(FMap):
    """fp32 engine: a feature map that exists only as the [hi | lo] f16 operand pair of the convolutions that read it
    (``buf[n, h+2, w+2, 2c]``; include/cm2.h "Split precision"): written by a convolution epilogue or the GroupNorm apply,
    never materialised in fp32.  ``c`` is the logical channel count."""

    __slots__ = ()

    @property
    def c(self):
        return self.buf.shape[3] // 2


class SplitPhaseMap(PhaseMap):
    """The same for a feature map stored as four stride-2 phase planes."""

    __slots__ = ()

    @property
    def c(self):
        return self.buf.shape[4] // 2


class SegList(list):
    """Segment table of a SegMap: (row0, n, h, w) per map; ``halo`` = 1 when the images of every map share their zero frame
    (``cm2_seg.halo``, the layout of SharedHaloFMap)."""

    halo = 0


class SegMap(object):
    """Several halo feature maps of different extent stored back to back in one flat ``[rows, c]`` buffer
    (``include/cm2.h``: ``cm2_seg``).  Used for the FPN levels the shared-weight FCOS towers run on, so that one
    convolution / GroupNorm launch covers all levels.  Every level is also reachable as an ordinary ``FMap``."""

    ALIGN = 256

    def __init__(self, shapes, c, dtype, device, alloc=None, shared_halo=False):
        """shapes: list of (n, h, w).  ``shared_halo``: the images of a map share their zero frame (SharedHaloFMap layout):
        on the five FCOS levels of an 800x1344 batch 2.4 % of the GEMM rows are frame instead of 9.7 %."""
        self.segs, row = SegList(), 0
        self.segs.halo = 1 if shared_halo else 0
        for n, h, w in shapes:
            self.segs.append((row, n, h, w))
            row += SharedHaloFMap.rows(n, h, w) if shared_halo else n * (h + 2) * (w + 2)
            row = (row + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.rows, self.c = max(row, self.ALIGN), c
        self.flat = (alloc or (lambda shape: torch.zeros(shape, dtype=dtype, device=device)))((self.rows, c))

    def level(self, i):
        row0, n, h, w = self.segs[i]
        if self.segs.halo:
            return SharedHaloFMap(self.flat[row0:row0 + SharedHaloFMap.rows(n, h, w)], n, h, w)
        return FMap(self.flat[row0:row0 + n * (h + 2) * (w + 2)].view(n, h + 2, w + 2, self.c), 1)

    is_split = False                                   # True: ``flat`` is the [hi | lo] f16 operand form [rows, 2c]

    def like(self, c, dtype, alloc, split=False):
        out = SegMap.__new__(SegMap)
        out.segs, out.rows, out.c = self.segs, self.rows, c
        out.flat = alloc((self.rows, 2 * c if split else c))
        out.is_split = split
        return out


def as_fmap(t, dtype, device):
    """Accept what a detectron2 caller hands over: one of our own tensors (zero copy) or any NCHW
    tensor (copied once into a halo buffer)."""
    fm = getattr(t, "_cm2_fmap", None)
    if fm is not None and fm.buf.dtype == dtype:
        return fm
    n, c, h, w = t.shape
    buf = torch.zeros((n, h + 2, w + 2, c), dtype=dtype, device=device)
    buf[:, 1:-1, 1:-1, :] = t.permute(0, 2, 3, 1).to(device=device, dtype=dtype)
    return FMap(buf, 1)


class Engine(object):
    """Buffer cache + layer launchers for one (cfg, precision, device)."""

    def __init__(self, cfg, precision="fp32", device="cuda"):
        assert precision in ("fp32", "fp32_simt", "bf16"), precision
        self.cfg = cfg
        self.precision = precision
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        # tcgen05 convolutions: bf16 operands ("bf16"), or fp32 activations whose convolutions run on split f16 hi/lo
        # operands ("fp32": three MMAs per product term, fp32-grade accuracy); "fp32_simt" keeps everything on CUDA cores
        self.tc = precision in ("bf16", "fp32")
        self.split = precision == "fp32"
        self.stem_variant = int(os.environ.get("CM2_STEM_VARIANT", "1"))      # 1: fused stem_1 (csrc/stem.cu); 0: im2col pass + K = 32 GEMM
        self.shared_halo_roi = self.tc and os.environ.get("CM2_SHARED_HALO_ROI", "1") != "0"
        self.tower_overlap = int(os.environ.get("CM2_TOWER_OVERLAP", "0")) if precision == "bf16" else 0
        self.shared_halo_seg = self.tc and os.environ.get("CM2_SHARED_HALO_SEG", "1") != "0"
        self.splitk_on = os.environ.get("CM2_SPLITK", "1") != "0"               # split-K for small-M layers (cm2_conv_desc.splitk)
        self.split_out_all = os.environ.get("CM2_SPLIT_OUT_ALL") == "1"       # [hi | lo] epilogue store on every eligible layer (tests)
        self._split_cache = {}
        self.device = torch.device(device)
        self._bufs = collections.OrderedDict()              # least recently used first
        self._buf_bytes = 0
        self._graphs = collections.OrderedDict()            # key -> (graph, result, launches, kept-alive objects)
        self._graph_seen = collections.OrderedDict()        # key -> number of eager runs so far
        self._recording = None                              # buffers touched while a plan is being captured
        self.records_slot_free = {}                         # GeneralizedRCNN.inference_records: pipeline slot -> "inputs consumed" event
        self._det_gen = 0                                   # bumped by every run_fcos_post (see roi_heads._det_from_instances)
        self.use_graphs = os.environ.get("CM2_GRAPH", "1") != "0"
        self.graph_after = max(1, int(os.environ.get("CM2_GRAPH_AFTER", "1")))      # eager runs of a key before it is captured
        self.graph_cache = max(1, int(os.environ.get("CM2_GRAPH_CACHE", "16")))     # captured graphs kept (LRU)
        self.buffer_budget = int(float(os.environ.get("CM2_BUFFER_BUDGET_GB", "96")) * (1 << 30))
        # independent branches of the launch plan (FCOS cls / bbox towers, the P6 / P7 chain) go to a second stream: when
        # a launch does not fill the 148 SMs (small batches, small maps) the other branch runs beside it
        # (measured, profiles/r2_small_batch.txt: -3.6 % step time at 2 images per GPU, -2 % on the Lite config, nothing at 16
        # images where every launch already runs ~20 waves of tiles -- so it is used below BRANCH_MAX_ROWS GEMM rows)
        self.branch_streams = os.environ.get("CM2_BRANCH_STREAMS", "1") != "0"
        lib.load()

    # -- buffers ---------------------------------------------------------------------------------
    def buffer(self, name, shape, dtype, zero=True):
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
            self._buf_bytes += t.numel() * t.element_size()
        else:
            self._bufs.move_to_end(key)
        if self._recording is not None:
            self._recording.append(t)
        return t

    def image_buffers(self, prefix, sig):
        """Engine-owned buffers for a batch of input images (``sig``: one (shape, dtype) per image).  Images of one shape and
        dtype live back to back in ONE buffer, so that a staging -> input move is a single device-to-device copy instead of
        one per image.  Returns (list of per-image tensors, the whole batch tensor or None)."""
        if len(sig) > 1 and len(set(sig)) == 1:
            whole = self.buffer(prefix + "_all", (len(sig),) + tuple(sig[0][0]), sig[0][1], zero=False)
            return [whole[i] for i in range(len(sig))], whole
        return [self.buffer("{}{}".format(prefix, i), tuple(shp), dt, zero=False) for i, (shp, dt) in enumerate(sig)], None

    def const(self, key, make):
        """Small cached device constant (or tuple of them) under a free-form ``key``; built by ``make()`` on first use."""
        t = self._bufs.get(key)
        if t is None:
            t = make()
            self._bufs[key] = t
            for x in (t if isinstance(t, tuple) else (t,)):
                self._buf_bytes += x.numel() * x.element_size()
        else:
            self._bufs.move_to_end(key)
        if self._recording is not None:
            self._recording.append(t)
        return t

    def trim(self):
        """Between steps: when the cached buffers exceed the budget (a data set with many padded shapes keeps one
        activation set per shape), drop the least recently used ones.  A buffer a captured graph refers to stays alive
        through that graph's entry, so dropping it here never invalidates a replay."""
        if self._buf_bytes <= self.buffer_budget:
            return
        for key in list(self._bufs.keys()):
            if self._buf_bytes <= self.buffer_budget // 2:
                break
            t = self._bufs[key]
            if not isinstance(t, torch.Tensor) or t.is_pinned():
                continue
            del self._bufs[key]
            self._buf_bytes -= t.numel() * t.element_size()

    _stage_free = None                                  # event: the staging input buffers have been consumed

    def copy_stream(self):
        """Side stream for host->device input copies that overlap the previous batch (GeneralizedRCNN.inference_stream)."""
        st = self._bufs.get(("copy_stream",))
        if st is None:
            st = torch.cuda.Stream(device=self.device)
            self._bufs[("copy_stream",)] = st
        return st

    def side_stream(self, name, priority=0):
        """A named extra stream (result read-back of the pipelined entry points; ``priority`` < 0: higher than the default)."""
        key = ("stream", name)
        st = self._bufs.get(key)
        if st is None:
            st = torch.cuda.Stream(device=self.device, priority=priority)
            self._bufs[key] = st
        return st

    BRANCH_MAX_ROWS = 8 * 148 * 128          # a tower launch of fewer than ~8 waves of 128-row tiles leaves SMs idle at its tail

    def fork(self):
        """Start a branch: returns a side stream ordered after everything enqueued so far on the current stream (works
        eagerly and under CUDA-graph capture, where it becomes a parallel branch of the graph)."""
        side = self.side_stream("branch")
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        return side

    def join(self, side):
        """End a branch: the current stream waits for everything enqueued on ``side``."""
        ev = torch.cuda.Event()
        ev.record(side)
        torch.cuda.current_stream().wait_event(ev)

    def pinned(self, name, shape, dtype):
        """Page-locked host staging buffer (async D2H of the small result-size tensors)."""
        key = ("pinned", name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
            self._bufs[key] = t
        return t

    def fmap(self, name, n, h, w, c, dtype=None, halo=1):
        return FMap(self.buffer(name, (n, h + 2 * halo, w + 2 * halo, c), dtype or self.dtype), halo)

    def shared_halo_fmap(self, name, n, h, w, c, dtype=None):
        return SharedHaloFMap(self.buffer(name, (SharedHaloFMap.rows(n, h, w), c), dtype or self.dtype), n, h, w)

    def phasemap(self, name, n, h, w, c):
        """Phase planes for a full-resolution [n, h, w, c] map."""
        return PhaseMap(self.buffer(name, (4, n, (h + 1) // 2 + 2, (w + 1) // 2 + 2, c), self.dtype))

    def phase_split(self, name, x, relu=False):
        out = self.phasemap(name, x.n, x.h, x.w, x.c)
        lib.phase_split(x.view, out.view, relu)
        return out

    def release(self):
        self._graphs.clear()
        self._graph_seen.clear()
        self._bufs.clear()
        self._buf_bytes = 0
        self.records_slot_free.clear()

    def drop_graphs(self, owner=None):
        """Forget captured graphs: all of them, or those whose key carries a ``graph_token`` of module ``owner`` (its
        packed weights -- whose addresses the capture baked in -- are about to be rebuilt)."""
        def owned(key):
            return any(isinstance(k, tuple) and any(isinstance(t, tuple) and len(t) == 3 and t[0] == "cm2w" and t[1] == owner
                                                    for t in k) for k in key)
        for d in (self._graphs, self._graph_seen):
            for key in [k for k in d if owner is None or owned(k)]:
                del d[key]

    # -- one convolution -------------------------------------------------------------------------
    def conv(self, name, srcs, w, out_dtype=None, residual=None, res_mode=0, out_mode=0, in_relu=False,
             out_halo=1, stats=None, stats_mode=0, out=None, pred=None, to_conv=False):
        """One convolution.  ``srcs`` are FMaps (virtual concat) or, for a stride-2 conv on the TC engine,
        PhaseMaps.  ``out_mode`` 1 = deconv scatter, 2 = write the result as a PhaseMap.  ``to_conv=True`` promises that
        only convolutions read the result: the fp32 engine then stores it directly as their [hi | lo] f16 operands
        (a ``SplitFMap`` / ``SplitPhaseMap``; no fp32 copy, no split pass); the other engines ignore the hint."""
        x0 = srcs[0]
        src_phase = isinstance(x0, PhaseMap)
        if src_phase:
            ho, wo = x0.h, x0.w
        else:
            ho = (x0.h + 2 * w.pad - w.k) // w.stride + 1
            wo = (x0.w + 2 * w.pad - w.k) // w.stride + 1
        # measured per layer (profiles/r2_layers_v39_fp32*.txt): the [hi | lo] store pays where one N tile covers the layer
        # (cout <= 128: stem_2 -> stem_3, OSA2 3x3 +10 %); with two N tiles per row block (OSA3/4, the ROI heads) the longer
        # epilogue delays the accumulator drains of the next tile by as much as the saved split pass took
        split_out = (to_conv and self.split and w.split and out is None and out_mode in (0, 2) and residual is None and stats is None
                     and w.cout % 16 == 0 and (w.cout <= 128 or self.split_out_all) and not in_relu and out_halo in (0, 1)
                     and all(s.c % 16 == 0 for s in srcs))
        if split_out:
            if out_mode == 0:
                out = SplitFMap(self.buffer(name, (x0.n, ho + 2 * out_halo, wo + 2 * out_halo, 2 * w.cout), torch.float16), out_halo)
            else:
                out = SplitPhaseMap(self.buffer(name, (4, x0.n, (ho + 1) // 2 + 2, (wo + 1) // 2 + 2, 2 * w.cout), torch.float16))
        if out is None:
            if out_mode == 0 and out_halo == 1 and isinstance(x0, SharedHaloFMap) and w.stride == 1 and out_dtype in (None, self.dtype):
                out = self.shared_halo_fmap(name, x0.n, ho, wo, w.cout)      # a chain of convolutions keeps the shared-halo layout
            elif out_mode == 0:
                out = self.fmap(name, x0.n, ho, wo, w.cout, out_dtype, out_halo)
            elif out_mode == 1:
                out = self.fmap(name, x0.n, 2 * ho, 2 * wo, w.cout // 4, out_dtype, out_halo)
            else:
                out = self.phasemap(name, x0.n, ho, wo, w.cout)
        presplit = [isinstance(s, (SplitFMap, SplitPhaseMap)) for s in srcs]
        views = [s.view for s in srcs]
        for s_, c in zip(srcs, w.src_c):
            assert s_.c == c, (name, s_.c, w.src_c)
        kw = dict(shift=w.shift, relu=w.relu, in_relu=in_relu, src_phase=src_phase,
                  residual=None if residual is None else residual.view, res_mode=res_mode, out_mode=out_mode)
        if self.tc and w.w_tc is not None:
            tc_views = views
            if w.split:
                # split precision: fp32 activations -> [hi | lo] f16 tensors (cm2_split_f16x2), one per source
                if in_relu or any(s.c % 16 for s in srcs) or not (split_out or out.view.dtype == torch.float32):
                    tc_views = None
                else:
                    tc_views = [s.view if pre else self.split_of(s).view for s, pre in zip(srcs, presplit)]
            if tc_views is not None:
                ks = 0
                if (self.splitk_on and not w.split and out_mode == 0 and residual is None and stats is None and pred is None and not in_relu
                        and w.cout % 8 == 0):
                    ks = self._splitk_slices(x0, ho, wo, w, src_phase)
                if ks >= 2:
                    # few output tiles, long K loop: K slices as separate tiles + a fixed-order reduction (cm2_conv_desc.splitk)
                    ob = out.buf
                    ws = self.buffer(name + "_splitk", (ks, ob.shape[0] * ob.stride(0)), torch.float32, zero=False)
                    if lib.conv2d(tc_views, w.w_tc, out.view, w.cout, w.k, w.stride, w.pad, scale=w.scale_tc, engine=lib.ENGINE_TC, probe=True,
                                  splitk=ks, splitk_ws=ws, **kw):
                        lib._count()                                  # the reduction kernel
                        return out
                if lib.conv2d(tc_views, w.w_tc, out.view, w.cout, w.k, w.stride, w.pad, scale=w.scale_tc,
                              engine=lib.ENGINE_TC, stats=stats, stats_mode=stats_mode, probe=True, pred=pred, **kw):
                    return out
        assert stats is None and pred is None, "fused epilogues require the tensor-core engine: " + lib.last_error()
        assert not split_out and not any(presplit), "split-precision operands require the tensor-core engine: " + lib.last_error()
        lib.conv2d(views, w.w_simt, out.view, w.cout, w.k, w.stride, w.pad, scale=w.scale, engine=lib.ENGINE_SIMT, **kw)
        return out

    @staticmethod
    def _splitk_slices(x0, ho, wo, w, src_phase, sms=148):
        """K slices for a layer whose output tiles (128 GEMM rows x up to 256 channels) fill at most half of the SMs: as many
        as fit on the SMs, at least 4 K-blocks (of 64 channels) each.  0: leave the layer alone."""
        halo = getattr(x0, "halo", 1)
        rows = x0.n * ((ho + 2) * (wo + 2) if (src_phase or halo) else ho * wo)
        tiles = -(-rows // 128) * -(-((w.cout + 15) // 16 * 16) // 256)
        kb = w.k * w.k * sum(-(-c // 64) for c in w.src_c)
        if 2 * tiles > sms or kb < 32:
            return 0
        return min(sms // tiles, kb // 4, 16)

    # -- split-precision operands ------------------------------------------------------------------
    def begin_pass(self):
        """Start of a run_* call: split tensors made from here on are valid until the next call (their sources are
        engine buffers that the following pass overwrites)."""
        self._split_cache = {}

    def split_of(self, x):
        """[hi | lo] f16 companion of an fp32 FMap / PhaseMap / flat segmented buffer (``cm2_split_f16x2``), made once per
        pass however many convolutions read ``x``.  The companion lives in an engine buffer named after the source's
        address, so a CUDA-graph capture sees the same addresses every step."""
        buf = x if isinstance(x, torch.Tensor) else x.buf
        key = (buf.data_ptr(), tuple(buf.shape))
        hit = self._split_cache.get(key)
        if hit is not None:
            return hit
        assert buf.dtype == torch.float32 and buf.is_contiguous(), (buf.dtype, buf.stride())
        sp = self.buffer(("split",) + key, tuple(buf.shape[:-1]) + (2 * buf.shape[-1],), torch.float16, zero=False)
        lib.split_f16x2(buf, sp)
        if isinstance(x, torch.Tensor):
            out = sp
        elif isinstance(x, SharedHaloFMap):
            out = SharedHaloFMap(sp, x.n, x.h, x.w)
        else:
            out = PhaseMap(sp) if isinstance(x, PhaseMap) else FMap(sp, x.halo)
        self._split_cache[key] = out
        return out

    def dw_unit(self, name, x, w9c, pw, stride):
        """Depthwise 3x3 (stride 1 / 2, no activation) -> pointwise 1x1 + FrozenBN + ReLU, vovnet.py:110-130."""
        mid = self.fmap(name + "_dw", x.n, (x.h - 1) // stride + 1, (x.w - 1) // stride + 1, x.c)
        lib.dwconv3x3(x.view, mid.view, w9c, stride)
        return self.conv(name, [mid], pw)

    def segmap(self, name, shapes, c, dtype=None):
        dt = dtype or self.dtype
        # bf16 engine, tower width a multiple of 256 (GroupNorm statistics from the conv epilogue + cm2_groupnorm_apply_seg): the
        # levels share their zero frames
        shared = self.shared_halo_seg and dt == self.dtype and c % 256 == 0 and 256 % (c // 8) == 0
        return SegMap(shapes, c, dt, self.device, alloc=lambda shape: self.buffer(name + ("_sh" if shared else ""), shape, dt),
                      shared_halo=shared)

    def conv_seg(self, name, x, w, out_dtype=None, stats=None, stats_mode=0):
        """One stride-1 convolution over all maps of a SegMap (TC engine); returns a SegMap of the same geometry."""
        dt = out_dtype or self.dtype
        cout_pad = (w.cout + 15) // 16 * 16
        assert cout_pad == w.cout, "segmented conv needs cout % 16 == 0"
        out = x.like(w.cout, dt, lambda shape: self.buffer(name, shape, dt))
        srcs = [x.flat if getattr(x, "is_split", False) else self.split_of(x.flat)] if w.split else [x.flat]
        lib.conv2d(srcs, w.w_tc, out.flat, w.cout, w.k, w.stride, w.pad, scale=w.scale_tc, shift=w.shift, relu=w.relu,
                   engine=lib.ENGINE_TC, segs=x.segs, stats=stats, stats_mode=stats_mode)
        return out

    # =============================================================================================
    # backbone: VoVNetV2-eSE + FPN + P6/P7
    # =============================================================================================
    def pack_backbone(self, sd, prefix=""):
        """``sd`` keys relative to the backbone module (``bottom_up.*``, ``fpn_*``, ``top_block.*``)."""
        cfg, dt, dev, tc = self.cfg, self.dtype, self.device, self.tc
        stem, blocks = vovnet_blocks(cfg.MODEL.VOVNET.CONV_BODY)
        P = {"stem": [], "blocks": [], "reduction": {}}
        dw = vovnet_is_depthwise(cfg.MODEL.VOVNET.CONV_BODY)
        cin = 3
        for i, (c, s) in enumerate(zip(stem, (2, 1, 2))):
            k = prefix + "bottom_up.stem.stem_{}".format(i + 1)
            if dw and i > 0:                                   # vovnet.py:408-411: (depthwise weights, pointwise unit, stride)
                P["stem"].append(packing.dw_pw_bn_relu(sd, k, c, dt, dev, tc) + (s,))
            else:
                P["stem"].append(packing.conv_bn_relu(sd, k, [cin], s, 1, dt, dev, tc))
            cin = c
        if tc and self.split and sd[prefix + "bottom_up.stem.stem_1/conv.weight"].shape[0] == 64:
            # fp32 engine: fused stem on f16 hi / lo fragments (csrc/stem.cu, SPLIT): weights pre-scaled per output channel by a
            # power of two (far from the half subnormals), undone exactly by the epilogue scale -- as packing.ConvW does
            k1 = prefix + "bottom_up.stem.stem_1"
            w1 = sd[k1 + "/conv.weight"].detach().float()
            sc, sh = packing.fold_frozen_bn(sd[k1 + "/norm.weight"], sd[k1 + "/norm.bias"], sd[k1 + "/norm.running_mean"],
                                            sd[k1 + "/norm.running_var"])
            w30 = torch.zeros((64, 32))
            w30[:, :30] = torch.nn.functional.pad(w1.permute(0, 2, 3, 1).reshape(64, 3, 9), (0, 1)).reshape(64, 30)
            pre = torch.exp2(torch.floor(torch.log2(256.0 / w30.abs().amax(dim=1).clamp(min=1e-30)))).clamp(max=2.0 ** 40)
            ws = w30 * pre.view(-1, 1)
            hi = ws.to(torch.float16)
            lo = (ws - hi.float()).to(torch.float16)
            P["stem1_fused"] = (torch.stack([hi, lo]).to(dev).contiguous(), (sc.float() / pre).to(dev).contiguous(), sh.float().to(dev).contiguous())
        if tc and not self.split:
            # stem_1 as a 1x1 conv over the fused normalise+im2col input (cm2_preprocess_im2col): K = 27 -> 32
            k1 = prefix + "bottom_up.stem.stem_1"
            w1 = sd[k1 + "/conv.weight"].detach().float()
            w32 = torch.zeros((w1.shape[0], 32, 1, 1))
            w32[:, :27, 0, 0] = w1.permute(0, 2, 3, 1).reshape(w1.shape[0], 27)
            sc, sh = packing.fold_frozen_bn(sd[k1 + "/norm.weight"], sd[k1 + "/norm.bias"], sd[k1 + "/norm.running_mean"],
                                            sd[k1 + "/norm.running_var"])
            P["stem1_im2col"] = packing.ConvW(w32, [32], 1, 0, sc, sh, True, dt, dev, tc)
            if w1.shape[0] == 64:
                # fused stem (csrc/stem.cu): K order ky * 10 + kx * 3 + c, the filter rows padded from 9 to 10 values
                w30 = torch.zeros((64, 32))
                w30[:, :30] = torch.nn.functional.pad(w1.permute(0, 2, 3, 1).reshape(64, 3, 9), (0, 1)).reshape(64, 30)
                P["stem1_fused"] = (w30.to(device=dev, dtype=torch.bfloat16).contiguous(), P["stem1_im2col"].scale, P["stem1_im2col"].shift)
        for b in blocks:
            convs = []
            c = b.in_ch
            if b.reduced:                                      # vovnet.py:278-283
                P["reduction"][b.name] = packing.conv_bn_relu(sd, prefix + "bottom_up." + b.reduction_key(), [c], 1, 0, dt, dev, tc)
                c = b.mid_ch
            for i in range(b.n_conv):
                if b.dw:
                    convs.append(packing.dw_pw_bn_relu(sd, prefix + "bottom_up." + b.key(i), c, dt, dev, tc))
                else:
                    convs.append(packing.conv_bn_relu(sd, prefix + "bottom_up." + b.key(i), [c], 1, 1, dt, dev, tc))
                c = b.mid_ch
            cat = packing.conv_bn_relu(sd, prefix + "bottom_up." + b.key("concat"), [b.in_ch] + [b.mid_ch] * b.n_conv, 1, 0, dt, dev, tc)
            ek = prefix + "bottom_up." + b.ese_key()
            ese_w = sd[ek + ".weight"].detach().reshape(b.out_ch, b.out_ch).to(device=dev, dtype=torch.float32).contiguous()
            ese_b = sd[ek + ".bias"].detach().to(device=dev, dtype=torch.float32).contiguous()
            P["blocks"].append((b, convs, cat, ese_w, ese_b))
        out_ch = {"stage{}".format(b.stage): b.out_ch for b in blocks}
        fc = cfg.MODEL.FPN.OUT_CHANNELS
        P["fpn"] = {}
        for f in cfg.MODEL.FPN.IN_FEATURES:
            lvl = int(f[-1])
            lat = packing.conv_bias(sd, prefix + "fpn_lateral{}".format(lvl), [out_ch[f]], 1, 0, False, dt, dev, tc)
            outc = packing.conv_bias(sd, prefix + "fpn_output{}".format(lvl), [fc], 1, 1, False, dt, dev, tc)
            P["fpn"][f] = (lvl, lat, outc)
        P["top"] = [packing.conv_bias(sd, prefix + "top_block.p{}".format(6 + i), [fc], 2, 1, False, dt, dev, tc)
                    for i in range(cfg.MODEL.FCOS.TOP_LEVELS)]
        return P

    @staticmethod
    def _pool_extent(h, w):
        """Output extent of MaxPool2d(3, 2, ceil_mode=True) (vovnet.py:349-350)."""
        ho, wo = -(-(h - 3) // 2) + 1, -(-(w - 3) // 2) + 1
        if (ho - 1) * 2 >= h:
            ho -= 1
        if (wo - 1) * 2 >= w:
            wo -= 1
        return ho, wo

    def run_backbone(self, x, P):
        """x: FMap [N, Hp, Wp, 3] (normalised, padded to /32).  Returns {"p3": FMap, ...}."""
        cfg = self.cfg
        self.begin_pass()
        dw_body = isinstance(P["stem"][1], tuple)
        if isinstance(x, RawInput):
            x = self.stem1_fused("stem1", x, P)
        elif self.tc and x.c == 32:
            x = self.conv("stem1", [x], P["stem1_im2col"])      # stem_1 = 1x1 over the im2col'd input
        else:
            x = self.conv("stem1", [x], P["stem"][0])
        if dw_body:
            for i in (1, 2):
                w9c, pw, stride = P["stem"][i]
                x = self.dw_unit("stem{}".format(i + 1), x, w9c, pw, stride)
        elif self.tc:
            # tensor-core path: stem_2 writes phase planes, stem_3 (stride 2) reads them
            x = self.conv("stem2", [x], P["stem"][1], out_mode=2, to_conv=True)
            x = self.conv("stem3", [x], P["stem"][2])
        else:
            for i in (1, 2):
                x = self.conv("stem{}".format(i + 1), [x], P["stem"][i])
        stage = 2
        stage_out = {}
        blocks = P["blocks"]
        fpn_in = set(cfg.MODEL.FPN.IN_FEATURES)
        pooled_next = None                                   # next stage's input when the eSE pass already pooled it
        for bi, (b, convs, cat, ese_w, ese_b) in enumerate(blocks):
            if b.stage != stage:
                if pooled_next is not None:
                    x, pooled_next = pooled_next, None
                else:
                    # MaxPool2d(3, 2, ceil_mode=True), vovnet.py:349-350
                    ho, wo = self._pool_extent(x.h, x.w)
                    pooled = self.fmap("pool{}".format(b.stage), x.n, ho, wo, x.c)
                    lib.maxpool3x3s2_ceil(x.view, pooled.view)
                    x = pooled
                stage = b.stage
            identity = x
            feats = [x]
            y = x
            if b.reduced:
                y = self.conv(b.name + "_reduction", [y], P["reduction"][b.name])
            for i, w in enumerate(convs):
                if b.dw:
                    y = self.dw_unit("{}_{}".format(b.name, i), y, w[0], w[1], 1)
                else:
                    y = self.conv("{}_{}".format(b.name, i), [y], w, to_conv=True)     # read by the next 3x3 and the aggregation
                feats.append(y)
            n, c = x.n, cat.cout
            gate = self.buffer(b.name + "_gate", (n, c), torch.float32, zero=False)
            last_of_stage = bi + 1 == len(blocks) or blocks[bi + 1][0].stage != b.stage
            if self.tc and cat.w_tc is not None:
                # eSE (vovnet.py:247-260 / :327-330): channel sums from the conv epilogue, gate, then x * gate (+ identity)
                # fused with the max-pool that opens the next stage
                sums = self.buffer(b.name + "_sums", (n, c), torch.float64, zero=False)
                agg = self.conv(b.name + "_cat", feats, cat, stats=sums, stats_mode=1)      # virtual concat, vovnet.py:324-325
                lib.ese_gate_f64(sums, 1.0 / (agg.h * agg.w), ese_w, ese_b, gate, n, c)
                want_pool = last_of_stage and bi + 1 < len(blocks) and agg.h >= 3 and agg.w >= 3
                want_full = not last_of_stage or "stage{}".format(b.stage) in fpn_in or not want_pool
                out = self.fmap(b.name + "_out", n, agg.h, agg.w, c) if want_full else None
                if want_pool:
                    ho, wo = self._pool_extent(agg.h, agg.w)
                    pooled_next = self.fmap("pool{}".format(b.stage + 1), n, ho, wo, c)
                lib.ese_apply_pool(agg.view, gate, identity.view if b.identity else None,
                                   out.view if out is not None else None, pooled_next.view if want_pool else None)
            else:
                agg = self.conv(b.name + "_cat", feats, cat)
                hw = agg.h * agg.w
                wsp = self.buffer(b.name + "_esews", (n * lib.ese_pool_chunks(hw) * c,), torch.float32, zero=False)
                pooled = self.buffer(b.name + "_pool", (n, c), torch.float32, zero=False)
                lib.ese_pool(agg.view, wsp, pooled)
                lib.ese_gate(pooled, 1.0, ese_w, ese_b, gate, n, c)
                out = self.fmap(b.name + "_out", n, agg.h, agg.w, c)
                lib.ese_apply(agg.view, gate, identity.view if b.identity else None, out.view)
            x = out
            if out is not None:
                stage_out["stage{}".format(b.stage)] = x
        # FPN top-down [d2] (constructed at vovnet.py:547-554)
        res = {}
        prev = None
        in_feats = list(cfg.MODEL.FPN.IN_FEATURES)
        names = ["p{}".format(P["fpn"][f][0]) for f in in_feats] + ["p{}".format(P["fpn"][in_feats[-1]][0] + 1 + i)
                                                                    for i in range(len(P["top"]))]
        pyramid = None
        if self.tc:
            # all pyramid levels live in one segmented buffer so that the FCOS towers run as single launches
            shapes = [(stage_out[f].n, stage_out[f].h, stage_out[f].w) for f in in_feats]
            hh, ww = shapes[-1][1], shapes[-1][2]
            for _ in P["top"]:
                hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                shapes.append((shapes[0][0], hh, ww))
            pyramid = self.segmap("pyramid", shapes, cfg.MODEL.FPN.OUT_CHANNELS)
        slot = {nme: i for i, nme in enumerate(names)}

        def top_levels():
            # LastLevelP6P7 / LastLevelP6, fpn.py:32-35, :50-53 (P7 = conv(relu(P6)))
            top = res[names[len(in_feats) - 1]]
            for i, w in enumerate(P["top"]):
                nme = names[len(in_feats) + i]
                if self.tc:
                    src = self.phase_split(nme + "_src_phase", top, relu=(i == 1))
                    top = self.conv(nme, [src], w, out=pyramid.level(slot[nme]))
                else:
                    top = self.conv(nme, [top], w, in_relu=(i == 1))
                res[nme] = top

        side = None
        for fi, f in enumerate(reversed(in_feats)):
            lvl, lat, outc = P["fpn"][f]
            prev = self.conv("fpn_inner{}".format(lvl), [stage_out[f]], lat, residual=prev, res_mode=2 if prev is not None else 0)
            nme = "p{}".format(lvl)
            res[nme] = self.conv(nme, [prev], outc, out=pyramid.level(slot[nme]) if pyramid is not None else None)
            if fi == 0 and P["top"] and self.tc and self.branch_streams and len(in_feats) > 1 and pyramid.rows <= self.BRANCH_MAX_ROWS:
                # the extra levels hang off the coarsest output only: their small launches run beside the finer levels
                side = self.fork()
                with torch.cuda.stream(side):
                    top_levels()
        if side is not None:
            self.join(side)
        elif P["top"]:
            top_levels()
        out = {k: res[k] for k in sorted(res)}
        if pyramid is not None:
            self._pyramid = (pyramid, [out[k] for k in names])
        return out

    # =============================================================================================
    # FCOS head + post-process
    # =============================================================================================
    def pack_fcos(self, sd, prefix=""):
        cfg, dt, dev, tc = self.cfg, self.dtype, self.device, self.tc
        fc = sd[prefix + "fcos_head.cls_logits.weight"].shape[1]
        use_gn = cfg.MODEL.FCOS.NORM == "GN"
        per = 3 if use_gn else 2
        P = {"towers": {}, "use_gn": use_gn}
        for name, n in (("share", cfg.MODEL.FCOS.NUM_SHARE_CONVS), ("cls", cfg.MODEL.FCOS.NUM_CLS_CONVS),
                        ("bbox", cfg.MODEL.FCOS.NUM_BOX_CONVS)):
            units = []
            for i in range(n):
                p = prefix + "fcos_head.{}_tower.{}".format(name, per * i)
                conv = packing.conv_bias(sd, p, [fc], 1, 1, not use_gn, dt, dev, tc)
                gn = None
                if use_gn:
                    q = prefix + "fcos_head.{}_tower.{}".format(name, per * i + 1)
                    gn = (sd[q + ".weight"].detach().to(device=dev, dtype=torch.float32).contiguous(),
                          sd[q + ".bias"].detach().to(device=dev, dtype=torch.float32).contiguous())
                units.append((conv, gn))
            P["towers"][name] = units
        P["cls"] = packing.conv_bias(sd, prefix + "fcos_head.cls_logits", [fc], 1, 1, False, dt, dev, tc)
        # bbox_pred (4) and ctrness (1) merged into one 5-column conv; the per-level Scale (fcos.py:19-25,
        # :233-238) is folded into per-level epilogue vectors, ReLU is applied by the decode kernel.
        wb = sd[prefix + "fcos_head.bbox_pred.weight"].detach().float()
        wc = sd[prefix + "fcos_head.ctrness.weight"].detach().float()
        bb = sd[prefix + "fcos_head.bbox_pred.bias"].detach().float()
        bc = sd[prefix + "fcos_head.ctrness.bias"].detach().float()
        # 16 output columns: (l, t, r, b, ctr, 11 x zero) -- keeps the vectorised epilogue / segmented path
        w16 = torch.zeros((16,) + tuple(wb.shape[1:]))
        w16[:4], w16[4:5] = wb, wc
        b16 = torch.zeros(16)
        b16[:4], b16[4:5] = bb, bc
        P["regctr"] = packing.ConvW(w16, [fc], 1, 1, None, b16, False, dt, dev, tc)
        P["reg_scale"] = [float(sd[prefix + "fcos_head.scales.{}.scale".format(l)].detach().float().reshape(()))
                          if cfg.MODEL.FCOS.USE_SCALE else 1.0 for l in range(len(cfg.MODEL.FCOS.FPN_STRIDES))]
        return P

    def run_fcos_head(self, feats, P):
        """feats: list of FMap (p3..p7).  Returns per level (logits f32 FMap [N,H,W,ncls], regctr f32 FMap [N,H,W,5])."""
        self.begin_pass()
        pyr = getattr(self, "_pyramid", None)
        if self.tc and pyr is not None and len(pyr[1]) == len(feats) and all(a is b for a, b in zip(pyr[1], feats)) \
                and P["cls"].cout % 16 == 0:
            return self._run_fcos_head_seg(pyr[0], P)
        out = []
        for l, f in enumerate(feats):
            def tower(x, units, tag):
                for i, (conv, gn) in enumerate(units):
                    x = self.conv("fcos_{}{}_l{}".format(tag, i, l), [x], conv)
                    if gn is not None:
                        wsp = self.buffer("fcos_gnws_l{}".format(l), (lib.gn_workspace_floats(x.n, x.h * x.w, x.c, 32),),
                                          torch.float32, zero=False)
                        lib.groupnorm_relu(x.view, 32, gn[0], gn[1], 1e-5, True, wsp)
                return x
            x = tower(f, P["towers"]["share"], "share")
            ct = tower(x, P["towers"]["cls"], "cls")
            bt = tower(x, P["towers"]["bbox"], "bbox")
            logits = self.conv("fcos_logits_l{}".format(l), [ct], P["cls"], out_dtype=torch.float32, out_halo=0)
            regctr = self.conv("fcos_regctr_l{}".format(l), [bt], P["regctr"], out_dtype=torch.float32, out_halo=0)
            out.append((logits, regctr))
        return out

    def _run_fcos_head_seg(self, pyramid, P):
        """All pyramid levels per launch: the towers' weights are shared across levels (fcos.py:227-238)."""
        n_img = sum(n for _, n, _, _ in pyramid.segs)

        def tower(x, units, tag):
            for i, (conv, gn) in enumerate(units):
                if gn is not None and conv.cout % 256 == 0:
                    # GroupNorm statistics come out of the conv epilogue (fp64 sums per image and 8-channel chunk)
                    st = self.buffer("fcos_gnstats_seg_" + tag, (n_img, conv.cout // 8, 2), torch.float64, zero=False)
                    x = self.conv_seg("fcos_{}{}_seg".format(tag, i), x, conv, stats=st, stats_mode=2)
                    if self.split and conv.cout // 8 <= 256 and 256 % (conv.cout // 8) == 0:
                        # every tower output feeds convolutions only: normalise straight into their [hi | lo] operands
                        nm = "fcos_{}{}_seg_split".format(tag, i)
                        sp = x.like(conv.cout, torch.float16, lambda shape: self.buffer(nm, shape, torch.float16), split=True)
                        lib.groupnorm_apply_seg_split(x.flat, sp.flat, x.segs, 32, gn[0], gn[1], 1e-5, True, st)
                        x = sp
                    else:
                        lib.groupnorm_apply_seg(x.flat, x.segs, 32, gn[0], gn[1], 1e-5, True, st)
                    continue
                x = self.conv_seg("fcos_{}{}_seg".format(tag, i), x, conv)
                if gn is not None:
                    wsp = self.buffer("fcos_gnws_seg_" + tag, (lib.gn_seg_workspace_floats(x.segs, x.c, 32),), torch.float32, zero=False)
                    lib.groupnorm_relu_seg(x.flat, x.segs, 32, gn[0], gn[1], 1e-5, True, wsp)
            return x
        x = tower(pyramid, P["towers"]["share"], "share")
        if self.tower_overlap == 2 and not self.split and pyramid.rows > self.BRANCH_MAX_ROWS:
            # experiment: convolutions of both towers on HIGH-priority streams, the GroupNorm apply passes on default-priority
            # ones -- when an SM frees up, a pending convolution CTA is dispatched before any apply CTA
            hi = [self.side_stream("tower_hi0", -1), self.side_stream("tower_hi1", -1)]
            lo = [self.side_stream("tower_lo0"), self.side_stream("tower_lo1")]
            ev0 = torch.cuda.Event()
            ev0.record()
            outs = []
            for b, (tag, head, hname) in enumerate((("cls", P["cls"], "fcos_logits_seg"), ("bbox", P["regctr"], "fcos_regctr_seg"))):
                hi[b].wait_event(ev0)
                y = x
                for i, (conv, gn) in enumerate(P["towers"][tag]):
                    st = self.buffer("fcos_gnstats_seg_" + tag, (n_img, conv.cout // 8, 2), torch.float64, zero=False)
                    with torch.cuda.stream(hi[b]):
                        y = self.conv_seg("fcos_{}{}_seg".format(tag, i), y, conv, stats=st, stats_mode=2)
                        e1 = torch.cuda.Event()
                        e1.record()
                    lo[b].wait_event(e1)
                    with torch.cuda.stream(lo[b]):
                        lib.groupnorm_apply_seg(y.flat, y.segs, 32, gn[0], gn[1], 1e-5, True, st)
                        e2 = torch.cuda.Event()
                        e2.record()
                    hi[b].wait_event(e2)
                with torch.cuda.stream(hi[b]):
                    outs.append(self.conv_seg(hname, y, head, out_dtype=torch.float32))
            for b in range(2):
                self.join(hi[b])
            logits, regctr = outs
        elif self.branch_streams and (pyramid.rows <= self.BRANCH_MAX_ROWS or self.tower_overlap == 1):
            # the classification and the box branch are independent (fcos.py:227-238): two streams.  Small problems: the launches
            # fill each other's tails.  Large ones (bf16): the GroupNorm apply pass of one tower (HBM-bound, 128-thread CTAs that
            # fit beside a convolution CTA) runs under the other tower's convolution (tensor-pipe-bound)
            if self.split:
                self.split_of(x.flat)                   # both branches read it: make the operand split before they part
            side = self.fork()
            with torch.cuda.stream(side):
                bt = tower(x, P["towers"]["bbox"], "bbox")
                regctr = self.conv_seg("fcos_regctr_seg", bt, P["regctr"], out_dtype=torch.float32)
            ct = tower(x, P["towers"]["cls"], "cls")
            logits = self.conv_seg("fcos_logits_seg", ct, P["cls"], out_dtype=torch.float32)
            self.join(side)
        else:
            ct = tower(x, P["towers"]["cls"], "cls")
            bt = tower(x, P["towers"]["bbox"], "bbox")
            logits = self.conv_seg("fcos_logits_seg", ct, P["cls"], out_dtype=torch.float32)
            regctr = self.conv_seg("fcos_regctr_seg", bt, P["regctr"], out_dtype=torch.float32)
        return [(logits.level(l), regctr.level(l)) for l in range(len(pyramid.segs))]

    def run_fcos_post(self, head_out, cand_cap=None, reg_scale=None):
        """fcos_outputs.py:372-495 on device.  Returns fixed-size detection buffers (dict of tensors):
        boxes [N,R,4], scores [N,R], classes [N,R] (int64), locations [N,R,2], count [N] (int32),
        cand_count [N,L] (int32; > cand_cap means overflow)."""
        cfg = self.cfg
        n = head_out[0][0].n
        L = len(head_out)
        ncls = head_out[0][0].c
        pre = cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST
        post = cfg.MODEL.FCOS.POST_NMS_TOPK_TEST
        cap = cand_cap or max(8192, 2 * pre)
        B = self.buffer
        cb = dict(boxes=B("cand_boxes", (n, L, cap, 4), torch.float32, False), score=B("cand_score", (n, L, cap), torch.float32, False),
                  cls=B("cand_cls", (n, L, cap), torch.int32, False), flat=B("cand_flat", (n, L, cap), torch.int32, False),
                  count=B("cand_count", (n, L), torch.int32))
        cb["count"].zero_()
        cand = lib.cand_buffers(cb["boxes"], cb["score"], cb["cls"], cb["flat"], cb["count"])
        strides = list(cfg.MODEL.FCOS.FPN_STRIDES)
        lib.fcos_decode_levels([lg.view for lg, _ in head_out], [rc.view for _, rc in head_out], strides[:L],
                               [1.0 if reg_scale is None else float(reg_scale[l]) for l in range(L)],
                               float(cfg.MODEL.FCOS.INFERENCE_TH_TEST), bool(cfg.MODEL.FCOS.THRESH_WITH_CTR), cap, cand)
        det = dict(boxes=B("det_boxes", (n, post, 4), torch.float32, False), scores=B("det_scores", (n, post), torch.float32, False),
                   classes=B("det_classes", (n, post), torch.int64, False), locations=B("det_locs", (n, post, 2), torch.float32, False),
                   count=B("det_count", (n,), torch.int32, False))
        level_w, level_s = self.const(("level_w", tuple(h[0].w for h in head_out), tuple(strides[:L])), lambda: (
            torch.tensor([h[0].w for h in head_out], dtype=torch.int32, device=self.device),
            torch.tensor(strides[:L], dtype=torch.int32, device=self.device)))
        wsp = B("select_ws", (lib.fcos_select_workspace(n, L, cap),), torch.uint8, False)
        lib.fcos_select(cand, n, L, cap, level_w, level_s, ncls, min(pre, cap), float(cfg.MODEL.FCOS.NMS_TH), post,
                        lib.det_buffers(det["boxes"], det["scores"], det["classes"], det["locations"], det["count"]), wsp)
        det["cand_count"] = cb["count"]
        det["cand_cap"] = cap
        self._det_gen += 1
        det["gen"] = self._det_gen
        return det

    # =============================================================================================
    # ROI heads: SAG-Mask + MaskIoU
    # =============================================================================================
    def pack_roi_heads(self, sd, prefix=""):
        cfg, dt, dev, tc = self.cfg, self.dtype, self.device, self.tc
        P = {}
        if cfg.MODEL.KEYPOINT_ON:
            # KRCNNConvDeconvUpsampleHead, keypoint_head.py:168-215
            kh = cfg.MODEL.ROI_KEYPOINT_HEAD
            c = sd[prefix + "keypoint_head.conv_fcn1.weight"].shape[1] if len(kh.CONV_DIMS) else sd[prefix + "keypoint_head.score_lowres.weight"].shape[0]
            P["kp_in_ch"] = c
            P["kp_fcn"] = []
            for k, dim in enumerate(kh.CONV_DIMS, 1):
                P["kp_fcn"].append(packing.conv_bias(sd, prefix + "keypoint_head.conv_fcn{}".format(k), [c], 1, 1, True, dt, dev, tc))
                c = dim
            P["kp_deconv"] = packing.deconv4x4s2(sd, prefix + "keypoint_head.score_lowres", dt, dev, tc)
        if not cfg.MODEL.MASK_ON:
            return P
        P.update(self.pack_mask_head(sd, prefix + "mask_head."))
        if cfg.MODEL.MASKIOU_ON:
            P.update(self.pack_maskiou_head(sd, prefix + "maskiou_head.", P["in_ch"], cfg.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION))
        return P

    def pack_mask_head(self, sd, prefix=""):
        """SpatialAttentionMaskHead (sam.py:41-90); keys ``prefix + mask_fcn{k}.*`` ..."""
        cfg, dt, dev, tc = self.cfg, self.dtype, self.device, self.tc
        mh = cfg.MODEL.ROI_MASK_HEAD
        P = {}
        c = sd[prefix + "mask_fcn1.weight"].shape[1] if mh.NUM_CONV > 0 else sd[prefix + "deconv.weight"].shape[0]
        P["in_ch"] = c
        P["mask_fcn"] = []
        for k in range(mh.NUM_CONV):
            P["mask_fcn"].append(packing.conv_bias(sd, prefix + "mask_fcn{}".format(k + 1), [c], 1, 1, True, dt, dev, tc))
            c = mh.CONV_DIM
        P["sam_w"] = sd[prefix + "spatialAtt.conv.weight"].detach().reshape(18).to(device=dev, dtype=torch.float32).contiguous()
        P["deconv"] = packing.deconv2x2(sd, prefix + "deconv", dt, dev, tc)
        pw = sd[prefix + "predictor.weight"].detach()
        P["pred_w"] = pw.reshape(pw.shape[0], pw.shape[1]).to(device=dev, dtype=torch.float32).contiguous()
        P["pred_b"] = sd[prefix + "predictor.bias"].detach().to(device=dev, dtype=torch.float32).contiguous()
        P["pred_key"] = prefix + "predictor"
        return P

    def pack_maskiou_head(self, sd, prefix, in_ch, resolution):
        """MaskIoUHead (maskiou_head.py:64-105); ``in_ch`` = channels of the pooled ROI feature."""
        cfg, dt, dev, tc = self.cfg, self.dtype, self.device, self.tc
        mi = cfg.MODEL.ROI_MASKIOU_HEAD
        P = {"iou_fcn": [], "iou_in_ch": in_ch}
        for k in range(mi.NUM_CONV):
            stride = 2 if k + 1 == mi.NUM_CONV else 1
            wk = prefix + "maskiou_fcn{}".format(k + 1)
            if k == 0:
                # input = cat(roi feature, pooled mask) (maskiou_head.py:109-112); the single mask channel
                # lives in a 16-channel zero-padded buffer so that the TC engine can read it
                w = sd[wk + ".weight"].detach().float()
                wpad = torch.zeros((w.shape[0], in_ch + 16, 3, 3))
                wpad[:, :in_ch + 1] = w
                P["iou_fcn"].append(packing.ConvW(wpad, [in_ch, 16], stride, 1, None, sd[wk + ".bias"], True, dt, dev, tc))
            else:
                P["iou_fcn"].append(packing.conv_bias(sd, wk, [mi.CONV_DIM], stride, 1, True, dt, dev, tc))
        r = resolution // 2
        P["iou_fc1"] = packing.linear(sd, prefix + "maskiou_fc1", True, dt, dev, tc, chw=(mi.CONV_DIM, r, r))
        P["iou_fc2"] = packing.linear(sd, prefix + "maskiou_fc2", True, dt, dev, tc)
        P["iou_out"] = packing.linear(sd, prefix + "maskiou", False, dt, dev, tc)
        return P

    # -- the two heads as stand-alone operators (modeling.roi_heads.SpatialAttentionMaskHead / MaskIoUHead) ----------
    def run_mask_head_logits(self, roi, P, pred_conv):
        """SpatialAttentionMaskHead.forward (sam.py:92-97): roi FMap [R, res, res, C] -> f32 logits [R, 2res, 2res, ncls] for
        ALL classes (the fused plan of ``run_roi_heads`` computes only the class of each ROI)."""
        self.begin_pass()
        x = roi
        for k, w in enumerate(P["mask_fcn"]):
            x = self.conv("sa_mask_fcn{}".format(k + 1), [x], w)
        att = self.fmap("sa_mask_att", x.n, x.h, x.w, x.c)
        lib.spatial_attention(x.view, att.view, P["sam_w"])
        up = self.conv("sa_mask_deconv", [att], P["deconv"], out_mode=1, out_halo=0)
        return self.conv("sa_mask_logits", [up], pred_conv, out_dtype=torch.float32, out_halo=0)

    def run_maskiou_head(self, roi, probs, P, fresh=True):
        """MaskIoUHead.forward (maskiou_head.py:107-120): roi FMap [R, res, res, C] + mask f32 [R, 1, 2res, 2res] -> f32 [R, ncls].
        ``fresh=False``: called inside a pass that already holds the split operands of ``roi`` (run_roi_heads)."""
        if fresh:
            self.begin_pass()
        R, res = roi.n, roi.h
        if isinstance(roi, SharedHaloFMap):
            pm = self.shared_halo_fmap("iou_mask_sh", R, res, res, 16)
        else:
            pm = self.fmap("iou_mask", R, res, res, 16)
        lib.maskiou_prep(probs, pm.view)
        y = None
        nconv = len(P["iou_fcn"])
        for k, w in enumerate(P["iou_fcn"]):
            last = k + 1 == nconv
            srcs = [roi, pm] if k == 0 else [y]
            # the last conv is stride 2: on the TC engine its producer stores phase planes
            # (also when the producer is the first conv over the [roi, mask] concat: NUM_CONV 2 of the Lite recipe -- its
            # stride-2 conv used to fall back to the CUDA-core engine, 0.40 of the 2.0 ms Lite step)
            to_phase = self.tc and k + 2 == nconv
            y = self.conv("iou_fcn{}".format(k + 1), srcs, w, out_halo=0 if last else 1, out_mode=2 if to_phase else 0,
                          to_conv=not last and (to_phase or k + 2 < nconv))
        flat = FMap(y.buf.reshape(R, 1, 1, -1), 0)
        y = self.conv("iou_fc1", [flat], P["iou_fc1"], out_halo=0, to_conv=True)
        y = self.conv("iou_fc2", [y], P["iou_fc2"], out_halo=0, to_conv=True)
        return self.conv("iou_out", [y], P["iou_out"], out_dtype=torch.float32, out_halo=0)

    def image_area(self, image_sizes):
        """Unpadded image areas (pooler.py:70-77) as a device vector, cached per size list."""
        return self.const(("img_area", tuple(image_sizes)), lambda: torch.tensor(
            [float(h * w) for h, w in image_sizes], dtype=torch.float32, device=self.device))

    def run_roi_heads(self, feats, strides, det, image_sizes, P):
        """center_heads.py:413-444 on fixed-size ROI slots [N*R].  feats: list of FMap (p3..p5).
        Returns (mask probs f32 [N*R, 1, 2*res, 2*res], mask_scores f32 [N*R] or None)."""
        cfg = self.cfg
        self.begin_pass()
        mh = cfg.MODEL.ROI_MASK_HEAD
        n, r_cap = det["boxes"].shape[0], det["boxes"].shape[1]
        R = n * r_cap
        res = mh.POOLER_RESOLUTION
        area = self.image_area(image_sizes)
        # bf16 engine: ROI maps with a shared zero frame -- 225 instead of 256 GEMM rows per 14x14 ROI in the seven 3x3 convolutions
        # of the mask / MaskIoU heads (the spatial attention output keeps its own frame: the fused deconv + predictor epilogue
        # wants one ROI per pair of 128-row tiles)
        if self.shared_halo_roi:
            roi = self.shared_halo_fmap("roi_feat_sh", R, res, res, P["in_ch"])
        else:
            roi = self.fmap("roi_feat", R, res, res, P["in_ch"])
        crit = 0 if mh.ASSIGN_CRITERION == "ratio" else 1
        lib.roialign_fpn([f.view for f in feats], strides, det["boxes"], det["count"], n, r_cap, area, crit,
                         int(mh.POOLER_SAMPLING_RATIO), roi.view,
                         workspace=self.buffer("roi_order", (max(R, 1),), torch.int32, False))
        x = roi
        for k, w in enumerate(P["mask_fcn"]):
            x = self.conv("mask_fcn{}".format(k + 1), [x], w, to_conv=k + 1 < len(P["mask_fcn"]))   # the last one feeds the SAM
        att = self.fmap("mask_att", R, res, res, x.c)
        lib.spatial_attention(x.view, att.view, P["sam_w"])
        probs = self.buffer("mask_probs", (R, 1, 2 * res, 2 * res), torch.float32, False)
        ncls = P["pred_w"].shape[0]
        classes = det["classes"].reshape(-1)
        dw = P["deconv"]
        if self.tc and not self.split and dw.w_tc is not None and ((res + 2) * (res + 2)) % 128 == 0 and (dw.cout // 4) % 32 == 0 and dw.cout // 4 <= 256:
            # deconv + ReLU + class-gathered predictor + sigmoid in one launch (the [R, 28, 28, C] tensor is never stored)
            self.conv("mask_deconv_predict", [att], dw, out_mode=3, out=FMap(probs.view(R, 2 * res, 2 * res, 1), 0),
                      pred=(P["pred_w"], P["pred_b"], classes, ncls))
        else:
            up = self.conv("mask_deconv", [att], dw, out_mode=1, out_halo=0)
            lib.mask_predict(up.view, P["pred_w"], P["pred_b"], classes, ncls, probs)
        mask_scores = None
        if cfg.MODEL.MASKIOU_ON:
            y = self.run_maskiou_head(roi, probs, P, fresh=False)
            mask_scores = self.buffer("mask_scores", (R,), torch.float32, False)
            lib.maskiou_score(y.buf, R, y.c, classes, det["scores"].reshape(-1), mask_scores)
        return probs, mask_scores

    def run_keypoint_head(self, feats, strides, det, image_sizes, P):
        """center_heads.py:551-553 + keypoint_head.py:95-120 on fixed-size ROI slots.  feats: list of FMap
        (ROI_KEYPOINT_HEAD.IN_FEATURES).  Returns f32 [N, R, K, 4] = (x, y, logit, score); pred_keypoints = [..., (0, 1, 3)]."""
        cfg = self.cfg
        self.begin_pass()
        kh = cfg.MODEL.ROI_KEYPOINT_HEAD
        n, r_cap = det["boxes"].shape[0], det["boxes"].shape[1]
        R = n * r_cap
        res, nk = kh.POOLER_RESOLUTION, kh.NUM_KEYPOINTS
        roi = self.fmap("kp_roi_feat", R, res, res, P["kp_in_ch"])
        lib.roialign_fpn([f.view for f in feats], strides, det["boxes"], det["count"], n, r_cap, self.image_area(image_sizes),
                         0 if kh.ASSIGN_CRITERION == "ratio" else 1, int(kh.POOLER_SAMPLING_RATIO), roi.view,
                         workspace=self.buffer("kp_roi_order", (max(R, 1),), torch.int32, False))
        x = roi
        for k, w in enumerate(P["kp_fcn"]):
            x = self.conv("kp_fcn{}".format(k + 1), [x], w)
        # score_lowres as a 3x3 conv to 4 * K phase columns, fp32 out (logits are not rounded to bf16)
        low = self.conv("kp_score_lowres", [x], P["kp_deconv"], out_dtype=torch.float32, out_halo=0)
        out = self.buffer("kp_out", (n, r_cap, nk, 4), torch.float32, False)
        lib.keypoints_decode(low.buf, det["boxes"], det["count"], n, r_cap, res, nk, out)
        return out

    # =============================================================================================
    # input / output side
    # =============================================================================================
    def preprocess(self, images, size_divisibility=32, fused_stem=True):
        """GeneralizedRCNN.preprocess_image [d2] (deploy_utils.py:76-98): normalise + pad to /32.
        ``images``: list of CHW device tensors (float32 or uint8, BGR)."""
        cfg = self.cfg
        sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in images]
        hp = max(s[0] for s in sizes)
        wp = max(s[1] for s in sizes)
        hp = (hp + size_divisibility - 1) // size_divisibility * size_divisibility
        wp = (wp + size_divisibility - 1) // size_divisibility * size_divisibility
        if self.tc and fused_stem and len(cfg.MODEL.PIXEL_MEAN) == 3:
            if self.stem_variant >= 1 and len({im.dtype for im in images}) == 1 and images[0].dtype in (torch.uint8, torch.float32):
                # normalise + pad + stem_1 in ONE pass (run_backbone -> stem1_fused): nothing to do here
                return RawInput([im.contiguous() for im in images], hp, wp), sizes
            if not self.split:
                return self._im2col(images, hp, wp), sizes
        x = self.fmap("input", len(images), hp, wp, 3)
        for i, im in enumerate(images):
            lib.preprocess_image(im.contiguous(), cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD, x.view, i)
        return x, sizes

    def _im2col(self, images, hp, wp):
        """normalise + pad + im2col of stem_1 in one pass (the TC engine then runs stem_1 as a 1x1 conv, K = 32)."""
        cfg = self.cfg
        x = self.fmap("input_im2col", len(images), hp // 2, wp // 2, 32)
        if len({im.dtype for im in images}) == 1:
            lib.preprocess_im2col_batch([im.contiguous() for im in images], cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD, hp, wp, x.view)
        else:
            for i, im in enumerate(images):
                lib.preprocess_im2col(im.contiguous(), cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD, hp, wp, x.view, i)
        return x

    def stem1_fused(self, name, raw, P):
        """stem_1 straight from the raw images (csrc/stem.cu); falls back to im2col + 1x1 GEMM when the packed fused weights
        do not exist (stem_1 with other than 64 output channels)."""
        if "stem1_fused" not in P or (self.split and isinstance(P["stem"][1], tuple)):
            if self.split:                                    # fp32 engine, stem_1 with other than 64 channels or a depthwise stem_2
                                                              # (reads fp32, not the [hi | lo] pair): the unfused path
                x = self.fmap("input", raw.n, raw.hp, raw.wp, 3)
                for i, im in enumerate(raw.images):
                    lib.preprocess_image(im, self.cfg.MODEL.PIXEL_MEAN, self.cfg.MODEL.PIXEL_STD, x.view, i)
                return self.conv(name, [x], P["stem"][0])
            return self.conv(name, [self._im2col(raw.images, raw.hp, raw.wp)], P["stem1_im2col"])
        w30, scale, shift = P["stem1_fused"]
        if self.split:
            out = SplitFMap(self.buffer(name, (raw.n, raw.hp // 2 + 2, raw.wp // 2 + 2, 128), torch.float16), 1)
        else:
            out = self.fmap(name, raw.n, raw.hp // 2, raw.wp // 2, 64)
        lib.stem1_fused_batch(raw.images, self.cfg.MODEL.PIXEL_MEAN, self.cfg.MODEL.PIXEL_STD, raw.hp, raw.wp, w30, scale, shift, True, out.view)
        return out

    def paste(self, probs, boxes, out_h, out_w, image_size, threshold=0.5):
        """detector_postprocess [d2] for the slots of one image: returns (boxes', valid u8, masks u8 [R,H,W])."""
        r = boxes.shape[0]
        sx, sy = out_w / image_size[1], out_h / image_size[0]
        b2 = torch.empty_like(boxes)
        valid = torch.empty((r,), dtype=torch.uint8, device=self.device)
        lib.scale_clip_boxes(boxes, b2, valid, r, sx, sy, float(out_w), float(out_h))
        masks = torch.empty((r, out_h, out_w), dtype=torch.uint8, device=self.device)
        m = probs.shape[-1]
        lib.paste_masks(probs, b2, valid, masks, r, m, out_h, out_w, threshold)
        return b2, valid, masks

    def rescale_boxes(self, det_boxes, sizes, out_sizes, det_count=None):
        """Box half of detector_postprocess for the whole batch (one launch): det_boxes [n, r_cap, 4] ->
        (boxes' [n, r_cap, 4], valid u8 [n, r_cap]) in engine-owned buffers.  With ``det_count`` the slots that hold no
        detection come out invalid (no mask is pasted into them)."""
        n, r_cap = det_boxes.shape[0], det_boxes.shape[1]
        boxes = self.buffer("pp_boxes", (n, r_cap, 4), torch.float32, False)
        valid = self.buffer("pp_valid", (n, r_cap), torch.uint8, False)
        lib.scale_clip_boxes_batch(det_boxes, boxes, valid, n, r_cap, self.pp_params(sizes, out_sizes), det_count)
        return boxes, valid

    def pp_params(self, sizes, out_sizes):
        """Per image (scale_x, scale_y, out_w, out_h) of detector_postprocess [d2], cached per (sizes, out_sizes)."""
        return self.const(("pp_params", tuple(sizes), tuple(out_sizes)), lambda: torch.tensor(
            [[ow / sz[1], oh / sz[0], float(ow), float(oh)] for (oh, ow), sz in zip(out_sizes, sizes)],
            dtype=torch.float32, device=self.device))

    def paste_batch(self, probs, boxes, valid, out_sizes, masks=None, threshold=0.5, dtype=torch.uint8):
        """Mask half of detector_postprocess: probs [n*r_cap, 1, m, m] -> list of [r_cap, oh, ow] 0/1 byte masks
        (one launch when every image has the same output size).  ``masks``: optional preallocated
        [n, r_cap, oh, ow] buffer (uniform sizes only)."""
        n, r_cap = boxes.shape[0], boxes.shape[1]
        m = probs.shape[-1]
        if len(set(out_sizes)) == 1 and n * r_cap <= 65535:
            oh, ow = out_sizes[0]
            if masks is None:
                masks = torch.empty((n, r_cap, oh, ow), dtype=dtype, device=self.device)
            lib.paste_masks(probs, boxes, valid, masks, n * r_cap, m, oh, ow, threshold)
            return [masks[i] for i in range(n)]
        out = []
        for i, (oh, ow) in enumerate(out_sizes):
            mk = torch.empty((r_cap, oh, ow), dtype=dtype, device=self.device)
            lib.paste_masks(probs[i * r_cap:(i + 1) * r_cap], boxes[i], valid[i], mk, r_cap, m, oh, ow, threshold)
            out.append(mk)
        return out

    # =============================================================================================
    # CUDA-graph replay of a launch plan
    # =============================================================================================
    def graphed(self, key, fn, keep=None):
        """Run ``fn`` (a sequence of C-ABI launches on engine-owned buffers, no host sync) through a CUDA graph.

        The first ``graph_after`` calls with a given ``key`` run eagerly (they also allocate every buffer and fill the
        lazily cached constants); the next one captures the plan, and every later call replays the capture -- the ~100
        launches of a step cost one graph launch on the host.  A key that is seen only once (a data set whose images
        all differ in shape) therefore never pays for a capture.  Returns whatever ``fn`` returned (the same engine-owned
        buffers every call).  Captured graphs are kept least-recently-used up to ``graph_cache``; an entry keeps alive
        every buffer the plan touched plus ``keep`` (the packed weights whose addresses are baked into it), and
        ``key`` must carry the ``graph_token`` of every module whose weights the plan reads so that re-packed weights
        can never be replayed (``drop_graphs``).  ``CM2_GRAPH=0`` disables."""
        if not self.use_graphs:
            return fn()
        ent = self._graphs.get(key)
        if ent is None:
            seen = self._graph_seen.get(key, 0)
            if seen < self.graph_after:
                self._graph_seen[key] = seen + 1
                self._graph_seen.move_to_end(key)
                while len(self._graph_seen) > 64 * self.graph_cache:
                    self._graph_seen.popitem(last=False)
                return fn()
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            c0 = lib.launch_count
            self._recording = []
            try:
                with torch.cuda.graph(g):
                    out = fn()
                touched = self._recording
            finally:
                self._recording = None
            ent = (g, out, lib.launch_count - c0, (touched, keep))
            self._graphs[key] = ent
            self._graph_seen.pop(key, None)
            while len(self._graphs) > self.graph_cache:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        g, out, launches = ent[0], ent[1], ent[2]
        g.replay()
        lib._count(launches)
        return out
