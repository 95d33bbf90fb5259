"""ctypes binding of ``libcm2.so`` (``include/cm2.h``).

There is no CPU fallback and no torch-op fallback: if the library is missing, or a call fails, a
``RuntimeError`` is raised.  All functions enqueue on ``torch.cuda.current_stream()``.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcm2.so")

F32, BF16, U8, F16 = 0, 1, 2, 3
ENGINE_SIMT, ENGINE_TC = 0, 1
MAX_SRC = 8

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.uint8: U8, torch.float16: F16}


class Act(C.Structure):
    """cm2_act: pitched NHWC view."""
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


MAX_SEG = 8


class Seg(C.Structure):
    """cm2_seg: one map of a segmented halo tensor."""
    _fields_ = [("row0", C.c_int64), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("halo", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("out_dtype", C.c_int32), ("engine", C.c_int32), ("num_src", C.c_int32),
                ("src", Act * MAX_SRC),
                ("cout", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("weight", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("relu", C.c_int32), ("in_relu", C.c_int32),
                ("residual", Act), ("res_mode", C.c_int32), ("out_mode", C.c_int32),
                ("out", Act), ("stats", C.c_void_p), ("src_phase", C.c_int32), ("num_seg", C.c_int32),
                ("seg", Seg * MAX_SEG), ("stats_mode", C.c_int32), ("pred_ncls", C.c_int32),
                ("pred_w", C.c_void_p), ("pred_b", C.c_void_p), ("pred_cls", C.c_void_p),
                ("splitk", C.c_int32), ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64)]


class CandBuffers(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("score", C.c_void_p), ("cls", C.c_void_p), ("flat", C.c_void_p),
                ("count", C.c_void_p)]


class DetBuffers(C.Structure):
    _fields_ = [("boxes", C.c_void_p), ("scores", C.c_void_p), ("classes", C.c_void_p), ("locations", C.c_void_p),
                ("count", C.c_void_p)]


# every symbol include/cm2.h declares: name -> (restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_AP = C.POINTER(Act)
SYMBOLS = {
    "cm2_version": (_I, []),
    "cm2_last_error": (C.c_char_p, []),
    "cm2_device_info": (_I, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cm2_conv2d": (_I, [C.POINTER(ConvDesc), _P]),
    "cm2_split_f16x2": (_I, [_P, _P, _L, _I, _P]),
    "cm2_conv_tc_klen": (_L, [_I, _I, _I, C.POINTER(_I)]),
    "cm2_conv_tc_supported": (_I, [C.POINTER(ConvDesc)]),
    "cm2_preprocess_image": (_I, [_P, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _AP, _I, _I, _P]),
    "cm2_preprocess_im2col": (_I, [_P, _I, _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _AP, _I, _P]),
    "cm2_stem1_fused_batch": (_I, [C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _P, _P, _P, _I,
                                    _AP, _I, _P]),
    "cm2_stem1_fused_split_batch": (_I, [C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _P, _P, _P,
                                          _I, _AP, _I, _P]),
    "cm2_preprocess_im2col_batch": (_I, [C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, _I, _I, _I, C.POINTER(_F), C.POINTER(_F), _AP,
                                         _I, _P]),
    "cm2_resize_pil_u8": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _I, _P]),
    "cm2_rle_count": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "cm2_rle_write": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "cm2_rle_encode": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P]),
    "cm2_phase_split": (_I, [_AP, _AP, _I, _I, _P]),
    "cm2_maxpool3x3s2_ceil": (_I, [_AP, _AP, _I, _P]),
    "cm2_dwconv3x3": (_I, [_AP, _AP, _I, _P, _I, _P]),
    "cm2_ese_pool_chunks": (_I, [_I]),
    "cm2_ese_pool": (_I, [_AP, _I, _P, _P, _P]),
    "cm2_ese_gate": (_I, [_P, _F, _P, _P, _P, _I, _I, _P]),
    "cm2_ese_apply": (_I, [_AP, _P, _AP, _AP, _I, _P]),
    "cm2_gn_workspace_floats": (_L, [_I, _I, _I, _I]),
    "cm2_groupnorm_relu": (_I, [_AP, _I, _I, _P, _P, _F, _I, _P, _P]),
    "cm2_gn_seg_workspace_floats": (_L, [_I, C.POINTER(Seg), _I, _I]),
    "cm2_groupnorm_relu_seg": (_I, [_P, _I, _I, _I, C.POINTER(Seg), _I, _P, _P, _F, _I, _P, _P]),
    "cm2_groupnorm_apply_seg": (_I, [_P, _I, _I, _I, C.POINTER(Seg), _I, _P, _P, _F, _I, _P, _P]),
    "cm2_groupnorm_apply_seg_split": (_I, [_P, _P, _I, _I, C.POINTER(Seg), _I, _P, _P, _F, _I, _P, _P]),
    "cm2_ese_gate_f64": (_I, [_P, C.c_double, _P, _P, _P, _I, _I, _P]),
    "cm2_ese_apply_pool": (_I, [_AP, _P, _AP, _AP, _AP, _I, _P]),
    "cm2_relu": (_I, [_AP, _AP, _I, _P]),
    "cm2_fcos_decode": (_I, [_AP, _AP, _I, _F, _F, _I, _I, _I, _I, C.POINTER(CandBuffers), _P]),
    "cm2_fcos_decode_levels": (_I, [_AP, _AP, C.POINTER(_I), C.POINTER(_F), _I, _F, _I, _I, C.POINTER(CandBuffers), _P]),
    "cm2_fcos_select_workspace": (_L, [_I, _I, _I]),
    "cm2_fcos_select": (_I, [C.POINTER(CandBuffers), _I, _I, _I, C.POINTER(_I), C.POINTER(_I), _I, _I, _F, _I,
                             C.POINTER(DetBuffers), _P, _P]),
    "cm2_roialign_fpn": (_I, [_AP, C.POINTER(_I), _I, _I, _P, _P, _I, _I, _P, _I, _I, _AP, _P, _P, _P]),
    "cm2_spatial_attention": (_I, [_AP, _AP, _I, _P, _P]),
    "cm2_mask_predict": (_I, [_AP, _I, _P, _P, _P, _I, _P, _P]),
    "cm2_maskiou_prep": (_I, [_P, _AP, _I, _P]),
    "cm2_maskiou_score": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "cm2_keypoints_decode": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "cm2_scale_clip_boxes": (_I, [_P, _P, _P, _I, _F, _F, _F, _F, _P]),
    "cm2_scale_clip_boxes_batch": (_I, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "cm2_paste_masks": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P]),
    "cm2_pack_records": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
}

_lib = None
launch_count = 0          # incremented by every kernel-launching call (bench.py reports it)


def load():
    """Load ``libcm2.so`` and bind every symbol of ``include/cm2.h``; raise if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("{} not found: run `python -m centermask2_b200.build` (there is no fallback path)".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise RuntimeError("libcm2.so does not export {}".format(name))
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().cm2_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise RuntimeError("{} failed ({}): {}".format(what, rc, last_error()))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t):
    return _DTYPES[t.dtype if isinstance(t, torch.Tensor) else t]


def act(t):
    """``cm2_act`` of a 4-D NHWC tensor/view [n, h, w, c] whose channel stride is 1."""
    if t is None:
        return Act()
    assert t.dim() == 4 and (t.stride(3) == 1 or t.shape[3] == 1), (t.shape, t.stride())
    return Act(t.data_ptr(), t.shape[0], t.shape[1], t.shape[2], t.shape[3], t.stride(0), t.stride(1), t.stride(2))


def flat_act(t):
    """``cm2_act`` of a flat [rows, c] buffer or a channel slice of one (segmented tensors: data, c and the row pitch
    ``sw`` are read)."""
    assert t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1], (t.shape, t.stride())
    pitch = t.stride(0)
    return Act(t.data_ptr(), 1, 1, t.shape[0], t.shape[1], t.shape[0] * pitch, t.shape[0] * pitch, pitch)


def seg_array(segs):
    """segs: (row0, n, h, w) per segment; ``segs.halo`` (engine.SegList) = 1 for segments whose images share their zero frame."""
    halo = int(getattr(segs, "halo", 0))
    return (Seg * len(segs))(*[Seg(r, n, h, w, halo) for r, n, h, w in segs])


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _count(k=1):
    global launch_count
    launch_count += k


# ------------------------------------------------------------------------------------------------
# thin wrappers (one per entry point)
# ------------------------------------------------------------------------------------------------
def conv2d(srcs, weight, out, cout, k, stride, pad, scale=None, shift=None, relu=False, in_relu=False,
           residual=None, res_mode=0, out_mode=0, engine=ENGINE_SIMT, stats=None, stats_mode=0, probe=False, src_phase=False,
           segs=None, pred=None, splitk=0, splitk_ws=None):
    """Enqueue one convolution.  ``pred`` = (pred_w, pred_b, classes, ncls) for out_mode 3.  With ``probe=True`` (TC engine) the descriptor is first checked with
    ``cm2_conv_tc_supported``; returns False without launching if the engine does not take it."""
    d = ConvDesc()
    d.dtype = dtype_code(srcs[0])
    d.out_dtype = dtype_code(out)
    d.engine = engine
    d.num_src = len(srcs)
    if segs is not None:
        # segmented halo tensors: srcs / out are flat [rows, c] buffers sharing the segment table
        d.num_seg = len(segs)
        for i, (row0, n, h, w) in enumerate(segs):
            d.seg[i] = Seg(row0, n, h, w, int(getattr(segs, "halo", 0)))
        for i, s in enumerate(srcs):
            d.src[i] = flat_act(s)
    else:
        for i, s in enumerate(srcs):
            d.src[i] = act(s)
    d.cout, d.kh, d.kw, d.stride, d.pad = cout, k, k, stride, pad
    d.weight = weight.data_ptr()
    d.scale = 0 if scale is None else scale.data_ptr()
    d.shift = 0 if shift is None else shift.data_ptr()
    d.relu, d.in_relu = int(relu), int(in_relu)
    d.residual = act(residual)
    d.res_mode = res_mode if residual is not None else 0
    d.out_mode = out_mode
    d.out = flat_act(out) if segs is not None else act(out)
    d.stats = 0 if stats is None else stats.data_ptr()
    d.stats_mode = stats_mode if stats is not None else 0
    d.src_phase = int(src_phase)
    if pred is not None:
        d.pred_w, d.pred_b, d.pred_cls, d.pred_ncls = pred[0].data_ptr(), pred[1].data_ptr(), pred[2].data_ptr(), int(pred[3])
    if splitk >= 2:
        d.splitk, d.splitk_ws, d.splitk_ws_bytes = int(splitk), splitk_ws.data_ptr(), splitk_ws.numel() * splitk_ws.element_size()
    if probe and not load().cm2_conv_tc_supported(C.byref(d)):
        return False
    check(load().cm2_conv2d(C.byref(d), stream()), "cm2_conv2d")
    _count()
    return True


def split_f16x2(x, out):
    """fp32 buffer [..., c] (contiguous, halo included) -> f16 [..., 2c] = [hi | lo] for the split-precision TC convolution."""
    assert x.dtype == torch.float32 and out.dtype == torch.float16 and x.is_contiguous() and out.is_contiguous()
    c = x.shape[-1]
    assert out.shape[-1] == 2 * c and out.numel() == 2 * x.numel(), (tuple(x.shape), tuple(out.shape))
    check(load().cm2_split_f16x2(ptr(x), ptr(out), x.numel() // c, c, stream()), "cm2_split_f16x2")
    _count()


def conv_tc_klen(k, src_c):
    arr = (C.c_int32 * len(src_c))(*src_c)
    return int(load().cm2_conv_tc_klen(k, k, len(src_c), arr))


def preprocess_image(img, mean, std, out, index):
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    a = act(out)
    check(load().cm2_preprocess_image(ptr(img), dtype_code(img), img.shape[1], img.shape[2], m, s, C.byref(a),
                                      dtype_code(out), index, stream()), "cm2_preprocess_image")
    _count()


def preprocess_im2col(img, mean, std, hp, wp, out, index):
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    a = act(out)
    check(load().cm2_preprocess_im2col(ptr(img), dtype_code(img), img.shape[1], img.shape[2], hp, wp, m, s, C.byref(a),
                                       index, stream()), "cm2_preprocess_im2col")
    _count()


def preprocess_im2col_batch(imgs, mean, std, hp, wp, out, index0=0):
    """All images (same dtype; [3, h, w] planar, contiguous) in one launch."""
    n = len(imgs)
    ptrs = (_P * n)(*[im.data_ptr() for im in imgs])
    hs = (_I * n)(*[im.shape[1] for im in imgs])
    ws = (_I * n)(*[im.shape[2] for im in imgs])
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    a = act(out)
    check(load().cm2_preprocess_im2col_batch(ptrs, hs, ws, n, dtype_code(imgs[0]), hp, wp, m, s, C.byref(a), index0, stream()),
          "cm2_preprocess_im2col_batch")
    _count((n + 31) // 32)


def stem1_fused_batch(imgs, mean, std, hp, wp, w30, scale, shift, relu, out, index0=0):
    """normalise + pad + stem_1 (3x3 / s2, 3 -> 64) + scale / shift + ReLU of all images in one pass (csrc/stem.cu)."""
    n = len(imgs)
    split = w30.dtype == torch.float16
    if split:        # fp32 engine: W_hi / W_lo, output = the [hi | lo] f16 operand pair of stem_2
        assert tuple(w30.shape) == (2, 64, 32) and w30.is_contiguous() and out.dtype == torch.float16 and out.shape[3] == 128
    else:
        assert w30.dtype == torch.bfloat16 and tuple(w30.shape) == (64, 32) and w30.is_contiguous() and out.dtype == torch.bfloat16
    ptrs = (_P * n)(*[im.data_ptr() for im in imgs])
    hs = (_I * n)(*[im.shape[1] for im in imgs])
    ws = (_I * n)(*[im.shape[2] for im in imgs])
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    a = act(out)
    fn = load().cm2_stem1_fused_split_batch if split else load().cm2_stem1_fused_batch
    check(fn(ptrs, hs, ws, n, dtype_code(imgs[0]), hp, wp, m, s, ptr(w30), ptr(scale), ptr(shift), int(relu), C.byref(a), index0, stream()),
          "cm2_stem1_fused_split_batch" if split else "cm2_stem1_fused_batch")
    _count((n + 31) // 32)


def resize_pil_u8(src, tmp, dst, h, w, c, oh, ow, bounds_x, kk_x, bounds_y, kk_y, chw):
    check(load().cm2_resize_pil_u8(ptr(src), ptr(tmp), ptr(dst), h, w, c, oh, ow, ptr(bounds_x), ptr(kk_x), kk_x.shape[1],
                                   ptr(bounds_y), ptr(kk_y), kk_y.shape[1], int(chw), stream()), "cm2_resize_pil_u8")
    _count(2)


def rle_count(masks, col_count, col_offset, total):
    r, h, w = masks.shape
    check(load().cm2_rle_count(ptr(masks), r, h, w, ptr(col_count), ptr(col_offset), ptr(total), stream()), "cm2_rle_count")
    _count(2)


def rle_write(masks, col_offset, total, mask_offset, positions, runs):
    r, h, w = masks.shape
    check(load().cm2_rle_write(ptr(masks), r, h, w, ptr(col_offset), ptr(total), ptr(mask_offset), ptr(positions), ptr(runs), stream()),
          "cm2_rle_write")
    _count(2)


def rle_encode(masks, col_count, col_offset, total, mask_offset, positions, runs, boxes=None, valid=None):
    """Count + device-side offsets + write in one call (no host sync); ``runs.numel()`` is the capacity.  ``boxes`` /
    ``valid``: what the masks were pasted with -- only the box windows are scanned then."""
    r, h, w = masks.shape
    assert mask_offset.dtype == torch.int64 and mask_offset.numel() >= r + 1 and positions.numel() >= runs.numel()
    check(load().cm2_rle_encode(ptr(masks), r, h, w, ptr(col_count), ptr(col_offset), ptr(total), ptr(mask_offset), ptr(positions),
                                ptr(runs), runs.numel(), ptr(boxes), ptr(valid), stream()), "cm2_rle_encode")
    _count(5)


def phase_split(x, out_plane0, relu=False):
    a, o = act(x), act(out_plane0)
    check(load().cm2_phase_split(C.byref(a), C.byref(o), dtype_code(x), int(relu), stream()), "cm2_phase_split")
    _count()


def maxpool3x3s2_ceil(x, out):
    a, b = act(x), act(out)
    check(load().cm2_maxpool3x3s2_ceil(C.byref(a), C.byref(b), dtype_code(x), stream()), "cm2_maxpool3x3s2_ceil")
    _count()


def dwconv3x3(x, out, w9c, stride):
    a, b = act(x), act(out)
    check(load().cm2_dwconv3x3(C.byref(a), C.byref(b), dtype_code(x), ptr(w9c), stride, stream()), "cm2_dwconv3x3")
    _count()


def ese_pool_chunks(hw):
    return int(load().cm2_ese_pool_chunks(hw))


def ese_pool(x, workspace, pooled):
    a = act(x)
    check(load().cm2_ese_pool(C.byref(a), dtype_code(x), ptr(workspace), ptr(pooled), stream()), "cm2_ese_pool")
    _count(2)


def ese_gate(pooled, inv_count, fc_w, fc_b, gate, n, c):
    check(load().cm2_ese_gate(ptr(pooled), inv_count, ptr(fc_w), ptr(fc_b), ptr(gate), n, c, stream()), "cm2_ese_gate")
    _count()


def ese_apply(x, gate, identity, out):
    a, i, o = act(x), act(identity), act(out)
    check(load().cm2_ese_apply(C.byref(a), ptr(gate), C.byref(i), C.byref(o), dtype_code(x), stream()), "cm2_ese_apply")
    _count()


def gn_workspace_floats(n, hw, c, groups):
    return int(load().cm2_gn_workspace_floats(n, hw, c, groups))


def groupnorm_relu(x, groups, gamma, beta, eps, relu, workspace):
    a = act(x)
    check(load().cm2_groupnorm_relu(C.byref(a), dtype_code(x), groups, ptr(gamma), ptr(beta), eps, int(relu),
                                    ptr(workspace), stream()), "cm2_groupnorm_relu")
    _count(3)


def gn_seg_workspace_floats(segs, c, groups):
    return int(load().cm2_gn_seg_workspace_floats(len(segs), seg_array(segs), c, groups))


def groupnorm_relu_seg(flat, segs, groups, gamma, beta, eps, relu, workspace):
    check(load().cm2_groupnorm_relu_seg(ptr(flat), dtype_code(flat), flat.shape[1], len(segs), seg_array(segs), groups,
                                        ptr(gamma), ptr(beta), eps, int(relu), ptr(workspace), stream()),
          "cm2_groupnorm_relu_seg")
    _count(3)


def groupnorm_apply_seg(flat, segs, groups, gamma, beta, eps, relu, stats):
    check(load().cm2_groupnorm_apply_seg(ptr(flat), dtype_code(flat), flat.shape[1], len(segs), seg_array(segs), groups,
                                         ptr(gamma), ptr(beta), eps, int(relu), ptr(stats), stream()),
          "cm2_groupnorm_apply_seg")
    _count()


def groupnorm_apply_seg_split(flat, out_split, segs, groups, gamma, beta, eps, relu, stats):
    """fp32 segmented tensor -> GroupNorm + ReLU written as the [hi | lo] f16 operand tensor [rows, 2c] of the next conv."""
    assert flat.dtype == torch.float32 and out_split.dtype == torch.float16 and out_split.shape[1] == 2 * flat.shape[1]
    check(load().cm2_groupnorm_apply_seg_split(ptr(flat), ptr(out_split), flat.shape[1], len(segs), seg_array(segs), groups,
                                               ptr(gamma), ptr(beta), eps, int(relu), ptr(stats), stream()),
          "cm2_groupnorm_apply_seg_split")
    _count()


def ese_gate_f64(sums, inv_count, fc_w, fc_b, gate, n, c):
    check(load().cm2_ese_gate_f64(ptr(sums), inv_count, ptr(fc_w), ptr(fc_b), ptr(gate), n, c, stream()), "cm2_ese_gate_f64")
    _count()


def ese_apply_pool(x, gate, identity, full, pool):
    a, i, f, p = act(x), act(identity), act(full), act(pool)
    check(load().cm2_ese_apply_pool(C.byref(a), ptr(gate), C.byref(i), C.byref(f), C.byref(p), dtype_code(x), stream()),
          "cm2_ese_apply_pool")
    _count()


def relu(x, out):
    a, o = act(x), act(out)
    check(load().cm2_relu(C.byref(a), C.byref(o), dtype_code(x), stream()), "cm2_relu")
    _count()


def cand_buffers(boxes, score, cls, flat, count):
    return CandBuffers(boxes.data_ptr(), score.data_ptr(), cls.data_ptr(), flat.data_ptr(), count.data_ptr())


def det_buffers(boxes, scores, classes, locations, count):
    return DetBuffers(boxes.data_ptr(), scores.data_ptr(), classes.data_ptr(), locations.data_ptr(), count.data_ptr())


def fcos_decode(logits, regctr, stride, reg_scale, thresh, thresh_with_ctr, level, num_levels, cap, cand):
    a, b = act(logits), act(regctr)
    check(load().cm2_fcos_decode(C.byref(a), C.byref(b), stride, reg_scale, thresh, int(thresh_with_ctr), level, num_levels, cap,
                                 C.byref(cand), stream()), "cm2_fcos_decode")
    _count()


def fcos_decode_levels(logits, regctrs, strides, reg_scales, thresh, thresh_with_ctr, cap, cand):
    L = len(logits)
    la = (Act * L)(*[act(t) for t in logits])
    ra = (Act * L)(*[act(t) for t in regctrs])
    st = (C.c_int32 * L)(*strides)
    rs = (C.c_float * L)(*reg_scales)
    check(load().cm2_fcos_decode_levels(la, ra, st, rs, L, thresh, int(thresh_with_ctr), cap, C.byref(cand), stream()),
          "cm2_fcos_decode_levels")
    _count()


def fcos_select_workspace(n, num_levels, cap):
    return int(load().cm2_fcos_select_workspace(n, num_levels, cap))


def fcos_select(cand, n, num_levels, cap, level_w, level_stride, ncls, pre_topk, nms_thresh, post_topk, det, workspace):
    check(load().cm2_fcos_select(C.byref(cand), n, num_levels, cap, C.cast(ptr(level_w), C.POINTER(C.c_int32)),
                                 C.cast(ptr(level_stride), C.POINTER(C.c_int32)), ncls, pre_topk, nms_thresh, post_topk,
                                 C.byref(det), ptr(workspace), stream()), "cm2_fcos_select")
    _count(2)


_roi_workspace = {}


def roialign_fpn(feats, strides, boxes, det_count, n, r_cap, image_area, crit, sampling_ratio, out, level_out=None,
                 workspace=None):
    """``workspace``: int32 device tensor of >= n * r_cap elements (the launch order of the column kernel); when omitted
    one is kept per (device, stream) -- a buffer shared between streams could be overwritten by a concurrent launch."""
    arr = (Act * len(feats))(*[act(f) for f in feats])
    st = (C.c_int32 * len(strides))(*strides)
    o = act(out)
    if workspace is None:
        key = (boxes.device, torch.cuda.current_stream().cuda_stream)
        workspace = _roi_workspace.get(key)
        if workspace is None or workspace.numel() < n * r_cap:
            workspace = _roi_workspace[key] = torch.empty((max(n * r_cap, 1024),), dtype=torch.int32, device=boxes.device)
    assert workspace.dtype == torch.int32 and workspace.numel() >= n * r_cap
    check(load().cm2_roialign_fpn(arr, st, len(feats), dtype_code(feats[0]), ptr(boxes), ptr(det_count), n, r_cap,
                                  ptr(image_area), crit, sampling_ratio, C.byref(o), ptr(level_out), ptr(workspace), stream()),
          "cm2_roialign_fpn")
    _count(2 if int(os.environ.get("CM2_ROIALIGN_VARIANT", "2")) >= 2 else 1)       # order kernel + ROIAlign kernel


def spatial_attention(x, out, w18):
    a, o = act(x), act(out)
    check(load().cm2_spatial_attention(C.byref(a), C.byref(o), dtype_code(x), ptr(w18), stream()), "cm2_spatial_attention")
    _count()


def mask_predict(x, wp, bp, classes, ncls, probs):
    a = act(x)
    check(load().cm2_mask_predict(C.byref(a), dtype_code(x), ptr(wp), ptr(bp), ptr(classes), ncls, ptr(probs), stream()),
          "cm2_mask_predict")
    _count()


def maskiou_prep(probs, out):
    o = act(out)
    check(load().cm2_maskiou_prep(ptr(probs), C.byref(o), dtype_code(out), stream()), "cm2_maskiou_prep")
    _count()


def maskiou_score(iou, r, ncls, classes, scores, mask_scores):
    check(load().cm2_maskiou_score(ptr(iou), dtype_code(iou), r, ncls, ptr(classes), ptr(scores), ptr(mask_scores),
                                   stream()), "cm2_maskiou_score")
    _count()


def keypoints_decode(lowres, boxes, det_count, n, r_cap, res, num_keypoints, out):
    """lowres f32 [n*r_cap, res, res, 4*k] (phase layout), boxes f32 [n, r_cap, 4] -> out f32 [n*r_cap, k, 4]."""
    assert lowres.dtype == torch.float32 and lowres.is_contiguous() and lowres.numel() == n * r_cap * res * res * 4 * num_keypoints
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and out.dtype == torch.float32 and out.is_contiguous()
    check(load().cm2_keypoints_decode(ptr(lowres), ptr(boxes), ptr(det_count), n, r_cap, res, num_keypoints, ptr(out),
                                      stream()), "cm2_keypoints_decode")
    _count()


def scale_clip_boxes(boxes_in, boxes_out, valid, r, sx, sy, out_w, out_h):
    check(load().cm2_scale_clip_boxes(ptr(boxes_in), ptr(boxes_out), ptr(valid), r, sx, sy, out_w, out_h, stream()),
          "cm2_scale_clip_boxes")
    _count()


def scale_clip_boxes_batch(boxes_in, boxes_out, valid, n, r_cap, params, det_count=None):
    check(load().cm2_scale_clip_boxes_batch(ptr(boxes_in), ptr(boxes_out), ptr(valid), n, r_cap, ptr(params), ptr(det_count), stream()),
          "cm2_scale_clip_boxes_batch")
    _count()


def paste_masks(probs, boxes, valid, out, r, m, out_h, out_w, threshold):
    check(load().cm2_paste_masks(ptr(probs), ptr(boxes), ptr(valid), ptr(out), r, m, out_h, out_w, threshold, stream()),
          "cm2_paste_masks")
    _count()


def pack_records(boxes, scores, classes, mask_scores, locations, valid, det_count, records):
    n, r_cap = boxes.shape[0], boxes.shape[1]
    assert records.dtype == torch.float32 and records.is_contiguous() and tuple(records.shape) == (n, r_cap, 11)
    assert classes.dtype == torch.int64 and det_count.dtype == torch.int32
    check(load().cm2_pack_records(ptr(boxes), ptr(scores), ptr(classes), ptr(mask_scores), ptr(locations), ptr(valid), ptr(det_count),
                                  n, r_cap, ptr(records), stream()), "cm2_pack_records")
    _count()
