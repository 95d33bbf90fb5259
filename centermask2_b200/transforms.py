"""Input-side transform on the GPU: ``ResizeShortestEdge`` + bilinear resize with Pillow's arithmetic.

The reference resizes on the CPU before the model (``/root/reference/deploy_utils.py:60-73``:
``T.ResizeShortestEdge([800, 800], 1333).get_transform(img).apply_image(img)``; detectron2's ``ResizeTransform``
calls ``PIL.Image.resize(.., BILINEAR)`` for uint8 HWC images).  Pillow's resampler (``src/libImaging/Resample.c``,
version pinned in this image: Pillow 12.2.0) is a separable two-pass filter on uint8 with 22-bit fixed-point
coefficients; ``pil_bilinear_coeffs`` restates its coefficient table exactly, and ``cm2_resize_pil_u8`` runs the two
passes on the device, so the result is bit-identical to the reference's resized image.
"""
import math

import numpy as np
import torch

from . import lib

PRECISION_BITS = 32 - 8 - 2          # Resample.c


def pil_bilinear_coeffs(in_size, out_size):
    """``precompute_coeffs`` + ``normalize_coeffs_8bpc`` of Resample.c for the bilinear (triangle) filter, full-image box.
    Returns (bounds int32 [out, 2] = (first input index, tap count), coeffs int32 [out, ksize])."""
    scale = filterscale = float(in_size) / float(out_size)
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)          # (int) cast truncates toward zero
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(arg < 1.0, 1.0 - arg, 0.0)
    w = np.where(x < xmax[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]            # Resample.c accumulates the taps left to right in double
    k = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kk = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS))).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kk


def shortest_edge_output_shape(h, w, short_edge_length, max_size):
    """``ResizeShortestEdge.get_output_shape`` [d2] (deploy_utils.py:69)."""
    size = short_edge_length * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * scale, neww * scale
    return int(newh + 0.5), int(neww + 0.5)


class ResizeTransform(object):
    """detectron2's ``ResizeTransform`` (interp = bilinear) for uint8 images, executed on the device."""

    _cache = {}

    def __init__(self, h, w, new_h, new_w):
        self.h, self.w, self.new_h, self.new_w = h, w, new_h, new_w

    def _tables(self, device):
        key = (self.h, self.w, self.new_h, self.new_w, str(device))
        t = self._cache.get(key)
        if t is None:
            bx, kx = pil_bilinear_coeffs(self.w, self.new_w)
            by, ky = pil_bilinear_coeffs(self.h, self.new_h)
            t = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in (bx, kx, by, ky))
            self._cache[key] = t
        return t

    def apply_image(self, img, chw=False):
        """img: uint8 HWC (numpy array or torch tensor, any device) -> uint8 CUDA tensor, HWC (``chw=False``, what
        ``apply_image`` returns in detectron2) or CHW (``chw=True``: the layout ``batched_inputs[i]["image"]`` takes)."""
        t = torch.as_tensor(img) if not isinstance(img, torch.Tensor) else img
        assert t.dtype == torch.uint8 and t.dim() == 3 and t.shape[0] == self.h and t.shape[1] == self.w, (t.shape, t.dtype)
        if not t.is_cuda:
            t = t.pin_memory().cuda(non_blocking=True) if torch.cuda.is_available() else t
        if not t.is_cuda:
            raise RuntimeError("centermask2_b200.transforms needs a CUDA device (there is no CPU fallback)")
        t = t.contiguous()
        c = t.shape[2]
        if (self.h, self.w) == (self.new_h, self.new_w):
            return t.permute(2, 0, 1).contiguous() if chw else t
        bx, kx, by, ky = self._tables(t.device)
        tmp = torch.empty((self.h, self.new_w, c), dtype=torch.uint8, device=t.device)
        out = torch.empty((c, self.new_h, self.new_w) if chw else (self.new_h, self.new_w, c), dtype=torch.uint8, device=t.device)
        lib.resize_pil_u8(t, tmp, out, self.h, self.w, c, self.new_h, self.new_w, bx, kx, by, ky, chw)
        return out


class ResizeShortestEdge(object):
    """detectron2's ``T.ResizeShortestEdge(short_edge_length, max_size)`` with ``sample_style="choice"`` and a single
    length (the inference configuration: ``deploy_utils.py:69``, ``INPUT.MIN_SIZE_TEST`` / ``MAX_SIZE_TEST``)."""

    def __init__(self, short_edge_length, max_size=1333, sample_style="choice"):
        if isinstance(short_edge_length, int):
            short_edge_length = (short_edge_length, short_edge_length)
        assert sample_style == "choice" and len(set(short_edge_length)) == 1, "inference-time resize takes one length"
        self.short_edge_length, self.max_size = short_edge_length[0], max_size

    def get_transform(self, image):
        h, w = image.shape[:2]
        new_h, new_w = shortest_edge_output_shape(h, w, self.short_edge_length, self.max_size)
        return ResizeTransform(h, w, new_h, new_w)


def get_sample_inputs(image, short_edge_length=800, max_size=1333):
    """``deploy_utils.get_sample_inputs`` (:60-73) from an already decoded BGR uint8 HWC image:
    -> ``[{"image": uint8 CHW CUDA tensor, "height": original h, "width": original w}]``."""
    h, w = image.shape[:2]
    tf = ResizeShortestEdge([short_edge_length, short_edge_length], max_size).get_transform(image)
    return [{"image": tf.apply_image(image, chw=True), "height": h, "width": w}]
