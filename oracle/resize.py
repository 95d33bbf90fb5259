"""Oracle for the input-side resize.  TEST INFRASTRUCTURE ONLY: imported by ``tests/`` alone.

The reference's resize is detectron2's ``ResizeShortestEdge`` + ``ResizeTransform.apply_image``
(``/root/reference/deploy_utils.py:60-73``), whose arithmetic lives in a third-party dependency that is present in this
image: **Pillow 12.2.0** (``PIL.Image.resize(size, BILINEAR)``; ``src/libImaging/Resample.c``).  Two checkers:

* ``pil_resize`` -- Pillow itself (the pin);
* ``restated_resize`` -- a plain-Python/numpy restatement of Resample.c's ``precompute_coeffs`` /
  ``normalize_coeffs_8bpc`` / ``ImagingResampleHorizontal_8bpc`` / ``...Vertical_8bpc`` written independently of
  ``centermask2_b200/transforms.py`` (scalar loops, one output column / row at a time).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def d2_output_shape(h, w, short_edge_length=800, max_size=1333):
    """``ResizeShortestEdge.get_output_shape`` [d2-memory]."""
    size = short_edge_length * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def pil_resize(img, oh, ow):
    from PIL import Image
    return np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))


def _coeffs(in_size, out_size):
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = filterscale                     # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds, kk = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        if ww != 0.0:
            k = [v / ww for v in k]
        k += [0.0] * (ksize - len(k))
        kk.append([int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS)) for v in k])
        bounds.append((xmin, xmax))
    return bounds, kk


def restated_resize(img, oh, ow):
    h, w, c = img.shape
    src = img.astype(np.int64)
    if ow != w:
        bounds, kk = _coeffs(w, ow)
        tmp = np.zeros((h, ow, c), dtype=np.int64)
        for xx in range(ow):
            x0, n = bounds[xx]
            ss = np.full((h, c), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for t in range(n):
                ss += src[:, x0 + t, :] * kk[xx][t]
            tmp[:, xx, :] = np.clip(ss >> PRECISION_BITS, 0, 255)
        src = tmp
    if oh != h:
        bounds, kk = _coeffs(h, oh)
        out = np.zeros((oh, src.shape[1], c), dtype=np.int64)
        for yy in range(oh):
            y0, n = bounds[yy]
            ss = np.full((src.shape[1], c), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for t in range(n):
                ss += src[y0 + t] * kk[yy][t]
            out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
        src = out
    return src.astype(np.uint8)
