"""Oracle for the COCO RLE encoding.  TEST INFRASTRUCTURE ONLY: imported by ``tests/`` alone.

The reference calls ``pycocotools.mask.encode`` (``/root/reference/centermask2/centermask/evaluation/coco_evaluation.py:388-391``);
pycocotools (pinned by detectron2 v0.5: ``pycocotools>=2.0.2``) is NOT installed in this image and not vendored in
/root/reference, so **parity is unpinned against the binary**.  This file restates the published algorithm of its
``common/maskApi.c`` in scalar Python -- ``rleEncode`` (column-major run lengths, first run counts zeros),
``rleToString`` / ``rleFrString`` (difference against the run two places back for i > 2, 5 data bits + continuation bit per
character, offset 48) and ``rleDecode`` -- and is anchored by hand-derived known answers in ``tests/test_rle.py`` plus the
encode -> string -> parse -> decode round trip.
"""
import numpy as np


def rle_encode(mask):
    """mask: [h, w] array of 0/1 -> list of run lengths (maskApi.c rleEncode on the Fortran-order pixels)."""
    flat = np.asarray(mask, dtype=np.uint8).flatten(order="F")
    cnts = []
    p, c = 0, 0
    for v in flat:
        v = int(v != 0)
        if v != p:
            cnts.append(c)
            c = 0
            p = v
        c += 1
    cnts.append(c)
    return cnts


def rle_to_string(cnts):
    s = []
    for i, x in enumerate(cnts):
        x = int(x)
        if i > 2:
            x -= int(cnts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            s.append(chr(c + 48))
    return "".join(s).encode("ascii")


def rle_from_string(s):
    if isinstance(s, bytes):
        s = s.decode("ascii")
    cnts = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(cnts) > 2:
            x += cnts[-2]
        cnts.append(x)
    return cnts


def rle_decode(cnts, h, w):
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in cnts:
        flat[pos:pos + c] = v
        pos += c
        v = 1 - v
    return flat.reshape((h, w), order="F")
