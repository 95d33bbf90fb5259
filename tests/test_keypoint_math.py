"""CPU: the arithmetic of the keypoint decode kernel (csrc/kp_math.cuh, shared by the kernel and this host harness)
against the oracle (torch CPU interpolate + the heatmaps_to_keypoints restatement).  The harness is test
infrastructure; the GPU tests (test_gpu_kernels.py) check the kernel itself."""
import ctypes
import os
import subprocess

import pytest
import torch
import torch.nn.functional as F

from centermask2_b200 import packing
from oracle import restate
from tests.helpers import pack_lowres, keypoint_boxes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("kp_host") / "kp_host.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                    os.path.join(HERE, "host", "kp_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.kp_host_decode.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_int] * 2
    return lib


@pytest.mark.parametrize("res,k,r", [(14, 17, 12), (7, 3, 9)])
def test_host_harness_matches_oracle(host_lib, res, k, r):
    g = torch.Generator().manual_seed(100 + res)
    low = torch.randn(r, k, 2 * res, 2 * res, generator=g) * 2.5
    boxes = keypoint_boxes(g, r)
    packed = pack_lowres(low)
    out = torch.zeros(r, k, 4)
    hi = torch.zeros(r, k, 4 * res, 4 * res)
    host_lib.kp_host_decode(packed.data_ptr(), boxes.data_ptr(), r, res, k, out.data_ptr(), hi.data_ptr(), 0, 0)
    ref_hi = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=False)
    assert torch.allclose(hi, ref_hi, rtol=0, atol=1e-6)
    ref = restate.heatmaps_to_keypoints(ref_hi, boxes)
    moved = ((out[..., :2] - ref[..., :2]).abs() > 1e-3).any(-1)
    assert moved.float().mean().item() <= 0.01, moved.nonzero()
    assert torch.allclose(out[..., 2], ref[..., 2], rtol=1e-5, atol=1e-5)
    assert torch.allclose(out[..., 3], ref[..., 3], rtol=1e-4)


def test_deconv4x4s2_as_phase_conv_equals_conv_transpose2d():
    """packing.deconv4x4s2: ConvTranspose2d(k 4, s 2, p 1) (keypoint_head.py:205-208) rewritten as a 3x3 convolution to
    4 * K phase columns; the phase layout is the one kp_lowres_offset reads."""
    g = torch.Generator().manual_seed(3)
    cin, k, r, res = 8, 3, 2, 5
    sd = {"d.weight": torch.randn(cin, k, 4, 4, generator=g), "d.bias": torch.randn(k, generator=g)}
    w = packing.deconv4x4s2(sd, "d", torch.float32, "cpu", False)
    x = torch.randn(r, cin, res, res, generator=g)
    ref = F.conv_transpose2d(x, sd["d.weight"], sd["d.bias"], stride=2, padding=1)
    w3 = w.w_simt.reshape(3, 3, cin, 4 * k).permute(3, 2, 0, 1)
    y = F.conv2d(x, w3, w.shift, 1, 1).permute(0, 2, 3, 1).reshape(r, res, res, 4, k)       # the engine's NHWC output
    assert torch.allclose(y, pack_lowres(ref), atol=1e-5)
    assert not w.relu and w.k == 3 and w.pad == 1 and w.cout == 4 * k


@pytest.mark.parametrize("threads,tab_rows", [(256, 1024), (256, 40), (32, 1024)])
def test_column_walk_decomposition_is_bit_identical_to_the_flat_loop(host_lib, threads, tab_rows):
    """kp_column_walk (the kernel's default work split: (column, row segment) items, x pass carried in registers down the
    column, y taps from a table; ROIs taller than the table fall back to the flat loop) against the flat per-pixel loop:
    same expression tree per pixel, first-index tie-break, so every output must be bit-identical."""
    g = torch.Generator().manual_seed(11)
    res, k = 14, 2
    boxes = torch.cat([keypoint_boxes(g, 10), torch.tensor([
        [0.0, 0.0, 640.0, 3.0],          # wide strip: wc >= threads, one short segment
        [5.0, 5.0, 6.0, 505.0],          # tall strip: one column, many segments
        [-3.5, 2.25, 500.5, 402.0],      # large: several items per thread
        [1.0, 1.0, 30.0, 1300.0],        # taller than any table: flat-loop fallback
        [10.2, 11.7, 10.5, 11.9],        # sub-pixel: one resized pixel
        [0.0, 0.0, 20.0, 20.0],          # downscale: the source row advances by more than one per resized row
        [0.0, 0.0, 9.0, 300.0]])])
    r = boxes.shape[0]
    low = torch.randn(r, k, 2 * res, 2 * res, generator=g) * 2.5
    low[1] = low[1].round()                 # plateaus / exact ties for the first-index rule
    packed = pack_lowres(low)
    flat, walk = torch.zeros(r, k, 4), torch.zeros(r, k, 4)
    host_lib.kp_host_decode(packed.data_ptr(), boxes.data_ptr(), r, res, k, flat.data_ptr(), None, 0, 0)
    host_lib.kp_host_decode(packed.data_ptr(), boxes.data_ptr(), r, res, k, walk.data_ptr(), None, threads, tab_rows)
    assert torch.equal(flat, walk), (flat - walk).abs().amax(dim=(1, 2))


@pytest.mark.parametrize("size", [2, 3, 14, 28, 48])
def test_bilinear_x2_constant_weights_equal_float_index_formulation(host_lib, size):
    """kp_bilinear2_at (weights 1/4, 3/4, 0 from the parity of the output index: what the kernel runs) against
    kp_bilinear_at (ATen's float source-index arithmetic): bit-identical on every pixel."""
    g = torch.Generator().manual_seed(size)
    low = (torch.randn(size, size, generator=g) * 3.0).contiguous()
    host_lib.kp_host_bilinear_mismatches.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert host_lib.kp_host_bilinear_mismatches(low.data_ptr(), size) == 0


def test_keypoint_decode_properties(host_lib):
    """Size-independent properties of heatmaps_to_keypoints that hold for any map: (1) translating the box translates
    the keypoints and leaves logit / score untouched (the resized map depends on the box size only); (2) every keypoint
    lies inside its (>= 1 px) box; (3) adding a constant to a map adds it to the logit and leaves location and score
    untouched up to rounding (softmax shift invariance)."""
    g = torch.Generator().manual_seed(21)
    res, k, r = 14, 5, 8
    low = torch.randn(r, k, 2 * res, 2 * res, generator=g) * 2.0
    boxes = keypoint_boxes(g, r)

    def run(lo, bx):
        out = torch.zeros(r, k, 4)
        packed, bx = pack_lowres(lo), bx.contiguous()                      # keep the buffers alive across the C call
        host_lib.kp_host_decode(packed.data_ptr(), bx.data_ptr(), r, res, k, out.data_ptr(), None, 256, 1024)
        return out

    base = run(low, boxes)
    shift = torch.tensor([64.0, -32.0, 64.0, -32.0])                       # exactly representable offsets
    moved = run(low, boxes + shift)
    assert torch.allclose(moved[..., 0], base[..., 0] + 64.0, atol=1e-4) and torch.allclose(moved[..., 1], base[..., 1] - 32.0, atol=1e-4)
    assert torch.equal(moved[..., 2:], base[..., 2:])
    w = (boxes[:, 2] - boxes[:, 0]).clamp(min=1).view(r, 1)
    h = (boxes[:, 3] - boxes[:, 1]).clamp(min=1).view(r, 1)
    assert bool(((base[..., 0] >= boxes[:, 0:1]) & (base[..., 0] <= boxes[:, 0:1] + w)).all())
    assert bool(((base[..., 1] >= boxes[:, 1:2]) & (base[..., 1] <= boxes[:, 1:2] + h)).all())
    lifted = run(low + 1.5, boxes)
    same = ((lifted[..., :2] - base[..., :2]).abs() < 1e-3).all(-1)
    assert same.float().mean().item() >= 0.95                              # rounding may move a near-tie
    assert torch.allclose(lifted[..., 2][same], base[..., 2][same] + 1.5, atol=1e-4)
    assert torch.allclose(lifted[..., 3][same], base[..., 3][same], rtol=1e-4)
