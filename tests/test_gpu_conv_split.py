"""GPU: the split-precision tensor-core convolution (fp32 activations, f16 hi/lo operands, three tcgen05 MMAs per
product term; include/cm2.h "Split precision") through ``Engine("fp32")`` against fp64 references.

The north star asks the fp32 variant for boxes <= 1e-2 px on coordinates of ~1e3 px after ~60 layers, i.e. fp32-grade
arithmetic in every layer: the gate here is a maximum error of SPLIT_TOL (relative to the largest output magnitude)
against the fp64 result -- the CUDA-core fp32 engine is measured on the same inputs and printed beside it.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from centermask2_b200 import lib, packing                       # noqa: E402
from centermask2_b200.engine import Engine, FMap, PhaseMap, SegMap   # noqa: E402

DEV = "cuda"
F32 = torch.float32
SPLIT_TOL = 1.5e-6          # max |err| / max |ref|; measured 4e-7 .. 1.0e-6 (the CUDA-core fp32 engine: 0.8 .. 1.9e-6)


@pytest.fixture(scope="module")
def eng():
    e = Engine(None, "fp32", DEV)
    e.use_graphs = False
    yield e
    e.release()


@pytest.fixture(scope="module")
def simt():
    e = Engine(None, "fp32_simt", DEV)
    e.use_graphs = False
    yield e
    e.release()


def halo(t_nchw):
    n, c, h, w = t_nchw.shape
    buf = torch.zeros((n, h + 2, w + 2, c), dtype=F32, device=DEV)
    buf[:, 1:-1, 1:-1] = t_nchw.permute(0, 2, 3, 1).to(DEV, F32)
    return FMap(buf, 1)


def nchw(view):
    return view.permute(0, 3, 1, 2).double().cpu()


def rel_err(got, ref):
    return ((got - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()


def check(got, ref, what, tol=SPLIT_TOL):
    e = rel_err(got, ref)
    print("{}: max err / max |ref| = {:.2e}".format(what, e))
    assert e <= tol, (what, e)
    return e


def conv_ref(xs, wt, scale, shift, relu, stride=1, pad=None):
    k = wt.shape[-1]
    y = F.conv2d(torch.cat(xs, 1).double(), wt.double(), None, stride, k // 2 if pad is None else pad)
    if scale is not None:
        y = y * scale.double().view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.double().view(1, -1, 1, 1)
    return F.relu(y) if relu else y


@pytest.mark.parametrize("srcs,cout,k,h,w,n", [
    ([64], 64, 3, 20, 24, 1),            # stem_2 shape class
    ([128], 128, 3, 17, 23, 2),          # OSA2 3x3, two images (tiles straddle image boundary)
    ([160], 160, 3, 13, 21, 1),          # K tail: 160 = 2.5 k-blocks (the hi box of the last block reads into the lo half)
    ([224], 224, 3, 9, 11, 1),
    ([256], 256, 3, 25, 42, 1),          # FCOS tower / FPN output
    ([256], 80, 3, 13, 21, 2),           # cls_logits (N = 80)
    ([256], 16, 3, 7, 11, 1),            # bbox_pred + ctrness merged (N padded to 16)
    ([128, 128, 128, 128, 128, 128], 256, 1, 12, 16, 1),   # OSA2 aggregation: virtual concat of 6
    ([256, 160, 160, 160, 160, 160], 512, 1, 10, 12, 2),   # OSA3 aggregation: two N tiles, K tails
    ([1024], 256, 1, 25, 42, 1),         # FPN lateral 5
    ([256, 16], 256, 3, 14, 14, 5),      # MaskIoU fcn1
    ([128], 128, 3, 100, 168, 2),        # large: CTA pairs / 256-row tiles
    ([224], 224, 3, 25, 42, 16),         # trimmed tile range + memset
])
def test_split_conv_matches_fp64(eng, simt, srcs, cout, k, h, w, n):
    g = torch.Generator().manual_seed(sum(srcs) + cout + k)
    cin = sum(srcs)
    xs = [torch.randn(n, c, h, w, generator=g) * 3.0 for c in srcs]
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    wt[0] *= 1e-3                                                # a channel of tiny weights (per-channel pre-scale)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    ref = conv_ref(xs, wt, scale, shift, True)
    cw = packing.ConvW(wt, srcs, 1, k // 2, scale, shift, True, F32, DEV, True)
    assert cw.split and cw.w_tc.dtype == torch.float16 and cw.w_tc.shape[0] == 2 * ((cout + 15) // 16 * 16)
    out = halo(torch.full((n, cout, h, w), 7.0))
    out.buf.fill_(7.0)                                           # the engine must rewrite the halo with zeros
    c0 = lib.launch_count
    eng.begin_pass()
    got = eng.conv("t", [halo(x) for x in xs], cw, out=out)
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == len(srcs) + 1, "expected one split per source + one TC launch (no SIMT fallback)"
    check(nchw(got.view), ref, "split {} -> {} k{} {}x{}x{}".format(srcs, cout, k, n, h, w))
    b = out.buf
    assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0
    cs = packing.ConvW(wt, srcs, 1, k // 2, scale, shift, True, F32, DEV, False)
    o2 = simt.conv("t", [halo(x) for x in xs], cs)
    torch.cuda.synchronize()
    print("   (CUDA-core fp32 engine on the same inputs: {:.2e})".format(rel_err(nchw(o2.view), ref)))


def test_split_conv_accuracy_on_long_k_and_wide_dynamic_range(eng, simt):
    """K = 9 * 256 with activations spanning 1e-4 .. 1e3 and a ReLU-like (non-negative, biased) distribution: the sum
    does not cancel, so a biased accumulator (round-toward-zero) would show as a systematic relative error."""
    g = torch.Generator().manual_seed(7)
    n, c, h, w, cout = 2, 256, 40, 56, 256
    x = torch.relu(torch.randn(n, c, h, w, generator=g)) * torch.exp(torch.randn(n, c, 1, 1, generator=g) * 2.0)
    wt = torch.rand(cout, c, 3, 3, generator=g) / (9 * c)        # positive weights: no cancellation
    ref = conv_ref([x], wt, None, None, False)
    cw = packing.ConvW(wt, [c], 1, 1, None, None, False, F32, DEV, True)
    eng.begin_pass()
    got = eng.conv("acc", [halo(x)], cw)
    torch.cuda.synchronize()
    d = (nchw(got.view) - ref) / ref.abs().clamp(min=1e-30)
    inner = d[:, :, 2:-2, 2:-2]
    print("split conv, positive sum of 2304 terms: mean signed rel err {:.2e}, max |rel err| {:.2e}".format(
        inner.mean().item(), inner.abs().max().item()))
    cs = packing.ConvW(wt, [c], 1, 1, None, None, False, F32, DEV, False)
    o2 = simt.conv("acc", [halo(x)], cs)
    torch.cuda.synchronize()
    d2 = ((nchw(o2.view) - ref) / ref.abs().clamp(min=1e-30))[:, :, 2:-2, 2:-2]
    print("CUDA-core fp32 engine:                   mean signed rel err {:.2e}, max |rel err| {:.2e}".format(
        d2.mean().item(), d2.abs().max().item()))
    assert inner.abs().max().item() <= 4e-6 and abs(inner.mean().item()) <= 2e-6


def test_split_conv_channel_sums_and_gate(eng):
    """stats_mode 1 on an fp32 output (eSE global pool, vovnet.py:254)."""
    g = torch.Generator().manual_seed(3)
    srcs, cout, h, w, n = [128, 128, 128], 256, 12, 16, 3
    xs = [torch.randn(n, c, h, w, generator=g) for c in srcs]
    wt = torch.randn(cout, sum(srcs), 1, 1, generator=g) / math.sqrt(sum(srcs))
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    cw = packing.ConvW(wt, srcs, 1, 0, scale, shift, True, F32, DEV, True)
    sums = torch.full((n, cout), 123.0, dtype=torch.float64, device=DEV)
    eng.begin_pass()
    out = eng.conv("s", [halo(x) for x in xs], cw, stats=sums, stats_mode=1)
    torch.cuda.synchronize()
    check(nchw(out.view), conv_ref(xs, wt, scale, shift, True), "split conv + channel sums")
    ref = out.view.double().sum(dim=(1, 2))
    assert torch.allclose(sums, ref, rtol=1e-6, atol=1e-4), (sums - ref).abs().max()


def test_split_segmented_conv_groupnorm_fused(eng):
    """All FPN levels in one launch with GroupNorm statistics from the epilogue (stats_mode 2), fp32 output."""
    g = torch.Generator().manual_seed(5)
    shapes = [(2, 13, 21), (2, 7, 11), (2, 4, 6), (2, 2, 3)]
    c = 256
    seg = SegMap(shapes, c, F32, DEV)
    xs = []
    for i, (n, h, w) in enumerate(shapes):
        x = torch.randn(n, c, h, w, generator=g)
        xs.append(x)
        seg.level(i).view.copy_(x.permute(0, 2, 3, 1).to(DEV))
    wt = torch.randn(c, c, 3, 3, generator=g) / 48
    bias = torch.randn(c, generator=g) * 0.1
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    cw = packing.ConvW(wt, [c], 1, 1, None, bias, False, F32, DEV, True)
    n_img = sum(s[0] for s in shapes)
    st = torch.zeros((n_img, c // 8, 2), dtype=torch.float64, device=DEV)
    eng.begin_pass()
    out = eng.conv_seg("segc", seg, cw, stats=st, stats_mode=2)
    torch.cuda.synchronize()
    img = 0
    for i, (n, h, w) in enumerate(shapes):
        ref = conv_ref([xs[i]], wt, None, bias, False)
        check(nchw(out.level(i).view), ref, "split segmented conv level {}".format(i))
        y = out.level(i).view.double()
        # fp32 warp partial sums in front of the fp64 atomics
        assert torch.allclose(st[img:img + n, :, 0], y.reshape(n, h * w, c // 8, 8).sum(dim=(1, 3)), rtol=1e-6, atol=1e-4)
        assert torch.allclose(st[img:img + n, :, 1], (y * y).reshape(n, h * w, c // 8, 8).sum(dim=(1, 3)), rtol=1e-6, atol=1e-4)
        img += n
        b = out.level(i).buf
        assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0
    lib.groupnorm_apply_seg(out.flat, out.segs, 32, gamma.to(DEV), beta.to(DEV), 1e-5, True, st)
    torch.cuda.synchronize()
    for i in range(len(shapes)):
        ref = F.relu(F.group_norm(conv_ref([xs[i]], wt, None, bias, False), 32, gamma.double(), beta.double(), 1e-5))
        check(nchw(out.level(i).view), ref, "split conv + GroupNorm level {}".format(i), tol=3e-6)


def test_split_fpn_lateral_with_fp32_upsample_add(eng):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 512, 12, 20, generator=g)
    low = torch.randn(2, 256, 6, 10, generator=g)
    wt = torch.randn(256, 512, 1, 1, generator=g) / 22
    bias = torch.randn(256, generator=g) * 0.1
    ref = conv_ref([x], wt, None, bias, False) + F.interpolate(low.double(), scale_factor=2.0, mode="nearest")
    cw = packing.ConvW(wt, [512], 1, 0, None, bias, False, F32, DEV, True)
    c0 = lib.launch_count
    eng.begin_pass()
    out = eng.conv("lat", [halo(x)], cw, residual=halo(low), res_mode=2)
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == 2
    check(nchw(out.view), ref, "split lateral + upsample add")


def test_split_deconv_scatter(eng):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 256, 14, 14, generator=g)
    wd = torch.randn(256, 256, 2, 2, generator=g) / 16
    bd = torch.randn(256, generator=g) * 0.1
    ref = F.relu(F.conv_transpose2d(x.double(), wd.double(), bd.double(), stride=2))
    cw = packing.deconv2x2({"d.weight": wd, "d.bias": bd}, "d", F32, DEV, True)
    c0 = lib.launch_count
    eng.begin_pass()
    out = eng.conv("dc", [halo(x)], cw, out_mode=1, out_halo=0)
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == 2
    check(nchw(out.view), ref, "split deconv scatter")


@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 128, 24, 32, 2), (256, 256, 25, 42, 1), (256, 256, 14, 14, 7)])
def test_split_stride2_chain_on_phase_planes(eng, cin, cout, h, w, n):
    """out_mode 2 (phase-split store, fp32) feeding a 3x3 / stride 2 convolution (stem_2 -> stem_3, maskiou_fcn3 -> 4),
    and cm2_phase_split (fp32, with ReLU) feeding one (P6 -> P7)."""
    g = torch.Generator().manual_seed(cin + h)
    x = torch.randn(n, cin, h, w, generator=g)
    w2 = torch.randn(cin, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    w3 = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    b3 = torch.randn(cout, generator=g) * 0.1
    c2 = packing.ConvW(w2, [cin], 1, 1, None, None, True, F32, DEV, True)
    c3 = packing.ConvW(w3, [cin], 2, 1, None, b3, False, F32, DEV, True)
    mid = conv_ref([x], w2, None, None, True)
    ref = conv_ref([mid.float()], w3, None, b3, False, stride=2, pad=1)
    eng.begin_pass()
    pm = eng.conv("p2", [halo(x)], c2, out_mode=2)
    assert isinstance(pm, PhaseMap)
    c0 = lib.launch_count
    out = eng.conv("p3", [pm], c3, out_halo=0)
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == 2, "stride-2 conv fell back to the CUDA-core engine"
    check(nchw(out.view), ref, "split stride-2 chain", tol=3e-6)       # two layers; mid went through a float() round trip
    # phase_split path with ReLU on load
    ref2 = conv_ref([F.relu(x)], w3, None, b3, False, stride=2, pad=1)
    eng.begin_pass()
    src = eng.phase_split("ps", halo(x), relu=True)
    out2 = eng.conv("p3b", [src], c3, out_halo=0)
    torch.cuda.synchronize()
    check(nchw(out2.view), ref2, "split stride-2 on cm2_phase_split planes")


def test_split_linear_dense_rows(eng):
    g = torch.Generator().manual_seed(4)
    r, kdim, cout = 150, 12544, 1024
    x = torch.randn(r, kdim, generator=g)
    wt = torch.randn(cout, kdim, generator=g) / math.sqrt(kdim)
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.linear(x.double(), wt.double(), bias.double()))
    cw = packing.linear({"l.weight": wt, "l.bias": bias}, "l", True, F32, DEV, True)
    xin = FMap(x.to(DEV).reshape(r, 1, 1, kdim).contiguous(), 0)
    c0 = lib.launch_count
    eng.begin_pass()
    out = eng.conv("fc", [xin], cw, out_halo=0)
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == 2
    check(out.view.reshape(r, cout).double().cpu(), ref, "split linear K=12544")


def test_split_is_made_once_per_pass_and_again_in_the_next(eng):
    g = torch.Generator().manual_seed(9)
    x = halo(torch.randn(1, 64, 8, 8, generator=g))
    wt = torch.randn(64, 64, 3, 3, generator=g) / 24
    cw = packing.ConvW(wt, [64], 1, 1, None, None, True, F32, DEV, True)
    eng.begin_pass()
    c0 = lib.launch_count
    a = eng.conv("o1", [x], cw)
    b = eng.conv("o2", [x], cw)
    assert lib.launch_count - c0 == 3                            # one split, two convolutions
    x.view.mul_(2.0)                                             # the source changes between passes
    eng.begin_pass()
    c = eng.conv("o3", [x], cw)
    torch.cuda.synchronize()
    assert torch.equal(a.view, b.view) and torch.allclose(c.view, 2.0 * a.view, rtol=1e-5, atol=1e-6)


def test_split_conv_kx_merged_slabs_give_the_same_result(eng):
    """CM2_TC3_MERGE=1: one 136-row {x_hi, x_lo} slab per filter row serves the three kx taps (MMA descriptors start 0 / 1 / 2
    rows into it); same products, another accumulation order."""
    import os
    g = torch.Generator().manual_seed(12)
    n, c, h, w, cout = 2, 160, 21, 30, 224
    x = torch.randn(n, c, h, w, generator=g)
    wt = torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c)
    ref = conv_ref([x], wt, None, None, True)
    cw = packing.ConvW(wt, [c], 1, 1, None, None, True, F32, DEV, True)
    outs = []
    for merge in ("0", "1"):
        os.environ["CM2_TC3_MERGE"] = merge
        try:
            eng.begin_pass()
            o = eng.conv("km" + merge, [halo(x)], cw)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CM2_TC3_MERGE", None)
        check(nchw(o.view), ref, "split 3x3 conv, CM2_TC3_MERGE=" + merge)
        outs.append(o.view.clone())
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-6 * ref.abs().max().item()


def test_split_output_feeds_the_next_convolution_without_an_fp32_copy(eng):
    """``to_conv=True``: the epilogue stores the [hi | lo] f16 operand pair of the next convolution directly (epilogue kind
    11) -- 3x3 chain on halo maps, the stride-2 chain on phase planes, and dense linear layers."""
    from centermask2_b200.engine import SplitFMap, SplitPhaseMap
    eng.split_out_all = True                                 # the engine itself uses the store for cout <= 128 only (measured)
    g = torch.Generator().manual_seed(21)
    n, c, h, w = 2, 160, 19, 26
    x = torch.randn(n, c, h, w, generator=g)
    w1 = torch.randn(c, c, 3, 3, generator=g) / math.sqrt(9 * c)
    w2 = torch.randn(224, 2 * c, 1, 1, generator=g) / math.sqrt(2 * c)
    sc, sh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    c1 = packing.ConvW(w1, [c], 1, 1, sc, sh, True, F32, DEV, True)
    c2 = packing.ConvW(w2, [c, c], 1, 0, None, None, False, F32, DEV, True)
    mid = conv_ref([x], w1, sc, sh, True)
    ref = conv_ref([x, mid.float()], w2, None, None, False)
    eng.begin_pass()
    xin = halo(x)
    c0 = lib.launch_count
    y = eng.conv("so1", [xin], c1, to_conv=True)
    assert isinstance(y, SplitFMap) and y.buf.dtype == torch.float16 and y.c == c and y.buf.shape[3] == 2 * c
    out = eng.conv("so2", [xin, y], c2)                     # virtual concat of an fp32 map and a split map
    torch.cuda.synchronize()
    assert lib.launch_count - c0 == 3                        # one split (x), two convolutions: none for y
    hi, lo = y.view[..., :c].float(), y.view[..., c:].float()
    check((hi + lo).permute(0, 3, 1, 2).double().cpu(), mid, "split output hi + lo", tol=1e-6)
    b = y.buf.float()
    assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0
    check(nchw(out.view), ref, "conv over [fp32 map, split map]", tol=3e-6)
    # stride-2 chain: split phase planes
    w3 = torch.randn(128, c, 3, 3, generator=g) / math.sqrt(9 * c)
    c3 = packing.ConvW(w3, [c], 2, 1, None, None, False, F32, DEV, True)
    ref3 = conv_ref([mid.float()], w3, None, None, False, stride=2, pad=1)
    eng.begin_pass()
    pm = eng.conv("so3", [xin], c1, out_mode=2, to_conv=True)
    assert isinstance(pm, SplitPhaseMap)
    o3 = eng.conv("so4", [pm], c3, out_halo=0)
    torch.cuda.synchronize()
    check(nchw(o3.view), ref3, "stride-2 conv on split phase planes", tol=3e-6)
    # dense rows (linear layers)
    r, k1, k2 = 70, 512, 256
    xl = torch.randn(r, k1, generator=g)
    wl1, wl2 = torch.randn(k2, k1, generator=g) / math.sqrt(k1), torch.randn(80, k2, generator=g) / math.sqrt(k2)
    bl1 = torch.randn(k2, generator=g) * 0.1
    l1 = packing.linear({"l.weight": wl1, "l.bias": bl1}, "l", True, F32, DEV, True)
    l2 = packing.linear({"l.weight": wl2, "l.bias": torch.zeros(80)}, "l", False, F32, DEV, True)
    refl = F.linear(F.relu(F.linear(xl.double(), wl1.double(), bl1.double())), wl2.double())
    eng.begin_pass()
    a = eng.conv("sl1", [FMap(xl.to(DEV).reshape(r, 1, 1, k1).contiguous(), 0)], l1, out_halo=0, to_conv=True)
    assert isinstance(a, SplitFMap) and a.halo == 0
    o = eng.conv("sl2", [a], l2, out_halo=0)
    torch.cuda.synchronize()
    check(o.view.reshape(r, 80).double().cpu(), refl, "linear chain through a split map", tol=3e-6)


def test_groupnorm_apply_writes_split_operands(eng):
    """cm2_groupnorm_apply_seg_split: normalise + ReLU of the fp32 tower tensor straight into [hi | lo] f16 (== the in-place
    fp32 apply followed by cm2_split_f16x2, bit for bit)."""
    g = torch.Generator().manual_seed(6)
    shapes = [(2, 13, 21), (2, 7, 11), (2, 4, 6)]
    c = 256
    seg = SegMap(shapes, c, F32, DEV)
    for i, (n, h, w) in enumerate(shapes):
        seg.level(i).view.copy_(torch.randn(n, h, w, c, generator=g).to(DEV))
    wt = torch.randn(c, c, 3, 3, generator=g) / 48
    bias = torch.randn(c, generator=g) * 0.1
    gamma, beta = (torch.rand(c, generator=g) + 0.5).to(DEV), (torch.randn(c, generator=g) * 0.1).to(DEV)
    cw = packing.ConvW(wt, [c], 1, 1, None, bias, False, F32, DEV, True)
    st = torch.zeros((6, c // 8, 2), dtype=torch.float64, device=DEV)
    eng.begin_pass()
    out = eng.conv_seg("gs", seg, cw, stats=st, stats_mode=2)
    sp = torch.zeros((out.rows, 2 * c), dtype=torch.float16, device=DEV)
    lib.groupnorm_apply_seg_split(out.flat, sp, out.segs, 32, gamma, beta, 1e-5, True, st)
    lib.groupnorm_apply_seg(out.flat, out.segs, 32, gamma, beta, 1e-5, True, st)          # in place, fp32
    want = torch.empty_like(sp)
    lib.split_f16x2(out.flat, want)
    torch.cuda.synchronize()
    # the two apply kernels evaluate v * a + b with the same coefficients; the fused one rounds once more nowhere
    diff = (sp.float() - want.float()).abs().max().item()
    assert diff <= 1e-3 * out.flat.abs().max().item(), diff
    tot = (sp[:, :c].float() + sp[:, c:].float() - out.flat).abs().max().item()
    assert tot <= 2e-6 * out.flat.abs().max().item(), tot
