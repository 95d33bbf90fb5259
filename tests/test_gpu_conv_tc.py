"""GPU: the tcgen05 implicit-GEMM convolution (CM2_ENGINE_TC) against torch.conv2d on the same
bf16-rounded operands (fp32 accumulation on both sides).  Tolerance: one bf16 rounding of the output
(rtol 2^-8) plus fp32 accumulation-order noise; f32 outputs are compared at 1e-3."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from centermask2_b200 import lib, packing              # noqa: E402
from centermask2_b200.engine import FMap               # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def halo(t_nchw, dtype=BF):
    n, c, h, w = t_nchw.shape
    buf = torch.zeros((n, h + 2, w + 2, c), dtype=dtype, device=DEV)
    buf[:, 1:-1, 1:-1] = t_nchw.permute(0, 2, 3, 1).to(DEV, dtype)
    return FMap(buf, 1)


def nchw(view):
    return view.permute(0, 3, 1, 2).float().cpu()


def rb(t):
    return t.to(BF).float()


def close(got, ref, out_bf16=True):
    tol = 1.0 / 128 if out_bf16 else 1e-3
    err = (got - ref).abs()
    bound = tol * ref.abs() + tol * ref.abs().mean() + 1e-3
    assert (err <= bound).all(), "max err {} at ref {}".format(err.max().item(), ref.flatten()[err.argmax()].item())


@pytest.mark.parametrize("srcs,cout,k,h,w,n", [
    ([64], 64, 3, 20, 24, 1),            # stem_2 shape class
    ([128], 128, 3, 17, 23, 2),          # OSA2 3x3, two images (tiles straddle image boundary)
    ([160], 160, 3, 13, 21, 1),          # K tail: 160 = 2.5 k-blocks, N = 160
    ([224], 224, 3, 9, 11, 1),           # K tail 224, N = 224
    ([256], 256, 3, 25, 42, 1),          # FCOS tower / FPN output
    ([256], 80, 3, 13, 21, 2),           # cls_logits (N = 80)
    ([256], 5, 3, 7, 11, 1),             # bbox_pred + ctrness merged (N padded to 16, scalar stores)
    ([128, 128, 128, 128, 128, 128], 256, 1, 12, 16, 1),   # OSA2 aggregation: virtual concat of 6
    ([256, 160, 160, 160, 160, 160], 512, 1, 10, 12, 2),   # OSA3 aggregation: two N tiles, K tails
    ([1024], 256, 1, 25, 42, 1),         # FPN lateral 5
    ([256, 16], 256, 3, 14, 14, 5),      # MaskIoU fcn1: roi feature + 16-channel padded mask plane
    ([224], 224, 3, 25, 42, 16),         # 149 row tiles on 148 SMs: leading / trailing halo rows trimmed (148 tiles) + memset
    ([1024], 256, 1, 25, 42, 16),        # same geometry, 1x1
])
def test_conv_tc_halo(srcs, cout, k, h, w, n):
    g = torch.Generator().manual_seed(sum(srcs) + cout + k)
    cin = sum(srcs)
    xs = [rb(torch.randn(n, c, h, w, generator=g)) for c in srcs]
    wt = rb(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(torch.cat(xs, 1), wt, None, 1, k // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    cw = packing.ConvW(wt, srcs, 1, k // 2, scale, shift, True, BF, DEV, True)
    out = halo(torch.full((n, cout, h, w), 7.0))
    out.buf.fill_(7.0)                                          # the engine must rewrite the halo with zeros
    ok = lib.conv2d([halo(x).view for x in xs], cw.w_tc, out.view, cout, k, 1, k // 2, scale=cw.scale, shift=cw.shift,
                    relu=True, engine=lib.ENGINE_TC, probe=True)
    assert ok, lib.last_error()
    torch.cuda.synchronize()
    close(nchw(out.view), ref)
    b = out.buf.float()
    assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0


def test_conv_tc_f32_dense_output_and_no_relu():
    g = torch.Generator().manual_seed(1)
    x = rb(torch.randn(2, 256, 13, 21, generator=g))
    wt = rb(torch.randn(80, 256, 3, 3, generator=g) / 48)
    bias = torch.randn(80, generator=g)
    ref = F.conv2d(x, wt, bias, 1, 1)
    cw = packing.ConvW(wt, [256], 1, 1, None, bias, False, BF, DEV, True)
    out = torch.full((2, 13, 21, 80), 3.0, device=DEV)
    assert lib.conv2d([halo(x).view], cw.w_tc, out, 80, 3, 1, 1, shift=cw.shift, engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(nchw(out), ref, out_bf16=False)
    # 5-channel f32 output (scalar store path)
    wt5 = rb(torch.randn(5, 256, 3, 3, generator=g) / 48)
    sc5, sh5 = torch.rand(5, generator=g) + 0.5, torch.randn(5, generator=g)
    ref5 = F.conv2d(x, wt5, None, 1, 1) * sc5.view(1, -1, 1, 1) + sh5.view(1, -1, 1, 1)
    cw5 = packing.ConvW(wt5, [256], 1, 1, sc5, sh5, False, BF, DEV, True)
    out5 = torch.zeros((2, 13, 21, 5), device=DEV)
    assert lib.conv2d([halo(x).view], cw5.w_tc, out5, 5, 3, 1, 1, scale=cw5.scale, shift=cw5.shift, engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(nchw(out5), ref5, out_bf16=False)


def test_conv_tc_fpn_lateral_with_upsample_add():
    g = torch.Generator().manual_seed(2)
    x = rb(torch.randn(2, 512, 12, 20, generator=g))
    low = rb(torch.randn(2, 256, 6, 10, generator=g))
    wt = rb(torch.randn(256, 512, 1, 1, generator=g) / 22)
    bias = torch.randn(256, generator=g) * 0.1
    ref = F.conv2d(x, wt, bias) + F.interpolate(low, scale_factor=2.0, mode="nearest")
    cw = packing.ConvW(wt, [512], 1, 0, None, bias, False, BF, DEV, True)
    out = halo(torch.zeros(2, 256, 12, 20))
    assert lib.conv2d([halo(x).view], cw.w_tc, out.view, 256, 1, 1, 0, shift=cw.shift, residual=halo(low).view,
                      res_mode=2, engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(nchw(out.view), ref)


def test_conv_tc_deconv_scatter():
    g = torch.Generator().manual_seed(3)
    x = rb(torch.randn(6, 256, 14, 14, generator=g))
    wd = rb(torch.randn(256, 256, 2, 2, generator=g) / 16)
    bd = torch.randn(256, generator=g) * 0.1
    ref = F.relu(F.conv_transpose2d(x, wd, bd, stride=2))
    cw = packing.deconv2x2({"d.weight": wd, "d.bias": bd}, "d", BF, DEV, True)
    out = torch.zeros((6, 28, 28, 256), dtype=BF, device=DEV)
    assert lib.conv2d([halo(x).view], cw.w_tc, out, 1024, 1, 1, 0, shift=cw.shift, relu=True, out_mode=1,
                      engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(nchw(out), ref)


def test_conv_tc_linear_dense_rows():
    g = torch.Generator().manual_seed(4)
    r, kdim, cout = 150, 12544, 1024
    x = rb(torch.randn(r, kdim, generator=g))
    wt = rb(torch.randn(cout, kdim, generator=g) / math.sqrt(kdim))
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.linear(x, wt, bias))
    cw = packing.linear({"l.weight": wt, "l.bias": bias}, "l", True, BF, DEV, True)
    xin = x.to(DEV, BF).reshape(r, 1, 1, kdim)
    out = torch.zeros((r, 1, 1, cout), dtype=BF, device=DEV)
    assert lib.conv2d([xin], cw.w_tc, out, cout, 1, 1, 0, shift=cw.shift, relu=True, engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(out.reshape(r, cout).float().cpu(), ref)


def test_conv_tc_large_multi_wave():
    """More tiles than SMs (persistent loop, both accumulator stages, ring wrap-around)."""
    g = torch.Generator().manual_seed(5)
    x = rb(torch.randn(2, 128, 100, 168, generator=g))
    wt = rb(torch.randn(128, 128, 3, 3, generator=g) / 34)
    ref = F.relu(F.conv2d(x, wt, None, 1, 1))
    cw = packing.ConvW(wt, [128], 1, 1, None, None, True, BF, DEV, True)
    out = halo(torch.zeros(2, 128, 100, 168))
    assert lib.conv2d([halo(x).view], cw.w_tc, out.view, 128, 3, 1, 1, relu=True, engine=lib.ENGINE_TC, probe=True)
    torch.cuda.synchronize()
    close(nchw(out.view), ref)


@pytest.mark.parametrize("c", [128, 160])
def test_conv_tc_cta_pair_layers(c):
    """Stride-1 3x3 layers with N <= 224 and enough tiles run as CTA pairs (cta_group::2 MMAs of M = 256, half a weight
    tile per CTA).  271 tiles of 256 rows: the last pair has only its leader half inside the matrix."""
    g = torch.Generator().manual_seed(50 + c)
    x = rb(torch.randn(4, c, 100, 168, generator=g))
    wt = rb(torch.randn(c, c, 3, 3, generator=g) / math.sqrt(9 * c))
    scale = torch.rand(c, generator=g) + 0.5
    shift = torch.randn(c, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, wt, None, 1, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    cw = packing.ConvW(wt, [c], 1, 1, scale, shift, True, BF, DEV, True)
    out = halo(torch.zeros(4, c, 100, 168))
    out.buf.fill_(3.0)
    assert lib.conv2d([halo(x).view], cw.w_tc, out.view, c, 3, 1, 1, scale=cw.scale, shift=cw.shift, relu=True,
                      engine=lib.ENGINE_TC, probe=True), lib.last_error()
    torch.cuda.synchronize()
    close(nchw(out.view), ref)
    b = out.buf.float()
    assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0


def test_conv_tc_every_shape_through_cta_pairs():
    """The whole halo-convolution shape table forced through the pair kernel (the variant switches are read once per
    process, hence the child process): CM2_TC_VARIANT=3 puts every layer on the 256-row kernel, CM2_TC_PAIR=2 pairs it."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, CM2_TC_PAIR="2", CM2_TC_VARIANT="3", CM2_TC_B_RESIDENT="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_conv_tc.py", "-m", "gpu", "-q", "-x", "-k",
                        "conv_tc_halo or multi_wave or cta_pair_layers"], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_unsupported_shapes_are_refused_not_miscomputed():
    x = halo(torch.zeros(1, 64, 8, 8))
    w = torch.zeros((64, 64 * 9), dtype=BF, device=DEV)
    out = halo(torch.zeros(1, 64, 4, 4))
    assert lib.conv2d([x.view], w, out.view, 64, 3, 2, 1, engine=lib.ENGINE_TC, probe=True) is False
    with pytest.raises(RuntimeError):
        lib.conv2d([x.view], w, out.view, 64, 3, 2, 1, engine=lib.ENGINE_TC)


@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 128, 24, 32, 2), (256, 256, 25, 42, 1), (256, 256, 13, 21, 2),
                                            (256, 256, 14, 14, 7)])
def test_conv_tc_stride2_on_phase_planes(cin, cout, h, w, n):
    """3x3 / stride 2 / pad 1 (stem_3, P6, P7, maskiou_fcn4) through cm2_phase_split + src_phase."""
    g = torch.Generator().manual_seed(cin + h)
    x = rb(torch.randn(n, cin, h, w, generator=g))
    wt = rb(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9))
    bias = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(F.relu(x), wt, bias, 2, 1)
    ho, wo = ref.shape[2], ref.shape[3]
    planes = torch.zeros((4, n, ho + 2, wo + 2, cin), dtype=BF, device=DEV)
    lib.phase_split(halo(x).view, planes[0, :, 1:-1, 1:-1, :], relu=True)
    torch.cuda.synchronize()
    # the four planes hold the (relu'd) input pixels by parity
    xr = F.relu(x)
    for q in range(4):
        py, px = q >> 1, q & 1
        sub = xr[:, :, py::2, px::2]
        got = planes[q, :, 1:1 + sub.shape[2], 1:1 + sub.shape[3], :].permute(0, 3, 1, 2).float().cpu()
        assert torch.equal(got, sub)
    cw = packing.ConvW(wt, [cin], 2, 1, None, bias, False, BF, DEV, True)
    out = torch.zeros((n, ho, wo, cout), dtype=BF, device=DEV)
    assert lib.conv2d([planes[0, :, 1:-1, 1:-1, :]], cw.w_tc, out, cout, 3, 2, 1, shift=cw.shift, engine=lib.ENGINE_TC,
                      probe=True, src_phase=True), lib.last_error()
    torch.cuda.synchronize()
    close(nchw(out), ref)


@pytest.mark.parametrize("kind", ["linear", "linear_f32", "conv3x3", "conv1x1_multi", "phase_s2"])
def test_conv_tc_split_k_matches_the_unsplit_convolution(kind):
    """cm2_conv_desc.splitk: K slices as separate tiles + fixed-order fp32 reduction (MaskIoU linear layers, P6 / P7, late
    stages at small batch) == F.conv2d, == the unsplit launch up to fp32 summation order, halo untouched, repeatable bit for bit."""
    g = torch.Generator().manual_seed(31)
    relu, out_f32, src_phase, stride = True, False, False, 1
    if kind.startswith("linear"):
        r, k, cout = 70, 2112, 144 if kind == "linear_f32" else 528            # K not a multiple of the slice, three N tiles of 176
        out_f32 = kind == "linear_f32"
        relu = not out_f32
        x = rb(torch.randn(r, k, 1, 1, generator=g))
        wt = rb(torch.randn(cout, k, 1, 1, generator=g) / math.sqrt(k))
        srcs, views, ksz = [k], [x.reshape(r, k).to(DEV, BF).reshape(r, 1, 1, k)], 1
        out = torch.zeros((r, 1, 1, cout), dtype=torch.float32 if out_f32 else BF, device=DEV)
        outs = [out, torch.zeros_like(out)]
        oview = lambda o: o
    elif kind == "conv3x3":
        n, c, h, w, cout = 2, 224, 13, 21, 224
        x = rb(torch.randn(n, c, h, w, generator=g))
        wt = rb(torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c))
        srcs, views, ksz = [c], [halo(x).view], 3
        outs = [halo(torch.zeros(n, cout, h, w)), halo(torch.zeros(n, cout, h, w))]
        oview = lambda o: o.view
    elif kind == "conv1x1_multi":
        n, h, w, cout = 1, 9, 14, 256
        srcs = [256, 160, 160]
        xs = [rb(torch.randn(n, c, h, w, generator=g)) for c in srcs]
        x = torch.cat(xs, dim=1)
        wt = rb(torch.randn(cout, sum(srcs), 1, 1, generator=g) / math.sqrt(sum(srcs)))
        views, ksz = [halo(t).view for t in xs], 1
        outs = [halo(torch.zeros(n, cout, h, w)), halo(torch.zeros(n, cout, h, w))]
        oview = lambda o: o.view
    else:
        n, c, h, w, cout = 3, 256, 13, 21, 256
        src_phase, stride = True, 2
        x = rb(torch.randn(n, c, h, w, generator=g))
        wt = rb(torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c))
        ho, wo = (h + 1) // 2, (w + 1) // 2
        planes = torch.zeros((4, n, ho + 2, wo + 2, c), dtype=BF, device=DEV)
        lib.phase_split(halo(x).view, planes[0, :, 1:-1, 1:-1, :], relu=False)
        srcs, views, ksz = [c], [planes[0, :, 1:-1, 1:-1, :]], 3
        outs = [halo(torch.zeros(n, cout, ho, wo)), halo(torch.zeros(n, cout, ho, wo))]
        oview = lambda o: o.view
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, wt, None, stride, ksz // 2) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    if relu:
        ref = torch.relu(ref)
    cw = packing.ConvW(wt, srcs, stride, ksz // 2, sc, sh, relu, BF, DEV, True)
    ob = outs[0] if isinstance(outs[0], torch.Tensor) else outs[0].buf
    ws = torch.full((6, ob.shape[0] * ob.stride(0)), float("nan"), dtype=torch.float32, device=DEV)
    kw = dict(scale=cw.scale_tc, shift=cw.shift, relu=relu, engine=lib.ENGINE_TC, probe=True, src_phase=src_phase)
    assert lib.conv2d(views, cw.w_tc, oview(outs[0]), cout, ksz, stride, ksz // 2, splitk=5, splitk_ws=ws, **kw), lib.last_error()
    assert lib.conv2d(views, cw.w_tc, oview(outs[1]), cout, ksz, stride, ksz // 2, **kw), lib.last_error()
    torch.cuda.synchronize()
    got, plain = nchw(oview(outs[0])), nchw(oview(outs[1]))
    close(got, ref, out_bf16=not out_f32)
    d = (got - plain).abs()
    assert (d <= plain.abs() / 64 + 1e-3).all() if not out_f32 else (d <= 1e-5 * plain.abs() + 1e-5).all(), d.max().item()
    if not isinstance(outs[0], torch.Tensor):
        b = outs[0].buf.float()
        assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0
    first = oview(outs[0]).clone()
    assert lib.conv2d(views, cw.w_tc, oview(outs[0]), cout, ksz, stride, ksz // 2, splitk=5, splitk_ws=ws, **kw)
    torch.cuda.synchronize()
    assert torch.equal(first, oview(outs[0]))
    # a workspace that is too small is refused, nothing is launched
    assert not lib.conv2d(views, cw.w_tc, oview(outs[0]), cout, ksz, stride, ksz // 2, splitk=5, splitk_ws=ws[:1, :64], **kw)


def test_conv_tc_on_shared_halo_maps():
    """ROI maps whose images share their zero frame (engine.SharedHaloFMap, csrc/conv_tc.cu halo_kind 2: 225 GEMM rows per
    14x14 ROI instead of 256): a chain of 3x3 convolutions, a two-source convolution (the MaskIoU head's [roi, mask] concat),
    the phase-plane store for a following stride-2 convolution -- all against F.conv2d; the shared frame stays zero."""
    from centermask2_b200.engine import Engine, SharedHaloFMap
    eng = Engine(None, "bf16", DEV)
    eng.use_graphs = False
    try:
        g = torch.Generator().manual_seed(41)
        r, c, res = 37, 256, 14                                       # 37 ROIs: 8325 rows = 65.04 tiles (a ragged last tile)
        x = rb(torch.randn(r, c, res, res, generator=g))
        m = rb(torch.randn(r, 16, res, res, generator=g))
        w1 = rb(torch.randn(c, c, 3, 3, generator=g) / 48)
        w2 = rb(torch.randn(c, c + 16, 3, 3, generator=g) / 49)
        w3 = rb(torch.randn(c, c, 3, 3, generator=g) / 48)
        b1 = torch.randn(c, generator=g) * 0.1
        y1 = rb(torch.relu(F.conv2d(x, w1, b1, 1, 1)))
        y2 = rb(torch.relu(F.conv2d(torch.cat([y1, m], 1), w2, None, 1, 1)))
        ref3 = torch.relu(F.conv2d(y2, w3, None, 2, 1))
        c1 = packing.ConvW(w1, [c], 1, 1, None, b1, True, BF, DEV, True)
        c2 = packing.ConvW(w2, [c, 16], 1, 1, None, None, True, BF, DEV, True)
        c3 = packing.ConvW(w3, [c], 2, 1, None, None, True, BF, DEV, True)
        xin = eng.shared_halo_fmap("t_x", r, res, res, c)
        min_ = eng.shared_halo_fmap("t_m", r, res, res, 16)
        xin.view.copy_(x.permute(0, 2, 3, 1).to(DEV, BF))
        min_.view.copy_(m.permute(0, 2, 3, 1).to(DEV, BF))
        eng.begin_pass()
        o1 = eng.conv("t_1", [xin], c1)
        assert isinstance(o1, SharedHaloFMap) and o1.buf.shape[0] == r * 225 + 16
        o2 = eng.conv("t_2", [o1, min_], c2)
        assert isinstance(o2, SharedHaloFMap)
        pm = eng.conv("t_2p", [o1, min_], c2, out_mode=2)             # the same layer, stored as phase planes
        o3 = eng.conv("t_3", [pm], c3, out_halo=0)
        torch.cuda.synchronize()
        close(nchw(o1.view), y1)
        close(nchw(o2.view), y2)
        close(nchw(o3.view), ref3)
        for o in (o1, o2):                                            # everything outside the interior view is still zero
            total = o.buf.float().abs().sum().item()
            inner = o.view.float().abs().sum().item()
            assert abs(total - inner) <= 1e-6 * max(1.0, inner), (total, inner)
        # and the same convolution on an ordinary halo map gives the same numbers (same tiles of K, other row order)
        std = halo(x)
        o1s = eng.conv("t_1s", [std], c1)
        torch.cuda.synchronize()
        assert not isinstance(o1s, SharedHaloFMap)
        assert torch.equal(o1s.view, o1.view)
    finally:
        eng.release()


def test_conv_tc_phase_split_store_feeds_stride2_conv():
    """stem_2 -> stem_3 chain: out_mode 2 (phase-split store) then a stride-2 conv on the planes."""
    g = torch.Generator().manual_seed(11)
    n, h, w = 2, 20, 28
    x = rb(torch.randn(n, 64, h, w, generator=g))
    w2 = rb(torch.randn(64, 64, 3, 3, generator=g) / 24)
    w3 = rb(torch.randn(128, 64, 3, 3, generator=g) / 24)
    mid = rb(F.relu(F.conv2d(x, w2, None, 1, 1)))
    ref = F.relu(F.conv2d(mid, w3, None, 2, 1))
    c2 = packing.ConvW(w2, [64], 1, 1, None, None, True, BF, DEV, True)
    c3 = packing.ConvW(w3, [64], 2, 1, None, None, True, BF, DEV, True)
    planes = torch.zeros((4, n, h // 2 + 2, w // 2 + 2, 64), dtype=BF, device=DEV)
    p0 = planes[0, :, 1:-1, 1:-1, :]
    assert lib.conv2d([halo(x).view], c2.w_tc, p0, 64, 3, 1, 1, relu=True, out_mode=2, engine=lib.ENGINE_TC, probe=True)
    out = halo(torch.zeros(n, 128, h // 2, w // 2))
    assert lib.conv2d([p0], c3.w_tc, out.view, 128, 3, 2, 1, relu=True, engine=lib.ENGINE_TC, probe=True, src_phase=True)
    torch.cuda.synchronize()
    for q in range(4):
        py, px = q >> 1, q & 1
        got = planes[q, :, 1:-1, 1:-1, :].permute(0, 3, 1, 2).float().cpu()
        close(got, mid[:, :, py::2, px::2])
        assert planes[q, :, 0].abs().max() == 0 and planes[q, :, :, 0].abs().max() == 0      # plane halos stay zero
    close(nchw(out.view), ref)


def test_fused_preprocess_im2col_stem1():
    """cm2_preprocess_im2col + 1x1 TC conv == normalise/pad + 3x3 s2 conv (vovnet.py:409)."""
    g = torch.Generator().manual_seed(12)
    h, w, hp, wp = 45, 61, 64, 64
    img = (torch.rand(3, h, w, generator=g) * 255).floor()
    mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
    xn = torch.zeros(1, 3, hp, wp)
    xn[0, :, :h, :w] = rb(img - torch.tensor(mean).view(3, 1, 1))
    w1 = rb(torch.randn(64, 3, 3, 3, generator=g) / 5)
    ref = F.conv2d(xn, w1, None, 2, 1)
    w32 = torch.zeros(64, 32, 1, 1)
    w32[:, :27, 0, 0] = w1.permute(0, 2, 3, 1).reshape(64, 27)
    cw = packing.ConvW(w32, [32], 1, 0, None, None, False, BF, DEV, True)
    for src in (img, img.to(torch.uint8)):
        cols = halo(torch.zeros(2, 32, hp // 2, wp // 2))
        lib.preprocess_im2col(src.to(DEV), mean, std, hp, wp, cols.view, 1)
        out = halo(torch.zeros(2, 64, hp // 2, wp // 2))
        assert lib.conv2d([cols.view], cw.w_tc, out.view, 64, 1, 1, 0, engine=lib.ENGINE_TC, probe=True)
        torch.cuda.synchronize()
        close(nchw(out.view)[1:], ref)
        assert nchw(out.view)[0].abs().max() == 0


@pytest.mark.parametrize("std", [[1.0, 1.0, 1.0], [57.375, 57.12, 58.395]])
def test_stem1_fused_matches_conv_and_the_im2col_path(std):
    """cm2_stem1_fused_batch (normalise + pad + 3x3 / s2 conv + FrozenBN + ReLU in one pass, csrc/stem.cu) against
    F.conv2d on the bf16-rounded normalised input (vovnet.py:409) and against the two-pass path it replaces (im2col +
    K = 32 GEMM on the tcgen05 engine): images of different extents (tile edges: 300 columns = 2 full tiles + 44, odd
    widths, an image smaller than the padded extent), uint8 and float inputs, halo left untouched."""
    g = torch.Generator().manual_seed(14)
    hp, wp = 96, 608
    mean = [103.53, 116.28, 123.675]
    extents = ((45, 61), (96, 608), (33, 599), (90, 257))
    imgs = [(torch.rand(3, h, w, generator=g) * 255).floor().to(torch.uint8) for h, w in extents]
    w1 = rb(torch.randn(64, 3, 3, 3, generator=g) / 5)
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    xn = torch.zeros(len(imgs), 3, hp, wp)
    for i, (im, (h, w)) in enumerate(zip(imgs, extents)):
        xn[i, :, :h, :w] = rb((im.float() - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1))
    ref = torch.relu(F.conv2d(xn, w1, None, 2, 1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    w30 = torch.zeros(64, 32)
    w30[:, :30] = F.pad(w1.permute(0, 2, 3, 1).reshape(64, 3, 9), (0, 1)).reshape(64, 30)
    w30 = w30.to(DEV, torch.bfloat16).contiguous()
    w32 = torch.zeros(64, 32, 1, 1)
    w32[:, :27, 0, 0] = w1.permute(0, 2, 3, 1).reshape(64, 27)
    cw = packing.ConvW(w32, [32], 1, 0, sc, sh, True, BF, DEV, True)
    cols = halo(torch.zeros(len(imgs) + 1, 32, hp // 2, wp // 2))
    lib.preprocess_im2col_batch([im.to(DEV) for im in imgs], mean, std, hp, wp, cols.view, 1)
    two = halo(torch.zeros(len(imgs) + 1, 64, hp // 2, wp // 2))
    assert lib.conv2d([cols.view], cw.w_tc, two.view, 64, 1, 1, 0, scale=cw.scale_tc, shift=cw.shift, relu=True, engine=lib.ENGINE_TC, probe=True)
    for conv_in in ([im.to(DEV) for im in imgs], [im.float().to(DEV) for im in imgs]):
        out = halo(torch.zeros(len(imgs) + 1, 64, hp // 2, wp // 2))
        lib.stem1_fused_batch(conv_in, mean, std, hp, wp, w30, cw.scale, cw.shift, True, out.view, 1)
        torch.cuda.synchronize()
        close(nchw(out.view)[1:], ref)
        assert nchw(out.view)[0].abs().max() == 0
        b = out.buf.float()
        assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0
        # same products, same fp32 accumulation up to the order of the sum: at most one bf16 ulp apart
        d = (out.view[1:].float() - two.view[1:].float()).abs()          # (slot 0 of the two-pass output is relu(shift): zero columns)
        assert (d <= two.view[1:].float().abs() / 64 + 1e-3).all(), d.max().item()
        assert (d > 0).float().mean().item() < 0.05


@pytest.mark.parametrize("std", [[1.0, 1.0, 1.0], [57.375, 57.12, 58.395]])
def test_stem1_fused_split_precision_matches_fp32_conv(std):
    """cm2_stem1_fused_split_batch (fp32 engine): f16 hi / lo fragments of the normalised input and of the weights, result
    stored as the [hi | lo] operand pair of stem_2 -- hi + lo against F.conv2d in float64 on the exact fp32 inputs."""
    g = torch.Generator().manual_seed(15)
    hp, wp = 96, 608
    mean = [103.53, 116.28, 123.675]
    extents = ((45, 61), (96, 608), (33, 599))
    imgs = [(torch.rand(3, h, w, generator=g) * 255).floor().to(torch.uint8) for h, w in extents]
    w1 = torch.randn(64, 3, 3, 3, generator=g) / 5
    w1[5] *= 1e-3                                             # a channel with tiny weights: the per-channel pre-scale matters
    sc, sh = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    xn = torch.zeros(len(imgs), 3, hp, wp, dtype=torch.float64)
    for i, (im, (h, w)) in enumerate(zip(imgs, extents)):
        v = im.float() - torch.tensor(mean).view(3, 1, 1)
        if std[0] != 1.0:
            v = v / torch.tensor(std).view(3, 1, 1)
        xn[i, :, :h, :w] = v.double()
    ref = torch.relu(F.conv2d(xn, w1.double(), None, 2, 1) * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    w30 = torch.zeros(64, 32)
    w30[:, :30] = F.pad(w1.permute(0, 2, 3, 1).reshape(64, 3, 9), (0, 1)).reshape(64, 30)
    pre = torch.exp2(torch.floor(torch.log2(256.0 / w30.abs().amax(dim=1))))
    ws = w30 * pre.view(-1, 1)
    hi = ws.to(torch.float16)
    lo = (ws - hi.float()).to(torch.float16)
    whl = torch.stack([hi, lo]).to(DEV).contiguous()
    scale_tc = (sc / pre).to(DEV).contiguous()
    for conv_in in ([im.to(DEV) for im in imgs], [im.float().to(DEV) for im in imgs]):
        buf = torch.zeros((len(imgs) + 1, hp // 2 + 2, wp // 2 + 2, 128), dtype=torch.float16, device=DEV)
        lib.stem1_fused_batch(conv_in, mean, std, hp, wp, whl, scale_tc, sh.to(DEV), True, buf[:, 1:-1, 1:-1, :], 1)
        torch.cuda.synchronize()
        got = (buf[:, 1:-1, 1:-1, :64].double() + buf[:, 1:-1, 1:-1, 64:].double()).permute(0, 3, 1, 2).cpu()
        err = (got[1:] - ref).abs().max().item()
        assert err <= 3e-6 * ref.abs().max().item(), (err, ref.abs().max().item())
        assert got[0].abs().max() == 0
        b = buf.float()
        assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0


@pytest.mark.parametrize("std", [[1.0, 1.0, 1.0], [57.375, 57.12, 58.395]])
def test_preprocess_im2col_batch_equals_per_image(std):
    """One launch for the whole batch (images of different extents) == the per-image entry point, bit for bit."""
    g = torch.Generator().manual_seed(13)
    hp, wp = 64, 96
    mean = [103.53, 116.28, 123.675]
    imgs = [(torch.rand(3, h, w, generator=g) * 255).floor().to(torch.uint8).to(DEV) for h, w in ((45, 61), (64, 96), (33, 90))]
    a, b = halo(torch.zeros(4, 32, hp // 2, wp // 2)), halo(torch.zeros(4, 32, hp // 2, wp // 2))
    a.view.fill_(9.0)
    b.view.fill_(9.0)
    for i, im in enumerate(imgs):
        lib.preprocess_im2col(im, mean, std, hp, wp, a.view, 1 + i)
    lib.preprocess_im2col_batch(imgs, mean, std, hp, wp, b.view, 1)
    torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)
    lib.preprocess_im2col_batch([im.float() for im in imgs], mean, std, hp, wp, b.view, 1)
    torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)


def test_segmented_conv_and_groupnorm_match_per_level_results():
    """All FPN levels of an FCOS tower in one launch (cm2_seg): conv3x3 + bias, then GroupNorm(32)+ReLU in place."""
    from centermask2_b200.engine import SegMap
    g = torch.Generator().manual_seed(21)
    n, c = 2, 256
    shapes = [(n, 25, 42), (n, 13, 21), (n, 7, 11), (n, 4, 6), (n, 2, 3)]
    xs = [rb(torch.randn(n, c, h, w, generator=g)) for _, h, w in shapes]
    wt = rb(torch.randn(c, c, 3, 3, generator=g) / 48)
    bias = torch.randn(c, generator=g) * 0.1
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.2
    seg = SegMap(shapes, c, BF, DEV)
    for i, x in enumerate(xs):
        seg.level(i).view.copy_(x.permute(0, 2, 3, 1).to(DEV, BF))
    cw = packing.ConvW(wt, [c], 1, 1, None, bias, False, BF, DEV, True)
    out = seg.like(c, BF, lambda shape: torch.full(shape, 5.0, dtype=BF, device=DEV))
    lib.conv2d([seg.flat], cw.w_tc, out.flat, c, 3, 1, 1, shift=cw.shift, engine=lib.ENGINE_TC, segs=seg.segs)
    torch.cuda.synchronize()
    conv_ref = [rb(F.conv2d(x, wt, bias, 1, 1)) for x in xs]
    for i, ref in enumerate(conv_ref):
        lv = out.level(i)
        close(nchw(lv.view), ref)
        b = lv.buf.float()
        assert b[:, 0].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, -1].abs().max() == 0       # halo zeroed
    ws = torch.empty(lib.gn_seg_workspace_floats(out.segs, c, 32), device=DEV)
    lib.groupnorm_relu_seg(out.flat, out.segs, 32, gamma.to(DEV), beta.to(DEV), 1e-5, True, ws)
    torch.cuda.synchronize()
    for i in range(len(shapes)):
        lv = out.level(i)
        got_conv = conv_ref[i]
        # reference GN on what the engine actually stored (bf16-rounded conv output)
        ref = F.relu(F.group_norm(got_conv, 32, gamma, beta, 1e-5))
        close(nchw(lv.view), ref)
        assert lv.buf.float()[:, 0].abs().max() == 0
    # f32 head output with 16 columns (bbox_pred + ctrness padded)
    w16 = torch.zeros(16, c, 3, 3)
    w16[:5] = rb(torch.randn(5, c, 3, 3, generator=g) / 48)
    b16 = torch.zeros(16)
    b16[:5] = torch.randn(5, generator=g)
    cw16 = packing.ConvW(w16, [c], 1, 1, None, b16, False, BF, DEV, True)
    o16 = seg.like(16, torch.float32, lambda shape: torch.zeros(shape, dtype=torch.float32, device=DEV))
    lib.conv2d([seg.flat], cw16.w_tc, o16.flat, 16, 3, 1, 1, shift=cw16.shift, engine=lib.ENGINE_TC, segs=seg.segs)
    torch.cuda.synchronize()
    for i, x in enumerate(xs):
        close(nchw(o16.level(i).view), F.conv2d(x, w16, b16, 1, 1), out_bf16=False)


# ---------------------------------------------------------------------------------------------------
# fused output statistics (cm2_conv_desc.stats) and their consumers
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("srcs,cout,k,h,w,n,variant", [
    ([128, 128, 128], 256, 1, 12, 16, 3, 0),      # OSA aggregation, tiles straddle images
    ([256, 160], 512, 1, 9, 11, 2, 0),            # two N tiles
    ([64, 64], 768, 1, 7, 9, 5, 0),               # three N tiles, many images per tile
    ([128], 128, 3, 17, 23, 2, 0),
])
def test_conv_tc_channel_sums(srcs, cout, k, h, w, n, variant):
    """stats_mode 1: per (image, channel) sums of the stored outputs (eSE global pool, vovnet.py:254)."""
    g = torch.Generator().manual_seed(cout + h)
    cin = sum(srcs)
    xs = [rb(torch.randn(n, c, h, w, generator=g)) for c in srcs]
    wt = rb(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    cw = packing.ConvW(wt, srcs, 1, k // 2, scale, shift, True, BF, DEV, True)
    out = halo(torch.zeros(n, cout, h, w))
    sums = torch.full((n, cout), 123.0, dtype=torch.float64, device=DEV)        # the call zeroes it
    assert lib.conv2d([halo(x).view for x in xs], cw.w_tc, out.view, cout, k, 1, k // 2, scale=cw.scale, shift=cw.shift,
                      relu=True, engine=lib.ENGINE_TC, probe=True, stats=sums, stats_mode=1), lib.last_error()
    torch.cuda.synchronize()
    ref = out.view.double().sum(dim=(1, 2))                                     # of the values as stored
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-3), (sums - ref).abs().max()


@pytest.mark.parametrize("shared_halo", [False, True])
def test_conv_tc_seg_groupnorm_fused(shared_halo):
    """stats_mode 2 + cm2_groupnorm_apply_seg == conv -> GroupNorm(32) -> ReLU per level (fcos.py:176-186); also on segments
    whose images share their zero frame (cm2_seg.halo 1, the layout the bf16 engine uses for the FCOS levels)."""
    from centermask2_b200.engine import SegMap
    g = torch.Generator().manual_seed(5)
    shapes = [(2, 13, 21), (2, 7, 11), (2, 4, 6), (2, 2, 3)]
    c = 256
    seg = SegMap(shapes, c, BF, DEV, shared_halo=shared_halo)
    assert seg.segs.halo == int(shared_halo)
    xs = []
    for i, (n, h, w) in enumerate(shapes):
        x = rb(torch.randn(n, c, h, w, generator=g))
        xs.append(x)
        seg.level(i).view.copy_(x.permute(0, 2, 3, 1).to(DEV, BF))
    wt = rb(torch.randn(c, c, 3, 3, generator=g) / 48)
    bias = torch.randn(c, generator=g) * 0.1
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    cw = packing.ConvW(wt, [c], 1, 1, None, bias, False, BF, DEV, True)
    out = seg.like(c, BF, lambda shape: torch.zeros(shape, dtype=BF, device=DEV))
    n_img = sum(s[0] for s in shapes)
    st = torch.zeros((n_img, c // 8, 2), dtype=torch.float64, device=DEV)
    assert lib.conv2d([seg.flat], cw.w_tc, out.flat, c, 3, 1, 1, shift=cw.shift, engine=lib.ENGINE_TC, segs=seg.segs,
                      stats=st, stats_mode=2)
    torch.cuda.synchronize()
    img = 0
    convs = []
    for i, (n, h, w) in enumerate(shapes):
        y = out.level(i).view.double()                                          # stored conv outputs [n, h, w, c]
        convs.append(y.clone())
        ref_s = y.reshape(n, h * w, c // 8, 8).sum(dim=(1, 3))
        ref_q = (y * y).reshape(n, h * w, c // 8, 8).sum(dim=(1, 3))
        assert torch.allclose(st[img:img + n, :, 0], ref_s, rtol=1e-6, atol=1e-4)
        assert torch.allclose(st[img:img + n, :, 1], ref_q, rtol=1e-6, atol=1e-4)
        img += n
    lib.groupnorm_apply_seg(out.flat, out.segs, 32, gamma.to(DEV), beta.to(DEV), 1e-5, True, st)
    torch.cuda.synchronize()
    for i, (n, h, w) in enumerate(shapes):
        ref = F.relu(F.group_norm(convs[i].float().permute(0, 3, 1, 2), 32, gamma.to(DEV), beta.to(DEV), 1e-5))
        got = out.level(i).view.permute(0, 3, 1, 2).float()
        close(got.cpu(), ref.cpu())
        lv = out.level(i)
        if shared_halo:                                                       # nothing but interior pixels is non-zero
            total, inner = lv.buf.float().abs().sum().item(), lv.view.float().abs().sum().item()
            assert abs(total - inner) <= 1e-6 * max(1.0, inner), (total, inner)
        else:
            b = lv.buf.float()
            assert b[:, 0].abs().max() == 0 and b[:, -1].abs().max() == 0 and b[:, :, 0].abs().max() == 0 and b[:, :, -1].abs().max() == 0


@pytest.mark.parametrize("h,w,c,identity,full", [(20, 33, 256, False, False), (25, 42, 512, True, True), (7, 8, 64, False, True),
                                                 (50, 84, 768, True, True)])
def test_ese_apply_pool(h, w, c, identity, full):
    """x * gate (+ identity) fused with MaxPool2d(3, 2, ceil_mode=True) (vovnet.py:258-260, :349-350)."""
    g = torch.Generator().manual_seed(h * w + c)
    n = 2
    x = rb(torch.randn(n, c, h, w, generator=g))
    idn = rb(torch.randn(n, c, h, w, generator=g)) if identity else None
    gate = torch.rand(n, c, generator=g)
    y = x * gate.view(n, c, 1, 1)
    if idn is not None:
        y = y + idn
    y = rb(y)
    ref_pool = F.max_pool2d(y, 3, 2, ceil_mode=True)
    ho, wo = ref_pool.shape[-2:]
    fx, fi = halo(x), (halo(idn) if idn is not None else None)
    out_full = halo(torch.zeros(n, c, h, w)) if full else None
    out_pool = halo(torch.zeros(n, c, ho, wo))
    lib.ese_apply_pool(fx.view, gate.to(DEV), fi.view if fi is not None else None, out_full.view if full else None, out_pool.view)
    torch.cuda.synchronize()
    assert torch.equal(nchw(out_pool.view), ref_pool)
    if full:
        assert torch.equal(nchw(out_full.view), y)
    # flat variant (no pooling)
    out2 = halo(torch.zeros(n, c, h, w))
    lib.ese_apply_pool(fx.view, gate.to(DEV), fi.view if fi is not None else None, out2.view, None)
    torch.cuda.synchronize()
    assert torch.equal(nchw(out2.view), y)
    assert out2.buf[:, 0].abs().max() == 0 and out2.buf[:, :, -1].abs().max() == 0


def test_ese_gate_f64_matches_reference():
    g = torch.Generator().manual_seed(9)
    n, c, hw = 3, 256, 1234
    sums = torch.randn(n, c, generator=g, dtype=torch.float64) * hw
    w, b = torch.randn(c, c, generator=g) / 16, torch.randn(c, generator=g)
    gate = torch.zeros(n, c, device=DEV)
    lib.ese_gate_f64(sums.to(DEV), 1.0 / hw, w.to(DEV), b.to(DEV), gate, n, c)
    torch.cuda.synchronize()
    ref = F.relu6((sums / hw).float() @ w.t() + b + 3.0) / 6.0
    assert torch.allclose(gate.cpu(), ref, atol=2e-5)


def test_conv_tc_deconv_with_fused_mask_predictor():
    """out_mode 3: ConvTranspose2d(2, 2) + ReLU + class-gathered 1x1 predictor + sigmoid in one launch
    (sam.py:74-83, :96-97; mask_head.py:196-216) == the unfused deconv (bf16 store) followed by cm2_mask_predict."""
    g = torch.Generator().manual_seed(31)
    r, c, s, ncls = 7, 256, 14, 80
    x = rb(torch.randn(r, c, s, s, generator=g))
    wd = rb(torch.randn(c, c, 2, 2, generator=g) / 16)
    bd = torch.randn(c, generator=g) * 0.1
    wp = torch.randn(ncls, c, generator=g) / 16
    bp = torch.randn(ncls, generator=g) * 0.1
    classes = torch.randint(0, ncls, (r,), generator=g)
    cw = packing.deconv2x2({"d.weight": wd, "d.bias": bd}, "d", BF, DEV, True)
    probs = torch.full((r, 2 * s, 2 * s, 1), -1.0, device=DEV)
    assert lib.conv2d([halo(x).view], cw.w_tc, probs, 4 * c, 1, 1, 0, shift=cw.shift, relu=True, out_mode=3, engine=lib.ENGINE_TC,
                      probe=True, pred=(wp.to(DEV), bp.to(DEV), classes.to(DEV), ncls)), lib.last_error()
    torch.cuda.synchronize()
    up = rb(F.relu(F.conv_transpose2d(x, wd, bd, stride=2)))                       # what the unfused deconv stores
    logits = torch.einsum("rchw,rc->rhw", up, wp[classes]) + bp[classes].view(r, 1, 1)
    ref = torch.sigmoid(logits)
    got = probs[..., 0].cpu()
    assert (got - ref).abs().max().item() <= 2e-3, (got - ref).abs().max().item()
    # and against the unfused kernels of the library itself
    up_dev = torch.zeros((r, 2 * s, 2 * s, c), dtype=BF, device=DEV)
    assert lib.conv2d([halo(x).view], cw.w_tc, up_dev, 4 * c, 1, 1, 0, shift=cw.shift, relu=True, out_mode=1, engine=lib.ENGINE_TC, probe=True)
    p2 = torch.zeros((r, 1, 2 * s, 2 * s), device=DEV)
    lib.mask_predict(up_dev, wp.to(DEV), bp.to(DEV), classes.to(DEV), ncls, p2)
    torch.cuda.synchronize()
    assert (probs[..., 0] - p2[:, 0]).abs().max().item() <= 1e-5
