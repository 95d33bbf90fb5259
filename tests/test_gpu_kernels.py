"""GPU: every libcm2 entry point against a plain torch / torchvision / oracle reference of the same op.

Tolerances are stated per test.  fp32 kernels accumulate in fp32 in a different order than ATen, so
dense ops are compared with rtol/atol ~1e-4; index / byte outputs are compared exactly."""
import math

import pytest
import torch
import torch.nn.functional as F
import torchvision

pytestmark = pytest.mark.gpu

from centermask2_b200 import lib, packing              # noqa: E402
from centermask2_b200.engine import FMap               # noqa: E402
from oracle import restate                             # noqa: E402

DEV = "cuda"


@pytest.fixture
def kernel_variant(monkeypatch):
    """Select the older implementation of a kernel family (CM2_<FAMILY>_VARIANT=0); the default (1) is the current one.
    The library reads the variable on every call."""
    def choose(family, variant):
        monkeypatch.setenv("CM2_{}_VARIANT".format(family), str(variant))
    return choose


def halo(t_nchw, dtype=torch.float32):
    n, c, h, w = t_nchw.shape
    buf = torch.zeros((n, h + 2, w + 2, c), dtype=dtype, device=DEV)
    buf[:, 1:-1, 1:-1] = t_nchw.permute(0, 2, 3, 1).to(DEV, dtype)
    return FMap(buf, 1)


def nchw(view):
    return view.permute(0, 3, 1, 2).float().cpu()


@pytest.mark.parametrize("cin,cout,k,stride,pad,h,w,n", [
    (3, 64, 3, 2, 1, 33, 47, 2), (64, 128, 3, 1, 1, 20, 24, 1), (160, 96, 3, 1, 1, 9, 13, 2),
    (256, 80, 3, 1, 1, 7, 11, 1), (128, 256, 1, 1, 0, 16, 20, 2), (256, 256, 3, 2, 1, 13, 21, 1),
    (257, 64, 3, 1, 1, 14, 14, 3), (512, 5, 3, 1, 1, 5, 5, 1)])
def test_conv_simt_f32(cin, cout, k, stride, pad, h, w, n):
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, wt, None, stride, pad) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    cw = packing.ConvW(wt, [cin], stride, pad, scale, shift, True, torch.float32, DEV, False)
    xin = halo(x)
    out = torch.zeros((n, ref.shape[2] + 2, ref.shape[3] + 2, cout), device=DEV)
    ov = out[:, 1:-1, 1:-1]
    lib.conv2d([xin.view], cw.w_simt, ov, cout, k, stride, pad, scale=cw.scale, shift=cw.shift, relu=True)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(ov), ref, rtol=1e-4, atol=1e-4)
    assert out[:, 0].abs().max() == 0 and out[:, :, 0].abs().max() == 0          # halo untouched


def test_conv_simt_virtual_concat_residual_upsample():
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(2, c, 8, 12, generator=g) for c in (32, 16, 24)]
    wt = torch.randn(40, 72, 1, 1, generator=g) / math.sqrt(72)
    bias = torch.randn(40, generator=g)
    low = torch.randn(2, 40, 4, 6, generator=g)
    ref = F.conv2d(torch.cat(xs, 1), wt, bias) + F.interpolate(low, scale_factor=2.0, mode="nearest")
    cw = packing.ConvW(wt, [32, 16, 24], 1, 0, None, bias, False, torch.float32, DEV, False)
    out = halo(torch.zeros(2, 40, 8, 12))
    lib.conv2d([halo(x).view for x in xs], cw.w_simt, out.view, 40, 1, 1, 0, shift=cw.shift,
               residual=halo(low).view, res_mode=2)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(out.view), ref, rtol=1e-4, atol=1e-4)


def test_conv_simt_deconv_scatter_and_in_relu():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 32, 7, 7, generator=g)
    wd = torch.randn(32, 24, 2, 2, generator=g) * 0.2
    bd = torch.randn(24, generator=g)
    ref = F.relu(F.conv_transpose2d(x, wd, bd, stride=2))
    cw = packing.deconv2x2({"d.weight": wd, "d.bias": bd}, "d", torch.float32, DEV, False)
    out = torch.zeros((3, 14, 14, 24), device=DEV)
    lib.conv2d([halo(x).view], cw.w_simt, out, 96, 1, 1, 0, shift=cw.shift, relu=True, out_mode=1)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(out), ref, rtol=1e-4, atol=1e-4)
    # in_relu: fpn.py:34
    wt = torch.randn(16, 32, 3, 3, generator=g) * 0.1
    ref2 = F.conv2d(F.relu(x), wt, None, 2, 1)
    cw2 = packing.ConvW(wt, [32], 2, 1, None, None, False, torch.float32, DEV, False)
    out2 = torch.zeros((3, 4, 4, 16), device=DEV)
    lib.conv2d([halo(x).view], cw2.w_simt, out2, 16, 3, 2, 1, in_relu=True)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(out2), ref2, rtol=1e-4, atol=1e-4)


def test_conv_bf16_simt_matches_bf16_rounded_reference():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 10, 10, generator=g).bfloat16().float()
    wt = (torch.randn(48, 64, 3, 3, generator=g) / 24).bfloat16().float()
    ref = F.conv2d(x, wt, None, 1, 1)
    cw = packing.ConvW(wt, [64], 1, 1, None, None, False, torch.bfloat16, DEV, False)
    out = halo(torch.zeros(1, 48, 10, 10), torch.bfloat16)
    lib.conv2d([halo(x, torch.bfloat16).view], cw.w_simt, out.view, 48, 3, 1, 1)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(out.view), ref, rtol=1e-2, atol=1e-2)          # one bf16 rounding of the output


@pytest.mark.parametrize("h,w", [(12, 16), (13, 17), (25, 42), (7, 7)])
def test_maxpool(h, w):
    x = torch.randn(2, 16, h, w)
    ref = F.max_pool2d(x, 3, 2, ceil_mode=True)
    out = halo(torch.zeros_like(ref))
    lib.maxpool3x3s2_ceil(halo(x).view, out.view)
    torch.cuda.synchronize()
    assert torch.equal(nchw(out.view), ref)


def test_ese_and_groupnorm():
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 64, 19, 23, generator=g)
    idn = torch.randn(2, 64, 19, 23, generator=g)
    fw = torch.randn(64, 64, generator=g) / 8
    fb = torch.randn(64, generator=g)
    gate = F.relu6(F.conv2d(F.adaptive_avg_pool2d(x, 1), fw[:, :, None, None], fb) + 3.0) / 6.0
    ref = x * gate + idn
    xin, out = halo(x), halo(torch.zeros_like(x))
    n, c, hw = 2, 64, 19 * 23
    ws = torch.empty(n * lib.ese_pool_chunks(hw) * c, device=DEV)
    pooled = torch.empty(n, c, device=DEV)
    gt = torch.empty(n, c, device=DEV)
    lib.ese_pool(xin.view, ws, pooled)
    lib.ese_gate(pooled, 1.0, fw.to(DEV), fb.to(DEV), gt, n, c)
    lib.ese_apply(xin.view, gt, halo(idn).view, out.view)
    torch.cuda.synchronize()
    assert torch.allclose(gt.cpu().view(2, 64, 1, 1), gate, atol=1e-5)
    assert torch.allclose(nchw(out.view), ref, rtol=1e-5, atol=1e-5)
    # GroupNorm(32) + ReLU, fcos.py:182-185
    x = torch.randn(2, 256, 13, 21, generator=g) * 3 + 1
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    ref = F.relu(F.group_norm(x, 32, gamma, beta, 1e-5))
    xin = halo(x)
    ws = torch.empty(lib.gn_workspace_floats(2, 13 * 21, 256, 32), device=DEV)
    lib.groupnorm_relu(xin.view, 32, gamma.to(DEV), beta.to(DEV), 1e-5, True, ws)
    torch.cuda.synchronize()
    assert torch.allclose(nchw(xin.view), ref, rtol=1e-4, atol=1e-4)
    assert xin.buf[:, 0].abs().max() == 0


def test_preprocess_u8_and_f32():
    img = (torch.rand(3, 37, 53) * 255).floor()
    mean, std = [103.53, 116.28, 123.675], [1.0, 1.0, 1.0]
    ref = torch.zeros(3, 64, 64)
    ref[:, :37, :53] = img - torch.tensor(mean).view(3, 1, 1)
    for t in (img, img.to(torch.uint8)):
        out = halo(torch.ones(2, 3, 64, 64))
        lib.preprocess_image(t.to(DEV), mean, std, out.view, 1)
        torch.cuda.synchronize()
        assert torch.allclose(nchw(out.view)[1], ref, atol=1e-5)
        assert nchw(out.view)[0].min() == 1                                     # other batch slot untouched


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("stride,h,w,c,n", [(1, 13, 17, 64, 2), (2, 13, 17, 64, 2), (2, 12, 16, 80, 1), (1, 1, 1, 8, 1), (1, 5, 2, 112, 3)])
def test_dwconv3x3_matches_grouped_conv2d(stride, h, w, c, n, dtype):
    """Depthwise 3x3 of the dw bodies (vovnet.py:110-130) against F.conv2d(groups=c) on halo and dense views."""
    g = torch.Generator().manual_seed(31 + stride + h)
    x = torch.randn(n, c, h, w, generator=g)
    wt = torch.randn(c, 1, 3, 3, generator=g) * 0.3
    if dtype == torch.bfloat16:
        x = x.to(torch.bfloat16).float()
    ref = F.conv2d(x, wt, None, stride, 1, 1, c)
    w9c = wt.reshape(c, 9).t().contiguous().to(DEV)
    out = halo(torch.full_like(ref, 7.0), dtype)
    lib.dwconv3x3(halo(x, dtype).view, out.view, w9c, stride)
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert torch.allclose(nchw(out.view), ref, rtol=tol, atol=tol)
    assert out.buf[:, 0].abs().max() == 0 and out.buf[:, :, 0].abs().max() == 0          # halo untouched
    with pytest.raises(RuntimeError):
        lib.dwconv3x3(halo(x, dtype).view, out.view, w9c, 3)


def _random_boxes(g, n, w, h):
    cx, cy = torch.rand(n, generator=g) * w, torch.rand(n, generator=g) * h
    bw, bh = torch.rand(n, generator=g) * w * 0.5 + 2, torch.rand(n, generator=g) * h * 0.5 + 2
    return torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)


# 4: x pass on the tensor cores (bf16 maps only: on the fp32 maps of this test it runs the column walk);
# 2: CTA per ROI, column walk (separable, rows carried in registers); 1: CTA per ROI, merged taps;
# 0: thread per (bin, 8 channels).  c = 256: a warp is one bin column (the production shape), c = 32: columns share warps
@pytest.mark.parametrize("c", [32, 256])
@pytest.mark.parametrize("variant", [4, 3, 2, 1, 0])
def test_roialign_fpn_matches_torchvision_and_reference_level_rule(variant, c, kernel_variant):
    kernel_variant("ROIALIGN", variant)
    g = torch.Generator().manual_seed(7)
    n, r_cap = 2, 24
    H, W = 96, 128
    feats = [torch.randn(n, c, H // s, W // s, generator=g) for s in (8, 16, 32)]
    counts = [24, 17]
    boxes = torch.zeros(n, r_cap, 4)
    for i in range(n):
        boxes[i, :counts[i]] = _random_boxes(g, counts[i], W, H)
    boxes[0, 0] = torch.tensor([0.0, 0.0, W, H])                    # whole image -> P5
    boxes[0, 1] = torch.tensor([10.0, 10.0, 10.0, 30.0])            # zero area -> P3
    sizes = [(H, W), (H - 6, W - 10)]
    dets = [{"pred_boxes": boxes[i, :counts[i]], "image_size": sizes[i]} for i in range(n)]

    class Cfg:                      # only what restate.roi_pool reads
        class MODEL:
            class ROI_HEADS:
                IN_FEATURES = ["p3", "p4", "p5"]
            class ROI_MASK_HEAD:
                POOLER_RESOLUTION = 14
                POOLER_SAMPLING_RATIO = 0
                ASSIGN_CRITERION = "ratio"
    ref, lv = restate.roi_pool({"p3": feats[0], "p4": feats[1], "p5": feats[2]}, dets, Cfg)
    out = halo(torch.zeros(n * r_cap, c, 14, 14))
    lvl_out = torch.full((n * r_cap,), -1, dtype=torch.int32, device=DEV)
    area = torch.tensor([float(h * w) for h, w in sizes], device=DEV)
    lib.roialign_fpn([halo(f).view for f in feats], [8, 16, 32], boxes.to(DEV), torch.tensor(counts, dtype=torch.int32, device=DEV),
                     n, r_cap, area, 0, 0, out.view, lvl_out)
    torch.cuda.synchronize()
    got = nchw(out.view)
    sel = torch.cat([torch.arange(counts[i]) + i * r_cap for i in range(n)])
    assert torch.equal(lvl_out.cpu()[sel].long(), lv)
    assert torch.allclose(got[sel], ref, rtol=1e-4, atol=1e-4)
    assert got[r_cap + counts[1]:].abs().max() == 0                  # invalid slots are zeroed


@pytest.mark.parametrize("variant", [4, 3, 2, 1, 0])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roialign_large_rois_separable_path(dtype, variant, kernel_variant):
    kernel_variant("ROIALIGN", variant)
    """Adaptive sampling grids of 2..20 samples per bin side: the separable (row / column weight) evaluation, boxes that
    stick out of the image, and bins wider than the register span (sample-loop fallback) against torchvision."""
    import torchvision
    g = torch.Generator().manual_seed(17)
    n, r_cap, c = 2, 16, 64
    H, W = 640, 1024
    strides = (4, 8, 16)                                      # P"3" = 160 x 256: bins up to 19 pixels wide
    feats = [torch.randn(n, c, H // s, W // s, generator=g) for s in strides]
    if dtype == torch.bfloat16:
        feats = [f.to(torch.bfloat16).float() for f in feats]
    boxes = torch.zeros(n, r_cap, 4)
    for i in range(n):
        cx, cy = torch.rand(r_cap, generator=g) * W, torch.rand(r_cap, generator=g) * H
        bw, bh = torch.rand(r_cap, generator=g) * W * 0.9 + 8, torch.rand(r_cap, generator=g) * H * 0.9 + 8
        boxes[i] = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)       # many stick out of the image
    boxes[0, 0] = torch.tensor([0.0, 0.0, W, H])
    boxes[0, 1] = torch.tensor([-300.0, -200.0, 90.0, 120.0])
    boxes[1, 0] = torch.tensor([W - 60.0, H - 40.0, W + 500.0, H + 300.0])
    counts = torch.tensor([r_cap, r_cap - 3], dtype=torch.int32)
    area = torch.tensor([float(H * W)] * n)
    out = halo(torch.zeros(n * r_cap, c, 14, 14), dtype)
    lvl = torch.full((n * r_cap,), -1, dtype=torch.int32, device=DEV)
    lib.roialign_fpn([halo(f, dtype).view for f in feats], list(strides), boxes.to(DEV), counts.to(DEV), n, r_cap, area.to(DEV), 0, 0,
                     out.view, lvl)
    torch.cuda.synchronize()
    got = nchw(out.view)
    lv = lvl.cpu()
    for i in range(n):
        for r in range(int(counts[i])):
            slot = i * r_cap + r
            li = int(lv[slot])
            roi = torch.cat([torch.tensor([float(i)]), boxes[i, r]])[None]
            ref = torchvision.ops.roi_align(feats[li], roi, 14, 1.0 / strides[li], 0, True)[0]
            tol = 1e-4 if dtype == torch.float32 else 1e-2
            assert torch.allclose(got[slot], ref, rtol=tol, atol=tol), (i, r, li, (got[slot] - ref).abs().max())
    assert got[r_cap + int(counts[1]):].abs().max() == 0


@pytest.mark.parametrize("split", [1, 0])
@pytest.mark.parametrize("stages", [3, 4])
@pytest.mark.parametrize("c,pitched", [(256, False), (256, True), (64, False), (32, False)])
def test_roialign_tensor_core_x_pass_against_fp32_sampling(c, pitched, stages, split, kernel_variant, monkeypatch):
    """Variant 4 (mma.sync x pass, bf16 maps) against torchvision's fp32 sampling of the same bf16 values: boxes from 1 pixel
    to the whole image (tap ranges of one and of several 32-column chunks, more than ROI_MMA_KMAX = 128 columns -> sample
    loop), boxes that stick out of the image, empty slots; channel-dense (compile-time stride for c = 256) and pitched maps.
    split = 1 (bf16 hi + lo x weights) has to agree with the fp32 evaluation to the output rounding (bf16, 2^-9 relative:
    2e-2 on values up to ~4); split = 0 (hi only) carries a 2^-9 weight error on top."""
    import torchvision
    kernel_variant("ROIALIGN", 4)
    monkeypatch.setenv("CM2_ROIALIGN_STAGES", str(stages))
    monkeypatch.setenv("CM2_ROIALIGN_SPLIT", str(split))
    g = torch.Generator().manual_seed(23 + c)
    n, r_cap = 2, 40
    H, W = 512, 1536
    strides = (8, 16, 32)                                     # P3 = 64 x 192: tap ranges up to the whole row (> 128 columns)
    feats = [torch.randn(n, c, H // s, W // s, generator=g).to(torch.bfloat16).float() for s in strides]
    boxes = torch.zeros(n, r_cap, 4)
    for i in range(n):
        area = torch.exp(torch.rand(r_cap, generator=g) * (math.log(0.9 * H * W) - math.log(8.0 * 8.0)) + math.log(8.0 * 8.0))
        ar = torch.exp((torch.rand(r_cap, generator=g) - 0.5) * 3.0)
        bw, bh = torch.sqrt(area * ar), torch.sqrt(area / ar)
        cx, cy = torch.rand(r_cap, generator=g) * W, torch.rand(r_cap, generator=g) * H
        boxes[i] = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)     # some stick out of the image
    boxes[0, 0] = torch.tensor([0.0, 0.0, W, H])                        # whole image
    boxes[0, 1] = torch.tensor([10.0, 10.0, 10.0, 30.0])                # zero area
    boxes[0, 2] = torch.tensor([3.0, 40.0, W - 5.0, 70.0])              # P3, 191 tap columns: wider than the walk handles
    boxes[0, 3] = torch.tensor([100.0, 100.0, 101.0, 101.0])            # one pixel: every bin taps the same two rows / columns
    boxes[0, 4] = torch.tensor([-400.0, -300.0, -100.0, -50.0])         # entirely outside
    boxes[1, 0] = torch.tensor([200.0, 8.0, 1190.0, 120.0])             # P3, 125 tap columns = four chunks
    boxes[1, 1] = torch.tensor([W - 60.0, H - 40.0, W + 500.0, H + 300.0])
    counts = torch.tensor([r_cap, r_cap - 7], dtype=torch.int32)
    area = torch.tensor([float(H * W)] * n)
    out = halo(torch.full((n * r_cap, c, 14, 14), 7.0), torch.bfloat16)
    lvl = torch.full((n * r_cap,), -1, dtype=torch.int32, device=DEV)
    fm = [halo(f, torch.bfloat16) for f in feats]
    if pitched:                                               # channel pitch 2 c: the run-time stride instantiation
        fm = []
        for f in feats:
            buf = torch.zeros((n, f.shape[2] + 2, f.shape[3] + 2, 2 * c), dtype=torch.bfloat16, device=DEV)
            buf[:, 1:-1, 1:-1, :c] = f.permute(0, 2, 3, 1).to(DEV, torch.bfloat16)
            fm.append(FMap(buf[..., :c], 1))
    lib.roialign_fpn([f.view for f in fm], list(strides), boxes.to(DEV), counts.to(DEV), n, r_cap, area.to(DEV), 0, 0, out.view, lvl)
    torch.cuda.synchronize()
    got = nchw(out.view)
    lv = lvl.cpu()
    worst = 0.0
    for i in range(n):
        for r in range(int(counts[i])):
            slot = i * r_cap + r
            li = int(lv[slot])
            roi = torch.cat([torch.tensor([float(i)]), boxes[i, r]])[None]
            ref = torchvision.ops.roi_align(feats[li], roi, 14, 1.0 / strides[li], 0, True)[0]
            tol = 1e-2 if split else 3e-2
            assert torch.allclose(got[slot], ref, rtol=tol, atol=tol), (i, r, li, boxes[i, r], (got[slot] - ref).abs().max())
            worst = max(worst, float((got[slot] - ref).abs().max()))
    print("variant 4 c={} split={}: max |diff| against fp32 sampling {:.3g}".format(c, split, worst))
    assert got[r_cap + int(counts[1]):].abs().max() == 0          # empty slots are zeroed
    assert out.buf[:, 0].abs().max() == 0 and out.buf[:, :, 0].abs().max() == 0          # halo untouched


def test_spatial_attention_mask_predict_maskiou_glue():
    g = torch.Generator().manual_seed(8)
    r, c = 5, 64
    x = torch.randn(r, c, 14, 14, generator=g)
    w = torch.randn(1, 2, 3, 3, generator=g) * 0.5
    att = torch.sigmoid(F.conv2d(torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1), w, None, 1, 1))
    out = halo(torch.zeros_like(x))
    lib.spatial_attention(halo(x).view, out.view, w.reshape(18).to(DEV))
    torch.cuda.synchronize()
    assert torch.allclose(nchw(out.view), x * att, rtol=1e-5, atol=1e-5)
    # class-gathered predictor + sigmoid
    y = torch.randn(r, c, 28, 28, generator=g)
    wp, bp = torch.randn(80, c, generator=g) * 0.1, torch.randn(80, generator=g) * 0.1
    cls = torch.randint(0, 80, (r,), generator=g)
    logits = F.conv2d(y, wp[:, :, None, None], bp)
    ref = logits[torch.arange(r), cls][:, None].sigmoid()
    yb = y.permute(0, 2, 3, 1).contiguous().to(DEV)
    probs = torch.empty(r, 1, 28, 28, device=DEV)
    lib.mask_predict(yb, wp.to(DEV), bp.to(DEV), cls.to(DEV), 80, probs)
    torch.cuda.synchronize()
    assert torch.allclose(probs.cpu(), ref, rtol=1e-5, atol=1e-5)
    # 2x2 max pool into channel 0
    pm = halo(torch.ones(r, 16, 14, 14))
    lib.maskiou_prep(probs, pm.view)
    torch.cuda.synchronize()
    got = nchw(pm.view)
    assert torch.equal(got[:, :1], F.max_pool2d(probs.cpu(), 2, 2)) and got[:, 1:].abs().max() == 0
    iou = torch.randn(r, 80, generator=g)
    sc = torch.rand(r, generator=g)
    ms = torch.empty(r, device=DEV)
    lib.maskiou_score(iou.to(DEV), r, 80, cls.to(DEV), sc.to(DEV), ms)
    torch.cuda.synchronize()
    assert torch.allclose(ms.cpu(), sc * iou[torch.arange(r), cls], atol=1e-6)


@pytest.mark.parametrize("c", [256, 64])
def test_spatial_attention_bf16_pipelined_equals_cta_per_roi(c, kernel_variant):
    """>= 296 ROIs select the persistent double-buffered kernel (bulk copies + mbarriers); it must equal the CTA-per-ROI
    kernel bit for bit (same reduction order) and torch within bf16 rounding (rtol 1e-2)."""
    g = torch.Generator().manual_seed(21 + c)
    r = 333
    x = torch.randn(r, c, 14, 14, generator=g).to(torch.bfloat16).float()
    w = torch.randn(1, 2, 3, 3, generator=g) * 0.5
    att = torch.sigmoid(F.conv2d(torch.cat([x.mean(1, keepdim=True), x.max(1, keepdim=True)[0]], 1), w, None, 1, 1))
    xin = halo(x, torch.bfloat16)
    outs = []
    for variant in (1, 0):
        kernel_variant("SAM", variant)
        out = halo(torch.zeros_like(x), torch.bfloat16)
        lib.spatial_attention(xin.view, out.view, w.reshape(18).to(DEV))
        torch.cuda.synchronize()
        outs.append(out.buf.clone())
    assert torch.equal(outs[0], outs[1])
    got = outs[0][:, 1:-1, 1:-1].permute(0, 3, 1, 2).float().cpu()
    assert torch.allclose(got, x * att, rtol=1e-2, atol=1e-2)
    assert outs[0][:, 0].abs().max() == 0 and outs[0][:, :, 0].abs().max() == 0        # halo untouched


# fused path (plane % 16 == 0: aligned rows / odd width with chunks straddling rows), word path, byte path
@pytest.mark.parametrize("out_h,out_w", [(96, 128), (80, 101), (75, 101), (76, 101), (40, 67)])
def test_paste_masks_and_box_rescale(out_h, out_w):
    g = torch.Generator().manual_seed(9)
    r = 12
    probs = torch.rand(r, 28, 28, generator=g)
    boxes = _random_boxes(g, r, 128, 96)
    boxes[0] = torch.tensor([-5.0, -5.0, 200.0, 200.0])
    boxes[1] = torch.tensor([20.0, 20.0, 20.0, 40.0])                # empty after clipping -> dropped
    det = {"image_size": (96, 128), "pred_boxes": boxes, "scores": torch.rand(r, generator=g),
           "pred_masks": probs[:, None]}
    ref = restate.detector_postprocess(det, out_h, out_w)
    b2 = torch.empty(r, 4, device=DEV)
    valid = torch.empty(r, dtype=torch.uint8, device=DEV)
    lib.scale_clip_boxes(boxes.to(DEV), b2, valid, r, out_w / 128, out_h / 96, float(out_w), float(out_h))
    masks = torch.empty(r, out_h, out_w, dtype=torch.uint8, device=DEV)
    lib.paste_masks(probs.to(DEV), b2, valid, masks, r, 28, out_h, out_w, 0.5)
    torch.cuda.synchronize()
    keep = valid.cpu().bool()
    assert keep.sum() == len(ref["scores"]) and not keep[1]
    assert torch.allclose(b2.cpu()[keep], ref["pred_boxes"], atol=1e-4)
    got = masks.cpu()[keep].bool()
    # bit-exact except pixels whose interpolated value is within float rounding of the 0.5 threshold
    diff = (got != ref["pred_masks"]).sum().item()
    assert diff <= 1e-4 * got.numel(), diff
    assert masks.cpu()[~keep].sum() == 0


@pytest.mark.parametrize("out_h,out_w", [(96, 128), (80, 101), (208, 333), (800, 1333)])
def test_paste_fused_equals_memset_window_path(out_h, out_w, kernel_variant):
    """The fused single-pass kernel (variant 1) writes every byte once and must equal the memset + window kernel
    (variant 0) bit for bit: boxes touching every image edge, full-image, sub-pixel, one-strip and invalid boxes;
    the output buffer is pre-filled with garbage."""
    g = torch.Generator().manual_seed(out_h * 7 + out_w)
    r = 24
    probs = torch.rand(r, 28, 28, generator=g)
    W, H = float(out_w), float(out_h)
    boxes = _random_boxes(g, r, out_w, out_h)
    boxes[0] = torch.tensor([0.0, 0.0, W, H])
    boxes[1] = torch.tensor([0.0, H * 0.3, W * 0.2, H * 0.6])            # touches the left edge
    boxes[2] = torch.tensor([W * 0.7, H * 0.1, W, H * 0.5])              # touches the right edge
    boxes[3] = torch.tensor([0.0, 0.0, W, 3.5])                          # top rows, full width
    boxes[4] = torch.tensor([0.0, H - 2.25, W, H])                       # bottom rows, full width
    boxes[5] = torch.tensor([W * 0.5, H * 0.5, W * 0.5 + 0.4, H * 0.5 + 0.3])   # sub-pixel
    boxes[6] = torch.tensor([3.0, 2.0, 9.5, H - 1.0])                    # narrow and tall (one strip)
    boxes[7] = torch.tensor([W - 1.5, 0.0, W, H])                        # last column
    boxes[8] = torch.tensor([0.0, 0.0, 1.0, H])                          # first column
    valid = torch.ones(r, dtype=torch.uint8)
    valid[9] = 0
    outs = []
    for variant in (1, 0):
        kernel_variant("PASTE", variant)
        masks = torch.full((r, out_h, out_w), 0xAB, dtype=torch.uint8, device=DEV)
        lib.paste_masks(probs.to(DEV), boxes.to(DEV), valid.to(DEV), masks, r, 28, out_h, out_w, 0.5)
        torch.cuda.synchronize()
        outs.append(masks.cpu())
    assert outs[0].max() <= 1
    assert torch.equal(outs[0], outs[1]), (outs[0] != outs[1]).sum().item()
    assert outs[0][9].sum() == 0 and outs[0][0].sum() > 0


def _fcos_post_case(g, n, sizes, ncls, target, h0, w0):
    """Random head outputs with ~target candidates per (image, level)."""
    strides = [8, 16, 32, 64, 128][:len(sizes)]
    logits, regs, ctrs = [], [], []
    for (h, w) in sizes:
        p = min(0.5, target / float(h * w * ncls))
        cut = math.log(0.05 / 0.95)
        lg = torch.randn(n, ncls, h, w, generator=g) * 2.0
        kth = torch.quantile(lg.flatten()[:200000], 1 - p).item() if p < 1 else -1e9
        logits.append(lg + (cut - kth))
        regs.append(torch.rand(n, 4, h, w, generator=g) * 6)
        ctrs.append(torch.randn(n, 1, h, w, generator=g))
    return logits, regs, ctrs, strides


@pytest.mark.parametrize("nms_variant", [1, 0])              # 1: multi-way merge of the sorted levels; 0: bitonic sort
@pytest.mark.parametrize("variant", [1, 0])                  # 1: flat 32 KB tiles; 0: one CTA per image row
@pytest.mark.parametrize("target,post", [(150, 50), (1500, 100)])
def test_fcos_postprocess_matches_oracle(target, post, variant, nms_variant, kernel_variant):
    kernel_variant("DECODE", variant)
    kernel_variant("NMS", nms_variant)
    from centermask2_b200.config import get_cfg
    from centermask2_b200.engine import Engine
    g = torch.Generator().manual_seed(10 + target)
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.FCOS.POST_NMS_TOPK_TEST", post])
    n, ncls = 2, 80
    sizes = [(24, 32), (12, 16), (6, 8), (3, 4), (2, 2)]
    logits, regs, ctrs, strides = _fcos_post_case(g, n, sizes, ncls, target, 192, 256)
    ref = restate.fcos_postprocess(logits, regs, ctrs, [(192, 256)] * n, cfg, pre_topk=True)
    eng = Engine(cfg, "fp32", DEV)
    head = []
    for lg, rg, ct in zip(logits, regs, ctrs):
        lv = FMap(lg.permute(0, 2, 3, 1).contiguous().to(DEV), 0)
        rc = FMap(torch.cat([rg, ct], 1).permute(0, 2, 3, 1).contiguous().to(DEV), 0)
        head.append((lv, rc))
    det = eng.run_fcos_post(head)
    torch.cuda.synchronize()
    assert (det["cand_count"] <= det["cand_cap"]).all()
    for i in range(n):
        k = int(det["count"][i])
        assert k == len(ref[i]["scores"]), (k, len(ref[i]["scores"]))
        assert torch.equal(det["classes"][i, :k].cpu(), ref[i]["pred_classes"])
        assert torch.equal(det["locations"][i, :k].cpu(), ref[i]["locations"])
        assert torch.allclose(det["boxes"][i, :k].cpu(), ref[i]["pred_boxes"], atol=1e-4)
        assert torch.allclose(det["scores"][i, :k].cpu(), ref[i]["scores"], atol=1e-6)
    if nms_variant == 1:                                          # both orderings keep exactly the same detections
        first = {k: det[k].clone() for k in ("boxes", "scores", "classes", "locations", "count")}
        kernel_variant("NMS", 0)
        det0 = eng.run_fcos_post(head)
        torch.cuda.synchronize()
        for k, v in first.items():
            assert torch.equal(v, det0[k]), k
