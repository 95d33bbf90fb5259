"""One-time repacking of reference ``state_dict`` tensors into the layouts ``libcm2.so`` consumes.

Packed buffers are derived data: they are rebuilt from the modules' parameters whenever those
change and are never serialised (SURVEY.md 8b "Weights").
"""
import torch

from . import lib

BN_EPS = 1e-5          # detectron2 FrozenBatchNorm2d eps [d2]


class ConvW(object):
    """Packed weights of one convolution / linear layer.

    ``w_simt``: [kh*kw*cin_total, cout] in the activation dtype (CM2_ENGINE_SIMT layout).
    ``w_tc``  : [cout_pad16, k_tc], K-major with every source's channels padded to 64 (CM2_ENGINE_TC layout):
                bf16 for bf16 models; for fp32 models (``dtype`` float32 with ``build_tc``) the *split-precision*
                f16 layout described below.
    ``scale`` / ``shift``: fp32 [cout] epilogue vectors (folded FrozenBN, or bias); ``scale_tc`` is what the TC launch
                uses (differs from ``scale`` only in the split layout).

    Split-precision layout (include/cm2.h, "Split precision"): ``w_tc`` = f16 [2 * cout_pad16, k_tc], rows
    [0, cout_pad) = W_hi = half(s * W), rows [cout_pad, 2 cout_pad) = W_lo = half(s * W - W_hi), s a per-output-channel
    power of two that lifts the largest |W| of the channel to [256, 512) -- far from the half subnormals, far from
    overflow -- and is undone exactly by ``scale_tc = scale / s``."""

    def __init__(self, weight, src_c, stride, pad, scale, shift, relu, dtype, device, build_tc):
        cout, cin, kh, kw = weight.shape
        assert cin == sum(src_c), (cin, src_c)
        self.k, self.stride, self.pad, self.cout, self.src_c, self.relu = kh, stride, pad, cout, list(src_c), relu
        w = weight.detach().to(torch.float32)
        self.w_simt = w.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout).contiguous().to(device=device, dtype=dtype)
        self.w_tc = None
        self.split = bool(build_tc and dtype == torch.float32)
        self.scale = None if scale is None else scale.detach().to(device=device, dtype=torch.float32).contiguous()
        self.shift = None if shift is None else shift.detach().to(device=device, dtype=torch.float32).contiguous()
        self.scale_tc = self.scale
        if build_tc:
            cout_pad = (cout + 15) // 16 * 16

            def pack(wf):
                # k = (tap, source, channel padded to 64): concatenate the sources inside each tap
                parts, off = [], 0
                for c in src_c:
                    cp = (c + 63) // 64 * 64
                    blk = torch.zeros((cout_pad, kh * kw, cp), dtype=torch.float32)
                    blk[:cout, :, :c] = wf[:, off:off + c].permute(0, 2, 3, 1).reshape(cout, kh * kw, c)
                    parts.append(blk)
                    off += c
                wt = torch.cat(parts, dim=2).reshape(cout_pad, -1)
                assert wt.shape[1] == lib.conv_tc_klen(kh, src_c)
                return wt

            if self.split:
                amax = w.abs().amax(dim=(1, 2, 3)).clamp(min=1e-30)
                pre = torch.exp2(torch.floor(torch.log2(256.0 / amax))).clamp(max=2.0 ** 40)   # power of two: exact rescale
                ws = w * pre.view(-1, 1, 1, 1)
                hi = ws.to(torch.float16).to(torch.float32)
                lo = (ws - hi).to(torch.float16).to(torch.float32)
                base = torch.ones(cout) if scale is None else scale.detach().to(torch.float32).cpu()
                self.scale_tc = (base / pre).to(device=device, dtype=torch.float32).contiguous()
                self.w_tc = torch.cat([pack(hi), pack(lo)], dim=0).contiguous().to(device=device, dtype=torch.float16)
            else:
                self.w_tc = pack(w).contiguous().to(device=device, dtype=torch.bfloat16)


def fold_frozen_bn(weight, bias, mean, var, eps=BN_EPS):
    """FrozenBatchNorm2d [d2]: y = (x - mean) * rsqrt(var + eps) * weight + bias  ->  scale, shift."""
    scale = weight.to(torch.float32) * torch.rsqrt(var.to(torch.float32) + eps)
    shift = bias.to(torch.float32) - mean.to(torch.float32) * scale
    return scale, shift


def conv_bn_relu(sd, prefix, src_c, stride, pad, dtype, device, tc):
    """conv (no bias) -> FrozenBN -> ReLU unit, vovnet.py:205-236."""
    scale, shift = fold_frozen_bn(sd[prefix + "/norm.weight"], sd[prefix + "/norm.bias"],
                                  sd[prefix + "/norm.running_mean"], sd[prefix + "/norm.running_var"])
    return ConvW(sd[prefix + "/conv.weight"], src_c, stride, pad, scale, shift, True, dtype, device, tc)


def dw_pw_bn_relu(sd, prefix, c, dtype, device, tc):
    """Depthwise unit, vovnet.py:110-130: (depthwise 3x3 weights as fp32 [9][c] for cm2_dwconv3x3, pointwise 1x1 ->
    FrozenBN -> ReLU as a ConvW).  The depthwise weights stay fp32 in every precision (c * 9 values)."""
    wd = sd[prefix + "/dw_conv3x3.weight"].detach().to(torch.float32)          # [c, 1, 3, 3]
    assert tuple(wd.shape) == (c, 1, 3, 3), (prefix, tuple(wd.shape))
    w9c = wd.reshape(c, 9).t().contiguous().to(device=device)
    scale, shift = fold_frozen_bn(sd[prefix + "/pw_norm.weight"], sd[prefix + "/pw_norm.bias"],
                                  sd[prefix + "/pw_norm.running_mean"], sd[prefix + "/pw_norm.running_var"])
    pw = ConvW(sd[prefix + "/pw_conv1x1.weight"], [c], 1, 0, scale, shift, True, dtype, device, tc)
    return w9c, pw


def conv_bias(sd, prefix, src_c, stride, pad, relu, dtype, device, tc):
    """conv with bias (no norm), optional ReLU."""
    return ConvW(sd[prefix + ".weight"], src_c, stride, pad, None, sd[prefix + ".bias"], relu, dtype, device, tc)


def deconv2x2(sd, prefix, dtype, device, tc):
    """ConvTranspose2d(k=2, s=2) + bias + ReLU (sam.py:74-80) as a 1x1 convolution to 4*cout columns:
    column j = (dy*2 + dx)*cout + co, scattered by the conv epilogue (out_mode 1)."""
    w = sd[prefix + ".weight"]                      # [cin, cout, 2, 2]
    cin, cout = w.shape[0], w.shape[1]
    w4 = w.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1, 1)
    bias = sd[prefix + ".bias"].repeat(4)
    return ConvW(w4, [cin], 1, 0, None, bias, True, dtype, device, tc)


def deconv4x4s2(sd, prefix, dtype, device, tc):
    """ConvTranspose2d(k=4, s=2, p=1) + bias, no activation (keypoint_head.py:205-208, :220) as a 3x3 / pad-1
    convolution to 4*cout columns: column j = (a*2 + b)*cout + co holds output pixel (2i + a, 2j + b).  Each phase
    uses a 2x2 subset of the 3x3 taps (the other five are zero): row phase a = 0 reads input rows i-1 (ky 3) and
    i (ky 1), a = 1 reads rows i (ky 2) and i+1 (ky 0); columns alike."""
    w = sd[prefix + ".weight"].detach().to(torch.float32)          # [cin, cout, 4, 4]
    cin, cout = w.shape[0], w.shape[1]
    assert tuple(w.shape[2:]) == (4, 4), tuple(w.shape)
    taps = {0: ((0, 3), (1, 1)), 1: ((1, 2), (2, 0))}              # phase -> ((tap index u+1, kernel index), ...)
    w3 = torch.zeros((4 * cout, cin, 3, 3), dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            blk = w3[(a * 2 + b) * cout:(a * 2 + b + 1) * cout]
            for u, ky in taps[a]:
                for v, kx in taps[b]:
                    blk[:, :, u, v] = w[:, :, ky, kx].t()
    bias = sd[prefix + ".bias"].detach().to(torch.float32).repeat(4)
    return ConvW(w3, [cin], 1, 1, None, bias, False, dtype, device, tc)


def linear(sd, prefix, relu, dtype, device, tc, chw=None):
    """nn.Linear as a 1x1 convolution over a [r, 1, 1, k] view.  ``chw`` = (c, h, w) re-orders the
    input features from the reference's flatten(C, H, W) order (maskiou_head.py:115) to NHWC."""
    w = sd[prefix + ".weight"]
    if chw is not None:
        c, h, ww = chw
        w = w.reshape(w.shape[0], c, h, ww).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    return ConvW(w[:, :, None, None], [w.shape[1]], 1, 0, None, sd[prefix + ".bias"], relu, dtype, device, tc)
