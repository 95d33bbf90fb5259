"""Deterministic synthetic weights and inputs (there is no network for checkpoints or datasets).

The recipe is independent of any model object: every ``state_dict`` entry is drawn from a CPU
generator seeded by ``(seed, crc32(key))`` with a distribution chosen per parameter kind so that
activations keep O(1) scale through the ~60 layers (SURVEY.md 8d "Synthetic inputs").  Stock
initialisation would give sigmoid(cls) ~ 0.01 < 0.05 and therefore *zero* detections
(SURVEY.md section 0), so ``cls_logits.bias`` is left for the caller to calibrate
(``calibrate_cls_bias``).  FrozenBN buffers are randomised so that the scale/shift fold is tested.
"""
import math
import zlib

import torch

from .arch import model_param_spec

__all__ = ["synthetic_state_dict", "synthetic_images", "calibrate_cls_bias"]


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode("utf-8"))) % (2 ** 63 - 1))
    return g


def _draw(kind, shape, g, key):
    def normal(std, mean=0.0):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean

    def uniform(lo, hi):
        return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    if kind == "conv_relu" or kind == "fc_relu":
        return normal(math.sqrt(2.0 / fan_in))
    if kind == "conv_relu_residual":
        return normal(0.35 * math.sqrt(2.0 / fan_in))
    if kind == "conv_linear":
        return normal(math.sqrt(1.0 / fan_in))
    if kind == "deconv":          # (Cin, Cout, 2, 2), stride 2: one tap per output pixel
        return normal(math.sqrt(2.0 / shape[0]))
    if kind == "kp_deconv":       # (Cin, K, 4, 4), stride 2, pad 1: four taps per output pixel.  The synthetic FPN
        # features are O(70), so the gain is small: logits of O(2-3) spread keep the scores exp(l - max) / sum in fp32 range
        return normal(0.02 * math.sqrt(4.0 / shape[0]))
    if kind == "bn_weight":
        return uniform(0.5, 1.5)
    if kind == "bn_var":
        return uniform(0.5, 1.5)
    if kind in ("bn_bias", "bn_mean", "bias", "gn_bias"):
        return normal(0.1)
    if kind == "gn_weight":
        return uniform(0.5, 1.5)
    if kind == "ese_weight":
        return normal(1.0 / math.sqrt(fan_in))
    if kind == "ese_bias":
        return normal(1.0)
    if kind == "cls_logits":
        return normal(0.05)
    if kind == "cls_bias":
        return torch.full(shape, -math.log(99.0))          # fcos.py:218-220 prior; calibrate later
    if kind == "bbox_pred":
        return normal(0.02)
    if kind == "bbox_bias":
        return torch.full(shape, 4.0)
    if kind == "ctrness":
        return normal(0.01)
    if kind == "zero":
        return torch.zeros(shape)
    if kind == "scale":
        lvl = int(key.split(".")[-2])
        return torch.full(shape, (1.0, 0.9, 1.1, 0.8, 1.2, 1.0, 1.0, 1.0)[lvl])
    if kind == "sam":
        return normal(0.5)
    if kind == "predictor":
        return normal(0.05)
    if kind == "maskiou_out":
        return normal(0.02)
    if kind == "maskiou_bias":
        return torch.full(shape, 0.5)
    raise KeyError(kind)


def synthetic_state_dict(cfg, seed=0):
    """Full ``GeneralizedRCNN`` ``state_dict`` (reference key names) of seeded random weights."""
    out = {}
    for key, (shape, kind) in model_param_spec(cfg).items():
        out[key] = _draw(kind, tuple(shape), _gen(seed, key), key)
    return out


def synthetic_images(n, height, width, seed=2):
    """``n`` BGR float images in [0, 255), CHW fp32, in the reference's ``batched_inputs`` format."""
    out = []
    for i in range(n):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + i)
        # smooth-ish content: low-res noise upsampled + fine noise, so features are not white
        base = torch.rand(1, 3, (height + 15) // 16, (width + 15) // 16, generator=g)
        base = torch.nn.functional.interpolate(base, size=(height, width), mode="bilinear", align_corners=False)[0]
        fine = torch.rand(3, height, width, generator=g)
        img = (0.7 * base + 0.3 * fine) * 255.0
        out.append({"image": img.contiguous(), "height": height, "width": width})
    return out


def calibrate_cls_bias(raw_logits_per_level, target_per_level, thresh=0.05):
    """Pick one ``cls_logits.bias`` value so that about ``target_per_level`` entries per (image,
    level) satisfy ``sigmoid(logit) > thresh`` on the *busiest* level.

    ``raw_logits_per_level``: list of tensors [N, ...] computed with bias 0.  Returns a float."""
    cut = math.log(thresh / (1.0 - thresh))
    best = None
    for t in raw_logits_per_level:
        n = t.shape[0]
        flat = t.reshape(n, -1).float()
        k = min(target_per_level, flat.shape[1])
        kth = torch.topk(flat, k, dim=1).values[:, -1].min().item()
        b = cut - kth
        best = b if best is None else min(best, b)
    return float(best) - 1e-3
