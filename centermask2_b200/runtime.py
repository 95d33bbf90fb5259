"""Process-wide runtime state: one ``Engine`` (buffer cache) per (cfg identity, precision, device).

Precision is a deployment knob, not part of the reference's cfg: ``set_precision("bf16" | "fp32")``
or the ``CM2_PRECISION`` environment variable (default ``fp32`` -- the strict-parity variant), or a
``MODEL.B200.PRECISION`` cfg key when present.
"""
import os

import torch

# "bf16": bf16 activations + tcgen05 convolutions (the benchmarked engine); "fp32": fp32 activations, convolutions on the
# tensor cores with fp16 hi/lo split operands (fp32-grade accuracy); "fp32_simt": fp32 activations, CUDA-core convolutions
PRECISIONS = ("fp32", "fp32_simt", "bf16")
_precision = os.environ.get("CM2_PRECISION", "fp32")
_engines = {}


def set_precision(p):
    global _precision
    assert p in PRECISIONS, p
    _precision = p


def precision_for(cfg):
    b200 = cfg.MODEL.get("B200") if hasattr(cfg.MODEL, "get") else None
    if b200 is not None and "PRECISION" in b200:
        return b200["PRECISION"]
    return _precision


def engine_for(cfg):
    from .engine import Engine
    if not torch.cuda.is_available():
        raise RuntimeError("centermask2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = str(getattr(cfg.MODEL, "DEVICE", "cuda"))
    if not dev.startswith("cuda"):
        raise RuntimeError("MODEL.DEVICE={!r}: centermask2_b200 runs on CUDA only".format(dev))
    device = torch.device(dev if ":" in dev else "cuda:{}".format(torch.cuda.current_device()))
    key = (id(cfg), precision_for(cfg), str(device))
    eng = _engines.get(key)
    if eng is None:
        eng = Engine(cfg, precision_for(cfg), device)
        _engines[key] = eng
        eng._cfg_ref = cfg             # keep cfg alive so id() stays unique
    return eng


def reset():
    for e in _engines.values():
        e.release()
    _engines.clear()
