"""Checkpoint ingestion: what ``DetectionCheckpointer(model).load(path)`` [d2] does for the reference
(``/root/reference/convert_model_into_onnx.py:66``, ``tester.py:160-163``), without detectron2 / fvcore.

Accepted files: a torch ``.pth`` written by detectron2 (``{"model": state_dict, "iteration": ..}``), a bare
``state_dict``, or a model-zoo style ``.pkl`` (``{"model": {name: ndarray}, "__author__": ..}``, e.g. the published
``centermask2-V-39-eSE-FPN-ms-3x.pth`` is of the first kind, ``README.md:251-255``).  Parameter names are the
reference's, so no key surgery is needed beyond what ``DetectionCheckpointer`` itself does:
stripping a ``module.`` prefix (DataParallel / DDP) and ignoring the meta-arch's ``pixel_mean`` / ``pixel_std`` buffers
(derived from the cfg here).  After loading, the packed device copies (``packing.py``) are rebuilt lazily on the next
forward.
"""
import logging
import pickle
from collections import OrderedDict, namedtuple

import numpy as np
import torch

IncompatibleKeys = namedtuple("IncompatibleKeys", ["missing_keys", "unexpected_keys", "incorrect_shapes"])
_IGNORED = ("pixel_mean", "pixel_std")
_LOG = logging.getLogger("centermask2_b200.checkpoint")


def read_checkpoint(path):
    """Return the raw checkpoint object (dict) stored at ``path``."""
    if str(path).endswith(".pkl"):
        with open(path, "rb") as f:
            data = pickle.load(f, encoding="latin1")
        return data
    return torch.load(path, map_location="cpu", weights_only=False)


def extract_state_dict(ckpt):
    """Checkpoint object -> ``OrderedDict[name, Tensor]`` under the module's own key names."""
    sd = ckpt
    if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict):
        sd = ckpt["model"]
    elif isinstance(ckpt, dict) and "state_dict" in ckpt and isinstance(ckpt["state_dict"], dict):
        sd = ckpt["state_dict"]
    out = OrderedDict()
    for k, v in sd.items():
        if not isinstance(k, str):
            continue
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        if not isinstance(v, torch.Tensor):
            continue
        if k.startswith("module."):
            k = k[len("module."):]
        if k in _IGNORED:
            continue
        out[k] = v
    return out


def load_checkpoint(model, path_or_obj, strict=False):
    """Load a reference checkpoint into ``model`` (``GeneralizedRCNN`` or any sub-module with reference key names).

    Returns ``IncompatibleKeys(missing_keys, unexpected_keys, incorrect_shapes)`` like detectron2's Checkpointer; with
    ``strict=True`` any entry raises ``RuntimeError`` instead."""
    ckpt = path_or_obj if isinstance(path_or_obj, dict) else read_checkpoint(path_or_obj)
    sd = extract_state_dict(ckpt)
    own = model.state_dict()
    bad_shape = []
    for k in list(sd.keys()):
        if k in own and tuple(own[k].shape) != tuple(sd[k].shape):
            bad_shape.append((k, tuple(sd[k].shape), tuple(own[k].shape)))
            sd.pop(k)
    matched = [k for k in sd if k in own]
    if own and not matched:
        # wrong file / wrong prefix: loading "successfully" would leave every weight at its random initial value
        raise RuntimeError("checkpoint has no key in common with the model (first checkpoint keys: {}; first model keys: {})".format(
            list(sd)[:3], list(own)[:3]))
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if k not in _IGNORED]
    inc = IncompatibleKeys(missing, list(res.unexpected_keys), bad_shape)
    if strict and (inc.missing_keys or inc.unexpected_keys or inc.incorrect_shapes):
        raise RuntimeError("checkpoint does not match the model: missing {}, unexpected {}, shape mismatches {}".format(
            inc.missing_keys[:5], inc.unexpected_keys[:5], inc.incorrect_shapes[:5]))
    # detectron2's Checkpointer logs these; silence here would hide a checkpoint whose names do not match
    if inc.missing_keys:
        _LOG.warning("%d model keys are missing from the checkpoint and keep their initial values: %s%s", len(inc.missing_keys),
                     inc.missing_keys[:8], " ..." if len(inc.missing_keys) > 8 else "")
    if inc.unexpected_keys:
        _LOG.warning("%d checkpoint keys are not used by the model: %s%s", len(inc.unexpected_keys), inc.unexpected_keys[:8],
                     " ..." if len(inc.unexpected_keys) > 8 else "")
    for k, got, want in inc.incorrect_shapes:
        _LOG.warning("shape mismatch, skipped: %s checkpoint %s vs model %s", k, got, want)
    for m in model.modules():                       # packed device copies are derived state: rebuild on next use
        if hasattr(m, "_invalidate"):
            m._invalidate()
    return inc


class DetectionCheckpointer(object):
    """Name-compatible subset of detectron2's ``DetectionCheckpointer``: ``DetectionCheckpointer(model).load(path)``."""

    def __init__(self, model, save_dir="", **_unused):
        self.model = model
        self.save_dir = save_dir

    def load(self, path, checkpointables=None):
        if not path:
            return {}
        ckpt = read_checkpoint(path)
        self.incompatible = load_checkpoint(self.model, ckpt)
        return {k: v for k, v in ckpt.items() if k != "model"} if isinstance(ckpt, dict) else {}

    def save(self, name, **extra):
        import os
        data = {"model": self.model.state_dict()}
        data.update(extra)
        path = os.path.join(self.save_dir, "{}.pth".format(name))
        torch.save(data, path)
        return path
