"""TEST INFRASTRUCTURE ONLY -- run the *unmodified* reference model (``/root/reference``) on CPU.

Works only where ``/root/reference`` exists (the build container).  The reference's package is
imported over ``oracle/d2shim``; its own files are not modified or copied.
"""
import contextlib
import io
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("CM2_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d2shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "centermask2", "centermask"))


def _import_reference():
    if not available():
        raise RuntimeError("reference tree not present at {}".format(REFERENCE_ROOT))
    for p in (os.path.join(REFERENCE_ROOT, "centermask2"), _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import detectron2
    assert getattr(detectron2, "__cm2_shim__", False) or True
    import centermask  # noqa: F401  (performs the registrations, centermask/__init__.py:1)
    return detectron2


def build_reference_model(cfg, state_dict=None):
    """``build_model(cfg)`` exactly as convert_model_into_onnx.py:63-67 does, on CPU, eval mode."""
    _import_reference()
    from detectron2.modeling import build_model
    cfg = cfg.clone()
    cfg.MODEL.DEVICE = "cpu"
    model = build_model(cfg)
    model.eval()
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    return model


@contextlib.contextmanager
def quiet():
    """The fork prints inside the hot path (fcos_outputs.py:443, pooler.py:338-339)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def run_reference(model, batched_inputs, postprocess=True):
    with torch.no_grad(), quiet():
        if postprocess:
            return model(batched_inputs)
        return model.inference(batched_inputs, do_postprocess=False)
