"""TEST INFRASTRUCTURE ONLY -- import-surface stubs of ``detectron2.data`` for ``/root/reference/deploy_utils.py:13-16``
(data loading is outside the hot path; nothing here is ever called by the oracle)."""
from . import detection_utils, transforms          # noqa: F401


def build_detection_test_loader(*args, **kwargs):
    raise NotImplementedError("detectron2.data is not part of the oracle shim")
