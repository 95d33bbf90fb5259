"""TEST INFRASTRUCTURE ONLY -- stub (see ``detectron2/data/__init__.py``)."""


def read_image(*args, **kwargs):
    raise NotImplementedError("detectron2.data is not part of the oracle shim")
