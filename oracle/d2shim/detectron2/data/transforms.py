"""TEST INFRASTRUCTURE ONLY -- stub (see ``detectron2/data/__init__.py``); the Pillow-exact resize oracle is oracle/resize.py."""


class ResizeShortestEdge(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("detectron2.data is not part of the oracle shim")
