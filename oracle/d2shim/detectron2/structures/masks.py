import torch
from ..layers.mask_ops import paste_masks_in_image


class PolygonMasks(object):
    def __init__(self, *a, **k):
        raise RuntimeError("training-only symbol; not available in the oracle shim")


class BitMasks(object):
    def __init__(self, tensor):
        self.tensor = tensor.to(torch.bool)

    def __len__(self):
        return self.tensor.shape[0]


class ROIMasks(object):
    """[N, M, M] soft masks attached to boxes (detectron2 >= 0.5)."""

    def __init__(self, tensor):
        assert tensor.dim() == 3
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]

    def to_bitmasks(self, boxes, height, width, threshold=0.5):
        bitmasks = paste_masks_in_image(self.tensor, boxes, (height, width), threshold=threshold)
        return BitMasks(bitmasks)
