from collections import namedtuple
import torch
import torch.nn.functional as F
from torch import nn
import torchvision

from .batch_norm import FrozenBatchNorm2d, get_norm
from .wrappers import Conv2d, ConvTranspose2d, cat, interpolate
from .mask_ops import paste_masks_in_image


class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
    def __new__(cls, channels=None, height=None, width=None, stride=None):
        return super().__new__(cls, channels, height, width, stride)


class ROIAlign(nn.Module):
    """detectron2.layers.ROIAlign == torchvision.ops.roi_align."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return torchvision.ops.roi_align(
            input, rois.to(dtype=input.dtype), self.output_size,
            self.spatial_scale, self.sampling_ratio, self.aligned,
        )


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2 v0.5 layers.nms.batched_nms: torchvision below 40k boxes, per-class loop above."""
    assert boxes.shape[-1] == 4
    if len(boxes) < 40000:
        return torchvision.ops.batched_nms(boxes.float(), scores, idxs, iou_threshold)
    result_mask = scores.new_zeros(scores.size(), dtype=torch.bool)
    for id in torch.unique(idxs).cpu().tolist():
        mask = (idxs == id).nonzero().view(-1)
        keep = torchvision.ops.nms(boxes[mask], scores[mask], iou_threshold)
        result_mask[mask[keep]] = True
    keep = result_mask.nonzero().view(-1)
    keep = keep[scores[keep].argsort(descending=True)]
    return keep


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise RuntimeError("deformable / rotated ops are out of scope (SURVEY section 2 rows 1, 7, 15)")


DeformConv = ModulatedDeformConv = ROIAlignRotated = _Unavailable
