"""Importing this package performs the registrations (as ``centermask/modeling/__init__.py:1-3`` does)."""
from .compat import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY, Boxes,
                     ImageList, Instances, ShapeSpec, build_model)
from .backbone import VoVNetFPN, build_fcos_vovnet_fpn_backbone
from .fcos import FCOS
from .roi_heads import (ROI_MASK_HEAD_REGISTRY, ROI_MASKIOU_HEAD_REGISTRY, CenterROIHeads, MaskIoUHead,
                        SpatialAttentionMaskHead)
from .rcnn import GeneralizedRCNN
