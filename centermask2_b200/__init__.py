"""centermask2_b200 -- B200-native CenterMask2 inference behind the reference's detectron2 registry surface.

``import centermask2_b200`` registers ``build_fcos_vovnet_fpn_backbone``, ``FCOS``, ``CenterROIHeads`` and
``GeneralizedRCNN`` (see ``modeling/``); all compute runs in ``libcm2.so`` (``csrc/``, ``include/cm2.h``).
"""
from .config import get_cfg, lite_overrides  # noqa: F401
from .runtime import set_precision  # noqa: F401
from . import modeling  # noqa: F401
from .modeling import build_model  # noqa: F401
from .checkpoint import DetectionCheckpointer, load_checkpoint  # noqa: F401

__version__ = "0.1.0"
