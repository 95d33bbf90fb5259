from .build import BACKBONE_REGISTRY, build_backbone
from .backbone import Backbone
from .fpn import FPN, LastLevelMaxPool


def build_resnet_backbone(*a, **k):
    raise RuntimeError("ResNet backbone is out of scope (SURVEY section 2 row 2)")
