from .boxes import Boxes
from .instances import Instances
from .image_list import ImageList
from .masks import PolygonMasks, ROIMasks, BitMasks
from .keypoints import heatmaps_to_keypoints


def pairwise_iou(*a, **k):
    raise RuntimeError("training-only symbol; not available in the oracle shim")


