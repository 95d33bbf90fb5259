from torch import nn
from ...layers import ShapeSpec


class Backbone(nn.Module):
    def __init__(self):
        super().__init__()

    @property
    def size_divisibility(self):
        return 0

    def output_shape(self):
        return {
            name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
            for name in self._out_features
        }
