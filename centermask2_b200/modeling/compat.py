"""The slice of detectron2's public surface the plug-ins are written against.

When a real detectron2 is importable its registries and structures are used, so that
``import centermask2_b200`` makes ``build_model(cfg)`` of an unmodified detectron2 pick up the
B200 plug-ins under the reference's names.  detectron2 cannot be installed in this environment
(SURVEY.md section 0), so equivalent minimal classes are provided here; they implement only the
behaviour the hot path and its callers use.
"""
import itertools
from collections import namedtuple

import torch

try:                                                              # pragma: no cover - not installable offline
    from detectron2.layers import ShapeSpec
    from detectron2.modeling import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY,
                                     ROI_HEADS_REGISTRY)
    from detectron2.structures import Boxes, ImageList, Instances
    from detectron2.utils.registry import Registry
    HAVE_DETECTRON2 = True
except ImportError:
    HAVE_DETECTRON2 = False

    class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
        def __new__(cls, channels=None, height=None, width=None, stride=None):
            return super().__new__(cls, channels, height, width, stride)

    class Registry(object):
        def __init__(self, name):
            self._name, self._map = name, {}

        def register(self, obj=None):
            def add(o):
                if o.__name__ in self._map:
                    raise KeyError("'{}' already registered in {}".format(o.__name__, self._name))
                self._map[o.__name__] = o
                return o
            return add if obj is None else add(obj)

        def get(self, name):
            if name not in self._map:
                raise KeyError("no object named '{}' in the {} registry".format(name, self._name))
            return self._map[name]

        def __contains__(self, name):
            return name in self._map

    BACKBONE_REGISTRY = Registry("BACKBONE")
    PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
    ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
    META_ARCH_REGISTRY = Registry("META_ARCH")

    class Boxes(object):
        """[M, 4] xyxy float32."""

        def __init__(self, tensor):
            if not isinstance(tensor, torch.Tensor):
                tensor = torch.as_tensor(tensor, dtype=torch.float32)
            self.tensor = tensor.to(torch.float32).reshape(-1, 4)

        def __len__(self):
            return self.tensor.shape[0]

        def __getitem__(self, item):
            return Boxes(self.tensor[item].reshape(-1, 4))

        def area(self):
            t = self.tensor
            return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

        def scale(self, sx, sy):
            self.tensor[:, 0::2] *= sx
            self.tensor[:, 1::2] *= sy

        def clip(self, size):
            h, w = size
            self.tensor[:, 0::2].clamp_(min=0, max=w)
            self.tensor[:, 1::2].clamp_(min=0, max=h)

        def nonempty(self, threshold=0.0):
            t = self.tensor
            return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

        def clone(self):
            return Boxes(self.tensor.clone())

        def to(self, *a, **k):
            return Boxes(self.tensor.to(*a, **k))

        @property
        def device(self):
            return self.tensor.device

        @staticmethod
        def cat(lst):
            return Boxes(torch.cat([b.tensor for b in lst], 0) if lst else torch.zeros((0, 4)))

    class Instances(object):
        """Per-image container: ``image_size`` plus equal-length named fields."""

        def __init__(self, image_size, **fields):
            object.__setattr__(self, "_image_size", image_size)
            object.__setattr__(self, "_fields", {})
            for k, v in fields.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, k, v):
            if k.startswith("_"):
                object.__setattr__(self, k, v)
            else:
                self.set(k, v)

        def __getattr__(self, k):
            f = object.__getattribute__(self, "_fields")
            if k not in f:
                raise AttributeError("Cannot find field '{}' in the given Instances!".format(k))
            return f[k]

        def set(self, k, v):
            if self._fields:
                assert len(v) == len(self), "field '{}' has length {} != {}".format(k, len(v), len(self))
            self._fields[k] = v

        def has(self, k):
            return k in self._fields

        def get(self, k):
            return self._fields[k]

        def remove(self, k):
            del self._fields[k]

        def get_fields(self):
            return self._fields

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            raise NotImplementedError("Empty Instances does not support __len__!")

        def __getitem__(self, item):
            if isinstance(item, int):
                item = slice(item, item + 1) if item >= 0 else slice(len(self) + item, len(self) + item + 1)
            out = Instances(self._image_size)
            for k, v in self._fields.items():
                out.set(k, v[item])
            return out

        def to(self, *a, **kw):
            out = Instances(self._image_size)
            for k, v in self._fields.items():
                out.set(k, v.to(*a, **kw) if hasattr(v, "to") else v)
            return out

        @staticmethod
        def cat(lst):
            if len(lst) == 1:
                return lst[0]
            out = Instances(lst[0].image_size)
            for k in lst[0]._fields:
                vals = [i.get(k) for i in lst]
                if isinstance(vals[0], torch.Tensor):
                    out.set(k, torch.cat(vals, 0))
                elif isinstance(vals[0], list):
                    out.set(k, list(itertools.chain(*vals)))
                else:
                    out.set(k, type(vals[0]).cat(vals))
            return out

    class ImageList(object):
        """Padded batch tensor + the unpadded (h, w) of every image."""

        def __init__(self, tensor, image_sizes):
            self.tensor, self.image_sizes = tensor, [tuple(s) for s in image_sizes]

        def __len__(self):
            return len(self.image_sizes)

        @property
        def device(self):
            return self.tensor.device

        @staticmethod
        def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
            sizes = [(t.shape[-2], t.shape[-1]) for t in tensors]
            mh, mw = max(s[0] for s in sizes), max(s[1] for s in sizes)
            if size_divisibility > 1:
                d = size_divisibility
                mh, mw = (mh + d - 1) // d * d, (mw + d - 1) // d * d
            out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (mh, mw), pad_value)
            for t, o in zip(tensors, out):
                o[..., :t.shape[-2], :t.shape[-1]].copy_(t)
            return ImageList(out, sizes)


def build_model(cfg):
    """``detectron2.modeling.build_model``: META_ARCH_REGISTRY lookup by ``cfg.MODEL.META_ARCHITECTURE``."""
    return META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
