"""Parameter trees with the reference's ``state_dict`` key names.

The reference's modules are deep trees of tiny ``nn.Module``s (one per conv / norm / relu, some with
"/" in their names, e.g. ``stem.stem_1/conv``).  The B200 plug-ins keep the *names* -- so that a
checkpoint written by the reference loads with ``load_state_dict`` -- but the leaves are plain
parameters/buffers hung on empty container modules; all compute happens in ``libcm2.so`` on packed
copies (``packing.py``).
"""
import torch
from torch import nn

_BUFFER_KINDS = ("bn_weight", "bn_bias", "bn_mean", "bn_var")        # FrozenBatchNorm2d holds buffers [d2]


def _default(kind, shape):
    """Identity-like defaults (real values come from ``load_state_dict``)."""
    if kind in ("bn_weight", "gn_weight", "scale"):
        return torch.ones(shape)
    if kind == "bn_var":
        return torch.ones(shape) - 1e-5                              # FrozenBatchNorm2d init [d2]
    if len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return torch.randn(shape) * (1.0 / fan_in) ** 0.5
    return torch.zeros(shape)


def attach_params(root, spec):
    """Create ``root.<a>.<b>.<leaf>`` for every key ``a.b.leaf`` of ``spec`` ({key: (shape, kind)})."""
    for key, (shape, kind) in spec.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        value = _default(kind, tuple(shape))
        if kind in _BUFFER_KINDS:
            mod.register_buffer(parts[-1], value)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))


class PackedModule(nn.Module):
    """Base of the plug-ins: owns reference-named parameters and a lazily built packed copy."""

    def __init__(self):
        super().__init__()
        self.training = False
        self._packed = None
        self._engine = None
        self._pack_gen = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self):
        """The parameters changed (``load_state_dict``, ``load_checkpoint``, ``.to()``): drop the packed device copies AND
        every CUDA graph that was captured with their addresses baked in -- a replay would read freed / stale weights."""
        self._packed = None
        self._pack_gen += 1
        if self._engine is not None:
            self._engine.drop_graphs(id(self))

    def graph_token(self):
        """Identity of this module's packed weights for CUDA-graph keys: (module, weight generation)."""
        return ("cm2w", id(self), self._pack_gen)

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("centermask2_b200 implements the inference path only (training is out of scope)")
        return super().train(False)
