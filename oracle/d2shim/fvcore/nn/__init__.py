from . import weight_init


def sigmoid_focal_loss_jit(*a, **k):
    raise RuntimeError("training-only symbol; not available in the oracle shim")
