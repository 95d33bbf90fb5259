"""TEST INFRASTRUCTURE ONLY -- ``detectron2.config`` stand-in: the yacs-like node of ``centermask2_b200.config`` (same
attribute / merge / clone behaviour) and detectron2 v0.5's defaults for the keys the path reads.  Lets the reference's
own ``centermask/config`` (defaults.py:1-86) and ``deploy_utils.py`` / ``modified_class.py`` import unchanged."""
from centermask2_b200.config import CfgNode, _d2_defaults


def get_cfg():
    return CfgNode(_d2_defaults())
