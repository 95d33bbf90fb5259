"""TEST INFRASTRUCTURE ONLY -- ``detectron2.config.defaults._C`` (the node centermask/config/defaults.py extends)."""
from centermask2_b200.config import CfgNode, _d2_defaults

_C = CfgNode(_d2_defaults())
