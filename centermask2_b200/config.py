"""Configuration surface of the drop-in.

The reference reads a yacs ``CfgNode`` (detectron2 defaults + ``centermask/config/defaults.py:9-86``
+ ``configs/centermask/*.yaml``).  yacs / detectron2 are not installable offline, so this module
provides a small attribute-dict with the same keys and the same ``merge_from_file`` /
``merge_from_list`` / ``_BASE_`` behaviour.  Every plug-in in ``centermask2_b200.modeling`` only
does attribute reads (``cfg.MODEL.FCOS.NMS_TH`` ...), so a real yacs node works equally well.
"""
import ast
import copy
import os

import yaml

__all__ = ["CfgNode", "get_cfg", "lite_overrides", "validate_cfg"]


class CfgNode(dict):
    """dict with attribute access; nested dicts become nested nodes."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
        object.__setattr__(self, "_frozen", False)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError("attempt to modify frozen CfgNode ({}={})".format(name, value))
        self[name] = value

    def freeze(self):
        object.__setattr__(self, "_frozen", True)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        object.__setattr__(self, "_frozen", False)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def clone(self):
        out = copy.deepcopy(self)
        return out

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    # -- merging ---------------------------------------------------------------------------
    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    dict.__setitem__(self, k, CfgNode())
                self[k]._merge(v)
            else:
                if isinstance(v, str):
                    v = _maybe_literal(v)
                if k in self and isinstance(self[k], tuple) and isinstance(v, list):
                    v = tuple(v)
                dict.__setitem__(self, k, v)

    def merge_from_file(self, path):
        self._merge(_load_yaml_with_base(path))

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0, "opts must be KEY VALUE pairs"
        for full_key, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    dict.__setitem__(node, p, CfgNode())
                node = node[p]
            if isinstance(v, str):
                v = _maybe_literal(v)
            dict.__setitem__(node, parts[-1], v)


def _maybe_literal(s):
    try:
        return ast.literal_eval(s)
    except (ValueError, SyntaxError):
        return s


def _load_yaml_with_base(path):
    with open(path, "r") as f:
        data = yaml.safe_load(f) or {}
    base = data.pop("_BASE_", None)
    if base is None:
        return data
    if not os.path.isabs(base):
        base = os.path.join(os.path.dirname(path), base)
    merged = CfgNode(_load_yaml_with_base(base))
    merged._merge(data)
    return merged


def _d2_defaults():
    """The subset of detectron2 v0.5 ``config/defaults.py`` that the hot path reads."""
    return {
        "VERSION": 2,
        "MODEL": {
            "META_ARCHITECTURE": "GeneralizedRCNN",
            "DEVICE": "cuda",
            "WEIGHTS": "",
            "MASK_ON": False,
            "KEYPOINT_ON": False,
            "LOAD_PROPOSALS": False,
            "PIXEL_MEAN": [103.530, 116.280, 123.675],
            "PIXEL_STD": [1.0, 1.0, 1.0],
            "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
            "FPN": {"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
            "PROPOSAL_GENERATOR": {"NAME": "RPN", "MIN_SIZE": 0},
            "ROI_HEADS": {
                "NAME": "Res5ROIHeads",
                "NUM_CLASSES": 80,
                "IN_FEATURES": ["res4"],
                "IOU_THRESHOLDS": [0.5],
                "IOU_LABELS": [0, 1],
                "BATCH_SIZE_PER_IMAGE": 512,
                "POSITIVE_FRACTION": 0.25,
                "SCORE_THRESH_TEST": 0.05,
                "NMS_THRESH_TEST": 0.5,
                "PROPOSAL_APPEND_GT": True,
            },
            "ROI_MASK_HEAD": {
                "NAME": "MaskRCNNConvUpsampleHead",
                "POOLER_RESOLUTION": 14,
                "POOLER_SAMPLING_RATIO": 0,
                "NUM_CONV": 0,
                "CONV_DIM": 256,
                "NORM": "",
                "CLS_AGNOSTIC_MASK": False,
                "POOLER_TYPE": "ROIAlignV2",
            },
            "ROI_KEYPOINT_HEAD": {
                "NAME": "KRCNNConvDeconvUpsampleHead",
                "POOLER_RESOLUTION": 14,
                "POOLER_SAMPLING_RATIO": 0,
                "POOLER_TYPE": "ROIAlignV2",
                "CONV_DIMS": (512,) * 8,
                "NUM_KEYPOINTS": 17,
                "MIN_KEYPOINTS_PER_IMAGE": 1,
                "NORMALIZE_LOSS_BY_VISIBLE_KEYPOINTS": True,
                "LOSS_WEIGHT": 1.0,
            },
        },
        "INPUT": {"MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333, "FORMAT": "BGR",
                  "MIN_SIZE_TRAIN": (800,)},
        "DATASETS": {"TRAIN": (), "TEST": ()},
        "DATALOADER": {"NUM_WORKERS": 4},
        "SOLVER": {"IMS_PER_BATCH": 16, "BASE_LR": 0.001, "STEPS": (30000,), "MAX_ITER": 40000,
                   "CHECKPOINT_PERIOD": 5000},
        "TEST": {"DETECTIONS_PER_IMAGE": 100},
        "OUTPUT_DIR": "./output",
    }


def _centermask_defaults():
    """Keys added by the reference at ``centermask/config/defaults.py:9-86`` (same names, same values)."""
    return {
        "MODEL": {
            "MOBILENET": False,
            "FCOS": {
                "NUM_CLASSES": 80,
                "IN_FEATURES": ["p3", "p4", "p5", "p6", "p7"],
                "FPN_STRIDES": [8, 16, 32, 64, 128],
                "PRIOR_PROB": 0.01,
                "INFERENCE_TH_TRAIN": 0.05,
                "INFERENCE_TH_TEST": 0.05,
                "NMS_TH": 0.6,
                "PRE_NMS_TOPK_TRAIN": 1000,
                "PRE_NMS_TOPK_TEST": 1000,
                "POST_NMS_TOPK_TRAIN": 100,
                "POST_NMS_TOPK_TEST": 100,
                "TOP_LEVELS": 2,
                "NORM": "GN",
                "USE_SCALE": True,
                "THRESH_WITH_CTR": False,
                "LOSS_ALPHA": 0.25,
                "LOSS_GAMMA": 2.0,
                "SIZES_OF_INTEREST": [64, 128, 256, 512],
                "USE_RELU": True,
                "USE_DEFORMABLE": False,
                "NUM_CLS_CONVS": 4,
                "NUM_BOX_CONVS": 4,
                "NUM_SHARE_CONVS": 0,
                "CENTER_SAMPLE": True,
                "POS_RADIUS": 1.5,
                "LOC_LOSS_TYPE": "giou",
            },
            "VOVNET": {
                "CONV_BODY": "V-39-eSE",
                "OUT_FEATURES": ["stage2", "stage3", "stage4", "stage5"],
                "NORM": "FrozenBN",
                "OUT_CHANNELS": 256,
                "BACKBONE_OUT_CHANNELS": 256,
                "STAGE_WITH_DCN": (False, False, False, False),
                "WITH_MODULATED_DCN": False,
                "DEFORMABLE_GROUPS": 1,
            },
            "ROI_MASK_HEAD": {"ASSIGN_CRITERION": "area"},
            "MASKIOU_ON": False,
            "MASKIOU_LOSS_WEIGHT": 1.0,
            "ROI_MASKIOU_HEAD": {"NAME": "MaskIoUHead", "CONV_DIM": 256, "NUM_CONV": 4},
            "ROI_KEYPOINT_HEAD": {"IN_FEATURES": ["p2", "p3", "p4", "p5"], "ASSIGN_CRITERION": "ratio"},
        }
    }


_CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def get_cfg(config_file=None, opts=None):
    """Defaults (detectron2 + centermask), optionally merged with a yaml file and KEY VALUE opts.

    ``config_file`` may be a path or the bare name of a yaml shipped in ``centermask2_b200/configs``.
    Mirrors ``centermask.config.get_cfg`` + ``deploy_utils.setup_cfg`` (deploy_utils.py:46-57).
    """
    cfg = CfgNode(_d2_defaults())
    cfg._merge(_centermask_defaults())
    if config_file is not None:
        if not os.path.exists(config_file):
            config_file = os.path.join(_CONFIG_DIR, config_file)
        cfg.merge_from_file(config_file)
    if opts:
        cfg.merge_from_list(list(opts))
    return cfg


def lite_overrides():
    """CenterMask-Lite recipe (upstream's Lite yaml; absent from the fork, SURVEY 8d cfg 2)."""
    return [
        "MODEL.VOVNET.CONV_BODY", "V-19-eSE",
        "MODEL.FPN.OUT_CHANNELS", 128,
        "MODEL.FCOS.NUM_CLS_CONVS", 2,
        "MODEL.FCOS.NUM_BOX_CONVS", 2,
        "MODEL.FCOS.POST_NMS_TOPK_TEST", 50,
        "MODEL.ROI_MASK_HEAD.CONV_DIM", 128,
        "MODEL.ROI_MASK_HEAD.NUM_CONV", 2,
        "MODEL.ROI_MASKIOU_HEAD.CONV_DIM", 128,
        "MODEL.ROI_MASKIOU_HEAD.NUM_CONV", 2,
        "INPUT.MIN_SIZE_TEST", 512,
        "INPUT.MAX_SIZE_TEST", 853,
    ]


def validate_cfg(cfg, part):
    """Refuse, at build time, cfg values the reference honours but this path does not implement -- a silently ignored
    option would give plausible but different results.  ``part``: "backbone", "fcos" or "roi_heads"."""
    def need(cond, msg):
        if not cond:
            raise NotImplementedError("centermask2_b200: " + msg)

    m = cfg.MODEL
    if part == "backbone":
        need(m.FPN.FUSE_TYPE == "sum", "MODEL.FPN.FUSE_TYPE={!r} (only 'sum', the reference's setting at vovnet.py:553)".format(m.FPN.FUSE_TYPE))
        need(m.FPN.NORM in ("", None), "MODEL.FPN.NORM={!r} (only '' -- FPN convs carry a bias, vovnet.py:551)".format(m.FPN.NORM))
        need(m.VOVNET.NORM == "FrozenBN", "MODEL.VOVNET.NORM={!r} (inference folds FrozenBN; other norms are out of scope)".format(m.VOVNET.NORM))
        need(0 <= m.FCOS.TOP_LEVELS <= 2, "MODEL.FCOS.TOP_LEVELS={} (0, 1 or 2: vovnet.py:541-546)".format(m.FCOS.TOP_LEVELS))
    elif part == "fcos":
        need(0 < m.FCOS.POST_NMS_TOPK_TEST <= 256, "MODEL.FCOS.POST_NMS_TOPK_TEST={} (the NMS kernel keeps at most 256 detections "
             "per image)".format(m.FCOS.POST_NMS_TOPK_TEST))
        need(0 < m.FCOS.PRE_NMS_TOPK_TEST and len(m.FCOS.IN_FEATURES) * m.FCOS.PRE_NMS_TOPK_TEST <= 16384,
             "MODEL.FCOS.PRE_NMS_TOPK_TEST={} x {} levels exceeds the in-CTA merge (16384 candidates per image)".format(
                 m.FCOS.PRE_NMS_TOPK_TEST, len(m.FCOS.IN_FEATURES)))
        need(m.FCOS.NORM in ("GN", "", None, "none"), "MODEL.FCOS.NORM={!r} (GN or none)".format(m.FCOS.NORM))
    elif part == "roi_heads":
        heads = [("ROI_MASK_HEAD", m.MASK_ON)] + ([("ROI_KEYPOINT_HEAD", True)] if m.KEYPOINT_ON else [])
        for name, on in heads:
            if not on:
                continue
            h = m[name]
            need(h.POOLER_TYPE == "ROIAlignV2", "MODEL.{}.POOLER_TYPE={!r} (only ROIAlignV2, i.e. aligned=True: pooler.py:249-255)".format(
                name, h.POOLER_TYPE))
            need(h.ASSIGN_CRITERION in ("ratio", "area"), "MODEL.{}.ASSIGN_CRITERION={!r}".format(name, h.ASSIGN_CRITERION))
        if m.MASK_ON:
            need(m.ROI_MASK_HEAD.NORM in ("", None), "MODEL.ROI_MASK_HEAD.NORM={!r} (only '': sam.py:59-68 with bias)".format(m.ROI_MASK_HEAD.NORM))
    else:
        raise KeyError(part)
