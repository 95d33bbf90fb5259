"""CPU: checkpoint ingestion -- detectron2-style .pth / .pkl files load into the plug-in modules
under the reference's key names (README.md:251-255 checkpoint format; convert_model_into_onnx.py:66 call site)."""
import pickle

import pytest
import torch

import centermask2_b200 as cm
from centermask2_b200 import checkpoint
from centermask2_b200.config import get_cfg
from centermask2_b200.synth import synthetic_state_dict


@pytest.fixture(scope="module")
def case():
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml")
    sd = synthetic_state_dict(cfg, seed=3)
    return cfg, sd


def _same(model, sd):
    own = model.state_dict()
    assert set(own) == set(sd)
    return all(torch.equal(own[k], sd[k]) for k in sd)


def test_pth_with_model_key_module_prefix_and_pixel_buffers(tmp_path, case):
    cfg, sd = case
    model = cm.build_model(cfg)
    assert not _same(model, sd)
    data = {"model": {"module." + k: v for k, v in sd.items()}, "iteration": 269999}
    data["model"]["module.pixel_mean"] = torch.zeros(3, 1, 1)
    data["model"]["module.pixel_std"] = torch.ones(3, 1, 1)
    path = tmp_path / "model_final.pth"
    torch.save(data, path)
    extra = checkpoint.DetectionCheckpointer(model).load(str(path))
    assert extra == {"iteration": 269999}
    assert _same(model, sd)


def test_pkl_numpy_arrays_and_incompatible_keys(tmp_path, case):
    cfg, sd = case
    model = cm.build_model(cfg)
    blob = {"model": {k: v.numpy() for k, v in sd.items()}, "__author__": "test"}
    dropped = "roi_heads.maskiou_head.maskiou.bias"
    blob["model"].pop(dropped)
    blob["model"]["roi_heads.unknown.weight"] = blob["model"]["backbone.fpn_lateral3.bias"]
    blob["model"]["backbone.fpn_lateral3.bias"] = blob["model"]["backbone.fpn_lateral3.bias"][:7]
    path = tmp_path / "zoo.pkl"
    with open(path, "wb") as f:
        pickle.dump(blob, f)
    inc = checkpoint.load_checkpoint(model, str(path))
    assert inc.missing_keys.count(dropped) == 1 and "backbone.fpn_lateral3.bias" in inc.missing_keys
    assert inc.unexpected_keys == ["roi_heads.unknown.weight"]
    assert inc.incorrect_shapes[0][0] == "backbone.fpn_lateral3.bias"
    own = model.state_dict()
    assert torch.equal(own["backbone.bottom_up.stem.stem_1/conv.weight"], sd["backbone.bottom_up.stem.stem_1/conv.weight"])
    with pytest.raises(RuntimeError):
        checkpoint.load_checkpoint(model, str(path), strict=True)


def test_bare_state_dict_invalidates_packed_weights(case):
    cfg, sd = case
    model = cm.build_model(cfg)
    model.backbone._packed = object()
    inc = checkpoint.load_checkpoint(model, dict(sd), strict=True)
    assert not inc.missing_keys and model.backbone._packed is None
    assert _same(model, sd)


def test_unrelated_checkpoint_is_refused_and_mismatches_are_logged(caplog):
    """A file whose names do not match must not leave random-init weights behind silently (detectron2 logs these)."""
    import logging
    import pytest
    import centermask2_b200 as cm
    from centermask2_b200.checkpoint import load_checkpoint
    cfg = cm.get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.VOVNET.CONV_BODY", "V-19-eSE"])
    model = cm.modeling.FCOS(cfg, {k: cm.modeling.ShapeSpec(channels=256, stride=2 ** int(k[1])) for k in cfg.MODEL.FCOS.IN_FEATURES})
    with pytest.raises(RuntimeError, match="no key in common"):
        load_checkpoint(model, {"model": {"some.other.net.weight": torch.zeros(3)}})
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    dropped = sorted(sd)[0]
    sd.pop(dropped)
    sd["extra.key"] = torch.zeros(1)
    with caplog.at_level(logging.WARNING, logger="centermask2_b200.checkpoint"):
        inc = load_checkpoint(model, {"model": sd})
    assert inc.missing_keys == [dropped] and inc.unexpected_keys == ["extra.key"]
    text = caplog.text
    assert "missing from the checkpoint" in text and "not used by the model" in text


def test_unsupported_cfg_values_are_refused_at_build_time():
    import pytest
    import centermask2_b200 as cm
    from centermask2_b200.config import validate_cfg
    base = ["MODEL.VOVNET.CONV_BODY", "V-19-eSE"]
    for part in ("backbone", "fcos", "roi_heads"):
        validate_cfg(cm.get_cfg("centermask_V_39_eSE_FPN.yaml", base), part)
    for part, opts in (("backbone", ["MODEL.FPN.FUSE_TYPE", "avg"]), ("backbone", ["MODEL.FPN.NORM", "GN"]),
                       ("backbone", ["MODEL.VOVNET.NORM", "BN"]), ("fcos", ["MODEL.FCOS.POST_NMS_TOPK_TEST", 300]),
                       ("roi_heads", ["MODEL.ROI_MASK_HEAD.POOLER_TYPE", "ROIAlign"]),
                       ("roi_heads", ["MODEL.ROI_MASK_HEAD.NORM", "GN"])):
        with pytest.raises(NotImplementedError):
            validate_cfg(cm.get_cfg("centermask_V_39_eSE_FPN.yaml", base + opts), part)
