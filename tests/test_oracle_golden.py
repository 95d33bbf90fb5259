"""CPU: the fp32 restatement (oracle/restate.py) against the golden vectors produced by the
unmodified reference (oracle/gen_golden.py).  This is the pin of the oracle."""
import pytest
import torch

from oracle import restate
from oracle.cases import CASES
from tests.helpers import (load_golden, build_case, weights_checksum, unpack_masks, mask_iou,
                           assert_detections_match, assert_keypoints_match, MASK_IOU_MIN)


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    gold = load_golden(request.param)
    cfg, sd, inputs = build_case(request.param, gold)
    trace = {}
    # the reference (fork) has the per-level pre-NMS top-k disabled (fcos_outputs.py:444-449)
    raw = restate.run_model(inputs, sd, cfg, postprocess=False, pre_topk=False, trace=trace)
    return request.param, gold, cfg, sd, inputs, raw, trace


def test_weights_reproduce(case):
    name, gold, cfg, sd, *_ = case
    assert abs(weights_checksum(sd) - gold["weights_checksum"]) <= 1e-6 * gold["weights_checksum"]
    assert set(sd.keys()) == set(gold["keys"])


def test_head_tensors(case):
    name, gold, cfg, sd, inputs, raw, trace = case
    if "features" not in gold:
        pytest.skip("lean golden (variant case): end results only")
    for k, v in gold["features"].items():
        assert torch.allclose(trace["features"][k], v, rtol=1e-4, atol=1e-4), k
    for a, b in zip(trace["logits"], gold["logits"]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    for a, b in zip(trace["regs"], gold["regs"]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)
    for a, b in zip(trace["ctrs"], gold["ctrs"]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-4)


def test_raw_detections(case):
    name, gold, cfg, sd, inputs, raw, trace = case
    for i, (g, r) in enumerate(zip(raw, gold["raw"])):
        assert_detections_match(g, r, what="{}[{}]".format(name, i))
        if "pred_keypoints" in r:                      # keypoint_head.py:95-120 (x, y, score) per keypoint
            assert_keypoints_match(g["pred_keypoints"], r["pred_keypoints"], what="{}[{}]".format(name, i), score_rtol=1e-4)
        if len(r["scores"]):
            assert torch.allclose(g["pred_masks"], r["pred_masks"], atol=1e-4)
            assert "mask_scores" in g
        else:
            # center_heads.py:511-513: MaskIoU is skipped on an empty batch -> no mask_scores field
            assert "mask_scores" not in r and "mask_scores" not in g


def test_postprocessed(case):
    name, gold, cfg, sd, inputs, raw, trace = case
    post = [restate.detector_postprocess(d, b["height"], b["width"]) for d, b in zip(raw, inputs)]
    for i, (g, r) in enumerate(zip(post, gold["post"])):
        assert_detections_match(g, r, what="{}[{}] post".format(name, i))
        if "pred_keypoints" in r:
            assert_keypoints_match(g["pred_keypoints"], r["pred_keypoints"], what="{}[{}] post".format(name, i), score_rtol=1e-4)
        if len(r["scores"]):
            ref_masks = unpack_masks(r)
            assert g["pred_masks"].shape == ref_masks.shape
            assert mask_iou(g["pred_masks"], ref_masks).min().item() >= MASK_IOU_MIN


def test_pre_topk_semantics_only_differ_when_crowded(case):
    name, gold, cfg, sd, inputs, raw, trace = case
    capped = restate.run_model(inputs, sd, cfg, postprocess=False, pre_topk=True)
    crowded = max(gold["candidates_per_level"]) > cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST * len(inputs)
    if not crowded and max(gold["candidates_per_level"]) <= cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST:
        for a, b in zip(capped, raw):
            assert torch.equal(a["pred_boxes"], b["pred_boxes"])


def test_tensor_in_tuple_out_variant_of_the_fork():
    """modified_class.GeneralizedRCNN.forward (modified_class.py:27-40) restated with the oracle's pieces: normalised +
    padded tensor in, (locations, mask_scores, pred_boxes, pred_classes, pred_masks, scores) out, image_sizes fixed at
    1344 x 1344 by FakeImageList (modified_class.py:11-24) -- against the tuple the reference's own class produced."""
    name = "v19_tensor_in"
    gold = load_golden(name)
    cfg, sd, inputs = build_case(name, gold)
    x = gold["tensor_in"]["input"]
    with torch.no_grad():
        feats = restate.fpn_forward(restate.vovnet_forward(x, sd, cfg), sd, cfg)
        logits, regs, ctrs = restate.fcos_head_forward(feats, sd, cfg)
        dets = restate.fcos_postprocess(logits, regs, ctrs, [(1344, 1344)] * x.shape[0], cfg, pre_topk=False)
        dets = restate.roi_heads_forward(feats, dets, sd, cfg)
    d = dets[0]
    want = dict(zip(gold["tensor_in"]["names"], gold["tensor_in"]["outputs"]))
    assert_detections_match(d, want, what=name)
    assert torch.allclose(d["pred_masks"], want["pred_masks"], atol=1e-4)


def test_config_defaults_equal_the_reference_defaults():
    """centermask/config/defaults.py:9-86 imported unchanged (needs /root/reference: build container only)."""
    from oracle import refrun
    if not refrun.available():
        pytest.skip("reference tree not present")
    refrun._import_reference()
    from centermask.config import get_cfg as ref_get_cfg
    from centermask2_b200.config import get_cfg

    def flat(node, prefix=""):
        out = {}
        for k, v in node.items():
            if isinstance(v, dict):
                out.update(flat(v, prefix + k + "."))
            else:
                out[prefix + k] = list(v) if isinstance(v, (list, tuple)) else v
        return out
    assert flat(ref_get_cfg()) == flat(get_cfg())
