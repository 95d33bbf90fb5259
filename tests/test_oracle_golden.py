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
