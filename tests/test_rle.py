"""COCO RLE.  CPU: the oracle restatement of maskApi.c against hand-derived known answers and the
round trip; the product's vectorised ``runs_to_string`` against the oracle.  GPU: device run lengths against the oracle."""
import numpy as np
import pytest
import torch

from centermask2_b200 import rle
from oracle import rle as oracle_rle


def test_known_answers_derived_by_hand_from_maskapi():
    # all zeros, 2x3: one run of 6 zeros -> '6'
    assert oracle_rle.rle_encode(np.zeros((2, 3))) == [6] and oracle_rle.rle_to_string([6]) == b"6"
    # single foreground pixel: zero-length run of zeros first -> "01"
    assert oracle_rle.rle_encode(np.ones((1, 1))) == [0, 1] and oracle_rle.rle_to_string([0, 1]) == b"01"
    # column-major order: [[0,1],[1,1]] is read 0,1,1,1
    assert oracle_rle.rle_encode(np.array([[0, 1], [1, 1]])) == [1, 3]
    # 40 = 0b01000 + 32 * 1: low chunk 8 with continuation bit (0x20) -> chr(48+40) = 'X', then '1'
    assert oracle_rle.rle_to_string([40]) == b"X1"
    # difference coding starts at index 3 (against index 1): 4 - 2 = 2; 1 - 1 = 0
    assert oracle_rle.rle_to_string([1, 2, 3, 4]) == b"1232"
    assert oracle_rle.rle_to_string([5, 1, 1, 1]) == b"5110"
    # negative difference 2 - 5 = -3: low five bits 29, sign bit set and x >> 5 == -1 -> single char chr(48+29) = 'M'
    assert oracle_rle.rle_to_string([1, 5, 1, 2]) == b"151M"


@pytest.mark.parametrize("h,w,p", [(7, 5, 0.5), (40, 67, 0.3), (33, 31, 0.02), (16, 16, 0.98), (64, 48, 0.5)])
def test_oracle_round_trip_and_product_string_codec(h, w, p):
    rng = np.random.default_rng(h * w)
    mask = (rng.random((h, w)) < p).astype(np.uint8)
    if p == 0.3:
        mask[5:30, 10:50] = 1                        # a blob: long runs, large positive / negative differences
    cnts = oracle_rle.rle_encode(mask)
    s = oracle_rle.rle_to_string(cnts)
    assert oracle_rle.rle_from_string(s) == cnts
    assert np.array_equal(oracle_rle.rle_decode(cnts, h, w), mask)
    assert rle.runs_to_string(np.array(cnts, dtype=np.uint32)) == s


def test_string_codec_long_runs():
    cnts = [1066400]                                  # an empty 800x1333 mask
    assert rle.runs_to_string(cnts) == oracle_rle.rle_to_string(cnts)
    cnts = [500000, 3, 566397, 7, 1, 90000, 2]
    assert rle.runs_to_string(cnts) == oracle_rle.rle_to_string(cnts)
    assert oracle_rle.rle_from_string(oracle_rle.rle_to_string(cnts)) == cnts


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", [(40, 67), (97, 131), (800, 1333)])
def test_gpu_rle_matches_oracle(h, w):
    g = torch.Generator().manual_seed(h + w)
    r = 6
    masks = torch.zeros((r, h, w), dtype=torch.bool)
    masks[0] = torch.rand(h, w, generator=g) < 0.5                       # noise: ~h*w/2 runs
    masks[1, h // 4:h // 2, w // 5:w // 2] = True                        # a box
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    masks[2] = ((yy - h / 2) ** 2 / (h / 3) ** 2 + (xx - w / 2) ** 2 / (w / 4) ** 2) < 1      # an ellipse
    masks[3] = True                                                       # all foreground: [0, h*w]
    masks[5, 0, 0] = True
    masks[5, h - 1, w - 1] = True                                         # first and last pixel
    got = rle.encode(masks.cuda())
    for i in range(r):
        m = masks[i].numpy().astype(np.uint8)
        cnts = oracle_rle.rle_encode(m) if h * w < 200000 or i != 0 else None
        if cnts is None:                                                  # the scalar oracle is slow on 1M noise pixels: numpy restatement
            flat = m.flatten(order="F")
            idx = np.flatnonzero(np.diff(np.concatenate([[0], flat])))
            cnts = np.diff(np.concatenate([[0], idx, [h * w]])).tolist()
        assert got[i]["size"] == [h, w]
        assert got[i]["counts"] == oracle_rle.rle_to_string(cnts), i
        assert np.array_equal(oracle_rle.rle_decode(oracle_rle.rle_from_string(got[i]["counts"]), h, w), m)


@pytest.mark.gpu
def test_gpu_instances_to_coco_json():
    from centermask2_b200.modeling.compat import Boxes, Instances
    inst = Instances((20, 30))
    inst.pred_boxes = Boxes(torch.tensor([[2.0, 3.0, 12.0, 9.0]]).cuda())
    inst.scores = torch.tensor([0.9]).cuda()
    inst.pred_classes = torch.tensor([17]).cuda()
    inst.mask_scores = torch.tensor([0.5]).cuda()
    m = torch.zeros((1, 20, 30), dtype=torch.bool)
    m[0, 3:9, 2:12] = True
    inst.pred_masks = m.cuda()
    out = rle.instances_to_coco_json(inst, 42)
    assert out[0]["image_id"] == 42 and out[0]["category_id"] == 17 and out[0]["bbox"] == [2.0, 3.0, 10.0, 6.0]
    assert out[0]["mask_score"] == 0.5 and isinstance(out[0]["segmentation"]["counts"], str)
    back = oracle_rle.rle_decode(oracle_rle.rle_from_string(out[0]["segmentation"]["counts"]), 20, 30)
    assert np.array_equal(back, m[0].numpy().astype(np.uint8))


def test_instances_to_coco_json_keypoints():
    """coco_evaluation.py:418-425: flat [x, y, score] * K with x, y shifted by -0.5 (no masks -> no device work)."""
    from centermask2_b200.modeling.compat import Boxes, Instances
    inst = Instances((20, 30))
    inst.pred_boxes = Boxes(torch.tensor([[2.0, 3.0, 12.0, 9.0], [0.0, 0.0, 4.0, 4.0]]))
    inst.scores = torch.tensor([0.9, 0.8])
    inst.pred_classes = torch.tensor([0, 0])
    kp = torch.tensor([[[3.5, 4.5, 0.7], [10.0, 8.25, 0.1]], [[1.5, 2.5, 0.3], [0.5, 0.5, 0.2]]])
    inst.pred_keypoints = kp.clone()
    out = rle.instances_to_coco_json(inst, 7)
    assert out[0]["keypoints"] == pytest.approx([3.0, 4.0, 0.7, 9.5, 7.75, 0.1])
    assert out[1]["keypoints"] == pytest.approx([1.0, 2.0, 0.3, 0.0, 0.0, 0.2])
    assert torch.equal(inst.pred_keypoints, kp)                   # the caller's tensor is not shifted
    assert "segmentation" not in out[0]


def test_batch_result_to_coco_json_and_mask_score_aware_segm_scoring():
    """BatchResult (records + run lengths on the host) -> the json of coco_evaluation.py:362-427, then the segm branch of
    _evaluate_predictions_on_coco (:551-563): bbox dropped, score replaced by mask_score.  CPU only."""
    import numpy as np
    import torch
    from centermask2_b200 import parallel as P
    from centermask2_b200 import rle
    from oracle import rle as orle
    h, w, r_cap = 7, 5, 3
    g = torch.Generator().manual_seed(0)
    rec = torch.zeros((2, r_cap, P.RECORD_FIELDS))
    masks, runs, offs = {}, [], [0]
    counts = [2, 1]
    for i in range(2):
        rec[i, :, P.F_COUNT] = counts[i]
        for j in range(r_cap):
            live = j < counts[i]
            m = (torch.rand(h, w, generator=g) > 0.5).numpy() if live else np.zeros((h, w), dtype=bool)
            masks[(i, j)] = m
            c = orle.rle_encode(m)
            runs += c
            offs.append(offs[-1] + len(c))
            if live:
                rec[i, j, :4] = torch.tensor([1.0 + j, 2.0, 4.0 + j, 6.0])
                rec[i, j, P.F_SCORE], rec[i, j, P.F_CLASS], rec[i, j, P.F_MASK_SCORE], rec[i, j, P.F_VALID] = 0.9 - 0.1 * j, 17 + j, 0.5 + 0.1 * j, 1.0
    rec[0, 1, P.F_VALID] = 0.0                                        # dropped by detector_postprocess (empty after clipping)
    res = P.BatchResult(rec, torch.tensor(offs), torch.tensor(runs, dtype=torch.int32), (h, w))
    js = rle.results_to_coco_json(res, [101, 102])
    assert [(d["image_id"], d["category_id"]) for d in js] == [(101, 17), (102, 17)]
    d = js[0]
    assert d["bbox"] == [1.0, 2.0, 3.0, 4.0] and abs(d["score"] - 0.9) < 1e-6 and abs(d["mask_score"] - 0.5) < 1e-6
    assert d["segmentation"]["size"] == [h, w]
    back = orle.rle_decode(orle.rle_from_string(d["segmentation"]["counts"]), h, w)
    assert np.array_equal(back.astype(bool), masks[(0, 0)]) and np.array_equal(res.mask(0, 0), masks[(0, 0)])
    segm = rle.prepare_segm_results(js)
    assert all("bbox" not in c and "mask_score" not in c for c in segm) and abs(segm[0]["score"] - 0.5) < 1e-6
    assert "bbox" in js[0] and "mask_score" in js[0]                 # the caller's list is untouched (deep copy, :553)
    no_ms = [{k: v for k, v in c.items() if k != "mask_score"} for c in js]
    assert abs(rle.prepare_segm_results(no_ms)[0]["score"] - 0.9) < 1e-6
    assert rle.prepare_segm_results([]) == []
