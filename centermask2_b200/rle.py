"""COCO RLE encoding of pasted masks with the run lengths computed on the device.

Mirrors what ``instances_to_coco_json`` does with pycocotools
(``/root/reference/centermask2/centermask/evaluation/coco_evaluation.py:362-427``):
``mask_util.encode(np.array(mask[:, :, None], order="F", dtype="uint8"))[0]`` per mask, then
``rle["counts"].decode("utf-8")``.  pycocotools' ``rleEncode`` (column-major run lengths, first run counts zeros) runs in
``libcm2.so`` (``csrc/rle.cu``); only the runs are copied to the host, where ``rleToString`` (maskApi.c: difference
coding against the run two places back, 5 bits per character + continuation bit, offset 48) is applied with numpy.
"""
import numpy as np
import torch

from . import lib


def runs_to_string(cnts):
    """maskApi.c ``rleToString``: uint32 run lengths -> ASCII ``bytes``."""
    x = np.asarray(cnts, dtype=np.int64).copy()
    if x.size > 3:
        x[3:] -= np.asarray(cnts, dtype=np.int64)[1:-2]
    chars, valid = [], []
    alive = np.ones(x.shape, dtype=bool)
    while alive.any():
        c = x & 0x1f
        x = x >> 5                                    # arithmetic shift, as C's >> on a signed long
        more = np.where((c & 0x10) != 0, x != -1, x != 0)
        c = np.where(more, c | 0x20, c) + 48
        chars.append(c.astype(np.uint8))
        valid.append(alive.copy())
        alive &= more
    ch = np.stack(chars, axis=1)
    va = np.stack(valid, axis=1)
    return ch[va].tobytes()


def encode_runs(masks):
    """masks: bool / uint8 CUDA tensor [R, H, W] -> list of R uint32 numpy arrays of run lengths (pycocotools ``rleEncode``)."""
    if masks.dim() != 3:
        raise ValueError("masks must be [R, H, W]")
    if not masks.is_cuda:
        raise RuntimeError("centermask2_b200.rle needs masks on a CUDA device (there is no CPU fallback)")
    r, h, w = masks.shape
    if r == 0:
        return []
    m8 = masks.contiguous().view(torch.uint8) if masks.dtype == torch.bool else masks.contiguous()
    dev = m8.device
    col_count = torch.empty((r, w), dtype=torch.int32, device=dev)
    col_offset = torch.empty((r, w), dtype=torch.int32, device=dev)
    total = torch.empty((r,), dtype=torch.int32, device=dev)
    lib.rle_count(m8, col_count, col_offset, total)
    total_h = total.cpu()                                          # one small D2H copy sizes the output
    nruns = total_h.to(torch.int64) + 1
    offs = torch.zeros((r + 1,), dtype=torch.int64)
    offs[1:] = torch.cumsum(nruns, 0)
    n_all = int(offs[-1])
    mask_offset = offs[:-1].to(dev)
    positions = torch.empty((n_all,), dtype=torch.int32, device=dev)
    runs = torch.empty((n_all,), dtype=torch.int32, device=dev)
    lib.rle_write(m8, col_offset, total, mask_offset, positions, runs)
    flat = runs.cpu().numpy().view(np.uint32)
    o = offs.numpy()
    return [flat[o[i]:o[i + 1]] for i in range(r)]


def encode(masks):
    """pycocotools ``mask.encode`` for a stack of masks: -> list of ``{"size": [h, w], "counts": bytes}``."""
    h, w = int(masks.shape[1]), int(masks.shape[2])
    return [{"size": [h, w], "counts": runs_to_string(c)} for c in encode_runs(masks)]


def instances_to_coco_json(instances, img_id):
    """``coco_evaluation.py:362-427`` for the fields this path produces (boxes XYXY -> XYWH, scores, classes, RLE
    segmentation with utf-8 ``counts``, ``mask_score``, and -- with the keypoint branch -- ``keypoints`` as the flat
    [x, y, score] * K list with x, y shifted by -0.5 to COCO's pixel-index convention, ``:418-425``; the reference
    shifts ``instances.pred_keypoints`` in place, this function leaves the caller's tensor untouched)."""
    n = len(instances)
    if n == 0:
        return []
    boxes = instances.pred_boxes.tensor.detach().float().cpu().clone()
    boxes[:, 2] -= boxes[:, 0]
    boxes[:, 3] -= boxes[:, 1]
    boxes = boxes.tolist()
    scores = instances.scores.tolist()
    classes = instances.pred_classes.tolist()
    has_mask = instances.has("pred_masks")
    has_ms = instances.has("mask_scores")
    if has_mask:
        rles = encode(instances.pred_masks)
        for rle in rles:
            rle["counts"] = rle["counts"].decode("utf-8")
        if has_ms:
            mask_scores = instances.mask_scores.tolist()
    has_kp = instances.has("pred_keypoints")
    if has_kp:
        kp = instances.pred_keypoints.detach().float().cpu().clone()
        kp[:, :, :2] -= 0.5
        kp = kp.reshape(n, -1).tolist()
    out = []
    for k in range(n):
        res = {"image_id": img_id, "category_id": classes[k], "bbox": boxes[k], "score": scores[k]}
        if has_mask:
            res["segmentation"] = rles[k]
            if has_ms:
                res["mask_score"] = mask_scores[k]
        if has_kp:
            res["keypoints"] = kp[k]
        out.append(res)
    return out


def results_to_coco_json(result, image_ids):
    """``coco_evaluation.py:362-427`` for a ``parallel.BatchResult`` (records + run lengths already on the host, as
    ``GeneralizedRCNN.inference_records`` yields them): one dict per valid detection with ``bbox`` XYWH, ``score``,
    ``category_id``, ``mask_score`` and the ``segmentation`` RLE with its ``counts`` compressed by ``rleToString`` and
    decoded to utf-8 -- the same json ``instances_to_coco_json`` builds from ``Instances``, without the masks ever
    crossing PCIe."""
    from . import parallel as P
    rec = result.records
    h, w = result.size
    out = []
    for i, img_id in enumerate(image_ids):
        k = int(rec[i, 0, P.F_COUNT])
        for j in range(k):
            r = rec[i, j]
            if r[P.F_VALID] != 1:
                continue
            x0, y0, x1, y1 = [float(v) for v in r[P.F_BOX:P.F_BOX + 4]]
            out.append({"image_id": img_id, "category_id": int(r[P.F_CLASS]), "bbox": [x0, y0, x1 - x0, y1 - y0],
                        "score": float(r[P.F_SCORE]),
                        "segmentation": {"size": [h, w], "counts": runs_to_string(result.runs(i, j)).decode("utf-8")},
                        "mask_score": float(r[P.F_MASK_SCORE])})
    return out


def prepare_segm_results(coco_results):
    """The mask-score-aware half of ``_evaluate_predictions_on_coco`` for ``iou_type == "segm"``
    (``coco_evaluation.py:551-563``): on a deep copy, drop ``bbox`` (so that COCOeval takes the instance area from the mask)
    and, when the results carry a ``mask_score`` (MaskIoU head: score * predicted mask IoU, maskiou_head.py:50-60), rank the
    masks by it instead of by the box score.  What follows in the reference (``coco_gt.loadRes`` + ``COCOeval``) needs
    pycocotools and the COCO annotations and is out of scope."""
    import copy
    results = copy.deepcopy(coco_results)
    if not results:
        return results
    has_mask_scores = "mask_score" in results[0]
    for c in results:
        c.pop("bbox", None)
        if has_mask_scores:
            c["score"] = c["mask_score"]
            del c["mask_score"]
    return results
