"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the CenterMask2 inference path.

An independent, reference-free restatement of the hot path in plain ``torch`` functional ops
(leaf ops = the installed torch 2.11 / torchvision 0.26 CPU kernels).  It takes a plain
``state_dict`` with the reference's key names plus a cfg and returns every intermediate, so the
CUDA path can be diffed layer by layer.  It is the checker for ``tests/`` (``-m gpu`` parity),
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``;
the product path never imports it.

Parity pin: ``tests/golden/*.pt`` hold outputs of the *unmodified reference* run in the build
container over ``oracle/d2shim`` (``oracle/gen_golden.py``); ``tests/test_oracle_golden.py`` checks
this restatement against them.  The reference itself ships no golden vectors.

Each function cites the reference lines it follows (paths relative to
``/root/reference/centermask2/centermask/``; ``[d2]`` = un-vendored detectron2 v0.5 semantics as
described in SURVEY.md Appendix A).
"""
import math

import torch
import torch.nn.functional as F
import torchvision

from centermask2_b200.arch import vovnet_blocks, vovnet_is_depthwise          # static layer tables only (no compute)

BN_EPS = 1e-5

# ----------------------------------------------------------------------------------------------
# bf16 simulation.  The tensor-core engine keeps activations and weights in bf16 (fp32 accumulation,
# fp32 epilogue, one rounding when a layer's output is stored).  With ``bf16_sim(True)`` this
# restatement rounds at the same places -- conv / linear weights, and every tensor the engine stores as
# bf16 -- so that the CUDA bf16 path can be checked against a reference that differs only by
# accumulation order.  Head outputs (logits, reg, ctr, mask probabilities, MaskIoU) stay fp32 there.
# ----------------------------------------------------------------------------------------------
_BF16 = False


class bf16_sim(object):
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _BF16
        self.prev, _BF16 = _BF16, self.on

    def __exit__(self, *a):
        global _BF16
        _BF16 = self.prev


def _q(x):
    """Round to bf16 (and back to fp32) when simulating the tensor-core engine."""
    return x.to(torch.bfloat16).to(torch.float32) if _BF16 else x


# ----------------------------------------------------------------------------------------------
# backbone
# ----------------------------------------------------------------------------------------------
def _conv_bn_relu(x, sd, prefix, stride, pad):
    """conv (no bias) -> FrozenBN -> ReLU.  modeling/backbone/vovnet.py:205-236; FrozenBN [d2]."""
    x = F.conv2d(x, _q(sd[prefix + "/conv.weight"]), None, stride, pad)
    x = F.batch_norm(x, sd[prefix + "/norm.running_mean"], sd[prefix + "/norm.running_var"],
                     sd[prefix + "/norm.weight"], sd[prefix + "/norm.bias"], False, 0.0, BN_EPS)
    return _q(F.relu(x))


def _dw_pw_bn_relu(x, sd, prefix, stride):
    """depthwise 3x3 (groups = channels, no bias / norm / ReLU) -> pointwise 1x1 -> FrozenBN -> ReLU.
    modeling/backbone/vovnet.py:110-130 ("dw_conv3x3")."""
    x = _q(F.conv2d(x, sd[prefix + "/dw_conv3x3.weight"], None, stride, 1, 1, x.shape[1]))
    x = F.conv2d(x, _q(sd[prefix + "/pw_conv1x1.weight"]))
    x = F.batch_norm(x, sd[prefix + "/pw_norm.running_mean"], sd[prefix + "/pw_norm.running_var"],
                     sd[prefix + "/pw_norm.weight"], sd[prefix + "/pw_norm.bias"], False, 0.0, BN_EPS)
    return _q(F.relu(x))


def _ese(x, sd, prefix):
    """eSE gate: x * relu6(fc(mean_hw(x)) + 3) / 6.  vovnet.py:238-260."""
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.conv2d(g, sd[prefix + ".weight"], sd[prefix + ".bias"])
    g = F.relu6(g + 3.0) / 6.0
    return x * g          # rounded by the caller (after the optional identity add)


def vovnet_forward(x, sd, cfg, prefix="backbone.bottom_up.", trace=None):
    """VoVNet.forward (vovnet.py:471-481) with _OSA_stage (:335-376) and _OSA_module.forward (:310-332)."""
    stem, blocks = vovnet_blocks(cfg.MODEL.VOVNET.CONV_BODY)
    dw = vovnet_is_depthwise(cfg.MODEL.VOVNET.CONV_BODY)            # vovnet.py:408: conv_type of stem_2 / stem_3
    x = _conv_bn_relu(x, sd, prefix + "stem.stem_1", 2, 1)          # vovnet.py:409
    if dw:
        x = _dw_pw_bn_relu(x, sd, prefix + "stem.stem_2", 1)        # :410
        x = _dw_pw_bn_relu(x, sd, prefix + "stem.stem_3", 2)        # :411
    else:
        x = _conv_bn_relu(x, sd, prefix + "stem.stem_2", 1, 1)      # :410
        x = _conv_bn_relu(x, sd, prefix + "stem.stem_3", 2, 1)      # :411
    if trace is not None:
        trace["stem"] = x
    outs = {}
    stage = 2
    for b in blocks:
        if b.stage != stage:
            # MaxPool2d(3, stride 2, ceil_mode=True) in front of stages 3..5 (:349-350)
            x = F.max_pool2d(x, kernel_size=3, stride=2, ceil_mode=True)
            stage = b.stage
        identity = x
        feats = [x]
        y = x
        if b.reduced:                                                  # :315-316 (the concat still takes the input x)
            y = _conv_bn_relu(y, sd, prefix + b.reduction_key(), 1, 0)
        for i in range(b.n_conv):
            y = _dw_pw_bn_relu(y, sd, prefix + b.key(i), 1) if b.dw else _conv_bn_relu(y, sd, prefix + b.key(i), 1, 1)
            feats.append(y)
        y = torch.cat(feats, dim=1)                                    # :324
        y = _conv_bn_relu(y, sd, prefix + b.key("concat"), 1, 0)        # :325
        y = _ese(y, sd, prefix + b.ese_key())                           # :327 (unconditional)
        if b.identity:
            y = y + identity                                           # :329-330
        x = _q(y)
        outs["stage{}".format(b.stage)] = x
        if trace is not None:
            trace[b.name] = x
    return outs


def fpn_forward(feats, sd, cfg, prefix="backbone."):
    """detectron2 FPN.forward [d2] as constructed at vovnet.py:547-554 + LastLevelP6P7 (fpn.py:32-35)."""
    in_features = list(cfg.MODEL.FPN.IN_FEATURES)
    results = {}
    prev = None
    for f in reversed(in_features):
        lvl = int(f[-1])
        lat = F.conv2d(feats[f], _q(sd[prefix + "fpn_lateral{}.weight".format(lvl)]),
                       sd[prefix + "fpn_lateral{}.bias".format(lvl)])
        if prev is not None:
            lat = lat + F.interpolate(prev, scale_factor=2.0, mode="nearest")
        prev = _q(lat)
        results["p{}".format(lvl)] = _q(F.conv2d(prev, _q(sd[prefix + "fpn_output{}.weight".format(lvl)]),
                                                 sd[prefix + "fpn_output{}.bias".format(lvl)], 1, 1))
    top = cfg.MODEL.FCOS.TOP_LEVELS
    if top >= 1:
        p6 = _q(F.conv2d(results["p5"], _q(sd[prefix + "top_block.p6.weight"]), sd[prefix + "top_block.p6.bias"], 2, 1))
        results["p6"] = p6
        if top == 2:
            # P7 consumes relu(P6); P6 itself is returned without the ReLU (fpn.py:33-35)
            results["p7"] = _q(F.conv2d(F.relu(p6), _q(sd[prefix + "top_block.p7.weight"]),
                                        sd[prefix + "top_block.p7.bias"], 2, 1))
    return {k: results[k] for k in sorted(results)}


# ----------------------------------------------------------------------------------------------
# FCOS head + post-process
# ----------------------------------------------------------------------------------------------
def compute_locations(h, w, stride):
    """fcos/fcos.py:132-144: (x, y) = (col*s + s//2, row*s + s//2), row-major."""
    xs = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
    ys = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((xx.reshape(-1), yy.reshape(-1)), dim=1) + stride // 2


def fcos_head_forward(features, sd, cfg, prefix="proposal_generator.fcos_head."):
    """FCOSHead.forward (fcos/fcos.py:222-240); towers :169-186; Scale :19-25."""
    use_gn = cfg.MODEL.FCOS.NORM == "GN"
    per_unit = 3 if use_gn else 2

    def tower(x, name, n):
        for i in range(n):
            p = prefix + "{}_tower.{}".format(name, per_unit * i)
            x = _q(F.conv2d(x, _q(sd[p + ".weight"]), sd[p + ".bias"], 1, 1))   # engine stores the conv output, then GN in place
            if use_gn:
                q = prefix + "{}_tower.{}".format(name, per_unit * i + 1)
                x = F.group_norm(x, 32, sd[q + ".weight"], sd[q + ".bias"], 1e-5)
            x = _q(F.relu(x))
        return x

    logits, regs, ctrs = [], [], []
    for l, f in enumerate(cfg.MODEL.FCOS.IN_FEATURES):
        x = tower(features[f], "share", cfg.MODEL.FCOS.NUM_SHARE_CONVS)
        ct = tower(x, "cls", cfg.MODEL.FCOS.NUM_CLS_CONVS)
        bt = tower(x, "bbox", cfg.MODEL.FCOS.NUM_BOX_CONVS)
        logits.append(F.conv2d(ct, _q(sd[prefix + "cls_logits.weight"]), sd[prefix + "cls_logits.bias"], 1, 1))
        ctrs.append(F.conv2d(bt, _q(sd[prefix + "ctrness.weight"]), sd[prefix + "ctrness.bias"], 1, 1))
        reg = F.conv2d(bt, _q(sd[prefix + "bbox_pred.weight"]), sd[prefix + "bbox_pred.bias"], 1, 1)
        if cfg.MODEL.FCOS.USE_SCALE:
            reg = reg * sd[prefix + "scales.{}.scale".format(l)]
        regs.append(F.relu(reg))
    return logits, regs, ctrs


def nms_per_class(boxes, scores, classes, thresh):
    """Greedy class-aware NMS on original coordinates, descending score order.

    layers/ml_nms.py:93-96 -> detectron2 batched_nms -> torchvision [d2].  IoU arithmetic is
    torchvision's CPU kernel: inter / (area_i + area_j - inter), areas (x2-x1)*(y2-y1), suppress
    when IoU > thresh.  Returns kept indices sorted by descending score (ties: lower index first)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.int64)
    order = torch.sort(scores, descending=True, stable=True).indices
    keep_mask = torch.zeros(n, dtype=torch.bool)
    for c in torch.unique(classes).tolist():
        sel = order[classes[order] == c]
        k = torchvision.ops.nms(boxes[sel], scores[sel], thresh)
        keep_mask[sel[k]] = True
    return order[keep_mask[order]]


def fcos_postprocess(logits, regs, ctrs, image_sizes, cfg, pre_topk=True):
    """FCOSOutputs.predict_proposals (fcos/fcos_outputs.py:372-394), forward_for_single_feature_map
    (:396-466) and select_over_all_levels (:468-495).

    ``pre_topk=True`` applies the upstream per-(image, level) ``PRE_NMS_TOPK`` selection that the fork
    commented out (:444-449); with <= PRE_NMS_TOPK candidates per level both semantics coincide
    (SURVEY.md row A12).  Returns per image a dict of tensors sorted by descending score."""
    strides = cfg.MODEL.FCOS.FPN_STRIDES
    thr = cfg.MODEL.FCOS.INFERENCE_TH_TEST
    pre_n = cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST
    post_n = cfg.MODEL.FCOS.POST_NMS_TOPK_TEST
    nms_th = cfg.MODEL.FCOS.NMS_TH
    n_img = logits[0].shape[0]
    per_image = [[] for _ in range(n_img)]
    for lvl, (o, r, c, s) in enumerate(zip(logits, regs, ctrs, strides)):
        N, C, H, W = o.shape
        loc = compute_locations(H, W, s)
        r = r * s                                                        # :383
        box_cls = o.permute(0, 2, 3, 1).reshape(N, -1, C).sigmoid()      # :404-405
        box_reg = r.permute(0, 2, 3, 1).reshape(N, -1, 4)
        ctr = c.permute(0, 2, 3, 1).reshape(N, -1).sigmoid()            # :408-409
        if cfg.MODEL.FCOS.THRESH_WITH_CTR:
            box_cls = box_cls * ctr[:, :, None]
        cand = box_cls > thr                                             # :415
        if not cfg.MODEL.FCOS.THRESH_WITH_CTR:
            box_cls = box_cls * ctr[:, :, None]                          # :419-420
        for i in range(N):
            nz = cand[i].nonzero()                                       # row-major (loc, class)
            li, ci = nz[:, 0], nz[:, 1]
            sc = box_cls[i][li, ci]
            if pre_topk and sc.numel() > pre_n:                          # upstream :444-449
                sc, top = sc.topk(pre_n, sorted=False)
                li, ci = li[top], ci[top]
            rg = box_reg[i][li]
            lc = loc[li]
            boxes = torch.stack([lc[:, 0] - rg[:, 0], lc[:, 1] - rg[:, 1],
                                 lc[:, 0] + rg[:, 2], lc[:, 1] + rg[:, 3]], dim=1)   # :451-456
            per_image[i].append((boxes, torch.sqrt(sc), ci, lc))         # :460
    results = []
    for i in range(n_img):
        boxes = torch.cat([p[0] for p in per_image[i]])
        scores = torch.cat([p[1] for p in per_image[i]])
        classes = torch.cat([p[2] for p in per_image[i]])
        locs = torch.cat([p[3] for p in per_image[i]])
        keep = nms_per_class(boxes, scores, classes, nms_th)             # :473
        keep = keep[:post_n]                                             # topk of a descending list (:476-482)
        results.append({"pred_boxes": boxes[keep], "scores": scores[keep],
                        "pred_classes": classes[keep], "locations": locs[keep],
                        "image_size": tuple(image_sizes[i]),
                        "num_candidates": int(boxes.shape[0])})
    return results


# ----------------------------------------------------------------------------------------------
# ROI heads
# ----------------------------------------------------------------------------------------------
def assign_levels_by_ratio(boxes, image_area, min_level, max_level):
    """centermask/pooler.py:80-118 (CenterMask Eq. 2); eps = sys.float_info.epsilon."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    img = torch.full_like(area, float(image_area))
    lv = torch.ceil(max_level - torch.log2(img / area + 2.220446049250313e-16))
    lv = torch.clamp(lv, min=min_level, max=max_level)
    return lv.to(torch.int64) - min_level


def assign_levels_by_area(boxes, min_level, max_level, canonical_box_size=224, canonical_level=4):
    """centermask/pooler.py:121-152 (FPN Eq. 1)."""
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    lv = torch.floor(canonical_level + torch.log2(torch.sqrt(area) / canonical_box_size + 2.220446049250313e-16))
    lv = torch.clamp(lv, min=min_level, max=max_level)
    return lv.to(torch.int64) - min_level


def roi_pool(features, dets, cfg, head="mask"):
    """ROIPooler.forward eager branch (centermask/pooler.py:320-366); ROIAlign [d2] = torchvision
    roi_align(output 14, scale 1/stride, sampling_ratio 0, aligned=True).  ``head``: "mask" (center_heads.py:334-356)
    or "keypoint" (center_heads.py:358-379, its own IN_FEATURES / resolution / criterion)."""
    hc = cfg.MODEL.ROI_MASK_HEAD if head == "mask" else cfg.MODEL.ROI_KEYPOINT_HEAD
    names = list(cfg.MODEL.ROI_HEADS.IN_FEATURES if head == "mask" else hc.IN_FEATURES)
    res = hc.POOLER_RESOLUTION
    ratio = hc.POOLER_SAMPLING_RATIO
    strides = [2 ** int(n[-1]) for n in names]
    min_l, max_l = int(math.log2(strides[0])), int(math.log2(strides[-1]))
    rois, lvls = [], []
    for i, d in enumerate(dets):
        b = d["pred_boxes"]
        rois.append(torch.cat([torch.full((b.shape[0], 1), float(i)), b], dim=1))
        if hc.ASSIGN_CRITERION == "ratio":
            lvls.append(assign_levels_by_ratio(b, d["image_size"][0] * d["image_size"][1], min_l, max_l))
        else:
            lvls.append(assign_levels_by_area(b, min_l, max_l))
    rois, lvls = torch.cat(rois), torch.cat(lvls)
    c = features[names[0]].shape[1]
    out = torch.zeros((rois.shape[0], c, res, res), dtype=torch.float32)
    for li, (n, s) in enumerate(zip(names, strides)):
        inds = (lvls == li).nonzero().squeeze(1)
        if len(names) == 1:
            inds = torch.arange(rois.shape[0])
        out[inds] = torchvision.ops.roi_align(features[n], rois[inds], res, 1.0 / s, ratio, True)
    return _q(out), lvls


def mask_head_forward(x, sd, cfg, prefix="roi_heads.mask_head."):
    """SpatialAttentionMaskHead.forward (centermask/sam.py:92-97) with SpatialAttention (:23-28)."""
    for k in range(cfg.MODEL.ROI_MASK_HEAD.NUM_CONV):
        p = prefix + "mask_fcn{}".format(k + 1)
        x = _q(F.relu(F.conv2d(x, _q(sd[p + ".weight"]), sd[p + ".bias"], 1, 1)))
    avg = torch.mean(x, dim=1, keepdim=True)
    mx = torch.max(x, dim=1, keepdim=True)[0] if x.numel() else x.new_empty((x.shape[0], 1) + x.shape[2:])
    att = F.conv2d(torch.cat([avg, mx], dim=1), sd[prefix + "spatialAtt.conv.weight"], None, 1, 1)
    x = _q(x * torch.sigmoid(att))
    feat = x
    x = _q(F.relu(F.conv_transpose2d(x, _q(sd[prefix + "deconv.weight"]), sd[prefix + "deconv.bias"], stride=2)))
    return F.conv2d(x, sd[prefix + "predictor.weight"], sd[prefix + "predictor.bias"]), feat      # fp32 predictor weights


def maskiou_head_forward(roi_feat, mask, sd, cfg, prefix="roi_heads.maskiou_head."):
    """MaskIoUHead.forward (centermask/maskiou_head.py:107-120)."""
    x = torch.cat((roi_feat, _q(F.max_pool2d(mask, kernel_size=2, stride=2))), 1)
    n_conv = cfg.MODEL.ROI_MASKIOU_HEAD.NUM_CONV
    for k in range(n_conv):
        p = prefix + "maskiou_fcn{}".format(k + 1)
        x = _q(F.relu(F.conv2d(x, _q(sd[p + ".weight"]), sd[p + ".bias"], 2 if k + 1 == n_conv else 1, 1)))
    x = torch.flatten(x, 1)
    x = _q(F.relu(F.linear(x, _q(sd[prefix + "maskiou_fc1.weight"]), sd[prefix + "maskiou_fc1.bias"])))
    x = _q(F.relu(F.linear(x, _q(sd[prefix + "maskiou_fc2.weight"]), sd[prefix + "maskiou_fc2.bias"])))
    return F.linear(x, _q(sd[prefix + "maskiou.weight"]), sd[prefix + "maskiou.bias"])


def keypoint_head_forward(x, sd, cfg, prefix="roi_heads.keypoint_head."):
    """KRCNNConvDeconvUpsampleHead.layers (centermask/keypoint_head.py:217-222): conv3x3 + ReLU per CONV_DIMS entry,
    ConvTranspose2d(k 4, s 2, p 1) to K maps, bilinear x2 (align_corners=False) -> [R, K, 4 * res, 4 * res] logits."""
    for k in range(len(cfg.MODEL.ROI_KEYPOINT_HEAD.CONV_DIMS)):
        p = prefix + "conv_fcn{}".format(k + 1)
        x = _q(F.relu(F.conv2d(x, _q(sd[p + ".weight"]), sd[p + ".bias"], 1, 1)))
    x = F.conv_transpose2d(x, _q(sd[prefix + "score_lowres.weight"]), sd[prefix + "score_lowres.bias"], stride=2, padding=1)
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def heatmaps_to_keypoints(maps, boxes):
    """heatmaps_to_keypoints [d2-memory] (detectron2 v0.5 structures/keypoints.py; called at keypoint_head.py:113).
    maps [R, K, S, S] logits, boxes [R, 4] -> [R, K, 4] = (x, y, logit, score).  Per ROI: bicubic resize
    (align_corners=False, A = -0.75, torch's upsample_bicubic2d) to (ceil(h), ceil(w)) with h, w clamped to >= 1;
    first arg-max pixel per map; x = (x_int + 0.5) * w / ceil(w) + x0 (same for y); logit = resized map there;
    score = exp(logit - max) / sum_{S x S}(exp(map - max)) = 1 / sum_{S x S}(exp(map - max))."""
    r, k = maps.shape[:2]
    out = maps.new_zeros((r, k, 4))
    for i in range(r):
        x0, y0, x1, y1 = [boxes[i, j] for j in range(4)]
        w = (x1 - x0).clamp(min=1)
        h = (y1 - y0).clamp(min=1)
        wc, hc = w.ceil(), h.ceil()
        big = F.interpolate(maps[i:i + 1], size=(int(hc), int(wc)), mode="bicubic", align_corners=False)[0]
        flat = big.reshape(k, -1)
        mx, pos = flat.max(dim=1)
        pos = flat.argmax(dim=1)
        xi = pos % int(wc)
        yi = (pos - xi) // int(wc)
        pool = (maps[i] - mx.view(k, 1, 1)).exp().sum(dim=(1, 2))
        out[i, :, 0] = (xi.float() + 0.5) * (w / wc) + x0
        out[i, :, 1] = (yi.float() + 0.5) * (h / hc) + y0
        out[i, :, 2] = mx
        out[i, :, 3] = (mx - mx).exp() / pool
    return out


def keypoints_forward(features, dets, sd, cfg, trace=None):
    """CenterROIHeads._forward_keypoint inference branch (center_heads.py:551-553) + keypoint_rcnn_inference
    (keypoint_head.py:95-120): adds pred_keypoints [R, K, 3] = (x, y, score)."""
    roi_feat, _ = roi_pool(features, dets, cfg, head="keypoint")
    logits = keypoint_head_forward(roi_feat, sd, cfg)
    boxes = torch.cat([d["pred_boxes"] for d in dets])
    res = heatmaps_to_keypoints(logits, boxes)
    if trace is not None:
        trace.update(kp_roi_feat=roi_feat, kp_logits=logits, kp_full=res)
    counts = [d["pred_classes"].shape[0] for d in dets]
    for d, kp in zip(dets, res[:, :, [0, 1, 3]].split(counts, dim=0)):
        d["pred_keypoints"] = kp
    return dets


def roi_heads_forward(features, dets, sd, cfg, trace=None):
    """CenterROIHeads.forward_with_given_boxes (centermask/center_heads.py:413-444): _forward_mask
    (:480-489) -> mask_rcnn_inference (mask_head.py:196-216) -> _forward_maskiou (:511-517) ->
    mask_iou_inference (maskiou_head.py:50-60).  Adds pred_masks [R,1,28,28] and, when the batch has
    at least one detection, mask_scores [R]."""
    roi_feat, lvls = roi_pool(features, dets, cfg)
    logits, feat = mask_head_forward(roi_feat, sd, cfg)
    classes = torch.cat([d["pred_classes"] for d in dets])
    idx = torch.arange(logits.shape[0])
    if logits.shape[1] == 1:
        probs = logits.sigmoid()
    else:
        probs = logits[idx, classes][:, None].sigmoid()
    counts = [d["pred_classes"].shape[0] for d in dets]
    if trace is not None:
        trace.update(roi_feat=roi_feat, roi_levels=lvls, mask_feat=feat, mask_logits=logits)
    for d, p in zip(dets, probs.split(counts, dim=0)):
        d["pred_masks"] = p
    if cfg.MODEL.MASKIOU_ON and probs.shape[0] > 0:
        # NB: the reference feeds the *pooled ROI feature* (mask_features = mask_pooler output,
        # center_heads.py:480,487) to the MaskIoU head, not the mask-head activations.
        iou = maskiou_head_forward(roi_feat, probs, sd, cfg)
        sel = iou[idx, classes]
        if trace is not None:
            trace["maskiou"] = iou
        for d, m in zip(dets, sel.split(counts, dim=0)):
            d["mask_scores"] = d["scores"] * m
    return dets


# ----------------------------------------------------------------------------------------------
# pre / post-processing and the whole model
# ----------------------------------------------------------------------------------------------
def preprocess(batched_inputs, cfg, size_divisibility=32):
    """GeneralizedRCNN.preprocess_image [d2]; mean/std as deploy_utils.py:81-82; right/bottom zero pad."""
    mean = torch.tensor(cfg.MODEL.PIXEL_MEAN, dtype=torch.float32).view(-1, 1, 1)
    std = torch.tensor(cfg.MODEL.PIXEL_STD, dtype=torch.float32).view(-1, 1, 1)
    imgs = [(b["image"].to(torch.float32) - mean) / std for b in batched_inputs]
    sizes = [(im.shape[-2], im.shape[-1]) for im in imgs]
    mh = max(s[0] for s in sizes)
    mw = max(s[1] for s in sizes)
    mh = (mh + size_divisibility - 1) // size_divisibility * size_divisibility
    mw = (mw + size_divisibility - 1) // size_divisibility * size_divisibility
    out = torch.zeros((len(imgs), imgs[0].shape[0], mh, mw), dtype=torch.float32)
    for i, im in enumerate(imgs):
        out[i, :, : im.shape[-2], : im.shape[-1]] = _q(im)
    return out, sizes


def paste_masks(masks, boxes, out_h, out_w, threshold=0.5):
    """paste_masks_in_image [d2]: bilinear grid_sample of each
    28x28 probability map at output-pixel centres mapped into its box, zero padding,
    align_corners=False, then ``>= threshold``.  One mask at a time inside the integer window
    [floor(x0)-1, ceil(x1)+1) clipped to the image, as detectron2 does on CPU."""
    n = masks.shape[0]
    out = torch.zeros((n, out_h, out_w), dtype=torch.bool)
    for i in range(n):
        x0, y0, x1, y1 = boxes[i].tolist()
        xa = int(max(math.floor(x0) - 1, 0))
        ya = int(max(math.floor(y0) - 1, 0))
        xb = int(min(math.ceil(x1) + 1, out_w))
        yb = int(min(math.ceil(y1) + 1, out_h))
        if xb <= xa or yb <= ya:
            continue
        b = boxes[i]
        img_y = torch.arange(ya, yb, dtype=torch.float32) + 0.5
        img_x = torch.arange(xa, xb, dtype=torch.float32) + 0.5
        img_y = (img_y - b[1]) / (b[3] - b[1]) * 2 - 1
        img_x = (img_x - b[0]) / (b[2] - b[0]) * 2 - 1
        gx = img_x[None, :].expand(img_y.numel(), img_x.numel())
        gy = img_y[:, None].expand(img_y.numel(), img_x.numel())
        grid = torch.stack([gx, gy], dim=2)[None]
        smp = F.grid_sample(masks[i][None, None].float(), grid, mode="bilinear",
                            padding_mode="zeros", align_corners=False)
        out[i, ya:yb, xa:xb] = smp[0, 0] >= threshold
    return out


def detector_postprocess(det, out_h, out_w, mask_threshold=0.5):
    """detector_postprocess [d2]: scale, clip, drop empty, paste masks."""
    ih, iw = det["image_size"]
    sx, sy = out_w / iw, out_h / ih
    boxes = det["pred_boxes"].clone()
    boxes[:, 0::2] *= sx
    boxes[:, 1::2] *= sy
    boxes[:, 0::2] = boxes[:, 0::2].clamp(min=0, max=out_w)
    boxes[:, 1::2] = boxes[:, 1::2].clamp(min=0, max=out_h)
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    out = {"image_size": (out_h, out_w), "pred_boxes": boxes[keep]}
    for k in ("scores", "pred_classes", "locations", "mask_scores"):
        if k in det:
            out[k] = det[k][keep]
    if "pred_keypoints" in det:                          # detector_postprocess [d2]: x *= scale_x, y *= scale_y
        kp = det["pred_keypoints"][keep].clone()
        kp[:, :, 0] *= sx
        kp[:, :, 1] *= sy
        out["pred_keypoints"] = kp
    if "pred_masks" in det:
        out["pred_masks"] = paste_masks(det["pred_masks"][keep][:, 0], out["pred_boxes"], out_h, out_w, mask_threshold)
    return out


def run_model(batched_inputs, sd, cfg, postprocess=True, pre_topk=True, trace=None):
    """GeneralizedRCNN.inference [d2] (call sequence mirrored in-tree at /root/reference/tester.py:24-75)."""
    with torch.no_grad():
        x, sizes = preprocess(batched_inputs, cfg)
        x = _q(x)          # the tensor-core engine stores the normalised input as bf16 (fused normalise + im2col)
        stages = vovnet_forward(x, sd, cfg, trace=trace)
        feats = fpn_forward(stages, sd, cfg)
        logits, regs, ctrs = fcos_head_forward(feats, sd, cfg)
        if trace is not None:
            trace.update(image=x, features=feats, logits=logits, regs=regs, ctrs=ctrs)
        dets = fcos_postprocess(logits, regs, ctrs, sizes, cfg, pre_topk=pre_topk)
        if cfg.MODEL.MASK_ON:
            dets = roi_heads_forward(feats, dets, sd, cfg, trace=trace)
        if cfg.MODEL.KEYPOINT_ON:                       # center_heads.py:441-442: masks first, then keypoints
            dets = keypoints_forward(feats, dets, sd, cfg, trace=trace)
        if not postprocess:
            return dets
        return [detector_postprocess(d, b.get("height", s[0]), b.get("width", s[1]))
                for d, b, s in zip(dets, batched_inputs, sizes)]
