"""Static description of the CenterMask2 networks: layer tables and ``state_dict`` layouts.

Everything here is plain Python (no tensors): the layer tables drive both the parameter trees
of the plug-in modules (``modeling/``) and the kernel launch plans (``engine.py``).  Key names are
the reference's ``state_dict`` keys (e.g. ``stem.stem_1/conv.weight``) so that a
checkpoint written by the reference loads unchanged.

Reference: ``centermask/modeling/backbone/vovnet.py:60-108`` (stage tables), ``:205-236``
(conv-norm-relu units), ``:263-376`` (OSA module/stage), ``fpn.py:17-35``, ``fcos/fcos.py:147-220``,
``centermask/sam.py:31-90``, ``centermask/maskiou_head.py:63-105``.
"""
from collections import OrderedDict

# (stem widths, 3x3 width per stage, stage output width, 3x3 convs per OSA block, OSA blocks per stage, depthwise)
# vovnet.py:30-108.  Depthwise bodies (vovnet.py:110-130, :284-293): every 3x3 unit except stem_1 is a depthwise 3x3
# (no norm / ReLU) followed by a pointwise 1x1 -> FrozenBN -> ReLU, and a block whose input width differs from its
# 3x3 width starts with a 1x1 "reduction" unit.
VOVNET_BODIES = {
    "V-19-slim-dw-eSE": ((64, 64, 64), (64, 80, 96, 112), (112, 256, 384, 512), 3, (1, 1, 1, 1), True),
    "V-19-dw-eSE": ((64, 64, 64), (128, 160, 192, 224), (256, 512, 768, 1024), 3, (1, 1, 1, 1), True),
    "V-19-slim-eSE": ((64, 64, 128), (64, 80, 96, 112), (112, 256, 384, 512), 3, (1, 1, 1, 1), False),
    "V-19-eSE": ((64, 64, 128), (128, 160, 192, 224), (256, 512, 768, 1024), 3, (1, 1, 1, 1), False),
    "V-39-eSE": ((64, 64, 128), (128, 160, 192, 224), (256, 512, 768, 1024), 5, (1, 1, 2, 2), False),
    "V-57-eSE": ((64, 64, 128), (128, 160, 192, 224), (256, 512, 768, 1024), 5, (1, 1, 4, 3), False),
    "V-99-eSE": ((64, 64, 128), (128, 160, 192, 224), (256, 512, 768, 1024), 5, (1, 3, 9, 3), False),
}


class OSABlock(object):
    """One eSE-OSA block: ``n_conv`` chained 3x3 units, 1x1 aggregation over all of them, eSE gate."""

    def __init__(self, stage, index, in_ch, mid_ch, out_ch, n_conv, dw=False):
        self.stage, self.index = stage, index
        self.name = "OSA{}_{}".format(stage, index)
        self.in_ch, self.mid_ch, self.out_ch, self.n_conv = in_ch, mid_ch, out_ch, n_conv
        self.identity = index > 1          # vovnet.py:363-376: blocks >= 2 add their input
        self.cat_ch = in_ch + n_conv * mid_ch
        self.dw = dw                       # vovnet.py:284-293: depthwise 3x3 + pointwise 1x1 units
        self.reduced = dw and in_ch != mid_ch      # vovnet.py:278-283: 1x1 reduction in front of the chain

    def reduction_key(self):
        return "stage{s}.{n}.conv_reduction.{n}_reduction_0".format(s=self.stage, n=self.name)

    def key(self, unit):
        # unit: 0..n_conv-1 for the 3x3 chain, "concat" for the aggregation conv
        if unit == "concat":
            return "stage{s}.{n}.concat.{n}_concat".format(s=self.stage, n=self.name)
        return "stage{s}.{n}.layers.{i}.{n}_{i}".format(s=self.stage, n=self.name, i=unit)

    def ese_key(self):
        return "stage{s}.{n}.ese.fc".format(s=self.stage, n=self.name)


def vovnet_blocks(body):
    """Stem widths and the flat list of OSA blocks (stage 2..5) of a VoVNetV2 body."""
    if body not in VOVNET_BODIES:
        raise KeyError("unsupported MODEL.VOVNET.CONV_BODY '{}'".format(body))
    stem, mid, out, n_conv, per_stage, dw = VOVNET_BODIES[body]
    blocks = []
    in_ch = stem[2]
    for si in range(4):
        for bi in range(per_stage[si]):
            blocks.append(OSABlock(si + 2, bi + 1, in_ch if bi == 0 else out[si], mid[si], out[si], n_conv, dw))
        in_ch = out[si]
    return stem, blocks


def vovnet_is_depthwise(body):
    return VOVNET_BODIES[body][5]


def _conv_bn(spec, prefix, cout, cin, k):
    spec[prefix + "/conv.weight"] = ((cout, cin, k, k), "conv_relu")
    spec[prefix + "/norm.weight"] = ((cout,), "bn_weight")
    spec[prefix + "/norm.bias"] = ((cout,), "bn_bias")
    spec[prefix + "/norm.running_mean"] = ((cout,), "bn_mean")
    spec[prefix + "/norm.running_var"] = ((cout,), "bn_var")


def _dw_pw_bn(spec, prefix, cout, cin):
    # vovnet.py:110-130: Conv2d(cin, cout, 3, groups=cout) needs cin == cout; no norm / ReLU between dw and pw
    assert cin == cout, (prefix, cin, cout)
    spec[prefix + "/dw_conv3x3.weight"] = ((cout, 1, 3, 3), "conv_linear")
    spec[prefix + "/pw_conv1x1.weight"] = ((cout, cin, 1, 1), "conv_relu")
    spec[prefix + "/pw_norm.weight"] = ((cout,), "bn_weight")
    spec[prefix + "/pw_norm.bias"] = ((cout,), "bn_bias")
    spec[prefix + "/pw_norm.running_mean"] = ((cout,), "bn_mean")
    spec[prefix + "/pw_norm.running_var"] = ((cout,), "bn_var")


def backbone_param_spec(cfg):
    """``state_dict`` layout of ``build_fcos_vovnet_fpn_backbone(cfg, ...)`` -> {key: (shape, kind)}."""
    stem, blocks = vovnet_blocks(cfg.MODEL.VOVNET.CONV_BODY)
    dw = vovnet_is_depthwise(cfg.MODEL.VOVNET.CONV_BODY)
    spec = OrderedDict()
    cin = 3
    for i, c in enumerate(stem):
        if dw and i > 0:                    # vovnet.py:408-411: stem_2 / stem_3 follow the body's conv type
            _dw_pw_bn(spec, "bottom_up.stem.stem_{}".format(i + 1), c, cin)
        else:
            _conv_bn(spec, "bottom_up.stem.stem_{}".format(i + 1), c, cin, 3)
        cin = c
    for b in blocks:
        c = b.in_ch
        if b.reduced:
            _conv_bn(spec, "bottom_up." + b.reduction_key(), b.mid_ch, b.in_ch, 1)
            c = b.mid_ch
        for i in range(b.n_conv):
            if b.dw:
                _dw_pw_bn(spec, "bottom_up." + b.key(i), b.mid_ch, c)
            else:
                _conv_bn(spec, "bottom_up." + b.key(i), b.mid_ch, c, 3)
            c = b.mid_ch
        _conv_bn(spec, "bottom_up." + b.key("concat"), b.out_ch, b.cat_ch, 1)
        if b.identity:
            # synthetic-weight hint only: residual blocks get a damped aggregation conv so that deep bodies
            # (V-99: 9 residual blocks in stage 4) keep O(1..100) activations instead of growing geometrically
            k = "bottom_up." + b.key("concat") + "/conv.weight"
            spec[k] = (spec[k][0], "conv_relu_residual")
        spec["bottom_up." + b.ese_key() + ".weight"] = ((b.out_ch, b.out_ch, 1, 1), "ese_weight")
        spec["bottom_up." + b.ese_key() + ".bias"] = ((b.out_ch,), "ese_bias")
    out_ch = {"stage{}".format(b.stage): b.out_ch for b in blocks}
    fpn_ch = cfg.MODEL.FPN.OUT_CHANNELS
    for f in cfg.MODEL.FPN.IN_FEATURES:
        lvl = int(f[-1])              # stage3 -> stride 8 -> log2 = 3
        spec["fpn_lateral{}.weight".format(lvl)] = ((fpn_ch, out_ch[f], 1, 1), "conv_linear")
        spec["fpn_lateral{}.bias".format(lvl)] = ((fpn_ch,), "bias")
        spec["fpn_output{}.weight".format(lvl)] = ((fpn_ch, fpn_ch, 3, 3), "conv_linear")
        spec["fpn_output{}.bias".format(lvl)] = ((fpn_ch,), "bias")
    for i in range(cfg.MODEL.FCOS.TOP_LEVELS):
        spec["top_block.p{}.weight".format(6 + i)] = ((fpn_ch, fpn_ch, 3, 3), "conv_linear")
        spec["top_block.p{}.bias".format(6 + i)] = ((fpn_ch,), "bias")
    return spec


def fcos_param_spec(cfg, in_channels):
    """``state_dict`` layout of the ``FCOS`` proposal generator (fcos.py:147-205)."""
    spec = OrderedDict()
    nclass = cfg.MODEL.FCOS.NUM_CLASSES
    use_gn = cfg.MODEL.FCOS.NORM == "GN"
    per_unit = 3 if use_gn else 2          # conv, [GN], ReLU occupy consecutive Sequential slots
    for tower, n in (("cls", cfg.MODEL.FCOS.NUM_CLS_CONVS), ("bbox", cfg.MODEL.FCOS.NUM_BOX_CONVS),
                     ("share", cfg.MODEL.FCOS.NUM_SHARE_CONVS)):
        for i in range(n):
            p = "fcos_head.{}_tower.{}".format(tower, per_unit * i)
            spec[p + ".weight"] = ((in_channels, in_channels, 3, 3), "conv_relu")
            spec[p + ".bias"] = ((in_channels,), "bias")
            if use_gn:
                q = "fcos_head.{}_tower.{}".format(tower, per_unit * i + 1)
                spec[q + ".weight"] = ((in_channels,), "gn_weight")
                spec[q + ".bias"] = ((in_channels,), "gn_bias")
    spec["fcos_head.cls_logits.weight"] = ((nclass, in_channels, 3, 3), "cls_logits")
    spec["fcos_head.cls_logits.bias"] = ((nclass,), "cls_bias")
    spec["fcos_head.bbox_pred.weight"] = ((4, in_channels, 3, 3), "bbox_pred")
    spec["fcos_head.bbox_pred.bias"] = ((4,), "bbox_bias")
    spec["fcos_head.ctrness.weight"] = ((1, in_channels, 3, 3), "ctrness")
    spec["fcos_head.ctrness.bias"] = ((1,), "zero")
    if cfg.MODEL.FCOS.USE_SCALE:
        for l in range(len(cfg.MODEL.FCOS.FPN_STRIDES)):
            spec["fcos_head.scales.{}.scale".format(l)] = ((1,), "scale")
    return spec


def mask_head_param_spec(cfg, in_channels):
    """``state_dict`` layout of ``SpatialAttentionMaskHead`` (sam.py:41-90), keys relative to the head."""
    spec = OrderedDict()
    dim = cfg.MODEL.ROI_MASK_HEAD.CONV_DIM
    n_conv = cfg.MODEL.ROI_MASK_HEAD.NUM_CONV
    if cfg.MODEL.ROI_MASK_HEAD.NORM:
        raise NotImplementedError("ROI_MASK_HEAD.NORM != '' is not configured by the reference")
    c = in_channels
    for k in range(n_conv):
        spec["mask_fcn{}.weight".format(k + 1)] = ((dim, c, 3, 3), "conv_relu")
        spec["mask_fcn{}.bias".format(k + 1)] = ((dim,), "bias")
        c = dim
    spec["spatialAtt.conv.weight"] = ((1, 2, 3, 3), "sam")
    spec["deconv.weight"] = ((c, dim, 2, 2), "deconv")
    spec["deconv.bias"] = ((dim,), "bias")
    ncls = 1 if cfg.MODEL.ROI_MASK_HEAD.CLS_AGNOSTIC_MASK else cfg.MODEL.ROI_HEADS.NUM_CLASSES
    spec["predictor.weight"] = ((ncls, dim, 1, 1), "predictor")
    spec["predictor.bias"] = ((ncls,), "zero")
    return spec


def maskiou_head_param_spec(cfg, in_channels, resolution):
    """``state_dict`` layout of ``MaskIoUHead`` (maskiou_head.py:64-105); ``in_channels`` = channels of the pooled ROI
    feature (the head concatenates one mask channel, :71), ``resolution`` = width of its input (:72 halves it)."""
    spec = OrderedDict()
    dim = cfg.MODEL.ROI_MASKIOU_HEAD.CONV_DIM
    n_conv = cfg.MODEL.ROI_MASKIOU_HEAD.NUM_CONV
    res = resolution // 2
    c = in_channels + 1
    for k in range(n_conv):
        spec["maskiou_fcn{}.weight".format(k + 1)] = ((dim, c, 3, 3), "conv_relu")
        spec["maskiou_fcn{}.bias".format(k + 1)] = ((dim,), "bias")
        c = dim
    spec["maskiou_fc1.weight"] = ((1024, dim * res * res), "fc_relu")
    spec["maskiou_fc1.bias"] = ((1024,), "bias")
    spec["maskiou_fc2.weight"] = ((1024, 1024), "fc_relu")
    spec["maskiou_fc2.bias"] = ((1024,), "bias")
    spec["maskiou.weight"] = ((cfg.MODEL.ROI_HEADS.NUM_CLASSES, 1024), "maskiou_out")
    spec["maskiou.bias"] = ((cfg.MODEL.ROI_HEADS.NUM_CLASSES,), "maskiou_bias")
    return spec


def keypoint_head_param_spec(cfg, in_channels):
    """KRCNNConvDeconvUpsampleHead, keypoint_head.py:168-215."""
    spec = OrderedDict()
    kh = cfg.MODEL.ROI_KEYPOINT_HEAD
    c = in_channels
    for k, dim in enumerate(kh.CONV_DIMS, 1):
        spec["conv_fcn{}.weight".format(k)] = ((dim, c, 3, 3), "conv_relu")
        spec["conv_fcn{}.bias".format(k)] = ((dim,), "bias")
        c = dim
    spec["score_lowres.weight"] = ((c, kh.NUM_KEYPOINTS, 4, 4), "kp_deconv")
    spec["score_lowres.bias"] = ((kh.NUM_KEYPOINTS,), "bias")
    return spec


def roi_heads_param_spec(cfg, in_channels):
    """``state_dict`` layout of ``CenterROIHeads``: its heads under ``mask_head.`` / ``maskiou_head.`` / ``keypoint_head.``
    (center_heads.py:337-383)."""
    spec = OrderedDict()
    if cfg.MODEL.MASK_ON:
        for k, v in mask_head_param_spec(cfg, in_channels).items():
            spec["mask_head." + k] = v
    if cfg.MODEL.MASKIOU_ON:
        for k, v in maskiou_head_param_spec(cfg, in_channels, cfg.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION).items():
            spec["maskiou_head." + k] = v
    if cfg.MODEL.KEYPOINT_ON:
        for k, v in keypoint_head_param_spec(cfg, in_channels).items():
            spec["keypoint_head." + k] = v
    return spec


def model_param_spec(cfg):
    """Full ``GeneralizedRCNN`` layout: ``backbone.*``, ``proposal_generator.*``, ``roi_heads.*``."""
    spec = OrderedDict()
    for k, v in backbone_param_spec(cfg).items():
        spec["backbone." + k] = v
    for k, v in fcos_param_spec(cfg, cfg.MODEL.FPN.OUT_CHANNELS).items():
        spec["proposal_generator." + k] = v
    for k, v in roi_heads_param_spec(cfg, cfg.MODEL.FPN.OUT_CHANNELS).items():
        spec["roi_heads." + k] = v
    return spec


def conv_gflop_per_image(cfg, height, width, rois):
    """Algorithmic dense FLOPs (2*MAC) of one padded ``height x width`` image with ``rois`` detections.

    Same convention as SURVEY.md 8(d) / BASELINE.md section 5; used for ``roofline.achieved``."""
    def conv(h, w, cin, cout, k):
        return 2.0 * h * w * cin * cout * k * k

    def half(x):
        return (x - 1) // 2 + 1          # 3x3 s2 p1 conv == ceil(x/2); also maxpool3 s2 ceil on even x

    stem, blocks = vovnet_blocks(cfg.MODEL.VOVNET.CONV_BODY)
    dw = vovnet_is_depthwise(cfg.MODEL.VOVNET.CONV_BODY)

    def unit(h, w, cin, cout):            # one 3x3 unit: dense, or depthwise 3x3 + pointwise 1x1
        return 2.0 * h * w * cin * 9 + conv(h, w, cin, cout, 1) if dw else conv(h, w, cin, cout, 3)

    f = 0.0
    h, w = half(height), half(width)
    f += conv(h, w, 3, stem[0], 3) + unit(h, w, stem[0], stem[1])
    h, w = half(h), half(w)
    f += unit(h, w, stem[1], stem[2])
    sizes = {}
    stage = 2
    for b in blocks:
        if b.stage != stage:
            h, w = -(-(h - 3) // 2) + 1, -(-(w - 3) // 2) + 1
            stage = b.stage
        c = b.in_ch
        if b.reduced:
            f += conv(h, w, c, b.mid_ch, 1)
            c = b.mid_ch
        for _ in range(b.n_conv):
            f += unit(h, w, c, b.mid_ch)
            c = b.mid_ch
        f += conv(h, w, b.cat_ch, b.out_ch, 1)
        sizes[b.stage] = (h, w, b.out_ch)
    fc = cfg.MODEL.FPN.OUT_CHANNELS
    levels = []
    for s in (3, 4, 5):
        hh, ww, cc = sizes[s]
        f += conv(hh, ww, cc, fc, 1) + conv(hh, ww, fc, fc, 3)
        levels.append((hh, ww))
    hh, ww = levels[-1]
    for _ in range(cfg.MODEL.FCOS.TOP_LEVELS):
        hh, ww = half(hh), half(ww)
        f += conv(hh, ww, fc, fc, 3)
        levels.append((hh, ww))
    px = sum(a * b for a, b in levels)
    n_tower = cfg.MODEL.FCOS.NUM_CLS_CONVS + cfg.MODEL.FCOS.NUM_BOX_CONVS + 2 * cfg.MODEL.FCOS.NUM_SHARE_CONVS
    f += px * 2.0 * fc * 9 * (fc * n_tower + cfg.MODEL.FCOS.NUM_CLASSES + 4 + 1)
    if cfg.MODEL.MASK_ON:
        r = cfg.MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION
        d = cfg.MODEL.ROI_MASK_HEAD.CONV_DIM
        per = 0.0
        c = fc
        for _ in range(cfg.MODEL.ROI_MASK_HEAD.NUM_CONV):
            per += conv(r, r, c, d, 3)
            c = d
        # the deconv, then the predictor for the ONE class of the ROI: mask_rcnn_inference (mask_head.py:196-216) reads a
        # single class plane, the engine computes only that one (class-gathered predictor) -- skipped work is not credited
        per += 2.0 * r * r * c * d * 4 + conv(2 * r, 2 * r, d, 1, 1)
        if cfg.MODEL.MASKIOU_ON:
            di = cfg.MODEL.ROI_MASKIOU_HEAD.CONV_DIM
            nc = cfg.MODEL.ROI_MASKIOU_HEAD.NUM_CONV
            c = d + 1
            for k in range(nc):
                rr = r // 2 if k == nc - 1 else r
                per += conv(rr, rr, c, di, 3)
                c = di
            per += 2.0 * (di * (r // 2) ** 2 * 1024 + 1024 * 1024 + 1024 * cfg.MODEL.ROI_HEADS.NUM_CLASSES)
        f += per * rois
    return f / 1e9
