#!/usr/bin/env python
"""Layer-isolated error of the MaskIoU head on the full-size workload: every layer's device output against an fp64
evaluation of the same layer on the DEVICE's own input (so errors do not compound).  Usage: python tools/debug/maskiou_layers.py [precision]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import centermask2_b200 as cm                      # noqa: E402
from centermask2_b200 import runtime              # noqa: E402
from centermask2_b200.config import get_cfg       # noqa: E402
from tests import fullsize                        # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    sd, imgs = fullsize.workload()
    cfg = get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.B200.PRECISION", prec])
    model = cm.build_model(cfg)
    model.load_state_dict(sd)
    eng = runtime.engine_for(cfg)
    global ENG
    ENG = eng
    eng.use_graphs = False
    model.inference(imgs, do_postprocess=False)
    torch.cuda.synchronize()

    def buf(name):
        for key, t in list(eng._bufs.items()):
            if len(key) == 3 and key[0] == name:
                return t
        raise KeyError(name)

    def dn(t):
        return t.double().cpu()

    roi = dn(buf("roi_feat")[:, 1:-1, 1:-1, :]).permute(0, 3, 1, 2)
    pm = dn(buf("iou_mask")[:, 1:-1, 1:-1, :1]).permute(0, 3, 1, 2)
    p = "roi_heads.maskiou_head."
    x = torch.cat([roi, pm], 1)
    print("roi_feat |max| {:.3e}  pm |max| {:.3e}".format(roi.abs().max().item(), pm.abs().max().item()))
    for k in range(1, 5):
        w, b = sd[p + "maskiou_fcn{}.weight".format(k)].double(), sd[p + "maskiou_fcn{}.bias".format(k)].double()
        ref = F.relu(F.conv2d(x, w, b, 2 if k == 4 else 1, 1))
        t = buf("iou_fcn{}".format(k))
        if t.dim() == 5:                       # phase planes [4, n, h/2+2, w/2+2, c] -> full resolution
            n, c = t.shape[1], t.shape[4]
            full = torch.zeros((n, 14, 14, c), dtype=torch.float64)
            for q in range(4):
                py, px = q >> 1, q & 1
                full[:, py::2, px::2] = dn(t[q, :, 1:8, 1:8, :])
            got = full.permute(0, 3, 1, 2)
        elif t.shape[1] == 16:
            got = dn(t[:, 1:-1, 1:-1, :]).permute(0, 3, 1, 2)
        else:
            got = dn(t).permute(0, 3, 1, 2)
        e = (got - ref).abs().max().item() / ref.abs().max().item()
        print("iou_fcn{}: |ref|max {:.3e}  max err / max|ref| {:.2e}".format(k, ref.abs().max().item(), e))
        x = got
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)          # device flatten order (h, w, c)
    w1 = sd[p + "maskiou_fc1.weight"].double()
    w1 = w1.reshape(w1.shape[0], 256, 7, 7).permute(0, 2, 3, 1).reshape(w1.shape[0], -1)
    for name, w, b, relu in (("iou_fc1", w1, sd[p + "maskiou_fc1.bias"].double(), True),
                             ("iou_fc2", sd[p + "maskiou_fc2.weight"].double(), sd[p + "maskiou_fc2.bias"].double(), True),
                             ("iou_out", sd[p + "maskiou.weight"].double(), sd[p + "maskiou.bias"].double(), False)):
        ref = F.linear(x, w, b)
        ref = F.relu(ref) if relu else ref
        got = dn(buf(name)).reshape(ref.shape)
        e = (got - ref).abs().max().item() / ref.abs().max().item()
        print("{}: |ref|max {:.3e}  max err / max|ref| {:.2e}".format(name, ref.abs().max().item(), e))
        x = got


def against_oracle(prec):
    """End-to-end: device tensors of the ROI stage against the fp32 oracle's trace (errors compound here)."""
    raw, post, tr = fullsize.oracle_outputs(False)
    sd, imgs = fullsize.workload()
    eng = ENG

    def buf(name):
        for key, t in list(eng._bufs.items()):
            if len(key) == 3 and key[0] == name:
                return t
        raise KeyError(name)

    roi = buf("roi_feat")[:, 1:-1, 1:-1, :].double().cpu().permute(0, 3, 1, 2)
    ref = tr["roi_feat"].double()
    print("roi_feat vs oracle: max err / max {:.2e}".format(((roi - ref).abs().max() / ref.abs().max()).item()))
    iou = buf("iou_out").double().cpu().reshape(ref.shape[0], -1)
    riou = tr["maskiou"].double()
    d = (iou - riou).abs()
    print("maskiou [R,80] vs oracle: max err {:.3e} (|ref| max {:.3e}); per-ROI max err: {}".format(
        d.max().item(), riou.abs().max().item(), [round(v, 4) for v in d.max(dim=1).values.tolist()[:100]]))
    from oracle import restate
    import torchvision
    worst = int((roi - ref).abs().reshape(ref.shape[0], -1).max(dim=1).values.argmax())
    boxes_dev = buf("det_boxes").cpu().reshape(-1, 4)
    boxes_ref = torch.cat([r["pred_boxes"] for r in raw])
    img_area = float(fullsize.H * fullsize.W)
    lv_dev = restate.assign_levels_by_ratio(boxes_dev, img_area, 3, 5)
    lv_ref = restate.assign_levels_by_ratio(boxes_ref, img_area, 3, 5)
    b = boxes_ref[worst]
    area = ((b[2] - b[0]) * (b[3] - b[1])).item()
    print("worst ROI {}: box ref {} dev {} area/img {:.9f} level ref {} (trace {}) dev-box level {}".format(
        worst, b.tolist(), boxes_dev[worst].tolist(), area / img_area, int(lv_ref[worst]), int(tr["roi_levels"][worst]), int(lv_dev[worst])))
    feats = tr["features"]
    for li, nme in enumerate(("p3", "p4", "p5")):
        rois = torch.cat([torch.tensor([[float(worst // 50)]]), boxes_dev[worst:worst + 1]], dim=1)
        cand = torchvision.ops.roi_align(feats[nme], rois, 14, 1.0 / (8 << li), 0, True).double()[0]
        print("   device roi_feat[{}] vs torchvision roi_align on oracle {}: max err / max {:.2e}".format(
            worst, nme, ((roi[worst] - cand).abs().max() / cand.abs().max()).item()))
    pm = buf("iou_mask")[:, 1:-1, 1:-1, :1].double().cpu().permute(0, 3, 1, 2)
    probs = torch.cat([r["pred_masks"] for r in raw]).double()
    rpm = F.max_pool2d(probs, 2, 2)
    print("pooled mask vs oracle: max err {:.3e}".format((pm - rpm).abs().max().item()))


if __name__ == "__main__":
    main()
    against_oracle(sys.argv[1] if len(sys.argv) > 1 else "fp32")
