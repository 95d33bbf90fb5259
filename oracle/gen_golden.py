"""TEST INFRASTRUCTURE ONLY -- generate ``tests/golden/*.pt`` by running the UNMODIFIED reference.

Run in the build container (needs ``/root/reference``):  ``python -m oracle.gen_golden [case ...]``

For each case in ``oracle/cases.py``: build seeded synthetic weights (``centermask2_b200.synth``),
calibrate ``cls_logits.bias`` so that a useful number of candidates survives the 0.05 threshold,
run the reference model (``oracle/refrun.py``) with and without ``detector_postprocess``, and
store the outputs plus the FCOS head tensors.  Weights are *not* stored (they are regenerated from
the seed); a checksum of the weights is stored so that RNG drift is detected.
"""
import os
import sys

import numpy as np
import torch

from centermask2_b200.synth import synthetic_state_dict, synthetic_images, calibrate_cls_bias
from oracle import refrun, restate
from oracle.cases import CASES, case_cfg, case_flags

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_case(name):
    overrides, sizes, wseed, iseed, target = CASES[name][:5]
    cfg = case_cfg(name)
    sd = synthetic_state_dict(cfg, seed=wseed)
    inputs = []
    for i, (h, w) in enumerate(sizes):
        inputs.extend(synthetic_images(1, h, w, seed=iseed + i))
    return cfg, sd, inputs, target


def weights_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


def calibrate(cfg, sd, inputs, target):
    key = "proposal_generator.fcos_head.cls_logits.bias"
    if target == 0:
        sd[key] = torch.full_like(sd[key], -20.0)
        return -20.0
    sd[key] = torch.zeros_like(sd[key])
    tr = {}
    restate.run_model(inputs, sd, cfg, postprocess=False, trace=tr)
    b = calibrate_cls_bias(tr["logits"], target)
    sd[key] = torch.full_like(sd[key], b)
    return b


def cfg_cpu(cfg):
    c = cfg.clone()
    c.MODEL.DEVICE = "cpu"
    return c


def fields_to_dict(inst):
    f = inst.get_fields()
    out = {"image_size": tuple(inst.image_size)}
    for k, v in f.items():
        out[k] = v.tensor.clone() if hasattr(v, "tensor") else v.clone()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]                   # optional case names; default: all
    for name in CASES:
        if only and name not in only:
            continue
        cfg, sd, inputs, target = build_case(name)
        bias = calibrate(cfg, sd, inputs, target)
        model = refrun.build_reference_model(cfg, sd)
        # head tensors straight from the reference modules
        with torch.no_grad(), refrun.quiet():
            images = model.preprocess_image(inputs)
            feats = model.backbone(images.tensor)
            flist = [feats[f] for f in cfg.MODEL.FCOS.IN_FEATURES]
            logits, regs, ctrs, _ = model.proposal_generator.fcos_head(flist)
        raw = refrun.run_reference(model, inputs, postprocess=False)
        raw = [fields_to_dict(r) for r in raw]
        post = refrun.run_reference(model, inputs, postprocess=True)
        post = [fields_to_dict(r["instances"]) for r in post]
        for p in post:
            if "pred_masks" in p:
                m = p.pop("pred_masks")
                p["pred_masks_shape"] = tuple(m.shape)
                p["pred_masks_packed"] = torch.from_numpy(np.packbits(m.numpy().reshape(-1)))
        n_cand = [int((l.sigmoid() > cfg.MODEL.FCOS.INFERENCE_TH_TEST).sum()) for l in logits]
        gold = {
            "case": name, "cls_bias": bias, "weights_checksum": weights_checksum(sd),
            "keys": list(model.state_dict().keys()),
            "features": {k: v.clone() for k, v in feats.items()},
            "logits": logits, "regs": regs, "ctrs": ctrs, "candidates_per_level": n_cand,
            "raw": raw, "post": post,
            "torch": torch.__version__,
        }
        flags = case_flags(name)
        if flags.get("lean"):                 # variant cases: the end results pin the behaviour, intermediates are dropped
            for k in ("features", "logits", "regs", "ctrs"):
                gold.pop(k)
        if flags.get("tensor_in"):
            # modified_class.GeneralizedRCNN.forward (the fork's export-friendly meta-arch): already normalised + padded
            # tensor in, 6-tuple out, image_sizes fixed by FakeImageList (modified_class.py:11-24)
            sys.path.insert(0, refrun.REFERENCE_ROOT)
            import modified_class
            m2 = modified_class.GeneralizedRCNN(cfg_cpu(cfg))
            m2.eval()
            m2.load_state_dict(sd, strict=True)
            with torch.no_grad(), refrun.quiet():
                x = model.preprocess_image(inputs).tensor.clone()
                out = m2(x)
            gold["tensor_in"] = {"input": x, "outputs": [o.clone() for o in out],
                                 "names": ["locations", "mask_scores", "pred_boxes", "pred_classes", "pred_masks", "scores"]}
        path = os.path.join(OUT, name + ".pt")
        torch.save(gold, path)
        print("{:16s} bias {:8.3f} cand/level {} dets {} -> {} ({:.0f} KB)".format(
            name, bias, n_cand, [len(r["scores"]) for r in raw], path, os.path.getsize(path) / 1024))


if __name__ == "__main__":
    sys.exit(main())
