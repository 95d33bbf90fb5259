"""GPU: the registered B200 plug-ins end to end against (a) the golden vectors produced by the
unmodified reference and (b) the fp32 oracle restatement, layer by layer.

Tolerances (BASELINE.json north_star, fp32 variant): boxes <= 1e-2 px, scores / mask_scores <= 1e-3,
mask IoU >= 0.99 per instance, kept-detection sets (classes + originating locations, in order) identical."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import centermask2_b200 as cm                                                    # noqa: E402
from centermask2_b200 import runtime                                             # noqa: E402
from oracle import restate                                                       # noqa: E402
from oracle.cases import CASES                                                   # noqa: E402
from tests.helpers import (load_golden, build_case, unpack_masks, mask_iou,     # noqa: E402
                           assert_detections_match, assert_masks_match, assert_keypoints_match, MASK_IOU_MIN)


def fields(inst):
    out = {}
    for k, v in inst.get_fields().items():
        out[k] = (v.tensor if hasattr(v, "tensor") else v).detach().cpu()
    return out


# both fp32 engines: "fp32" = tensor cores on split f16 hi/lo operands, "fp32_simt" = CUDA cores
@pytest.fixture(scope="module", params=[(c, p) for p in ("fp32", "fp32_simt") for c in sorted(CASES)],
                ids=lambda cp: "{}-{}".format(*cp))
def case(request):
    name, precision = request.param
    runtime.reset()
    runtime.set_precision(precision)
    gold = load_golden(name)
    cfg, sd, inputs = build_case(name, gold)
    model = cm.build_model(cfg)
    model.load_state_dict(sd, strict=True)
    yield name, gold, cfg, sd, inputs, model
    runtime.reset()
    runtime.set_precision("fp32")


def test_backbone_and_head_tensors_match_reference(case):
    name, gold, cfg, sd, inputs, model = case
    if "features" not in gold:
        pytest.skip("lean golden (variant case): end results only")
    images = model.preprocess_image(inputs)
    feats = model.backbone(images.tensor)
    for k, v in gold["features"].items():
        got = feats[k].float().cpu()
        assert got.shape == v.shape, (k, got.shape, v.shape)
        err = (got - v).abs().max().item()
        assert err <= 2e-3 * max(1.0, v.abs().max().item()), (k, err)


def test_raw_detections_match_reference(case):
    name, gold, cfg, sd, inputs, model = case
    crowded = max(gold["candidates_per_level"]) > cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST
    raw = model.inference(inputs, do_postprocess=False)
    if crowded:
        # the fork has no pre-NMS top-k (fcos_outputs.py:444-449); compare with the restated upstream semantics
        ref = restate.run_model(inputs, sd, cfg, postprocess=False, pre_topk=True)
    else:
        ref = gold["raw"]
    for i, (g, r) in enumerate(zip(raw, ref)):
        g = fields(g)
        assert_detections_match(g, r, what="{}[{}]".format(name, i))
        assert ("pred_keypoints" in g) == ("pred_keypoints" in r)
        if "pred_keypoints" in r:
            # fp32 accumulation-order noise (~1e-5 on the logits) may move an arg-max between two near-equal pixels
            assert_keypoints_match(g["pred_keypoints"], r["pred_keypoints"], what="{}[{}]".format(name, i), max_moved_frac=0.02)
        if len(r["scores"]):
            assert (g["pred_masks"] - r["pred_masks"]).abs().max().item() <= 1e-3
            assert "mask_scores" in g
        else:
            assert "mask_scores" not in g


def test_postprocessed_match_reference(case):
    name, gold, cfg, sd, inputs, model = case
    if max(gold["candidates_per_level"]) > cfg.MODEL.FCOS.PRE_NMS_TOPK_TEST:
        pytest.skip("crowded case is covered by the raw comparison")
    out = model(inputs)
    for i, (o, r) in enumerate(zip(out, gold["post"])):
        g = fields(o["instances"])
        assert tuple(o["instances"].image_size) == tuple(r["image_size"])
        assert_detections_match(g, r, what="{}[{}] post".format(name, i))
        if "pred_keypoints" in r:
            assert_keypoints_match(g["pred_keypoints"], r["pred_keypoints"], what="{}[{}] post".format(name, i), max_moved_frac=0.02)
        if len(r["scores"]):
            ref_masks = unpack_masks(r)
            assert g["pred_masks"].dtype == torch.bool and g["pred_masks"].shape == ref_masks.shape
            assert_masks_match(g["pred_masks"], ref_masks, what="{}[{}]".format(name, i))


def test_inference_stream_equals_forward_per_batch(case):
    """The pipelined API (next batch's host->device copy overlapped on a copy stream) returns exactly what
    ``forward`` returns batch by batch, in order, for batches with different pixels."""
    name, gold, cfg, sd, inputs, model = case
    batches = []
    for k in range(3):
        batch = []
        for b in inputs:
            img = b["image"].roll(shifts=7 * k, dims=-1).contiguous().pin_memory()
            batch.append(dict(b, image=img))
        batches.append(batch)
    want = []
    for batch in batches:
        want.append([fields(o["instances"]) for o in model(batch)])
    got = [[fields(o["instances"]) for o in out] for out in model.inference_stream(iter(batches))]
    assert len(got) == len(want) == 3
    for gb, wb in zip(got, want):
        assert len(gb) == len(wb)
        for g, w in zip(gb, wb):
            assert set(g) == set(w)
            for k in g:
                assert torch.equal(g[k], w[k]), k


def test_registry_level_modules_compose_like_the_reference(case):
    """backbone -> FCOS.forward -> CenterROIHeads.forward through the public module API (the sequence of
    modified_class.py:27-40), on arbitrary (non-engine) NCHW feature tensors."""
    name, gold, cfg, sd, inputs, model = case
    images = model.preprocess_image(inputs)
    feats = model.backbone(images.tensor)
    feats = {k: v.contiguous().clone() for k, v in feats.items()}             # drop the zero-copy tag
    props, _ = model.proposal_generator(images, feats, None)
    res, _ = model.roi_heads(images, feats, props, None)
    assert res is not None and len(res) == len(inputs)
    for p, r in zip(res, props):
        assert p is r                                                          # mutated in place, center_heads.py:433-444
    raw = model.inference(inputs, do_postprocess=False)
    for a, b in zip(res, raw):
        fa, fb = fields(a), fields(b)
        assert set(fa) == set(fb)
        for k in fa:
            # ``inference`` runs the FCOS towers as one segmented launch over all levels, the module-level call one launch
            # per level: other tile shapes, another accumulation order (~1e-6 relative per layer, amplified by the layers
            # behind it); mask probabilities and mask scores see the whole ROI stage on top.  The module-level backbone also takes
            # the NORMALISED tensor (stem_1 as an ordinary convolution), ``inference`` the raw images (fused stem, csrc/stem.cu)
            big = fb[k].float().abs().max().item() if fb[k].numel() else 0.0
            tol = 1e-3 * max(1.0, big) if k in ("pred_masks", "mask_scores") else 5e-6 * max(10.0, big)   # 1e-3: the north-star probability tolerance
            assert torch.allclose(fa[k].float(), fb[k].float(), atol=tol), k


def test_layerwise_against_oracle_trace():
    runtime.reset()
    runtime.set_precision("fp32")
    name = "v39_one_image"
    gold = load_golden(name)
    cfg, sd, inputs = build_case(name, gold)
    tr = {}
    restate.run_model(inputs, sd, cfg, postprocess=False, trace=tr)
    model = cm.build_model(cfg)
    model.load_state_dict(sd)
    eng = runtime.engine_for(cfg)
    x, sizes = eng.preprocess([b["image"].cuda() for b in inputs], fused_stem=False)
    assert torch.allclose(x.view.permute(0, 3, 1, 2).cpu(), tr["image"], atol=1e-4)
    x, sizes = eng.preprocess([b["image"].cuda() for b in inputs])         # the product path: raw images -> fused stem
    feats = model.backbone.forward_fmap(x)
    fcos = model.proposal_generator
    eng2, P = fcos._pack()
    head = eng2.run_fcos_head([feats[f] for f in fcos.in_features], P)
    for l, (lg, rc) in enumerate(head):
        got = lg.view.permute(0, 3, 1, 2).cpu()
        assert (got - tr["logits"][l]).abs().max().item() <= 2e-3, l
        reg = torch.relu(rc.view[..., :4] * P["reg_scale"][l]).permute(0, 3, 1, 2).cpu()      # Scale + ReLU live in the decode kernel
        assert (reg - tr["regs"][l]).abs().max().item() <= 2e-3 * max(1.0, tr["regs"][l].abs().max().item()), l
        ctr = rc.view[..., 4:5].permute(0, 3, 1, 2).cpu()
        assert (ctr - tr["ctrs"][l]).abs().max().item() <= 2e-3, l


BF16_SMALL_OVERLAP = 0.75          # the measured overlap is printed by the test (TODO_MEASURED)


def test_bf16_tensor_core_model_matches_bf16_rounding_oracle():
    """bf16 variant (tcgen05 engine).  bf16 through ~60 layers of a random-init network deviates from pure fp32
    by several percent (the same is true of the oracle when it rounds at the same places), so the reference
    here is ``restate.bf16_sim()``: the fp32 restatement with weights and stored activations rounded to bf16
    exactly where the engine rounds.  What remains is accumulation order plus the rare rounding flip.
    Gates: feature maps within 2% relative L2 of the bf16-rounding oracle; head outputs within 2% of their
    range; the post-processing applied to the engine's own head outputs keeps exactly the oracle's detections; end to
    end >= BF16_SMALL_OVERLAP of the oracle's kept detections (class + originating location) are kept -- this 128x160
    case has ~40 kept detections, many of them within a bf16 rounding flip of an NMS / top-k decision; the gate at the
    benchmark size is in tests/test_gpu_fullsize.py.  The raw deviation against the pure-fp32 oracle is printed."""
    runtime.reset()
    runtime.set_precision("bf16")
    try:
        name = "v39_one_image"
        gold = load_golden(name)
        cfg, sd, inputs = build_case(name, gold)
        tr = {}
        with restate.bf16_sim():
            ref = restate.run_model(inputs, sd, cfg, postprocess=False, trace=tr)
        model = cm.build_model(cfg)
        model.load_state_dict(sd)
        eng = runtime.engine_for(cfg)
        x, sizes = eng.preprocess([b["image"].cuda() for b in inputs])
        feats = model.backbone.forward_fmap(x)
        for k, v in tr["features"].items():
            got = feats[k].view.permute(0, 3, 1, 2).float().cpu()
            rel = ((got - v).norm() / v.norm()).item()
            raw = ((got - gold["features"][k]).norm() / gold["features"][k].norm()).item()
            print("bf16 {}: rel L2 vs bf16-rounding oracle {:.4f}, vs fp32 reference {:.4f}".format(k, rel, raw))
            assert rel <= 0.02, (k, rel)
        fcos = model.proposal_generator
        e2, P = fcos._pack()
        head = e2.run_fcos_head([feats[f] for f in fcos.in_features], P)
        for l, (lg, rc) in enumerate(head):
            got = lg.view.permute(0, 3, 1, 2).cpu()
            span = (tr["logits"][l].max() - tr["logits"][l].min()).item()
            assert (got - tr["logits"][l]).abs().max().item() <= 0.02 * span, l
        out = model.inference(inputs, do_postprocess=False)
        # (a) post-processing logic in isolation: the oracle's decode / top-k / NMS applied to the head outputs the
        #     bf16 engine itself produced must keep exactly the same detections in the same order.
        lgs = [lg.view.permute(0, 3, 1, 2).float().cpu() for lg, _ in head]
        regs = [torch.relu(rc.view[..., :4] * P["reg_scale"][l]).permute(0, 3, 1, 2).float().cpu() for l, (_, rc) in enumerate(head)]
        ctrs = [rc.view[..., 4:5].permute(0, 3, 1, 2).float().cpu() for _, rc in head]
        same_in = restate.fcos_postprocess(lgs, regs, ctrs, sizes, cfg)
        for o, r in zip(out, same_in):
            g = fields(o)
            assert g["pred_classes"].tolist() == r["pred_classes"].tolist()
            assert torch.equal(g["locations"], r["locations"])
            assert (g["pred_boxes"] - r["pred_boxes"]).abs().max().item() <= 1e-2
            assert (g["scores"] - r["scores"]).abs().max().item() <= 1e-3
        # (b) end to end against the bf16-rounding oracle: 1% feature noise moves boxes by ~1 px, which flips NMS
        #     decisions near IoU 0.6 and swaps detections near the top-k cut, so only a loose overlap is required here.
        for o, r in zip(out, ref):
            g = fields(o)
            keys_ref = {(int(c), float(l[0]), float(l[1])) for c, l in zip(r["pred_classes"], r["locations"])}
            keys_got = {(int(c), float(l[0]), float(l[1])) for c, l in zip(g["pred_classes"], g["locations"])}
            print("bf16 detections kept in common: {}/{}".format(len(keys_ref & keys_got), len(keys_ref)))
            assert len(keys_ref & keys_got) >= BF16_SMALL_OVERLAP * len(keys_ref)
    finally:
        runtime.reset()
        runtime.set_precision("fp32")


def _detections(model, inputs):
    return [fields(o["instances"]) for o in model(inputs)]


def _same_fields(a, b):
    return all(set(x) == set(y) and all(torch.equal(x[k], y[k]) for k in x) for x, y in zip(a, b))


def test_reloaded_weights_and_second_model_never_replay_a_stale_graph():
    """A captured CUDA graph bakes in the addresses of the packed weights.  (1) After ``load_state_dict`` with other
    weights the next call must compute with the new weights (== an eager run of a freshly built model), not replay the
    old capture; (2) two models built from ONE cfg (they share an engine and its buffers) must each replay their own."""
    runtime.reset()
    runtime.set_precision("fp32")
    name = "v19_two_images"
    gold = load_golden(name)
    cfg, sd, inputs = build_case(name, gold)
    from centermask2_b200.synth import synthetic_state_dict
    sd2 = synthetic_state_dict(cfg, seed=777)
    key = "proposal_generator.fcos_head.cls_logits.bias"
    sd2[key] = sd[key].clone()
    eng = runtime.engine_for(cfg)
    assert eng.use_graphs
    model = cm.build_model(cfg)
    model.load_state_dict(sd)
    a1 = _detections(model, inputs)            # eager (first sighting of the key)
    a2 = _detections(model, inputs)            # capture + replay
    a3 = _detections(model, inputs)            # replay
    assert len(eng._graphs) == 1 and _same_fields(a1, a2) and _same_fields(a1, a3)
    model.load_state_dict(sd2)                 # drops the capture
    assert len(eng._graphs) == 0
    b1 = _detections(model, inputs)
    b2 = _detections(model, inputs)
    b3 = _detections(model, inputs)
    eng.use_graphs = False
    try:
        fresh = cm.build_model(cfg)
        fresh.load_state_dict(sd2)
        want_b = _detections(fresh, inputs)
        fresh.load_state_dict(sd)
        want_a = _detections(fresh, inputs)
    finally:
        eng.use_graphs = True
    assert _same_fields(b1, want_b) and _same_fields(b2, want_b) and _same_fields(b3, want_b)
    assert not _same_fields(b3, want_a)
    # (2) a second model from the same cfg, other weights, interleaved with the first
    other = cm.build_model(cfg)
    other.load_state_dict(sd)
    for _ in range(3):
        assert _same_fields(_detections(other, inputs), want_a)
        assert _same_fields(_detections(model, inputs), want_b)
    assert len(eng._graphs) == 2
    runtime.reset()


def test_inference_records_hand_the_whole_result_to_the_host(case):
    """``inference_records``: detection records + masks as COCO run lengths on the host == what ``forward`` returns as
    ``Instances`` (boxes, scores, classes, locations, mask scores, pasted bool masks), batch after batch."""
    import numpy as np
    from centermask2_b200 import parallel
    name, gold, cfg, sd, inputs, model = case
    oh, ow = 100, 140                                                        # one output size per batch
    batch = [dict(b, image=b["image"].contiguous().pin_memory(), height=oh, width=ow) for b in inputs]
    want = [fields(o["instances"]) for o in model(batch)]
    results = [r.clone() for r in model.inference_records(iter([batch, batch, batch]), rle_capacity=64)]   # tiny capacity: grows
    assert len(results) == 3
    for res in results:
        assert tuple(res.size) == (oh, ow) and res.records.shape[0] == len(batch)
        r_cap = res.records.shape[1]
        assert res.rle_offsets.numel() == len(batch) * r_cap + 1 and int(res.rle_offsets[-1]) == res.rle_runs.numel()
        for i, w in enumerate(want):
            rec = res.records[i]
            k = int(rec[0, parallel.F_COUNT])
            assert (rec[k:, :parallel.F_COUNT] == 0).all()
            keep = [j for j in range(k) if rec[j, parallel.F_VALID] == 1]
            assert len(keep) == len(w["scores"])
            if not keep:
                continue
            sel = torch.tensor(keep)
            assert torch.equal(rec[sel, 0:4], w["pred_boxes"]) and torch.equal(rec[sel, parallel.F_SCORE], w["scores"])
            assert torch.equal(rec[sel, parallel.F_CLASS].long(), w["pred_classes"])
            assert torch.equal(rec[sel, parallel.F_LOC:parallel.F_LOC + 2], w["locations"])
            if "mask_scores" in w:
                assert torch.equal(rec[sel, parallel.F_MASK_SCORE], w["mask_scores"])
            for n_, j in enumerate(keep):
                runs = res.runs(i, j)
                assert int(runs.astype(np.int64).sum()) == oh * ow
                assert np.array_equal(res.mask(i, j), w["pred_masks"][n_].numpy()), (i, j)
            for j in range(r_cap):
                if j not in keep:                                            # empty / dropped slots: one run of zeros
                    assert res.runs(i, j).tolist() == [oh * ow]


@pytest.mark.parametrize("precision", ["fp32", "fp32_simt"])
def test_forward_tensor_matches_the_forks_tuple_out_meta_arch(precision):
    """``GeneralizedRCNN.forward_tensor`` against the 6-tuple the reference's own ``modified_class.GeneralizedRCNN.forward``
    (modified_class.py:27-40) returned for the same normalised + padded tensor (golden ``v19_tensor_in``)."""
    runtime.reset()
    runtime.set_precision(precision)
    try:
        name = "v19_tensor_in"
        gold = load_golden(name)
        cfg, sd, inputs = build_case(name, gold)
        model = cm.build_model(cfg)
        model.load_state_dict(sd)
        out = model.forward_tensor(gold["tensor_in"]["input"].cuda())
        got = dict(zip(gold["tensor_in"]["names"], [o.cpu() for o in out]))
        want = dict(zip(gold["tensor_in"]["names"], gold["tensor_in"]["outputs"]))
        assert_detections_match(got, want, what=name)
        assert (got["pred_masks"] - want["pred_masks"]).abs().max().item() <= 1e-3
    finally:
        runtime.reset()
        runtime.set_precision("fp32")


@pytest.mark.parametrize("precision", ["fp32", "fp32_simt", "bf16"])
def test_standalone_mask_heads_match_the_reference_modules(precision):
    """``build_mask_head(cfg, shape).forward(x)`` (sam.py:92-97: logits of ALL classes) and ``build_maskiou_head(cfg,
    shape).forward(x, mask)`` (maskiou_head.py:107-120) as stand-alone registry objects, against the oracle's restatement
    of the two reference modules; and the heads hanging off ``CenterROIHeads`` are these very classes."""
    from centermask2_b200.modeling import ShapeSpec
    from centermask2_b200.modeling.roi_heads import build_mask_head, build_maskiou_head, SpatialAttentionMaskHead, MaskIoUHead
    from centermask2_b200.synth import synthetic_state_dict
    runtime.reset()
    runtime.set_precision(precision)
    try:
        cfg = cm.get_cfg("centermask_V_39_eSE_FPN.yaml", ["MODEL.VOVNET.CONV_BODY", "V-19-eSE"])
        sd = synthetic_state_dict(cfg, seed=9)
        shape = ShapeSpec(channels=256, width=14, height=14)
        mh, ih = build_mask_head(cfg, shape), build_maskiou_head(cfg, shape)
        assert isinstance(mh, SpatialAttentionMaskHead) and isinstance(ih, MaskIoUHead)
        mh.load_state_dict({k[len("roi_heads.mask_head."):]: v for k, v in sd.items() if k.startswith("roi_heads.mask_head.")}, strict=True)
        ih.load_state_dict({k[len("roi_heads.maskiou_head."):]: v for k, v in sd.items() if k.startswith("roi_heads.maskiou_head.")},
                           strict=True)
        g = torch.Generator().manual_seed(3)
        x = torch.relu(torch.randn(37, 256, 14, 14, generator=g))
        mask = torch.rand(37, 1, 28, 28, generator=g)
        bf = precision == "bf16"
        with restate.bf16_sim(bf):
            ref_logits, _ = restate.mask_head_forward(restate._q(x), sd, cfg)
            ref_iou = restate.maskiou_head_forward(restate._q(x), mask, sd, cfg)
        got_logits = mh(x.cuda()).float().cpu()
        got_iou = ih(x.cuda(), mask.cuda()).float().cpu()
        assert got_logits.shape == ref_logits.shape == (37, 80, 28, 28) and got_iou.shape == ref_iou.shape == (37, 80)
        tol = 3e-2 if bf else 1e-4
        for got, ref, what in ((got_logits, ref_logits, "mask logits"), (got_iou, ref_iou, "maskiou")):
            err = ((got - ref).abs().max() / ref.abs().max()).item()
            print("{} [{}]: max err / max |ref| = {:.2e}".format(what, precision, err))
            assert err <= tol, (what, err)
        model = cm.build_model(cfg)
        assert type(model.roi_heads.mask_head) is SpatialAttentionMaskHead and type(model.roi_heads.maskiou_head) is MaskIoUHead
    finally:
        runtime.reset()
        runtime.set_precision("fp32")
