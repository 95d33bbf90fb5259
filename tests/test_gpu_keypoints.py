"""GPU: the keypoint branch -- cm2_keypoints_decode against the oracle's heatmaps_to_keypoints,
the ConvTranspose2d(4, 2, 1)-as-phase-convolution on both conv engines, and the whole branch in bf16 on the
tensor-core engine against the bf16-rounding oracle.  The fp32 end-to-end parity against the unmodified reference is
the golden case ``v19_keypoints`` in test_gpu_model.py.

Locations are arg-max positions of a bicubic-resized map (see helpers.assert_keypoints_match): they are compared
exactly up to a stated fraction of near-tie moves; logits within 1e-4, scores within 2e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from centermask2_b200 import lib, packing, runtime                                  # noqa: E402
from centermask2_b200.engine import FMap                                            # noqa: E402
from oracle import restate                                                          # noqa: E402
from tests.helpers import pack_lowres, keypoint_boxes, load_golden, build_case      # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("variant", [1, 0])          # 1: column walk (default); 0: flat per-pixel loop
@pytest.mark.parametrize("res,k,n,r_cap,counts", [(14, 17, 2, 8, (8, 5)), (7, 3, 3, 4, (4, 0, 1)), (14, 17, 1, 3, (3,))])
def test_keypoints_decode_matches_oracle(res, k, n, r_cap, counts, variant, monkeypatch):
    monkeypatch.setenv("CM2_KP_VARIANT", str(variant))
    g = torch.Generator().manual_seed(7 * res + n)
    r = n * r_cap
    low = torch.randn(r, k, 2 * res, 2 * res, generator=g) * 2.5
    boxes = keypoint_boxes(g, r)
    if r_cap == 3:                                   # a whole-image ROI, a sub-pixel one (clamped to 1 px), a 1-px-wide strip
        boxes[0] = torch.tensor([-3.5, 2.25, 1329.5, 802.0])
        boxes[1] = torch.tensor([10.2, 11.7, 10.5, 11.9])
        boxes[2] = torch.tensor([40.0, 5.0, 40.75, 300.0])
    out = torch.full((n, r_cap, k, 4), float("nan"), device=DEV)
    lib.keypoints_decode(pack_lowres(low).to(DEV), boxes.view(n, r_cap, 4).to(DEV),
                         torch.tensor(counts, dtype=torch.int32, device=DEV), n, r_cap, res, k, out)
    torch.cuda.synchronize()
    got = out.cpu().view(r, k, 4)
    hi = F.interpolate(low, scale_factor=2, mode="bilinear", align_corners=False)          # keypoint_head.py:221
    ref = restate.heatmaps_to_keypoints(hi, boxes)
    valid = torch.cat([torch.arange(r_cap) < c for c in counts])
    assert torch.equal(got[~valid], torch.zeros_like(got[~valid]))                          # empty slots: zeros
    got, ref = got[valid], ref[valid]
    moved = ((got[..., :2] - ref[..., :2]).abs() > 1e-3).any(-1)
    assert moved.float().mean().item() <= 0.01, (moved.nonzero().tolist(), got[moved], ref[moved])
    assert torch.allclose(got[..., 2], ref[..., 2], rtol=1e-5, atol=1e-4)
    assert torch.allclose(got[..., 3], ref[..., 3], rtol=2e-3)
    assert torch.equal(got[~moved][..., :2], ref[~moved][..., :2])                          # same pixel -> same fp32 expression
    if variant == 1:                                  # both work splits evaluate the same expression tree per pixel
        monkeypatch.setenv("CM2_KP_VARIANT", "0")
        flat = torch.full_like(out, float("nan"))
        lib.keypoints_decode(pack_lowres(low).to(DEV), boxes.view(n, r_cap, 4).to(DEV),
                             torch.tensor(counts, dtype=torch.int32, device=DEV), n, r_cap, res, k, flat)
        torch.cuda.synchronize()
        assert torch.equal(out, flat)


def test_keypoints_decode_rejects_bad_arguments():
    out = torch.zeros((1, 1, 1, 4), device=DEV)
    low = torch.zeros((1, 25, 25, 4), device=DEV)
    with pytest.raises(RuntimeError):
        lib.keypoints_decode(low, torch.zeros((1, 1, 4), device=DEV), torch.ones(1, dtype=torch.int32, device=DEV), 1, 1, 25, 1, out)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_score_lowres_phase_conv_on_both_engines(precision):
    """ConvTranspose2d(k 4, s 2, p 1) + bias through Engine.conv (SIMT fp32 / tcgen05 bf16) in the phase layout."""
    runtime.reset()
    runtime.set_precision(precision)
    try:
        from centermask2_b200.config import get_cfg
        eng = runtime.engine_for(get_cfg("centermask_V_39_eSE_FPN.yaml", []))
        g = torch.Generator().manual_seed(5)
        cin, k, r, res = 128, 17, 6, 14
        sd = {"d.weight": torch.randn(cin, k, 4, 4, generator=g) * 0.05, "d.bias": torch.randn(k, generator=g)}
        x = torch.randn(r, cin, res, res, generator=g)
        if precision == "bf16":
            x = x.bfloat16().float()
            sd["d.weight"] = sd["d.weight"].bfloat16().float()
        w = packing.deconv4x4s2(sd, "d", eng.dtype, eng.device, eng.tc)
        buf = torch.zeros((r, res + 2, res + 2, cin), dtype=eng.dtype, device=DEV)
        buf[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1).to(DEV, eng.dtype)
        low = eng.conv("t_score_lowres", [FMap(buf, 1)], w, out_dtype=torch.float32, out_halo=0)
        torch.cuda.synchronize()
        ref = pack_lowres(F.conv_transpose2d(x, sd["d.weight"], sd["d.bias"], stride=2, padding=1))
        got = low.buf.cpu().view(r, res, res, 4, k)
        assert low.buf.is_contiguous() and low.buf.dtype == torch.float32
        assert torch.allclose(got, ref, rtol=1e-4, atol=2e-4 if precision == "fp32" else 2e-3)
    finally:
        runtime.reset()
        runtime.set_precision("fp32")


def test_keypoint_branch_bf16_against_bf16_rounding_oracle():
    """The whole branch (ROIAlign -> conv tower -> score_lowres -> decode) on the tensor-core engine, on the golden
    case's detections, against the oracle rounding to bf16 where the engine does.  bf16 feature noise moves a share of the
    arg-max positions, so the gate is: >= 70 % of the keypoints within 1.5 px and their scores within 10 %."""
    import centermask2_b200 as cm
    runtime.reset()
    runtime.set_precision("bf16")
    try:
        gold = load_golden("v19_keypoints")
        cfg, sd, inputs = build_case("v19_keypoints", gold)
        tr = {}
        with restate.bf16_sim():
            restate.run_model(inputs, sd, cfg, postprocess=False, trace=tr)
            feats = {k: v.bfloat16().float() for k, v in tr["features"].items()}    # what the engine stores
            # the oracle's own detections as given boxes, so that only the keypoint branch is compared
            dets = [dict(image_size=r["image_size"], pred_boxes=r["pred_boxes"], pred_classes=r["pred_classes"],
                         scores=r["scores"]) for r in gold["raw"]]
            ref = restate.keypoints_forward(feats, dets, sd, cfg)
        model = cm.build_model(cfg)
        model.load_state_dict(sd)
        eng = runtime.engine_for(cfg)
        from centermask2_b200.modeling.compat import Boxes, Instances
        insts = []
        for r in gold["raw"]:
            i = Instances(tuple(r["image_size"]))
            i.pred_boxes = Boxes(r["pred_boxes"].to(DEV))
            i.pred_classes = r["pred_classes"].to(DEV)
            i.scores = r["scores"].to(DEV)
            insts.append(i)
        fm = {k: v.to(DEV) for k, v in feats.items()}
        out = model.roi_heads.forward_with_given_boxes(fm, insts)
        torch.cuda.synchronize()
        for o, r in zip(out, ref):
            got, want = o.pred_keypoints.cpu(), r["pred_keypoints"]
            assert got.shape == want.shape
            near = ((got[..., :2] - want[..., :2]).abs() <= 1.5).all(-1)
            print("bf16 keypoints within 1.5 px: {}/{}".format(int(near.sum()), near.numel()))
            assert near.float().mean().item() >= 0.70
            rel = ((got[..., 2] - want[..., 2]).abs() / want[..., 2].abs().clamp(min=1e-6))[near]
            assert rel.max().item() <= 0.10, rel.max().item()
    finally:
        runtime.reset()
        runtime.set_precision("fp32")
