"""Input-side resize: bit-exact against Pillow, the library detectron2's ResizeTransform calls
(/root/reference/deploy_utils.py:60-73).  CPU: the oracle restatement and the product's coefficient tables against
Pillow; GPU: cm2_resize_pil_u8 through centermask2_b200.transforms against Pillow."""
import numpy as np
import pytest
import torch

from centermask2_b200 import transforms
from oracle import resize as oracle_resize

CASES = [(480, 640), (1080, 1920), (37, 53), (300, 200), (800, 1333), (500, 375), (1200, 1600), (64, 2000)]


def _img(h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w", CASES)
def test_output_shape_rule(h, w):
    assert transforms.shortest_edge_output_shape(h, w, 800, 1333) == oracle_resize.d2_output_shape(h, w, 800, 1333)
    oh, ow = transforms.shortest_edge_output_shape(h, w, 800, 1333)
    assert max(oh, ow) <= 1333 and (min(oh, ow) == 800 or max(oh, ow) == 1333)


@pytest.mark.parametrize("h,w,oh,ow", [(48, 64, 80, 107), (108, 192, 75, 133), (37, 53, 80, 115), (30, 20, 12, 8), (50, 37, 107, 80)])
def test_oracle_restatement_matches_pillow(h, w, oh, ow):
    img = _img(h, w, h * w)
    assert np.array_equal(oracle_resize.restated_resize(img, oh, ow), oracle_resize.pil_resize(img, oh, ow))


@pytest.mark.parametrize("n_in,n_out", [(640, 1067), (1920, 1333), (53, 115), (200, 80), (64, 64), (7, 1000), (1000, 7)])
def test_product_coefficient_tables_match_oracle(n_in, n_out):
    bounds, kk = transforms.pil_bilinear_coeffs(n_in, n_out)
    rb, rk = oracle_resize._coeffs(n_in, n_out)
    assert bounds.tolist() == [list(b) for b in rb]
    assert kk.tolist() == rk


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", CASES)
def test_gpu_resize_shortest_edge_bit_exact_vs_pillow(h, w):
    img = _img(h, w, h + w)
    tf = transforms.ResizeShortestEdge([800, 800], 1333).get_transform(img)
    oh, ow = oracle_resize.d2_output_shape(h, w)
    assert (tf.new_h, tf.new_w) == (oh, ow)
    ref = oracle_resize.pil_resize(img, oh, ow)
    got = tf.apply_image(img)
    assert got.is_cuda and got.dtype == torch.uint8
    assert np.array_equal(got.cpu().numpy(), ref)
    chw = tf.apply_image(torch.from_numpy(img).cuda(), chw=True)
    assert np.array_equal(chw.cpu().numpy(), ref.transpose(2, 0, 1))
    sample = transforms.get_sample_inputs(img)
    assert sample[0]["height"] == h and sample[0]["width"] == w and tuple(sample[0]["image"].shape) == (3, oh, ow)


@pytest.mark.gpu
def test_gpu_resize_downscale_and_single_channel():
    img = _img(301, 203, 5)
    for oh, ow in ((120, 81), (301, 100), (77, 203)):
        got = transforms.ResizeTransform(301, 203, oh, ow).apply_image(img).cpu().numpy()
        assert np.array_equal(got, oracle_resize.pil_resize(img, oh, ow))
