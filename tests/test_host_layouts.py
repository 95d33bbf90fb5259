"""CPU: host-side layout logic of the engine that the GPU tests only see through results -- the shared-halo ROI map
(engine.SharedHaloFMap, csrc/conv_tc.cu halo_kind 2), the split-K slice heuristic and the input-buffer batching."""
import torch

from centermask2_b200.engine import Engine, FMap, SharedHaloFMap


def test_shared_halo_map_geometry_on_cpu_tensors():
    """Line pitch w + 1, image pitch (h + 1)(w + 1): every interior pixel maps to a distinct flat row, the frame rows are
    exactly the remaining ones, and the four neighbours of an interior pixel that leave the image land on frame rows."""
    n, h, w, c = 3, 14, 14, 8
    rows = SharedHaloFMap.rows(n, h, w)
    assert rows == n * 225 + 16
    flat = torch.zeros((rows, c))
    m = SharedHaloFMap(flat, n, h, w)
    assert (m.n, m.h, m.w, m.c, m.halo) == (n, h, w, c, 1)
    v = m.view
    assert tuple(v.shape) == (n, h, w, c) and v.stride() == (225 * c, 15 * c, c, 1)
    v.copy_(torch.arange(1, n * h * w + 1, dtype=torch.float32).view(n, h, w, 1).expand(n, h, w, c))
    used = flat[:, 0] != 0
    assert int(used.sum()) == n * h * w                                        # no two pixels share a row
    pitch, plane = w + 1, (h + 1) * (w + 1)
    row_of = lambda b, y, x: b * plane + (y + 1) * pitch + (x + 1)             # noqa: E731  (flat row of pixel (b, y, x); -1 = frame)
    for b in range(n):
        for y in range(h):
            for x in range(w):
                assert flat[row_of(b, y, x), 0] == v[b, y, x, 0]
    # a 3x3 tap is the flat row shift (ky - 1) * pitch + (kx - 1): neighbours outside the image are frame (zero) rows --
    # including the pixel right of a line (= left frame pixel of the next line) and the line under an image (= top frame line
    # of the next image)
    for b, y, x in ((0, 0, 0), (0, 0, w - 1), (1, h - 1, 0), (2, h - 1, w - 1), (1, 5, w - 1)):
        for ky in range(3):
            for kx in range(3):
                r = row_of(b, y, x) + (ky - 1) * pitch + (kx - 1)
                yy, xx = y + ky - 1, x + kx - 1
                inside = 0 <= yy < h and 0 <= xx < w
                assert 0 <= r < rows
                assert bool(used[r]) == inside
                if inside:
                    assert flat[r, 0] == v[b, yy, xx, 0]


def test_split_k_slice_heuristic():
    """Engine._splitk_slices: K slices only for layers whose output tiles fill at most half of the 148 SMs and whose K loop is
    long enough; as many slices as fit, at least four 64-channel blocks each."""
    class W(object):
        def __init__(self, k, src_c, cout):
            self.k, self.src_c, self.cout = k, src_c, cout

    def fmap(n, h, w, c, halo=1):
        return FMap(torch.zeros((n, h + 2 * halo, w + 2 * halo, c)), halo)

    f = Engine._splitk_slices
    assert f(fmap(800, 1, 1, 12544, halo=0), 1, 1, W(1, [12544], 1024), False) == 5      # iou_fc1: 7 x 4 tiles, 196 K-blocks
    assert f(fmap(800, 1, 1, 1024, halo=0), 1, 1, W(1, [1024], 1024), False) == 0        # iou_fc2: K loop too short
    assert f(fmap(16, 13, 21, 256), 13, 21, W(3, [256], 256), True) == 3                 # P6 at batch 16: 44 tiles, 36 K-blocks
    assert f(fmap(16, 25, 42, 224), 25, 42, W(3, [224], 224), False) == 0                # OSA5 at batch 16 fills the SMs
    assert f(fmap(2, 25, 42, 224), 25, 42, W(3, [224], 224), False) == 7                 # ... at 2 images it does not
    assert f(fmap(16, 200, 336, 128), 200, 336, W(3, [128], 128), False) == 0


def test_image_buffers_batch_equal_shapes_into_one_tensor():
    eng = Engine(None, "fp32_simt", "cpu")
    sig = tuple(((3, 8, 12), torch.uint8) for _ in range(4))
    imgs, whole = eng.image_buffers("t_in", sig)
    assert whole is not None and tuple(whole.shape) == (4, 3, 8, 12)
    assert all(im.data_ptr() == whole[i].data_ptr() and im.is_contiguous() for i, im in enumerate(imgs))
    again, whole2 = eng.image_buffers("t_in", sig)
    assert whole2.data_ptr() == whole.data_ptr()                               # static addresses: a captured graph reads them
    mixed = (((3, 8, 12), torch.uint8), ((3, 8, 10), torch.uint8))
    imgs, whole = eng.image_buffers("t_in", mixed)
    assert whole is None and [tuple(i.shape) for i in imgs] == [(3, 8, 12), (3, 8, 10)]
