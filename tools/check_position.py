"""Is a conv layer's output for an image independent of the image's slot in the batch?  (bit-exact check)"""
import math, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centermask2_b200 import lib, packing
from centermask2_b200.engine import FMap
BF = torch.bfloat16
def halo(t, dtype=BF):
    n, c, h, w = t.shape
    buf = torch.zeros((n, h + 2, w + 2, c), dtype=dtype, device="cuda")
    buf[:, 1:-1, 1:-1] = t.permute(0, 2, 3, 1).to("cuda", dtype)
    return FMap(buf, 1)
for (c, h, w, n) in [(128, 200, 336, 4), (160, 100, 168, 4), (192, 50, 84, 16), (256, 100, 168, 4), (128, 100, 168, 4)]:
    g = torch.Generator().manual_seed(c)
    imgs = [torch.randn(1, c, h, w, generator=g) for _ in range(n + 2)]
    wt = torch.randn(c, c, 3, 3, generator=g) / math.sqrt(9 * c)
    cw = packing.ConvW(wt, [c], 1, 1, None, None, True, BF, "cuda", True)
    def run(order):
        x = torch.cat([imgs[i] for i in order], 0)
        out = halo(torch.zeros(len(order), c, h, w))
        assert lib.conv2d([halo(x).view], cw.w_tc, out.view, c, 3, 1, 1, relu=True, engine=lib.ENGINE_TC, probe=True), lib.last_error()
        torch.cuda.synchronize()
        return out.view.clone()
    a = run(list(range(n)))
    order = [2, n, n + 1, 0] + list(range(4, n))
    b = run(order)
    d0 = (a[0].float() - b[3].float()).abs().max().item()
    d2 = (a[2].float() - b[0].float()).abs().max().item()
    print("c={} {}x{} n={}: image0 slot0->3 maxdiff {}  image2 slot2->0 maxdiff {}".format(c, h, w, n, d0, d2))
