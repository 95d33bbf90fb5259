"""CPU, world_size 2, gloo: the N>1 host logic (contiguous image shards + final result-record gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centermask2_b200 import parallel
from centermask2_b200.modeling.compat import Boxes, Instances


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_instances(image_index, r_cap):
    g = torch.Generator().manual_seed(1000 + image_index)
    k = int(torch.randint(0, r_cap + 1, (1,), generator=g))
    inst = Instances((64, 64))
    inst.pred_boxes = Boxes(torch.rand(k, 4, generator=g) * 64)
    inst.scores = torch.rand(k, generator=g)
    inst.pred_classes = torch.randint(0, 80, (k,), generator=g)
    inst.locations = torch.rand(k, 2, generator=g) * 64
    if k:
        inst.mask_scores = torch.rand(k, generator=g)
    return inst


def _fake_result(images, r_cap, h=6, w=5):
    """BatchResult of a shard with random masks encoded by the oracle's column-major RLE (run lengths only)."""
    import numpy as np
    from oracle import rle as orle
    rec = parallel.pack_records([_fake_instances(i, r_cap) for i in images], r_cap, device="cpu")
    runs, offs = [], [0]
    for i in images:
        g = torch.Generator().manual_seed(5000 + i)
        for k in range(r_cap):
            m = (torch.rand(h, w, generator=g) > 0.5).numpy()
            c = orle.rle_encode(m)
            runs.append(np.asarray(c, dtype=np.uint32))
            offs.append(offs[-1] + len(c))
    flat = np.concatenate(runs) if runs else np.zeros(0, dtype=np.uint32)
    return parallel.BatchResult(rec, torch.tensor(offs, dtype=torch.int64), torch.from_numpy(flat.astype(np.int32)), (h, w))


def _worker(rank, world, port, n_items, r_cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_range(n_items, rank, world)
        local = parallel.pack_records([_fake_instances(i, r_cap) for i in mine], r_cap, device="cpu")
        full = parallel.gather_records(local, n_items)
        res = parallel.gather_results(_fake_result(list(mine), r_cap), n_items)
        q.put((rank, list(mine), full, (res.records, res.rle_offsets, res.rle_runs)))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(parallel.shard_range(n, r, world))
            assert seen == list(range(n))
            sizes = [len(parallel.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_gather_records_world2_gloo():
    world, n_items, r_cap = 2, 5, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, r_cap, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = parallel.pack_records([_fake_instances(i, r_cap) for i in range(n_items)], r_cap, device="cpu")
    shards = sorted((r, m) for r, m, _, _ in results)
    assert shards[0][1] == [0, 1, 2] and shards[1][1] == [3, 4]
    whole = _fake_result(list(range(n_items)), r_cap)
    for _, _, full, (rec, offs, runs) in results:
        assert full.shape == (n_items, r_cap, parallel.RECORD_FIELDS)
        assert torch.equal(full, expect)
        # full results (records + ragged run lengths) reassembled in global image order
        assert torch.equal(rec, whole.records) and torch.equal(offs, whole.rle_offsets) and torch.equal(runs, whole.rle_runs)
    got = parallel.BatchResult(*results[0][3], (6, 5))
    import numpy as np
    g = torch.Generator().manual_seed(5000 + 3)
    want = [(torch.rand(6, 5, generator=g) > 0.5).numpy() for _ in range(r_cap)]
    assert np.array_equal(got.mask(3, 2), want[2])                     # decode of a gathered slot == the mask that was encoded


def test_gather_is_identity_without_process_group():
    local = torch.arange(2 * 3 * parallel.RECORD_FIELDS, dtype=torch.float32).reshape(2, 3, -1)
    assert parallel.gather_records(local, 2) is local


def test_pack_records_layout_against_a_plain_loop():
    """Record layout (x0, y0, x1, y1, score, class, mask_score, loc x, loc y, valid, count), zero padding, truncation at
    r_cap, images without detections and without a mask_scores field (center_heads.py:511-513), written out field by field."""
    r_cap = 6
    insts = [_fake_instances(i, 9) for i in range(7)]                 # up to 9 detections: some are truncated at r_cap
    empty = Instances((64, 64))
    empty.pred_boxes = Boxes(torch.zeros((0, 4)))
    empty.scores = torch.zeros((0,))
    empty.pred_classes = torch.zeros((0,), dtype=torch.int64)
    empty.locations = torch.zeros((0, 2))
    insts.insert(2, empty)
    no_ms = _fake_instances(50, 5)
    no_ms.remove("mask_scores") if no_ms.has("mask_scores") else None
    insts.append(no_ms)
    got = parallel.pack_records(insts, r_cap, device="cpu")
    want = torch.zeros((len(insts), r_cap, parallel.RECORD_FIELDS))
    for i, inst in enumerate(insts):
        k = min(len(inst), r_cap)
        want[i, :, 10] = k
        want[i, :k, :4] = inst.pred_boxes.tensor[:k]
        want[i, :k, 4] = inst.scores[:k]
        want[i, :k, 5] = inst.pred_classes[:k].float()
        want[i, :k, 7:9] = inst.locations[:k]
        want[i, :k, 9] = 1.0
        if inst.has("mask_scores"):
            want[i, :k, 6] = inst.mask_scores[:k]
    assert torch.equal(got, want)
    assert parallel.pack_records([empty, empty], r_cap, device="cpu").abs().sum().item() == 0
