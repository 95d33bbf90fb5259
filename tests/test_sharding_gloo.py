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
    if k:
        inst.mask_scores = torch.rand(k, generator=g)
    return inst


def _worker(rank, world, port, n_items, r_cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_range(n_items, rank, world)
        local = parallel.pack_records([_fake_instances(i, r_cap) for i in mine], r_cap, device="cpu")
        full = parallel.gather_records(local, n_items)
        q.put((rank, list(mine), full))
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
    shards = sorted((r, m) for r, m, _ in results)
    assert shards[0][1] == [0, 1, 2] and shards[1][1] == [3, 4]
    for _, _, full in results:
        assert full.shape == (n_items, r_cap, parallel.RECORD_FIELDS)
        assert torch.equal(full, expect)


def test_gather_is_identity_without_process_group():
    local = torch.arange(2 * 3 * parallel.RECORD_FIELDS, dtype=torch.float32).reshape(2, 3, -1)
    assert parallel.gather_records(local, 2) is local


def test_pack_records_layout_against_a_plain_loop():
    """Record layout (x0, y0, x1, y1, score, class, mask_score, count), zero padding, truncation at r_cap, images without
    detections and without a mask_scores field (center_heads.py:511-513), written out field by field."""
    r_cap = 6
    insts = [_fake_instances(i, 9) for i in range(7)]                 # up to 9 detections: some are truncated at r_cap
    empty = Instances((64, 64))
    empty.pred_boxes = Boxes(torch.zeros((0, 4)))
    empty.scores = torch.zeros((0,))
    empty.pred_classes = torch.zeros((0,), dtype=torch.int64)
    insts.insert(2, empty)
    no_ms = _fake_instances(50, 5)
    no_ms.remove("mask_scores") if no_ms.has("mask_scores") else None
    insts.append(no_ms)
    got = parallel.pack_records(insts, r_cap, device="cpu")
    want = torch.zeros((len(insts), r_cap, parallel.RECORD_FIELDS))
    for i, inst in enumerate(insts):
        k = min(len(inst), r_cap)
        want[i, :, 7] = k
        want[i, :k, :4] = inst.pred_boxes.tensor[:k]
        want[i, :k, 4] = inst.scores[:k]
        want[i, :k, 5] = inst.pred_classes[:k].float()
        if inst.has("mask_scores"):
            want[i, :k, 6] = inst.mask_scores[:k]
    assert torch.equal(got, want)
    assert parallel.pack_records([empty, empty], r_cap, device="cpu").abs().sum().item() == 0
