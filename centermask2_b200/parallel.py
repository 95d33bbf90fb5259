"""Image data-parallel plumbing: one process per GPU, contiguous image shards, weights
replicated, NO collective on the data path -- only a final gather of fixed-size result records.

Nothing here touches CUDA directly, so the same code runs under ``gloo`` on CPU (tests) and ``nccl`` on GPUs.
"""
import torch
import torch.distributed as dist

RECORD_FIELDS = 8          # x0, y0, x1, y1, score, class, mask_score, count


def shard_range(n_items, rank, world):
    """Contiguous block of ``range(n_items)`` owned by ``rank``: sizes differ by at most one, earlier ranks get
    the larger blocks (rank r owns images [r*B/G, (r+1)*B/G) when G divides B)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def pack_records(instances_list, r_cap, device=None):
    """list[Instances] -> float32 [len, r_cap, RECORD_FIELDS] (zero padded; field 7 = number of detections)."""
    rec = []
    for inst in instances_list:
        k = len(inst)
        dev = device if device is not None else inst.scores.device
        t = torch.zeros((r_cap, RECORD_FIELDS), dtype=torch.float32, device=dev)
        if k:
            t[:k, :4] = inst.pred_boxes.tensor
            t[:k, 4] = inst.scores
            t[:k, 5] = inst.pred_classes.to(torch.float32)
            if inst.has("mask_scores"):
                t[:k, 6] = inst.mask_scores
        t[:, 7] = float(k)
        rec.append(t)
    if not rec:
        return torch.zeros((0, r_cap, RECORD_FIELDS), dtype=torch.float32, device=device)
    return torch.stack(rec)


def gather_records(local, n_items, group=None):
    """all_gather the per-rank record blocks (shards from ``shard_range``) into global image order.

    ``local``: [len(shard), r_cap, F].  Returns [n_items, r_cap, F] on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [len(shard_range(n_items, r, world)) for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    cap = max(sizes)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
