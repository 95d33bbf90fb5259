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
    """list[Instances] -> float32 [len, r_cap, RECORD_FIELDS] (zero padded; field 7 = number of detections).
    A handful of batched ops for the whole list (one cat per field, one scatter), not a loop of small kernels."""
    n = len(instances_list)
    if n == 0:
        return torch.zeros((0, r_cap, RECORD_FIELDS), dtype=torch.float32, device=device)
    dev = device if device is not None else instances_list[0].scores.device
    counts = [min(len(inst), r_cap) for inst in instances_list]
    rec = torch.zeros((n, r_cap, RECORD_FIELDS), dtype=torch.float32, device=dev)
    rec[:, :, 7] = torch.tensor(counts, dtype=torch.float32).to(dev, non_blocking=True).view(n, 1)
    total = sum(counts)
    if total:
        rows = torch.cat([torch.arange(k, dtype=torch.int64) + i * r_cap for i, k in enumerate(counts) if k]).to(dev, non_blocking=True)
        live = [(inst, k) for inst, k in zip(instances_list, counts) if k]
        vals = torch.zeros((total, 7), dtype=torch.float32, device=dev)
        vals[:, :4] = torch.cat([inst.pred_boxes.tensor[:k] for inst, k in live]).to(dev)
        vals[:, 4] = torch.cat([inst.scores[:k] for inst, k in live]).to(dev)
        vals[:, 5] = torch.cat([inst.pred_classes[:k] for inst, k in live]).to(dev, torch.float32)
        if all(inst.has("mask_scores") for inst, _ in live):
            vals[:, 6] = torch.cat([inst.mask_scores[:k] for inst, k in live]).to(dev)
        else:
            off = 0
            for inst, k in live:
                if inst.has("mask_scores"):
                    vals[off:off + k, 6] = inst.mask_scores[:k].to(dev)
                off += k
        rec.view(n * r_cap, RECORD_FIELDS)[rows, :7] = vals
    return rec


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
