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

    This runs on the host once per step in the end-to-end loop, so it is written for few framework calls: one padded
    batch per field (``pad_sequence`` walks the list in C++) and one strided assignment each -- about a dozen ops for
    the whole list instead of several per image."""
    n = len(instances_list)
    if n == 0:
        return torch.zeros((0, r_cap, RECORD_FIELDS), dtype=torch.float32, device=device)
    dev = device if device is not None else instances_list[0].scores.device
    counts = [min(len(inst), r_cap) for inst in instances_list]
    rec = torch.zeros((n, r_cap, RECORD_FIELDS), dtype=torch.float32, device=dev)
    rec[:, :, 7] = torch.tensor(counts, dtype=torch.float32).to(dev, non_blocking=True).view(n, 1)
    m = max(counts)
    if m:
        def padded(tensors):
            return torch.nn.utils.rnn.pad_sequence(tensors, batch_first=True).to(dev)      # [n, m, ...], zero filled

        rec[:, :m, :4] = padded([inst.pred_boxes.tensor[:k] for inst, k in zip(instances_list, counts)])
        rec[:, :m, 4] = padded([inst.scores[:k] for inst, k in zip(instances_list, counts)])
        rec[:, :m, 5] = padded([inst.pred_classes[:k] for inst, k in zip(instances_list, counts)]).to(torch.float32)
        empty = None
        ms = []
        for inst, k in zip(instances_list, counts):
            if inst.has("mask_scores"):
                ms.append(inst.mask_scores[:k])
            else:                                   # center_heads.py:511-513: no field on a batch without detections
                if empty is None:
                    empty = instances_list[0].scores.new_zeros((0,))
                ms.append(empty.to(inst.scores.device))
        ms = padded(ms)                             # [n, longest present]; shorter than m when fields are missing
        rec[:, :ms.shape[1], 6] = ms
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
