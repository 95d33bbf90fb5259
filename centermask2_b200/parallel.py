"""Image data-parallel plumbing: one process per GPU, contiguous image shards, weights
replicated, NO collective on the data path -- only a final gather of the result records.

The record of one detection slot (box + score + class + location + mask):

    field   0..3   post-processed box x0, y0, x1, y1       4      score        5   class
            6      mask score (score * MaskIoU)            7, 8   location x, y (fcos_outputs.py:458-462)
            9      valid (1 = a detection that survived detector_postprocess)     10  detections of the image

plus the mask of every slot as COCO run lengths (``BatchResult``): the pasted full-resolution bool masks never leave the
device, their column-major run lengths (what ``instances_to_coco_json`` makes of them, coco_evaluation.py:388-391) do.

Nothing here launches kernels except ``pack_slots`` (one ``libcm2`` launch), so the gather logic runs under ``gloo``
on CPU (tests) and ``nccl`` on GPUs alike.
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_FIELDS = 11
F_BOX, F_SCORE, F_CLASS, F_MASK_SCORE, F_LOC, F_VALID, F_COUNT = 0, 4, 5, 6, 7, 9, 10


def shard_range(n_items, rank, world):
    """Contiguous block of ``range(n_items)`` owned by ``rank``: sizes differ by at most one, earlier ranks get
    the larger blocks (rank r owns images [r*B/G, (r+1)*B/G) when G divides B)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def pack_slots(rec, boxes, scores, classes, mask_scores, locations, valid, count):
    """Device side: fixed-size detection buffers of the engine -> ``rec`` float32 [n, r_cap, RECORD_FIELDS] in one launch
    (no host sync; slots beyond an image's count come out as zeros apart from the count field)."""
    from . import lib
    lib.pack_records(boxes, scores, classes, mask_scores, locations, valid, count, rec)
    return rec


def pack_records(instances_list, r_cap, device=None):
    """list[Instances] -> float32 [len, r_cap, RECORD_FIELDS] (zero padded), the same layout ``pack_slots`` writes.

    Host-side form for callers that hold ``Instances`` (``GeneralizedRCNN.forward``); written for few framework calls:
    one padded batch per field (``pad_sequence`` walks the list in C++) and one strided assignment each."""
    n = len(instances_list)
    if n == 0:
        return torch.zeros((0, r_cap, RECORD_FIELDS), dtype=torch.float32, device=device)
    dev = device if device is not None else instances_list[0].scores.device
    counts = [min(len(inst), r_cap) for inst in instances_list]
    rec = torch.zeros((n, r_cap, RECORD_FIELDS), dtype=torch.float32, device=dev)
    rec[:, :, F_COUNT] = torch.tensor(counts, dtype=torch.float32).to(dev, non_blocking=True).view(n, 1)
    m = max(counts)
    if m:
        def padded(tensors):
            return torch.nn.utils.rnn.pad_sequence(tensors, batch_first=True).to(dev)      # [n, m, ...], zero filled

        rec[:, :m, F_BOX:F_BOX + 4] = padded([inst.pred_boxes.tensor[:k] for inst, k in zip(instances_list, counts)])
        rec[:, :m, F_SCORE] = padded([inst.scores[:k] for inst, k in zip(instances_list, counts)])
        rec[:, :m, F_CLASS] = padded([inst.pred_classes[:k] for inst, k in zip(instances_list, counts)]).to(torch.float32)
        rec[:, :m, F_VALID] = padded([inst.scores.new_ones((k,)) for inst, k in zip(instances_list, counts)])
        if all(inst.has("locations") for inst in instances_list):
            rec[:, :m, F_LOC:F_LOC + 2] = padded([inst.locations[:k] for inst, k in zip(instances_list, counts)])
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
        rec[:, :ms.shape[1], F_MASK_SCORE] = ms
    return rec


class BatchResult(object):
    """Host-side result of one batch: ``records`` float32 [n, r_cap, RECORD_FIELDS], and the masks of all n * r_cap slots
    as COCO run lengths: ``rle_runs`` uint32-valued int32 [total runs], slot s = image * r_cap + k owns
    ``rle_runs[rle_offsets[s]:rle_offsets[s + 1]]`` (column-major, the first run counts zeros; an empty slot is one run
    of h * w zeros).  When produced by ``GeneralizedRCNN.inference_records`` the tensors are views of pinned staging
    buffers that are reused a few batches later: ``clone()`` to keep them."""

    __slots__ = ("records", "rle_offsets", "rle_runs", "size")

    def __init__(self, records, rle_offsets, rle_runs, size):
        self.records, self.rle_offsets, self.rle_runs, self.size = records, rle_offsets, rle_runs, tuple(size)

    @property
    def counts(self):
        return self.records[:, 0, F_COUNT].to(torch.int64).tolist()

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.records, self.rle_offsets, self.rle_runs))

    def clone(self):
        return BatchResult(self.records.clone(), self.rle_offsets.clone(), self.rle_runs.clone(), self.size)

    def runs(self, image, k):
        """uint32 numpy run lengths of detection ``k`` of image ``image`` (pycocotools ``rleEncode`` counts)."""
        s = image * self.records.shape[1] + k
        a, b = int(self.rle_offsets[s]), int(self.rle_offsets[s + 1])
        return self.rle_runs[a:b].numpy().view(np.uint32)

    def mask(self, image, k):
        """Decode the mask of a slot back to a bool [h, w] array (tests / spot checks)."""
        h, w = self.size
        runs = self.runs(image, k).astype(np.int64)
        vals = (np.arange(runs.size) & 1).astype(bool)
        return np.repeat(vals, runs).reshape(w, h).T


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


def gather_results(local, n_items, group=None, device=None):
    """The final result gather of SURVEY 8e for full results: ``local`` is this rank's ``BatchResult`` (its shard of
    ``n_items`` images); returns the ``BatchResult`` of all images in global order on every rank.  Records are fixed
    size; the run lengths are ragged, so their per-slot lengths travel first and the runs are padded to the longest
    shard.  ``device``: where the collective runs (the GPU for nccl; defaults to the records' device)."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else local.records.device
    r_cap = local.records.shape[1]
    records = gather_records(local.records.to(dev), n_items, group)
    lens = (local.rle_offsets[1:] - local.rle_offsets[:-1]).to(torch.int64).reshape(-1, r_cap, 1).to(dev)
    lens = gather_records(lens, n_items, group).reshape(-1)                          # runs per slot, global order
    sizes = [len(shard_range(n_items, r, world)) * r_cap for r in range(world)]
    bounds = np.cumsum([0] + sizes)
    totals = [int(lens[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    assert totals[rank] == local.rle_runs.numel(), (totals, local.rle_runs.numel())
    pad = torch.zeros((max(max(totals), 1),), dtype=local.rle_runs.dtype, device=dev)
    pad[:totals[rank]] = local.rle_runs.to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    runs = torch.cat([o[:t] for o, t in zip(out, totals)])
    offsets = torch.zeros((lens.numel() + 1,), dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(lens, 0)
    return BatchResult(records.cpu(), offsets.cpu(), runs.cpu(), local.size)
