"""Per-image sharding of the ROI hot path across the GPUs of one box.

The reference's only parallelism is batch-axis tower splitting (core/parallel_model.py:46-67): images are split
across GPUs, every op instance is per image (NMS runs inside utils.batch_slice, core/models.py:487-490; ROIs carry a
box_index that never crosses images, core/models.py:657).  So the path shards with NO data-path collective: image i
goes to rank i mod G together with its feature-map slices, boxes and grads; results are concatenated by the caller.
The only communication is the benchmark's barrier and max-over-ranks timing.
"""
import numpy as np


def images_of_rank(batch, world, rank):
    """Image ids owned by `rank` (image i -> rank i mod world)."""
    return [i for i in range(batch) if i % world == rank]


def shard_rois(boxes, box_index, batch, world, rank):
    """ROIs whose image lives on `rank`.

    Returns (boxes_local [n,6], box_index_local [n] into the rank's own image order, positions [n] in the input)."""
    mine = images_of_rank(batch, world, rank)
    remap = -np.ones(batch, np.int64)
    remap[mine] = np.arange(len(mine))
    box_index = np.asarray(box_index)
    pos = np.nonzero(remap[box_index] >= 0)[0]
    return np.ascontiguousarray(np.asarray(boxes)[pos]), remap[box_index[pos]].astype(np.int32), pos


def shard_volume(volume, world, rank):
    """The rank's images of a [B,...] volume, in local order."""
    return np.ascontiguousarray(np.asarray(volume)[images_of_rank(volume.shape[0], world, rank)])


def merge_by_position(parts, positions, n_total):
    """Inverse of shard_rois for per-ROI results: parts[r][k] belongs at positions[r][k]."""
    first = next(p for p in parts if p is not None and len(p))
    out = np.zeros((n_total,) + first.shape[1:], first.dtype)
    for part, pos in zip(parts, positions):
        if part is not None and len(pos):
            out[pos] = part
    return out


def max_over_ranks(value):
    """Max of a python float over all ranks (device timing rule: report the slowest rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_over_ranks(value):
    """The python float of every rank, in rank order (diagnostics: which GPU was the slowest)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x[0]) for x in out]
