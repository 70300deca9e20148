"""Batch-sharded data parallelism for inference: one process per GPU, images partitioned
contiguously over ranks, weights replicated, NO collective on the data path.  torch.distributed (NCCL on GPUs,
gloo in the CPU tests) is only plumbing: barriers, max-over-ranks of timings, and an optional host-side gather
of the per-rank detection lists when a caller wants one merged result.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def image_shard(num_images, rank=None, world_size=None):
    """Contiguous slice of the batch owned by ``rank``: [r*B/W, (r+1)*B/W) with the remainder spread over the
    first ranks, so every image is processed exactly once."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(num_images, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (multi-GPU timings are reported as the slowest rank)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(local_preds):
    """Merge per-rank lists of [M,6] detections (CPU tensors) into one list in global image order."""
    rank, ws = world()
    if ws == 1:
        return list(local_preds)
    buckets = [None] * ws
    dist.all_gather_object(buckets, [p.cpu() for p in local_preds])
    return [p for b in buckets for p in b]
