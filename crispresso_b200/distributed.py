"""Multi-GPU plumbing: reads shard across ranks, one process per GPU; every alignment
is independent, so the only exchange is a SUM all-reduce of the int64 accumulator block
(`Reductions.flat()`, ~80 KB) -- NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous, equal-count read range [lo, hi) of `rank`."""
    return rank * n // world, (rank + 1) * n // world


def shard_reads(buf, offsets, rank, world):
    """The (buffer, offsets) pair of this rank's shard (views where possible)."""
    n = len(offsets) - 1
    lo, hi = shard_range(n, rank, world)
    return buf[offsets[lo]:offsets[hi]], offsets[lo:hi + 1] - offsets[lo]


def allreduce_reductions(red, device=None):
    """In-place SUM over all ranks of a hotpath.Reductions (torch.distributed must be initialised).
    device: torch device for the staging tensor (cuda for NCCL, None/cpu for gloo)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return red
    t = torch.from_numpy(red.flat())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t)
    red.load_flat(t.cpu().numpy())
    return red


def run_chunked(ctx, amplicon, reads, chunk_reads=1 << 22, **kw):
    """run_hot_path over read chunks (for read sets larger than one call should hold in HBM),
    accumulating into one Reductions.  Returns the Reductions; per-read outputs are not kept."""
    from . import hotpath
    buf, offsets = reads
    n = len(offsets) - 1
    red = kw.pop("red", None) or hotpath.Reductions(len(amplicon))
    for lo in range(0, n, chunk_reads):
        hi = min(n, lo + chunk_reads)
        hotpath.run_hot_path(ctx, amplicon, (buf[offsets[lo]:offsets[hi]], (offsets[lo:hi + 1] - offsets[lo])), red=red, **kw)
    return red
