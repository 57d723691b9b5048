"""Target-row sharding across GPUs (one process per GPU, torch.distributed plumbing).

The reference has no multi-GPU path; its unit of parallelism is the target instance
(``prange``/one CUDA block per instance: MultiSURF.py:174, SURF.py:139, ReliefF.py:143).
Here the target rows of the distance matrix are split into ``world_size`` contiguous
ranges, every rank scores its own range against all samples, and the partial
per-feature weight sums are combined by ONE allreduce (NCCL over NVLink on GPUs,
gloo in the CPU tests).  There is no other exchange on the data path.
"""
from __future__ import annotations

import numpy as np


def shard_rows(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition of ``range(n)``: rank r gets ``[lo, hi)``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def dist_info() -> tuple[int, int]:
    """(rank, world_size) of the initialised torch.distributed group, else (0, 1)."""
    try:
        import torch.distributed as dist
    except Exception:  # torch absent: single process
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_numpy(partial: np.ndarray) -> np.ndarray:
    """Sum a float64 host vector over all ranks (gloo, or NCCL via a staging tensor)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(partial, np.float64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def score_sharded(n: int, n_kept: int, score_rows, device_buffers: bool):
    """Run ``score_rows(lo, hi, out_device_ptr)`` on this rank's target rows and return
    the allreduced float64 weight sums.

    ``score_rows`` returns a float64 numpy vector when ``out_device_ptr`` is None and
    writes the device buffer otherwise (fastselect_b200._native.Dataset.score)."""
    rank, world = dist_info()
    lo, hi = shard_rows(n, world, rank)
    if world == 1:
        return score_rows(lo, hi, None)
    if device_buffers:
        import torch
        import torch.distributed as dist

        buf = torch.empty(n_kept, dtype=torch.float64, device="cuda")
        score_rows(lo, hi, buf.data_ptr())
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)   # one NCCL allreduce over NVLink per fit
        return buf.cpu().numpy()
    return allreduce_sum_numpy(score_rows(lo, hi, None))
