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


def shard_rows(n: int, world_size: int, rank: int, align: int = 1) -> tuple[int, int]:
    """Contiguous, balanced partition of ``range(n)``: rank r gets ``[lo, hi)``.  With
    ``align`` > 1 every boundary but the last is rounded down to a multiple of it (the
    multi-GPU symmetric distance kernel wants shard starts that are multiples of 4)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n, world_size)

    def start(r):
        if r >= world_size:
            return n
        return (r * base + min(r, extra)) // align * align

    return start(rank), start(rank + 1)


def shard_starts(n: int, world_size: int, align: int = 1) -> list[int]:
    """``[lo_0, lo_1, ..., lo_{world-1}, n]`` of :func:`shard_rows`."""
    return [shard_rows(n, world_size, r, align)[0] for r in range(world_size)] + [n]


def shard_triangle(q: int, world_size: int, rank: int) -> tuple[int, int]:
    """Row range ``[lo, hi)`` of rank ``rank`` when the strict upper triangle of a ``q x q`` pair
    matrix is split into ``world_size`` contiguous row bands of (nearly) equal pair count: row c
    holds ``q - 1 - c`` pairs (the joint-count path computes the pairs (c, g), g > c, of its rows)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    work = np.arange(q - 1, -1, -1, dtype=np.int64)              # pairs in row c
    cum = np.concatenate([[0], np.cumsum(work)])                 # pairs above row c
    total = int(cum[-1])
    cuts = [int(np.searchsorted(cum, total * r / world_size, side="left")) for r in range(world_size)] + [q]
    cuts[0] = 0
    cuts = [min(max(c, 0), q) for c in cuts]
    for r in range(1, world_size + 1):                           # monotone
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


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


def setup_peers(ds, n: int, align: int = 4) -> bool:
    """Multi-GPU symmetric distances for one open data set: exchange the ranks' slab handles
    (one all_gather) and hand the library a cross-rank barrier.  Returns False (and leaves the
    plain row-sharded path in place) when there is a single rank, no NCCL, or no peer access."""
    rank, world = dist_info()
    if world == 1:
        return False
    import torch
    import torch.distributed as dist

    if dist.get_backend() != "nccl" or world > 16:
        return False
    starts = shard_starts(n, world, align)
    if any(a == b for a, b in zip(starts, starts[1:])):
        return False        # a rank without rows would skip the barrier inside fs_score
    ok = torch.ones(1, dtype=torch.int32, device="cuda")
    handles = torch.zeros(world * 64, dtype=torch.uint8, device="cuda")
    try:
        handle, _ = ds.peer_slab(starts[rank + 1] - starts[rank])
        mine = torch.from_numpy(handle).cuda()
    except Exception:
        ok.zero_()
        mine = torch.zeros(64, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(handles, mine)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        return False

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    try:
        ds.set_peers(rank, world, starts, handles=handles.cpu().numpy(), barrier=barrier)
    except Exception:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)      # all ranks or none: the barrier inside fs_score must match
    if int(ok.item()) == 0:
        # some rank could not map its peers: every rank falls back to the unsymmetric path
        try:
            ds.set_peers(rank, 1, [0, n], raw_ptrs=[0], barrier=lambda: None)
        except Exception:
            pass
        return False
    return True


def score_sharded(n: int, n_kept: int, score_rows, device_buffers: bool, align: int = 1):
    """Run ``score_rows(lo, hi, out_device_ptr)`` on this rank's target rows and return
    the allreduced float64 weight sums.

    ``score_rows`` returns a float64 numpy vector when ``out_device_ptr`` is None and
    writes the device buffer otherwise (fastselect_b200._native.Dataset.score)."""
    rank, world = dist_info()
    lo, hi = shard_rows(n, world, rank, align)
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


def joint_sharded(q: int, compute_band, device_buffers: bool):
    """Run ``compute_band(lo, hi, out_device_ptr)`` on this rank's row band of the ``q x q`` pair
    matrix (:func:`shard_triangle`) and return the allreduced float64 matrix.  A band's matrix is
    zero outside its own pairs, so the sum over ranks is the full matrix.

    ``compute_band`` returns a float64 ``[q, q]`` numpy array when ``out_device_ptr`` is None and
    writes the device buffer otherwise (fastselect_b200._native.Dataset.joint_matrix)."""
    rank, world = dist_info()
    if world == 1:
        return compute_band(0, q, None)
    import torch
    import torch.distributed as dist

    lo, hi = shard_triangle(q, world, rank)
    if device_buffers:
        buf = torch.empty((q, q), dtype=torch.float64, device="cuda")
        compute_band(lo, hi, buf.data_ptr())
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)       # one NCCL allreduce over NVLink per matrix
        return buf.cpu().numpy()
    t = torch.from_numpy(np.ascontiguousarray(compute_band(lo, hi, None), np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
