"""Multi-GPU plumbing of the estimators.

The reference has no multi-GPU path; its unit of parallelism is the target instance
(``prange``/one CUDA block per instance: MultiSURF.py:174, SURF.py:139, ReliefF.py:143).
Here the target rows of the distance matrix are split into ``world`` contiguous ranges.

Three ways to use more than one GPU, all OPT-IN (a plain ``fit`` uses one GPU and never
communicates):

* ``FASTSELECT_B200_GPUS=8`` (or ``set_gpus``): one process, the library drives one host
  thread per GPU (``_native.MultiDataset`` / ``fs_multi_*``); no torch involved.
* ``enable_distributed()`` (or ``FASTSELECT_B200_DISTRIBUTED=1``) inside a ``torchrun`` job
  with an initialised NCCL process group: one process per GPU, every rank calls ``fit`` with
  the SAME ``X, y``.  torch.distributed is used only to exchange the 64-byte CUDA IPC handles
  of the ranks' exchange arenas (once per arena size); the data path is the library's own:
  1 / world of X uploaded per rank and replicated over NVLink, symmetric distance tiles,
  neighbour masks and weight slices stored straight into the peers' arenas, device-side
  barriers (``include/fastselect_b200.h``, "Multi-GPU group").
* if the ranks cannot map each other's memory (no peer access) or the backend is gloo (CPU
  tests): plain row sharding, every rank uploads X, ONE allreduce of the partial weight sums.
"""
from __future__ import annotations

import os

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


def group_shard_starts(n: int, world_size: int) -> list[int]:
    """Shard starts of a multi-GPU group.  The distance GEMM works in super-blocks of 256 target rows, and a
    rank whose shard ends inside a super-block pays for a whole row of tiles (two GPUs on 5656 samples: 12 + 12
    row super-blocks and 150 tiles per rank = three waves of 74 clusters, against 23 super-blocks and 138 tiles
    = two waves).  So the ``ceil(n / 256)`` super-blocks are dealt to the ranks as evenly as whole super-blocks
    allow (the last ranks take the odd ones: the last super-block is the partial one); shards shorter than four super-blocks keep the plain balanced split (multiples of 4)."""
    blocks = -(-n // 256)
    if world_size < 1:
        raise ValueError(f"bad world_size {world_size}")
    if blocks < 4 * world_size:
        return shard_starts(n, world_size, 4)
    base, extra = divmod(blocks, world_size)
    # the LAST `extra` ranks take one super-block more: the last super-block is the partial one
    return [256 * (r * base + max(0, r - (world_size - extra))) for r in range(world_size)] + [n]


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


_DISTRIBUTED = None        # None: follow the environment variable


def enable_distributed(flag: bool = True) -> None:
    """Opt in to (or out of) sharding ``fit`` across the ranks of the initialised torch.distributed
    process group.  Every rank must then call ``fit`` with the same ``X, y``.  Without the opt-in a
    ``fit`` inside a torchrun job stays local to its process (per-rank data sets, CV folds ...)."""
    global _DISTRIBUTED
    _DISTRIBUTED = bool(flag)


def distributed_enabled() -> bool:
    if _DISTRIBUTED is not None:
        return _DISTRIBUTED
    return os.environ.get("FASTSELECT_B200_DISTRIBUTED", "").strip() not in ("", "0")


def set_gpus(devices) -> None:
    """Use several GPUs of this process: a count (devices 0..k-1), a list of CUDA ordinals, or None / 1
    for one GPU.  Same as the environment variable ``FASTSELECT_B200_GPUS``."""
    if devices is None:
        os.environ.pop("FASTSELECT_B200_GPUS", None)
    elif isinstance(devices, int):
        os.environ["FASTSELECT_B200_GPUS"] = str(devices)
    else:
        os.environ["FASTSELECT_B200_GPUS"] = ",".join(str(int(d)) for d in devices)


def local_devices() -> list[int]:
    """CUDA ordinals of FASTSELECT_B200_GPUS ("4" -> [0, 1, 2, 3]; "0,2,5" -> [0, 2, 5]); [] = one GPU."""
    v = os.environ.get("FASTSELECT_B200_GPUS", "").strip()
    if v == "":
        return []
    devs = [int(t) for t in v.split(",")] if "," in v else list(range(int(v)))
    return devs if len(devs) > 1 else []


def dist_info() -> tuple[int, int]:
    """(rank, world_size) of the initialised torch.distributed group when sharding is opted in, else (0, 1)."""
    if not distributed_enabled():
        return 0, 1
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


# --------------------------------------------------------------------------- #
# one process per GPU: the group's communicator (cached for the life of the process)
# --------------------------------------------------------------------------- #
_COMM = None            # _native.Comm of this process
_COMM_BROKEN = False    # the ranks cannot map each other: stay on the allreduce path


def _input_digest(n, p, y_enc):
    """A cheap fingerprint of the call's inputs; ranks that were handed different data must not be
    summed together."""
    y = np.ascontiguousarray(y_enc, np.int64)
    return [int(n), int(p), int(y.sum()), int((y * (np.arange(y.size, dtype=np.int64) % 1009 + 1)).sum())]


def group_comm(n, p, dtype, y_enc):
    """The connected communicator of this rank, with an arena large enough for an n x p data set of
    ``dtype`` (COLLECTIVE: one all_gather; re-allocation and re-connection happen on all ranks or none).
    Returns None when there is a single rank, the backend is not NCCL, or the ranks cannot map each
    other's memory (the caller then shards rows and allreduces)."""
    global _COMM, _COMM_BROKEN
    rank, world = dist_info()
    if world == 1 or _COMM_BROKEN:
        return None
    import torch
    import torch.distributed as dist

    from . import _native

    if dist.get_backend() != "nccl" or world > 16 or n < 8 * world:
        return None
    device = _native.default_device()
    if _COMM is None or _COMM.world != world or _COMM.rank != rank or _COMM.device != device:
        _COMM = _native.Comm(rank, world, device)
    need = _native.Comm.required_bytes(n, p, dtype, world, with_x=True)
    ok = 1
    handle = np.zeros(64, np.uint8)
    changed = False
    try:
        handle, changed = _COMM.reserve(need)
    except Exception:
        ok = 0
    # one exchange: IPC handle, "my arena moved", "still fine", and the input fingerprint
    mine = np.zeros(64 + 8 * 6, np.uint8)
    mine[:64] = handle
    mine[64:].view(np.int64)[:] = [int(changed or not _COMM.connected), ok] + _input_digest(n, p, y_enc)
    dev = torch.device("cuda", device)
    got = torch.empty(world * mine.size, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(got, torch.from_numpy(mine).to(dev))
    got = got.cpu().numpy().reshape(world, mine.size)
    meta = got[:, 64:].copy().view(np.int64).reshape(world, 6)
    if not (meta[:, 2:] == meta[0, 2:]).all():
        raise ValueError("fastselect_b200: the ranks of the process group were given different X / y; "
                         "distributed fitting (enable_distributed) needs the same data on every rank")
    if not meta[:, 1].all():
        _COMM_BROKEN = True
        return None
    if meta[:, 0].any():
        try:
            _COMM.connect(handles=np.ascontiguousarray(got[:, :64]).reshape(-1))
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # also the host-level barrier fs_comm_connect asks for
        if int(flag.item()) == 0:
            _COMM_BROKEN = True                          # all ranks or none: the collectives inside fs_score must match
            return None
    return _COMM


def open_dataset(x, y_enc, n_classes):
    """The device-resident data set of a ``fit``: several GPUs of this process (FASTSELECT_B200_GPUS),
    this rank's member of a multi-GPU group (enable_distributed + NCCL), or one plain GPU data set."""
    from . import _native

    devs = local_devices()
    rank, world = dist_info()
    if devs and world == 1:
        return _native.MultiDataset(x, y_enc, n_classes, devs)
    comm = group_comm(x.shape[0], x.shape[1], x.dtype, y_enc) if world > 1 else None
    if comm is None:
        return _native.Dataset(x, y_enc, n_classes)
    ds = _native.Dataset(x, y_enc, n_classes, comm=comm)
    try:
        ds.attach_comm(comm, group_shard_starts(ds.n, world))
    except BaseException:
        ds.close()
        raise
    return ds


def score_sharded(n: int, n_kept: int, score_rows, device_buffers: bool, align: int = 1, device: int | None = None):
    """Fallback without a multi-GPU group: run ``score_rows(lo, hi, out_device_ptr)`` on this rank's
    target rows and return the allreduced float64 weight sums.

    ``score_rows`` returns a float64 numpy vector when ``out_device_ptr`` is None and
    writes the device buffer otherwise (fastselect_b200._native.Dataset.score)."""
    rank, world = dist_info()
    lo, hi = shard_rows(n, world, rank, align)
    if world == 1:
        return score_rows(lo, hi, None)
    if device_buffers:
        import torch
        import torch.distributed as dist

        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        buf = torch.empty(n_kept, dtype=torch.float64, device=dev)
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
        from . import _native

        buf = torch.empty((q, q), dtype=torch.float64, device=torch.device("cuda", _native.default_device()))
        compute_band(lo, hi, buf.data_ptr())
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)       # one NCCL allreduce over NVLink per matrix
        return buf.cpu().numpy()
    t = torch.from_numpy(np.ascontiguousarray(compute_band(lo, hi, None), np.float64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()
