"""ctypes front-end of the CPU oracle (oracle/fs_oracle.c) plus a numpy restatement
of the reference estimators' ``fit`` preprocessing.

TEST INFRASTRUCTURE ONLY -- imported by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs; never by fastselect_b200/.

Citations are into /root/reference/src/fast_select/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfs_oracle.so")
_lib = None

NONE, NEAR_HIT, NEAR_MISS, FAR_MISS, FAR_HIT = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    """Compile oracle/fs_oracle.c with oracle/Makefile (gcc)."""
    src = os.path.join(_HERE, "fs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.fso_max_threads.restype = C.c_int
    return _lib


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(C.POINTER(typ))


def max_threads() -> int:
    return int(lib().fso_max_threads())


def set_threads(n: int) -> None:
    """OpenMP threads used by the oracle (torchrun exports OMP_NUM_THREADS=1)."""
    lib().fso_set_threads(C.c_int(int(n)))


# --------------------------------------------------------------------------- #
# preprocessing restated (a10)
# --------------------------------------------------------------------------- #
def is_discrete_columns(x: np.ndarray, discrete_limit: int) -> np.ndarray:
    """``np.unique(x[:, f]).size <= discrete_limit`` per column
    (MultiSURF.py:416-420, SURF.py:347-350, ReliefF.py:366-369), vectorised."""
    xs = np.sort(x, axis=0)
    n_unique = 1 + (xs[1:] != xs[:-1]).sum(axis=0)
    return n_unique <= discrete_limit


def multisurf_prep(x, discrete_limit):
    """MultiSURF.py:384-420: X -> float32; ranges from the float32 matrix;
    discrete ranges are left as they are (unused by the kernel)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    ranges = (x.max(axis=0) - x.min(axis=0)).astype(np.float32)
    ranges[ranges == 0] = 1
    recip = (1.0 / ranges).astype(np.float32)
    return x, recip, is_discrete_columns(x, discrete_limit)


def surf_prep(x, discrete_limit):
    """SURF.py:330-355: X stays float64; discrete -> range 1."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    isd = is_discrete_columns(x, discrete_limit)
    ranges = x.max(axis=0) - x.min(axis=0)
    ranges[isd] = 1.0
    ranges[ranges == 0] = 1.0
    recip = (1.0 / ranges).astype(np.float32)
    return x, recip, isd


def relieff_prep(x, y, discrete_limit):
    """ReliefF.py:343-380: ranges in float64 (discrete -> 1), X cast to float32
    for the kernel, class priors in float32, labels encoded 0..C-1."""
    x64 = np.ascontiguousarray(x, dtype=np.float64)
    isd = is_discrete_columns(x64, discrete_limit)
    labels, counts = np.unique(y, return_counts=True)
    class_probs = (counts / len(y)).astype(np.float32)
    y_enc = np.searchsorted(labels, y).astype(np.int32)
    ranges = x64.max(axis=0) - x64.min(axis=0)
    ranges[isd] = 1.0
    ranges[ranges == 0] = 1.0
    recip = (1.0 / ranges).astype(np.float32)
    return x64.astype(np.float32), y_enc, class_probs, recip, isd


# --------------------------------------------------------------------------- #
# kernels
# --------------------------------------------------------------------------- #
def _check(rc, name):
    if rc != 0:
        raise RuntimeError(f"{name} failed with status {rc}")


def multisurf_scores(x32, y, recip, isd, use_star=False):
    x32 = np.ascontiguousarray(x32, np.float32)
    n, p = x32.shape
    y = np.ascontiguousarray(y, np.int64)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    out = np.empty(p, np.float32)
    _check(lib().fso_multisurf_scores(_p(x32, C.c_float), C.c_int64(n), C.c_int64(p), _p(y, C.c_int64),
                                      _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int(int(use_star)),
                                      _p(out, C.c_float)), "fso_multisurf_scores")
    return out


def multisurf_targets(x32, y, recip, isd, use_star, targets, want_mask=True, want_dist=True):
    x32 = np.ascontiguousarray(x32, np.float32)
    n, p = x32.shape
    y = np.ascontiguousarray(y, np.int64)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    targets = np.ascontiguousarray(targets, np.int64)
    nt = targets.size
    wsum = np.empty(p, np.float64)
    thresh = np.empty(nt, np.float64)
    mask = np.empty((nt, n), np.int8) if want_mask else None
    dist = np.empty((nt, n), np.float64) if want_dist else None
    counts = np.empty((nt, 3), np.int64)
    _check(lib().fso_multisurf_targets(_p(x32, C.c_float), C.c_int64(n), C.c_int64(p), _p(y, C.c_int64),
                                       _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int(int(use_star)),
                                       _p(targets, C.c_int64), C.c_int64(nt), _p(wsum, C.c_double),
                                       _p(thresh, C.c_double), _p(mask, C.c_int8), _p(dist, C.c_double),
                                       _p(counts, C.c_int64)), "fso_multisurf_targets")
    return dict(wsum=wsum, thresh=thresh, mask=mask, dist=dist, counts=counts)


def multisurf_targets_bytes(x8, y, use_star, targets, cols=None, want_mask=True, want_dist=True):
    """multisurf_targets for a one-byte genotype matrix whose columns are all discrete, optionally
    restricted to the column subset ``cols`` (no float32 copy of the matrix is made)."""
    assert x8.dtype in (np.int8, np.uint8) and x8.ndim == 2 and x8.strides[1] == 1
    n, p_all = x8.shape
    ld = x8.strides[0]
    y = np.ascontiguousarray(y, np.int64)
    targets = np.ascontiguousarray(targets, np.int64)
    cols = None if cols is None else np.ascontiguousarray(cols, np.int64)
    p = p_all if cols is None else cols.size
    nt = targets.size
    wsum = np.empty(p, np.float64)
    thresh = np.empty(nt, np.float64)
    mask = np.empty((nt, n), np.int8) if want_mask else None
    dist = np.empty((nt, n), np.float64) if want_dist else None
    _check(lib().fso_multisurf_targets_u8(C.c_void_p(x8.ctypes.data), C.c_int64(n), C.c_int64(ld), _p(cols, C.c_int64),
                                          C.c_int64(p), _p(y, C.c_int64), C.c_int(int(use_star)),
                                          _p(targets, C.c_int64), C.c_int64(nt), _p(wsum, C.c_double),
                                          _p(thresh, C.c_double), _p(mask, C.c_int8), _p(dist, C.c_double)),
           "fso_multisurf_targets_u8")
    return dict(wsum=wsum, thresh=thresh, mask=mask, dist=dist)


def surf_scores(x64, y, recip, isd, use_star=False, sum_mode=0):
    x64 = np.ascontiguousarray(x64, np.float64)
    n, p = x64.shape
    y = np.ascontiguousarray(y, np.int32)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    out = np.empty(p, np.float32)
    _check(lib().fso_surf_scores(_p(x64, C.c_double), C.c_int64(n), C.c_int64(p), _p(y, C.c_int32),
                                 _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int(int(use_star)),
                                 C.c_int(sum_mode), _p(out, C.c_float)), "fso_surf_scores")
    return out


def surf_targets(x64, y, recip, isd, use_star, targets, sum_mode=1, want_mask=True, want_dist=True):
    x64 = np.ascontiguousarray(x64, np.float64)
    n, p = x64.shape
    y = np.ascontiguousarray(y, np.int32)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    targets = np.ascontiguousarray(targets, np.int64)
    nt = targets.size
    wsum = np.empty(p, np.float64)
    thresh = np.empty(nt, np.float64)
    mask = np.empty((nt, n), np.int8) if want_mask else None
    dist = np.empty((nt, n), np.float64) if want_dist else None
    _check(lib().fso_surf_targets(_p(x64, C.c_double), C.c_int64(n), C.c_int64(p), _p(y, C.c_int32),
                                  _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int(int(use_star)),
                                  C.c_int(sum_mode), _p(targets, C.c_int64), C.c_int64(nt),
                                  _p(wsum, C.c_double), _p(thresh, C.c_double), _p(mask, C.c_int8),
                                  _p(dist, C.c_double)), "fso_surf_targets")
    return dict(wsum=wsum, thresh=thresh, mask=mask, dist=dist)


def relieff_scores(x32, y_enc, recip, isd, k, class_probs, tie_mode=0):
    x32 = np.ascontiguousarray(x32, np.float32)
    n, p = x32.shape
    y_enc = np.ascontiguousarray(y_enc, np.int32)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    class_probs = np.ascontiguousarray(class_probs, np.float32)
    out = np.empty(p, np.float32)
    _check(lib().fso_relieff_scores(_p(x32, C.c_float), C.c_int64(n), C.c_int64(p), _p(y_enc, C.c_int32),
                                    _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int32(k),
                                    _p(class_probs, C.c_float), C.c_int32(class_probs.size),
                                    C.c_int(tie_mode), _p(out, C.c_float)), "fso_relieff_scores")
    return out


def relieff_targets(x32, y_enc, recip, isd, k, class_probs, targets, tie_mode=1,
                    want_mask=True, want_dist=True):
    x32 = np.ascontiguousarray(x32, np.float32)
    n, p = x32.shape
    y_enc = np.ascontiguousarray(y_enc, np.int32)
    recip = np.ascontiguousarray(recip, np.float32)
    isd = np.ascontiguousarray(isd, np.uint8)
    class_probs = np.ascontiguousarray(class_probs, np.float32)
    targets = np.ascontiguousarray(targets, np.int64)
    nt = targets.size
    wsum = np.empty(p, np.float64)
    mask = np.empty((nt, n), np.int8) if want_mask else None
    dist = np.empty((nt, n), np.float64) if want_dist else None
    _check(lib().fso_relieff_targets(_p(x32, C.c_float), C.c_int64(n), C.c_int64(p), _p(y_enc, C.c_int32),
                                     _p(recip, C.c_float), _p(isd, C.c_uint8), C.c_int32(k),
                                     _p(class_probs, C.c_float), C.c_int32(class_probs.size),
                                     C.c_int(tie_mode), _p(targets, C.c_int64), C.c_int64(nt),
                                     _p(wsum, C.c_double), _p(mask, C.c_int8), _p(dist, C.c_double)),
           "fso_relieff_targets")
    return dict(wsum=wsum, mask=mask, dist=dist)


def argsort_numba(a32):
    a32 = np.ascontiguousarray(a32, np.float32)
    out = np.empty(a32.size, np.int64)
    lib().fso_argsort_numba(_p(a32, C.c_float), C.c_int64(a32.size), _p(out, C.c_int64))
    return out


# --------------------------------------------------------------------------- #
# end-to-end "fit" restatements (scores + ranking, a11)
# --------------------------------------------------------------------------- #
def rank(scores, n_select):
    """``np.argsort(scores)[::-1][:n_select]`` (MultiSURF.py:443, SURF.py:375, ReliefF.py:406)."""
    return np.argsort(scores)[::-1][:n_select]


def fit_multisurf(x, y, discrete_limit=10, use_star=False):
    x32, recip, isd = multisurf_prep(x, discrete_limit)
    # MultiSURF.py:216 compares labels by equality only; any injective coding is equivalent
    y_code = np.unique(np.asarray(y), return_inverse=True)[1].astype(np.int64)
    return multisurf_scores(x32, y_code, recip, isd, use_star), isd


def fit_surf(x, y, discrete_limit=10, use_star=False, sum_mode=0):
    x64, recip, isd = surf_prep(x, discrete_limit)
    # SURF.py:363,371: y.astype(np.int32) truncates
    return surf_scores(x64, np.asarray(y).astype(np.int32), recip, isd, use_star, sum_mode), isd


def fit_relieff(x, y, discrete_limit=10, n_neighbors=3, tie_mode=0):
    x32, y_enc, class_probs, recip, isd = relieff_prep(x, y, discrete_limit)
    if class_probs.size < 2:  # ReliefF.py:351-356
        return np.zeros(x32.shape[1], np.float32), isd
    return relieff_scores(x32, y_enc, recip, isd, n_neighbors, class_probs, tie_mode), isd


# --------------------------------------------------------------------------- #
# joint-count tables, mutual information, symmetrical uncertainty
# --------------------------------------------------------------------------- #
def joint_counts(xa, xb, ka=None, kb=None):
    """Contingency table of two code vectors (mutual_information.py:31-33, CFS.py:51-53): int64 [ka, kb]."""
    xa = np.ascontiguousarray(xa, np.int32)
    xb = np.ascontiguousarray(xb, np.int32)
    ka = int(xa.max()) + 1 if ka is None else int(ka)
    kb = int(xb.max()) + 1 if kb is None else int(kb)
    table = np.empty((ka, kb), np.int64)
    _check(lib().fso_joint_counts(_p(xa, C.c_int32), _p(xb, C.c_int32), C.c_int64(xa.size), C.c_int32(ka),
                                  C.c_int32(kb), _p(table, C.c_int64)), "fso_joint_counts")
    return table


def _joint_matrices(x_codes, y_codes, kind, log_base, want_matrix):
    x_codes = np.ascontiguousarray(x_codes, np.int32)
    y_codes = np.ascontiguousarray(y_codes, np.int32)
    n, p = x_codes.shape
    vec = np.empty(p, np.float64)
    mat = np.empty((p, p), np.float64) if want_matrix else None
    _check(lib().fso_joint_matrices(_p(x_codes, C.c_int32), C.c_int64(n), C.c_int64(p), _p(y_codes, C.c_int32),
                                    C.c_int(kind), C.c_double(log_base), _p(vec, C.c_double), _p(mat, C.c_double)),
           "fso_joint_matrices")
    return vec, mat


def mi_matrices(x_codes, y_codes, unit="bit", want_matrix=True):
    """``calculate_mi_matrices`` (mutual_information.py:158-196, CPU path :49-63): (relevance [p], redundancy
    [p, p]) of non-negative integer codes."""
    return _joint_matrices(x_codes, y_codes, 0, np.log(2.0) if unit == "bit" else 1.0, want_matrix)


def su_matrices(x_codes, y_codes, want_matrix=True):
    """``_precompute_correlations_cpu`` (CFS.py:81-104) in float64: (r_cf [p], r_ff [p, p])."""
    return _joint_matrices(x_codes, y_codes, 1, 1.0, want_matrix)


def mrmr_codes(x, y):
    """mRMR.fit's value coding (mRMR.py:90-92): index into the sorted union of all values of X and y."""
    x = np.asarray(x)
    y = np.asarray(y)
    u = np.unique(np.concatenate([np.unique(x), np.unique(y)]))
    return np.searchsorted(u, x).astype(np.int32), np.searchsorted(u, y).astype(np.int32), u
