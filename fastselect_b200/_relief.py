"""scikit-learn style estimators over the B200 scoring library.

Host-side mirror of the reference's estimator interface for the Relief-family path:
same class names, constructor parameters, validation, messages and fitted attributes
as ``fast_select.MultiSURF`` (MultiSURF.py:273-489), ``fast_select.SURF``
(SURF.py:220-425) and ``fast_select.ReliefF`` (ReliefF.py:239-452).  Only the GPU
backend exists here: ``backend='gpu'``/``'auto'`` run hand-written sm_100a CUDA through
the C ABI, and there is no CPU fallback (``backend='cpu'`` raises).
"""
from __future__ import annotations

import warnings

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.utils.validation import check_is_fitted, validate_data

from . import _native
from ._shard import open_dataset, score_sharded

_NO_GPU_MULTISURF = ("backend='gpu' was selected, but no compatible "
                     "NVIDIA GPU was found or CUDA toolkit is not installed.")       # MultiSURF.py:399-403
_NO_GPU_SURF = "backend='gpu', but no CUDA-enabled GPU is available."                  # SURF.py:341-342
_NO_CPU = ("fastselect_b200 implements only the GPU backend of the Relief-family path; "
           "backend='cpu' is not available (use the reference package for CPU runs).")


def _validate_n_select(n_features_to_select, n_features):
    """MultiSURF.py:350-364 / SURF.py:296-310 / ReliefF.py:319-333."""
    if isinstance(n_features_to_select, float):
        if not 0.0 < n_features_to_select <= 1.0:
            raise ValueError("If n_features_to_select is a float, it must be in (0, 1].")
        return max(1, int(n_features_to_select * n_features))
    if isinstance(n_features_to_select, int):
        if not 0 < n_features_to_select <= n_features:
            raise ValueError(
                f"If n_features_to_select is an int ({n_features_to_select}), "
                f"it must be > 0 and <= n_features ({n_features}).")
        return n_features_to_select
    raise TypeError("n_features_to_select must be an int or a float.")


def _narrow_integers(x):
    """Wide integer matrices whose values fit one byte (``np.random.randint`` returns int64) are narrowed
    to int8 / uint8 BEFORE validation: sklearn would otherwise convert them to float32 on the host -- four
    times the bytes to convert and to upload -- and one-byte matrices are what the genotype path is built
    for.  Values are unchanged, so every result is."""
    if isinstance(x, np.ndarray) and x.ndim == 2 and x.dtype.kind in "iu" and x.dtype.itemsize > 1 and x.size:
        lo, hi = int(x.min()), int(x.max())
        if -128 <= lo and hi <= 127:
            return x.astype(np.int8)
        if 0 <= lo and hi <= 255:
            return x.astype(np.uint8)
    return x


def _count_distinct_host(x, cols):
    """Exact distinct counts for the few columns whose on-device scan overflowed."""
    return np.array([np.unique(x[:, f]).size for f in cols], dtype=np.int64)


def _is_discrete(x, n_distinct, discrete_limit):
    """``np.unique(x[:, f]).size <= discrete_limit`` (MultiSURF.py:416-420, SURF.py:347-350,
    ReliefF.py:366-369) from the GPU column scan; only columns with more than
    FS_DISTINCT_CAP distinct values AND a larger limit need a host recount."""
    n_distinct = n_distinct.astype(np.int64)
    over = np.flatnonzero(n_distinct > _native.FS_DISTINCT_CAP)
    if over.size and discrete_limit > _native.FS_DISTINCT_CAP:
        n_distinct[over] = _count_distinct_host(x, over)
    elif over.size:
        n_distinct[over] = np.iinfo(np.int64).max
    return n_distinct <= discrete_limit


class _Session:
    """One uploaded data set plus its per-column typing; ``score(feat_idx)`` returns the
    float32 scores ``sum_i W_i / n`` for a column subset (what the reference's host
    callers return).  Used by ``fit`` and, kept open across iterations, by TuRF."""

    def __init__(self, algo, x, y_enc, n_classes, is_discrete, recip, arith, use_star=False, k=0,
                 class_probs=None, dataset=None):
        self.algo, self.use_star, self.k, self.class_probs = algo, use_star, k, class_probs
        self.ds = dataset if dataset is not None else open_dataset(x, y_enc, n_classes)
        self.n, self.p = self.ds.n, self.ds.p
        self.is_discrete, self.recip, self.arith = is_discrete, recip, arith
        self.ds.set_features(is_discrete, recip, arith)
        self.last_stats = None
        # one process per GPU without a multi-GPU group (no peer access / gloo): shard starts are multiples of 4
        self.row_align = 4
        # inside a multi-GPU group (or on a MultiDataset) one score call returns the complete sums
        self.collective = isinstance(self.ds, _native.MultiDataset) or self.ds.comm is not None

    def score(self, feat_idx=None, want_stats=False):
        n_kept = self.p if feat_idx is None else len(feat_idx)

        def score_rows(lo, hi, out_ptr):
            res = self.ds.score(self.algo, self.use_star, self.k, self.class_probs, feat_idx, lo, hi,
                                out_device_ptr=out_ptr, want_stats=want_stats)
            if want_stats:
                res, self.last_stats = res
            return res

        if isinstance(self.ds, _native.MultiDataset):
            wsum = score_rows(0, self.n, None)
        elif self.collective:
            wsum = score_rows(self.ds.shard[0], self.ds.shard[1], None)      # COLLECTIVE: every rank, own shard
        else:
            wsum = score_sharded(self.n, n_kept, score_rows, device_buffers=True, align=self.row_align,
                                 device=getattr(self.ds, "device", None))
        # "/ n_samples" of the reference host callers (MultiSURF.py:162, SURF.py:128, ReliefF.py:134)
        return (wsum / self.n).astype(np.float32)

    def close(self):
        self.ds.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class _ReliefBase(TransformerMixin, BaseEstimator):
    _algo_label = "Relief"

    def _validate_parameters(self, n_samples, n_features):
        if self.backend not in ["auto", "gpu", "cpu"]:
            raise ValueError("backend must be one of 'auto', 'gpu', or 'cpu'")
        if n_samples < 2:
            raise ValueError(
                f"{self._algo_label} requires at least 2 samples, but got n_samples = {n_samples}")
        self._validate_extra(n_samples)
        return _validate_n_select(self.n_features_to_select, n_features)

    def _validate_extra(self, n_samples):
        pass

    def _resolve_backend(self, no_gpu_message):
        if self.backend == "cpu":
            raise NotImplementedError(_NO_CPU)
        if _native.device_count() < 1:
            # 'auto' falls back to the CPU in the reference; this package has no CPU path
            raise RuntimeError(no_gpu_message)
        return "gpu"

    def transform(self, x):
        """Reduce x to the selected features (MultiSURF.py:446-467)."""
        check_is_fitted(self)
        x = validate_data(self, x, reset=False, dtype=[np.float64, np.float32])
        return x[:, self.top_features_]

    def fit_transform(self, x, y=None, **fit_params):
        self.fit(x, y)
        return self.transform(x)

    @staticmethod
    def _top(scores, n_select):
        """``np.argsort(scores)[::-1][:n_select]`` (MultiSURF.py:443, SURF.py:375, ReliefF.py:406)
        without sorting all p scores: when the n_select largest scores are distinct and strictly
        above the rest, their descending order IS that prefix; ties fall back to the full argsort
        so that the reference's tie order is kept."""
        p = scores.size
        if n_select * 8 < p and not np.isnan(scores).any():
            part = np.argpartition(scores, p - n_select - 1)
            cand = part[p - n_select - 1:]                      # the n_select + 1 largest scores
            vals = scores[cand]
            order = np.argsort(vals)[::-1]
            sv = vals[order]
            if np.all(sv[:-1] > sv[1:]):                        # strictly decreasing: no ties anywhere near the cut
                return cand[order][:n_select]
        return np.argsort(scores)[::-1][:n_select]

    def _finish(self, scores, n_select):
        self.feature_importances_ = scores
        self.top_features_ = self._top(scores, n_select)
        return self


class MultiSURF(_ReliefBase):
    """MultiSURF / MultiSURF* on B200 (drop-in for ``fast_select.MultiSURF``, GPU backend).

    Parameters, attributes and error behaviour follow MultiSURF.py:273-444."""

    _algo_label = "MultiSURF"

    def __init__(self, n_features_to_select: int | float = 0.2, backend: str = "auto", use_star: bool = False,
                 discrete_limit: int = 10, n_jobs: int = -1, verbose: bool = False):
        self.n_features_to_select = n_features_to_select
        self.backend = backend
        self.use_star = use_star
        self.discrete_limit = discrete_limit
        self.n_jobs = n_jobs
        self.verbose = verbose

    def _open_session(self, x, y):
        """Validation and per-column preprocessing of ``fit`` (MultiSURF.py:384-420);
        int8/uint8 genotype matrices are kept as they are (their float32 images are exact)."""
        x, y = validate_data(self, _narrow_integers(x), y, y_numeric=True, dtype=[np.float32, np.int8, np.uint8],
                             ensure_2d=True)
        self.n_features_in_ = x.shape[1]
        n_select = self._validate_parameters(x.shape[0], self.n_features_in_)
        self.effective_backend_ = self._resolve_backend(_NO_GPU_MULTISURF)
        y_enc = np.unique(y, return_inverse=True)[1].astype(np.int32)    # labels are only compared (:216)
        ds = open_dataset(x, y_enc, int(y_enc.max()) + 1)
        try:
            cmin, cmax, cnt = ds.column_stats()
            # ranges from the float32 matrix, zero -> 1, reciprocal in float32 (:409-412)
            ranges = cmax.astype(np.float32) - cmin.astype(np.float32)
            ranges[ranges == 0] = 1
            recip = (1.0 / ranges).astype(np.float32)
            self.is_discrete_ = _is_discrete(x, cnt, self.discrete_limit)
            sess = _Session(_native.FS_MULTISURF, None, None, None, self.is_discrete_, recip,
                            _native.FS_ARITH_F32, use_star=self.use_star, dataset=ds)
        except BaseException:
            ds.close()
            raise
        return sess, n_select

    def fit(self, x, y):
        sess, n_select = self._open_session(x, y)
        with sess:
            if self.verbose and self.use_star:
                print("Running MultiSURF* on the GPU now...")
            elif self.verbose:
                print("Running MultiSURF on the GPU now...")
            scores = sess.score()
        return self._finish(scores, n_select)


class SURF(_ReliefBase):
    """SURF / SURF* on B200 (drop-in for ``fast_select.SURF``, GPU backend; SURF.py:220-380)."""

    _algo_label = "SURF"

    def __init__(self, n_features_to_select: int | float = 0.2, backend: str = "auto", use_star: bool = False,
                 discrete_limit: int = 10, n_jobs: int = -1, verbose: bool = False):
        self.n_features_to_select = n_features_to_select
        self.backend = backend
        self.use_star = use_star
        self.discrete_limit = discrete_limit
        self.n_jobs = n_jobs
        self.verbose = verbose

    def _open_session(self, X, y):
        # SURF.py:330-332 validates to float64; float32/int8/uint8 inputs are uploaded as
        # they are and widened on the device (their float64 images are exact)
        X, y = validate_data(self, _narrow_integers(X), y, y_numeric=True,
                             dtype=[np.float64, np.float32, np.int8, np.uint8], ensure_2d=True)
        self.n_features_in_ = X.shape[1]
        n_select = self._validate_parameters(X.shape[0], self.n_features_in_)
        self.effective_backend_ = self._resolve_backend(_NO_GPU_SURF)
        # SURF.py:363,371: y.astype(np.int32) truncates before the equality test
        y_enc = np.unique(np.asarray(y).astype(np.int32), return_inverse=True)[1].astype(np.int32)
        ds = open_dataset(X, y_enc, int(y_enc.max()) + 1)
        try:
            cmin, cmax, cnt = ds.column_stats()
            self.is_discrete_ = _is_discrete(X, cnt, self.discrete_limit)
            ranges = cmax - cmin                                   # float64 (:352)
            ranges[self.is_discrete_] = 1.0
            ranges[ranges == 0] = 1.0
            recip = (1.0 / ranges).astype(np.float32)
            sess = _Session(_native.FS_SURF, None, None, None, self.is_discrete_, recip,
                            _native.FS_ARITH_F64, use_star=self.use_star, dataset=ds)
        except BaseException:
            ds.close()
            raise
        return sess, n_select

    def fit(self, X, y):
        sess, n_select = self._open_session(X, y)
        with sess:
            algo_name = "SURF*" if self.use_star else "SURF"
            if self.verbose:
                print(f"Running {algo_name} on the {self.effective_backend_.upper()} now...")
            scores = sess.score()
        self._finish(scores, n_select)
        if self.verbose:
            print("Feature scoring completed.")
        return self


class ReliefF(_ReliefBase):
    """ReliefF on B200 (drop-in for ``fast_select.ReliefF``, GPU backend; ReliefF.py:239-407).

    Follows the reference's CPU semantics (k hits + k misses of every other class,
    prior-weighted), any ``n_neighbors`` and any number of classes.  Candidates tied at the
    k-th distance are taken in the order the reference takes them (numba's quicksort order,
    replayed on the GPU); ``FS_B200_RELIEFF_TIES=index`` switches to sample-index order."""

    _algo_label = "ReliefF"

    def __init__(self, n_features_to_select: int | float = 0.2, discrete_limit: int = 10, n_neighbors: int = 3,
                 backend: str = "auto", verbose: bool = False, n_jobs: int = -1):
        self.n_features_to_select = n_features_to_select
        self.discrete_limit = discrete_limit
        self.n_neighbors = n_neighbors
        self.backend = backend
        self.verbose = verbose
        self.n_jobs = n_jobs

    def _validate_extra(self, n_samples):
        if not (0 < self.n_neighbors < n_samples):
            raise ValueError(
                f"n_neighbors ({self.n_neighbors}) must be an integer "
                f"between 1 and n_samples - 1 ({n_samples - 1}).")

    def _open_session(self, x, y):
        x, y = validate_data(self, _narrow_integers(x), y, dtype=[np.float64, np.float32, np.int8, np.uint8],
                             ensure_2d=True, y_numeric=True)
        self.n_features_in_ = x.shape[1]
        n_select = self._validate_parameters(x.shape[0], self.n_features_in_)
        self.classes_, y_encoded = np.unique(y, return_inverse=True)
        if len(self.classes_) < 2:                                  # ReliefF.py:351-356
            self.feature_importances_ = np.zeros(self.n_features_in_, dtype=np.float32)
            self.top_features_ = np.arange(n_select)
            self.effective_backend_ = "cpu" if self.backend != "gpu" else "gpu"
            return None, n_select
        min_class_size = np.min(np.bincount(y_encoded))
        if self.n_neighbors >= min_class_size:                      # ReliefF.py:358-364
            warnings.warn(
                f"n_neighbors ({self.n_neighbors}) is greater than or equal to the "
                f"smallest class size ({min_class_size}).", UserWarning)
        self.effective_backend_ = self._resolve_backend(_NO_GPU_MULTISURF)
        class_counts = np.bincount(y_encoded)
        class_probs = (class_counts / len(y)).astype(np.float32)    # :371-373, cast at :401
        ds = open_dataset(x, y_encoded.astype(np.int32), len(self.classes_))
        try:
            cmin, cmax, cnt = ds.column_stats()
            self.is_discrete_ = _is_discrete(x, cnt, self.discrete_limit)
            ranges = cmax - cmin                                    # float64 (:377)
            ranges[self.is_discrete_] = 1.0
            ranges[ranges == 0] = 1.0
            recip = (1.0 / ranges).astype(np.float32)
            # the kernel runs on float32 X (:388,400): the device narrows float64 columns
            sess = _Session(_native.FS_RELIEFF, None, None, None, self.is_discrete_, recip,
                            _native.FS_ARITH_F32, k=int(self.n_neighbors), class_probs=class_probs, dataset=ds)
        except BaseException:
            ds.close()
            raise
        return sess, n_select

    def fit(self, x, y):
        sess, n_select = self._open_session(x, y)
        if sess is None:
            return self
        with sess:
            if self.verbose:
                print("Running ReliefF on the GPU now...")
            scores = sess.score()
        return self._finish(scores, n_select)
