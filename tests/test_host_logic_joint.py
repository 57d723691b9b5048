"""CPU: host side of the joint-count path (fastselect_b200/_mi.py) -- the greedy searches of mRMR and
CFS replayed on the REFERENCE's own matrices (tests/golden/joint_vectors.npz), the value coding, the
validation / error behaviour (ports of the behaviours in the reference's tests/test_mrmr.py:38-50,
:164-186 and tests/test_cfs.py:60-75, :205-212), the triangular band sharding and its two-rank gloo
combination."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
from sklearn.exceptions import NotFittedError

import fastselect_b200 as fsb
from fastselect_b200 import _mi, _native
from fastselect_b200._shard import shard_triangle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HAVE_GPU = _native.device_count() > 0


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))


@pytest.mark.parametrize("data", ["mrmr_fixture", "mrmr_dup", "geno", "states"])
@pytest.mark.parametrize("method", ["MID", "MIQ"])
def test_mrmr_selection_replays_the_reference(g, data, method):
    rel, red = g[f"mi_rel_bit_{data}"], g[f"mi_red_bit_{data}"]
    ref = g[f"mrmr_top_{method}_{data}"]
    top = fsb.mRMR._select(rel, red, len(ref), method)
    assert top.dtype == np.int32 and np.array_equal(top, ref)


def test_mrmr_prefers_the_less_redundant_feature(g):
    """tests/test_mrmr.py:104-150: feature 1 duplicates feature 0, feature 9 is the cleaner copy of y."""
    top = fsb.mRMR._select(g["mi_rel_bit_mrmr_dup"], g["mi_red_bit_mrmr_dup"], 2, "MID")
    assert set(top.tolist()) == {0, 9}


@pytest.mark.parametrize("data", ["cfs_fixture", "geno", "states", "mrmr_fixture"])
def test_cfs_search_replays_the_reference(g, data):
    r_cf, r_ff = g[f"cfs_rcf_{data}"], g[f"cfs_rff_{data}"]
    sel = np.sort(np.array(_mi._best_first_search(r_cf, r_ff), dtype=int))
    kept = np.sort(np.array(_mi._prune_redundant(sel, r_cf, r_ff), dtype=int))
    assert np.array_equal(kept, g[f"cfs_sel_{data}"])
    k = len(kept)
    merit = 0.0
    if k:
        merit = float(_mi._cfs_merit(np.float64(np.sum(r_cf[kept])), k,
                                     np.float64(np.sum(np.triu(r_ff[np.ix_(kept, kept)], k=1)))))
    assert merit == pytest.approx(float(g[f"cfs_merit_{data}"][0]), rel=1e-12, abs=1e-15)


def test_cfs_fixture_selects_features_0_and_2(g):
    """tests/test_cfs.py:77-105: the informative and the independent feature, not the redundant copy."""
    assert g["cfs_sel_cfs_fixture"].tolist() == [0, 2]


@pytest.mark.parametrize("data", ["cfs_fixture", "geno", "states"])
def test_cfs_value_coding_matches_the_reference(g, data):
    codes, n_states = fsb.CFS()._encode(g[f"X_{data}"])
    assert np.array_equal(codes, g[f"cfs_codes_{data}"])
    assert n_states.max() <= 16 and codes.max() < n_states.max()


def test_no_features_above_the_relevance_floor():
    """tests/test_cfs.py:145-160: nothing reaches r_cf >= 0.1 -> empty selection."""
    r_cf = np.array([0.01, 0.05, 0.0999], np.float32)
    assert _mi._best_first_search(r_cf, np.zeros((3, 3), np.float32)) == []


def test_validation_and_error_behaviour():
    with pytest.raises(ValueError, match="Method must be either 'MID' or 'MIQ'"):
        fsb.mRMR(n_features_to_select=5, method="INVALID_METHOD")
    with pytest.raises(ValueError, match="Backend must be either 'cpu' or 'gpu'"):
        fsb.mRMR(n_features_to_select=5, backend="tpu")
    with pytest.raises(NotImplementedError, match="only the GPU backend"):
        fsb.mRMR(n_features_to_select=5, backend="cpu")
    x = np.arange(12).reshape(6, 2) % 3
    y = np.arange(6) % 2
    mi = fsb.mutual_information
    with pytest.raises(ValueError, match="X must be 2-D and y 1-D"):
        mi.calculate_mi_matrices(x, y[:5])
    with pytest.raises(ValueError, match="integer-coded"):
        mi.calculate_mi_matrices(x.astype(float), y)
    with pytest.raises(ValueError, match="negative values"):
        mi.calculate_mi_matrices(x - 1, y)
    with pytest.raises(ValueError, match="1-D arrays of equal length"):
        mi.calculate_mi_single_pair(x[:, 0], y[:3])
    with pytest.raises(NotImplementedError, match="only the GPU backend"):
        mi.calculate_mi_matrices(x, y, backend="cpu")
    cfs = fsb.CFS(n_bins=5, strategy="quantile", backend="gpu", n_jobs=4)
    assert (cfs.n_bins, cfs.strategy, cfs.backend, cfs.n_jobs) == (5, "quantile", "gpu", 4)
    with pytest.raises(NotFittedError):
        fsb.CFS().transform(x)
    with pytest.raises(NotFittedError):
        fsb.CFS()._get_support_mask()
    with pytest.raises(NotImplementedError, match="only the GPU backend"):
        fsb.CFS(backend="cpu").fit(x.astype(float), y)


@pytest.mark.skipif(HAVE_GPU, reason="only where no GPU is usable")
def test_no_gpu_is_an_error_never_a_fallback():
    x = np.arange(12).reshape(6, 2) % 3
    y = np.arange(6) % 2
    with pytest.raises(RuntimeError, match="CUDA not available"):
        fsb.mutual_information.calculate_mi_matrices(x, y, backend="gpu")
    with pytest.raises(RuntimeError, match="GPU backend was selected"):
        fsb.mRMR(n_features_to_select=1, backend="gpu")
    with pytest.raises(RuntimeError, match="no CUDA-enabled GPU is available"):
        fsb.CFS(backend="gpu").fit(x.astype(float), y)


def test_upload_dtype_holds_every_value_exactly():
    y = np.array([0, 1, 0])
    for vals, dt in (([0, 255, 3], np.uint8), ([-1, 5, 7], np.float32), ([0, 1 << 24, 2], np.float32),
                     ([0, (1 << 24) + 1, 2], np.float64), ([-(1 << 40), 5, 7], np.float64)):
        x = np.array(vals, np.int64)[:, None]
        up = _mi._stack_for_upload(x, y)
        assert up.dtype == dt and np.array_equal(up[:, 0].astype(np.int64), x[:, 0])
    with pytest.raises(ValueError, match="2\\^53"):
        _mi._stack_for_upload(np.array([[1 << 60], [0], [1]]), y)


def test_triangle_bands_cover_every_pair_once_and_balance():
    for q in (1, 2, 5, 63, 1000):
        for w in (1, 2, 3, 8):
            cuts = [shard_triangle(q, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == q
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:])) and all(lo <= hi for lo, hi in cuts)
            work = [sum(q - 1 - c for c in range(lo, hi)) for lo, hi in cuts]
            assert sum(work) == q * (q - 1) // 2
            if q >= 1000:
                assert max(work) - min(work) <= 2 * q        # at most one row's worth off the mean


_WORKER = """
import ctypes as C, os, subprocess, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from fastselect_b200._shard import joint_sharded, shard_triangle
from oracle import ref_oracle as R

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
from fastselect_b200._shard import enable_distributed
enable_distributed()          # sharding across the process group is opt-in
rs = np.random.RandomState(5)
x = rs.randint(0, 4, (90, 23)); y = rs.randint(0, 3, 90)
xa = np.concatenate([x, y[:, None]], axis=1)
q = xa.shape[1]
full = np.zeros((q, q))
rel, red = R.mi_matrices(x, y)
full[:q - 1, :q - 1] = red; full[q - 1, :q - 1] = rel; full[:q - 1, q - 1] = rel
calls = []
def compute_band(lo, hi, out_ptr):          # the oracle stands in for the GPU kernels in this CPU test
    assert out_ptr is None
    calls.append((lo, hi))
    band = np.zeros((q, q))
    for c in range(lo, hi):
        band[c, c + 1:] = full[c, c + 1:]
        band[c + 1:, c] = full[c + 1:, c]
    return band
total = joint_sharded(q, compute_band, device_buffers=False)
assert calls == [shard_triangle(q, 2, dist.get_rank())], calls
assert np.array_equal(total, full)
print("rank", dist.get_rank(), "ok", calls)
dist.destroy_process_group()
"""


def test_two_rank_gloo_band_sum_is_the_full_matrix(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "tests")]))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "rank 0 ok" in outs[0] and "rank 1 ok" in outs[1]


@pytest.fixture
def oracle_backend(monkeypatch):
    """The estimators' host glue without a GPU: the C-ABI call is replaced by the oracle (test only)."""
    from oracle import ref_oracle as R

    def fake_joint_matrix(codes, kind, log_base=1.0, stats_out=None):
        codes = np.asarray(codes)
        xc = np.stack([np.unique(codes[:, f], return_inverse=True)[1] for f in range(codes.shape[1])], axis=1)
        q = xc.shape[1]
        if kind == _native.FS_JOINT_MI:
            vec, mat = R.mi_matrices(xc[:, :-1], xc[:, -1], unit="nat")
            vec, mat = vec / log_base, mat / log_base
        else:
            vec, mat = R.su_matrices(xc[:, :-1], xc[:, -1])
        full = np.zeros((q, q))
        full[:q - 1, :q - 1] = mat
        full[q - 1, :q - 1] = vec
        full[:q - 1, q - 1] = vec
        return full

    monkeypatch.setattr(_mi, "joint_matrix", fake_joint_matrix)
    monkeypatch.setattr(_native, "device_count", lambda: 1)


@pytest.mark.parametrize("method", ["MID", "MIQ"])
def test_mrmr_fit_glue_reproduces_the_reference(g, oracle_backend, method):
    """tests/test_mrmr.py:53-101, :164-186 through fit / transform with the oracle as the backend."""
    for data in ("mrmr_fixture", "mrmr_dup", "geno", "states"):
        x, y = g[f"X_{data}"], g[f"y_{data}"]
        ref = g[f"mrmr_top_{method}_{data}"]
        est = fsb.mRMR(n_features_to_select=len(ref), method=method, backend="gpu").fit(x, y)
        assert np.array_equal(est.top_features_, ref), (data, method)
        np.testing.assert_allclose(est.relevance_scores_, g[f"mi_rel_bit_{data}"], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(est.redundancy_matrix_, g[f"mi_red_bit_{data}"], rtol=1e-9, atol=1e-14)
        assert est.feature_importances_ is est.relevance_scores_ and est.n_features_in_ == x.shape[1]
        assert np.array_equal(est.transform(x), x[:, ref])
        assert np.array_equal(est.unique_vals_, np.unique(np.concatenate([np.unique(x), np.unique(y)])))
    with pytest.raises(ValueError, match="n_features_to_select must be a positive integer"):
        fsb.mRMR(n_features_to_select=x.shape[1] + 1, backend="gpu").fit(x, y)
    with pytest.raises(ValueError, match=f"X has {x.shape[1] - 1} features, but mRMR is expecting {x.shape[1]} features"):
        est.transform(x[:, 1:])
    with pytest.raises(ValueError, match="integer-coded"):
        fsb.mRMR(n_features_to_select=2, backend="gpu").fit(x.astype(float), y)


def test_cfs_fit_glue_reproduces_the_reference(g, oracle_backend):
    """tests/test_cfs.py:77-105, :126-175 through fit / transform with the oracle as the backend."""
    import pandas as pd

    for data in ("cfs_fixture", "geno", "states", "mrmr_fixture"):
        x, y = g[f"X_{data}"], g[f"y_{data}"]
        est = fsb.CFS(backend="gpu").fit(x, y)
        assert np.array_equal(est.selected_indices_, g[f"cfs_sel_{data}"]), data
        assert est.merit_ == pytest.approx(float(g[f"cfs_merit_{data}"][0]), rel=1e-5, abs=1e-6)
        assert est.r_cf_.dtype == np.float32 and est.r_ff_.dtype == np.float32
        np.testing.assert_allclose(est.r_cf_, g[f"cfs_rcf_{data}"], rtol=0, atol=1e-6)
        assert np.array_equal(est.transform(x), x[:, g[f"cfs_sel_{data}"]])
        assert np.array_equal(est.get_support(indices=True), g[f"cfs_sel_{data}"])
    x, y = g["X_cfs_fixture"], g["y_cfs_fixture"]
    df = pd.DataFrame(x, columns=[f"feature_{i}" for i in range(x.shape[1])])
    est = fsb.CFS(backend="auto").fit(df, y)
    assert list(est.feature_names_in_) == list(df.columns)
    out = est.transform(df)
    assert isinstance(out, pd.DataFrame) and list(out.columns) == ["feature_0", "feature_2"]
    noise = fsb.CFS(backend="gpu").fit(x[:, 3:5], y)
    assert len(noise.selected_indices_) == 0 and noise.merit_ == 0.0 and noise.transform(x[:, 3:5]).shape[1] == 0
    single = fsb.CFS(backend="gpu").fit(x[:, [0]], y)
    assert single.selected_indices_.tolist() == [0] and single.merit_ > 0
    with pytest.raises(ValueError, match="up to 16 unique states/bins"):
        fsb.CFS(backend="gpu", n_bins=20).fit(x, y)
    with pytest.raises(ValueError, match="backend must be one of"):
        fsb.CFS(backend="tpu").fit(x, y)


def test_mi_function_glue(g, oracle_backend):
    x, y = g["X_states"], g["y_states"]
    mi = fsb.mutual_information
    for unit in ("bit", "nat"):
        rel, red = mi.calculate_mi_matrices(x, y, backend="auto", unit=unit)
        np.testing.assert_allclose(rel, g[f"mi_rel_{unit}_states"], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(red, g[f"mi_red_{unit}_states"], rtol=1e-9, atol=1e-14)
        assert red.flags["C_CONTIGUOUS"] and rel.shape == (x.shape[1],)
    v = mi.calculate_mi_single_pair(x[:, 0], x[:, 1], backend="gpu", unit="bit")
    assert v == pytest.approx(g["mi_red_bit_states"][0, 1], rel=1e-9, abs=1e-14)


def test_integration_stubs_bind_the_built_library(tmp_path):
    """The reference-side ctypes stubs printed in INTEGRATION.md are real code: they load the built
    library and call it with the documented argument lists.  Without a GPU every compute entry point
    must come back with the library's 'no usable GPU' error (no crash, no fallback); with one, the stub
    of the joint-count path reproduces the oracle."""
    import re

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "_b200.py" in b.splitlines()[0]]
    assert len(blocks) == 3          # single GPU, several GPUs of one process, joint-count path
    src = "\n".join(blocks).replace('C.CDLL("libfastselect_b200.so")', f"C.CDLL({_native.LIB_PATH!r})")
    ns = {}
    exec(compile(src, "INTEGRATION.md", "exec"), ns)
    rs = np.random.RandomState(0)
    x = rs.randint(0, 3, (40, 6)).astype(np.uint8)
    y = rs.randint(0, 2, 40)
    if not ns["available"]():
        with pytest.raises(RuntimeError, match="no usable NVIDIA sm_100"):
            ns["score"](x, y, 2, np.ones(6, bool), np.ones(6, np.float32), algo=2)
        with pytest.raises(RuntimeError, match="no usable NVIDIA sm_100"):
            ns["score_multi"](x, y, 2, np.ones(6, bool), np.ones(6, np.float32), algo=2, devices=[0, 1])
        with pytest.raises(RuntimeError, match="no usable NVIDIA sm_100"):
            ns["joint_matrix"](np.column_stack([x, y]), 0, np.log(2.0))
        return
    from oracle import ref_oracle as R

    m = ns["joint_matrix"](np.column_stack([x, y]), 0, np.log(2.0))
    rel, red = R.mi_matrices(x, y)
    np.testing.assert_allclose(m[-1, :-1], rel, rtol=1e-11, atol=1e-15)
    np.testing.assert_allclose(m[:-1, :-1], red, rtol=1e-11, atol=1e-15)
