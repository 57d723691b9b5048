"""GPU: the joint-count path through the C ABI -- contingency tables
bit-exact against the oracle, mutual information / symmetrical uncertainty against the oracle and the
reference's own outputs (tests/golden/joint_vectors.npz), the mRMR / CFS estimators end to end, band
ranges (emulated ranks), forced multi-band runs, and size-independent properties at a larger shape.

Tolerances: tables exact; MI rtol 1e-11 / atol 1e-15 against both the float64 oracle and the
reference (float64 there too); SU rtol 1e-11 against the float64 oracle and atol 1e-6 against the
reference, whose float32 intermediates carry +-1e-7 of noise (oracle/fs_oracle.c header)."""
import json
import os

import numpy as np
import pytest

import fastselect_b200 as fsb
from fastselect_b200 import _mi, _native
from oracle import ref_oracle as R

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
MI_TOL = dict(rtol=1e-11, atol=1e-15)


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))


def open_codes(native, codes):
    """A data set whose every column is discrete (what fastselect_b200._mi.joint_matrix opens)."""
    codes = np.ascontiguousarray(codes)
    ds = native.Dataset(codes, np.zeros(codes.shape[0], np.int32), 1)
    arith = native.FS_ARITH_F64 if codes.dtype == np.float64 else native.FS_ARITH_F32
    ds.set_features(np.ones(codes.shape[1], np.uint8), np.ones(codes.shape[1], np.float32), arith)
    return ds


def per_column_codes(x):
    return np.stack([np.unique(x[:, f], return_inverse=True)[1] for f in range(x.shape[1])], axis=1).astype(np.int32)


def oracle_full(x, y, kind):
    """[q, q] matrix of [x | y] from the oracle (q = p + 1)."""
    xc = per_column_codes(np.concatenate([x, y[:, None]], axis=1))
    q = xc.shape[1]
    vec, mat = (R.mi_matrices if kind == 0 else R.su_matrices)(xc[:, :-1], xc[:, -1])
    full = np.zeros((q, q))
    full[:q - 1, :q - 1] = mat
    full[q - 1, :q - 1] = vec
    full[:q - 1, q - 1] = vec
    return full


def many_states(seed, n, p, dtype=np.uint8, spread=1):
    """2..16 states per column incl. constants, duplicates, non-contiguous values."""
    rs = np.random.RandomState(seed)
    x = np.empty((n, p), np.int64)
    for f in range(p):
        x[:, f] = rs.randint(0, 2 + f % 15, n) * spread + (f % 3)
    x[:, 3 % p] = 7
    if p > 9:
        x[:, 9] = x[:, 2] * 2 + 1
    y = rs.randint(0, 3, n)
    return x.astype(dtype), y.astype(dtype)


CASES = [
    ("geno_lean", lambda: (np.random.RandomState(1).randint(0, 3, (300, 64)).astype(np.uint8),
                           np.random.RandomState(2).randint(0, 2, 300).astype(np.uint8))),
    ("geno_ragged", lambda: (np.random.RandomState(3).randint(0, 3, (257, 70)).astype(np.int8),
                             np.random.RandomState(4).randint(0, 3, 257).astype(np.int8))),
    ("states_u8", lambda: many_states(5, 333, 50)),
    ("states_f32", lambda: many_states(6, 129, 40, np.float32, spread=1000)),
    ("states_f64", lambda: many_states(7, 200, 33, np.float64, spread=(1 << 30))),
    ("tiny", lambda: (np.array([[0, 1], [1, 1], [0, 0], [1, 0], [2, 1]], np.uint8), np.array([0, 1, 0, 1, 1], np.uint8))),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_tables_are_bit_exact(native, name, make):
    x, y = make()
    xa = np.concatenate([x, y[:, None]], axis=1)
    xc = per_column_codes(xa)
    q = xa.shape[1]
    rs = np.random.RandomState(0)
    pairs = np.array([(a, b) for a in range(q) for b in range(q) if a != b])
    if len(pairs) > 400:
        pairs = pairs[rs.choice(len(pairs), 400, replace=False)]
    with open_codes(native, xa) as ds:
        tables = ds.joint_tables(pairs)
    for (a, b), t in zip(pairs, tables):
        ref = R.joint_counts(xc[:, a], xc[:, b])
        assert np.array_equal(t[:ref.shape[0], :ref.shape[1]], ref), (name, a, b)
        assert t.sum() == x.shape[0]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("kind", [0, 1], ids=["mi", "su"])
def test_matrix_matches_oracle(native, name, make, kind):
    x, y = make()
    xa = np.concatenate([x, y[:, None]], axis=1)
    with open_codes(native, xa) as ds:
        m, st = ds.joint_matrix(kind, np.log(2.0), want_stats=True)
    ref = oracle_full(x, y, kind)
    np.testing.assert_allclose(m, ref, **MI_TOL)
    assert np.array_equal(m, m.T) and not m.diagonal().any()
    assert st["launches"] >= 3 and st["ops_dist_tensor"] > 0


def test_forced_bands_and_position_ranges(native, monkeypatch):
    """A 1 MB slab forces 128-row bands; two position ranges (emulated ranks) sum to the full matrix;
    a column subset in a different order gives the permuted matrix."""
    x, y = many_states(11, 400, 120)
    xa = np.concatenate([x, y[:, None]], axis=1)
    q = xa.shape[1]
    with open_codes(native, xa) as ds:
        full = ds.joint_matrix(0, 1.0)
        monkeypatch.setenv("FS_B200_JOINT_SLAB_MB", "1")
        banded, st = ds.joint_matrix(0, 1.0, want_stats=True)
        assert st["n_chunks"] > 1
        assert np.array_equal(banded, full)
        from fastselect_b200._shard import shard_triangle
        parts = [ds.joint_matrix(0, 1.0, pos_begin=lo, pos_end=hi)
                 for lo, hi in (shard_triangle(q, 3, r) for r in range(3))]
        assert np.array_equal(parts[0] + parts[1] + parts[2], full)
        monkeypatch.delenv("FS_B200_JOINT_SLAB_MB")
        idx = np.random.RandomState(1).permutation(q)[:57]
        sub = ds.joint_matrix(0, 1.0, feat_idx=idx)
        np.testing.assert_allclose(sub, full[np.ix_(idx, idx)], rtol=1e-13, atol=1e-16)
    np.testing.assert_allclose(full, oracle_full(x, y, 0) * np.log(2.0), **MI_TOL)


def test_mi_matches_the_reference_vectors(native, g):
    for data in ("mrmr_fixture", "mrmr_dup", "geno", "states"):
        x, y = g[f"X_{data}"], g[f"y_{data}"]
        for unit in ("bit", "nat"):
            rel, red = fsb.mutual_information.calculate_mi_matrices(x, y, backend="gpu", unit=unit)
            np.testing.assert_allclose(rel, g[f"mi_rel_{unit}_{data}"], **MI_TOL)
            np.testing.assert_allclose(red, g[f"mi_red_{unit}_{data}"], **MI_TOL)
        a = fsb.mutual_information.calculate_mi_single_pair(x[:, 0], x[:, 1], backend="gpu")
        assert a == pytest.approx(g[f"mi_red_bit_{data}"][0, 1], rel=1e-11, abs=1e-15)


def test_su_matches_the_reference_vectors(native, g):
    for data in ("cfs_fixture", "geno", "states", "mrmr_fixture"):
        codes = g[f"cfs_codes_{data}"]
        y = np.unique(g[f"y_{data}"], return_inverse=True)[1]
        p = codes.shape[1]
        su = _mi.joint_matrix(_mi._stack_for_upload(codes, y), native.FS_JOINT_SU)
        np.testing.assert_allclose(su[p, :p], g[f"cfs_rcf_{data}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(su[:p, :p], g[f"cfs_rff_{data}"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("method", ["MID", "MIQ"])
def test_mrmr_estimator_matches_reference(native, g, method):
    """tests/test_mrmr.py:53-101 (fit / transform / attributes) with the reference's selections."""
    for data in ("mrmr_fixture", "mrmr_dup", "geno", "states"):
        x, y = g[f"X_{data}"], g[f"y_{data}"]
        ref = g[f"mrmr_top_{method}_{data}"]
        est = fsb.mRMR(n_features_to_select=len(ref), method=method, backend="gpu").fit(x, y)
        assert np.array_equal(est.top_features_, ref), (data, method)
        assert est.relevance_scores_.shape == (x.shape[1],) and est.redundancy_matrix_.shape == (x.shape[1],) * 2
        assert est.feature_importances_ is est.relevance_scores_
        assert est.transform(x).shape == (x.shape[0], len(ref))
        assert np.array_equal(fsb.mRMR(len(ref), method, "gpu").fit_transform(x, y), x[:, ref])
    with pytest.raises(ValueError, match="n_features_to_select must be a positive integer"):
        fsb.mRMR(n_features_to_select=x.shape[1] + 1, backend="gpu").fit(x, y)
    with pytest.raises(ValueError, match="integer-coded"):
        fsb.mRMR(n_features_to_select=2, backend="gpu").fit(x.astype(float), y)


def test_cfs_estimator_matches_reference(native, g):
    """tests/test_cfs.py:77-105, :126-160 with the reference's selections and merit."""
    import pandas as pd

    for data in ("cfs_fixture", "geno", "states", "mrmr_fixture"):
        x, y = g[f"X_{data}"], g[f"y_{data}"]
        est = fsb.CFS(backend="gpu").fit(x, y)
        assert np.array_equal(est.selected_indices_, g[f"cfs_sel_{data}"]), data
        assert est.merit_ == pytest.approx(float(g[f"cfs_merit_{data}"][0]), rel=1e-5, abs=1e-6)
        assert est.n_features_in_ == x.shape[1] and est.support_mask_.sum() == len(est.selected_indices_)
        assert np.array_equal(est.transform(x), x[:, g[f"cfs_sel_{data}"]])
    x, y = g["X_cfs_fixture"], g["y_cfs_fixture"]
    assert fsb.CFS(backend="auto").fit(x, y).selected_indices_.tolist() == [0, 2]
    df = pd.DataFrame(x, columns=[f"feature_{i}" for i in range(x.shape[1])])
    est = fsb.CFS(backend="gpu").fit(df, y)
    assert list(est.transform(df).columns) == ["feature_0", "feature_2"] and hasattr(est, "feature_names_in_")
    noise = fsb.CFS(backend="gpu").fit(x[:, 3:5], y)                     # noise + constant column
    assert len(noise.selected_indices_) == 0 and noise.merit_ == 0.0 and noise.transform(x[:, 3:5]).shape[1] == 0
    assert fsb.CFS(backend="gpu").fit(x[:, [0]], y).selected_indices_.tolist() == [0]
    with pytest.raises(ValueError, match="up to 16 unique states/bins"):
        fsb.CFS(backend="gpu", n_bins=20).fit(x, y)


def test_too_many_states_is_an_error(native):
    x = np.arange(40, dtype=np.int64).reshape(20, 2) % 17
    with pytest.raises(ValueError, match="up to 16 unique states"):
        fsb.mutual_information.calculate_mi_matrices(x, np.arange(20) % 2, backend="gpu")
    with _native.Dataset(np.random.RandomState(0).rand(30, 3).astype(np.float32), np.zeros(30, np.int32), 1) as ds:
        ds.set_features(np.zeros(3, np.uint8), np.ones(3, np.float32), _native.FS_ARITH_F32)
        with pytest.raises(ValueError, match="not a discrete column"):
            ds.joint_matrix(0, 1.0)


def test_properties_at_a_larger_shape(native):
    """n = 5000, 1500 genotype columns (K = 3000 reduced rows, several 256-wide column tiles and row
    tiles): symmetry, zero diagonal, I(f; f') = H(f) for a duplicated column, I >= 0 up to the guard,
    every table sums to n, and 300 random pairs against the oracle."""
    rs = np.random.RandomState(21)
    n, p = 5000, 1500
    x = rs.randint(0, 3, (n, p)).astype(np.uint8)
    x[:, 700] = x[:, 5]
    y = ((x[:, 5] == 1) ^ (rs.random_sample(n) < 0.2)).astype(np.uint8)
    rel, red = fsb.mutual_information.calculate_mi_matrices(x, y, backend="gpu")
    assert np.array_equal(red, red.T) and not red.diagonal().any()
    assert red.min() > -1e-10 and rel.min() > -1e-10
    pr = np.bincount(x[:, 5]) / n
    assert red[5, 700] == pytest.approx(-(pr * np.log2(pr)).sum(), rel=1e-9)
    assert int(np.argmax(rel)) in (5, 700)
    pairs = rs.randint(0, p, (300, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    xi = x.astype(np.int32)
    for a, b in pairs:
        ref = R.mi_matrices(xi[:, [a]], xi[:, b], want_matrix=False)[0][0]
        assert red[a, b] == pytest.approx(ref, rel=1e-10, abs=1e-14)
    with open_codes(native, x) as ds:
        t = ds.joint_tables(pairs)
    assert (t.reshape(len(pairs), -1).sum(axis=1) == n).all()
