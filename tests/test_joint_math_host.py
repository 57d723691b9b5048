"""CPU: the per-pair arithmetic of the joint-count kernels (fastselect_b200/csrc/joint_math.cuh), built
for the host by tests/helpers/joint_math_host.cpp and fed with what the device pipeline hands it -- the
negated reduced one-hot count matrix -(At At^T), the reduced-row offsets and the marginals -- against
the oracle and the reference's golden vectors.  (The GPU suite checks the kernels themselves.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import ref_oracle as R

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def jm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("jm") / "libjm.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out,
                    os.path.join(HERE, "helpers", "joint_math_host.cpp")], check=True)
    return C.CDLL(out)


def reduced_onehot(x):
    """What onehot.cu builds: per column the indicator rows of its first V - 1 values (ascending);
    a constant column owns no row (an empty range).  Returns (A [K, n] int32, toff [p + 1], columns)."""
    rows, toff, kept = [], [0], []
    for f in range(x.shape[1]):
        vals = np.unique(x[:, f])
        kept.append(f)
        for v in vals[:-1]:
            rows.append((x[:, f] == v).astype(np.int32))
        toff.append(len(rows))
    return np.array(rows, np.int32), np.array(toff, np.int32), kept


def device_view(x, y):
    """Inputs of the finishing kernel for the matrix [x | y]."""
    xa = np.concatenate([x, y[:, None]], axis=1)
    a, toff, kept = reduced_onehot(xa)
    neg_c = np.ascontiguousarray(-(a @ a.T), np.int32)
    marg = np.ascontiguousarray(a.sum(axis=1), np.int32)
    return xa, neg_c, toff, marg, kept


def finish(jm, neg_c, toff, marg, n, kind, log_base):
    pt = toff.size - 1
    out = np.full((pt, pt), np.nan)
    ptr = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    jm.jm_finish_host(ptr(neg_c, C.c_int32), C.c_int64(neg_c.shape[1]), ptr(toff, C.c_int32), C.c_int64(pt),
                      ptr(marg, C.c_int32), C.c_int64(n), C.c_int(kind), C.c_double(log_base), ptr(out, C.c_double))
    return out


@pytest.mark.parametrize("data", ["states", "geno", "mrmr_fixture", "mrmr_dup"])
def test_mi_from_reduced_counts_matches_reference(jm, data):
    g = np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))
    x, y = g[f"X_{data}"], g[f"y_{data}"]
    xa, neg_c, toff, marg, kept = device_view(x, y)
    p = x.shape[1]
    for unit, lb in (("bit", np.log(2.0)), ("nat", 1.0)):
        m = finish(jm, neg_c, toff, marg, x.shape[0], 0, lb)
        full = np.zeros((p + 1, p + 1))
        full[np.ix_(kept, kept)] = m
        ref_rel, ref_red = g[f"mi_rel_{unit}_{data}"], g[f"mi_red_{unit}_{data}"]
        # constant columns included: the reference's value there is its 1e-12 guard residue
        np.testing.assert_allclose(full[p, :p], ref_rel, rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(full[:p, :p], ref_red, rtol=1e-11, atol=1e-15)


@pytest.mark.parametrize("data", ["states", "geno", "cfs_fixture", "mrmr_fixture"])
def test_su_from_reduced_counts_matches_reference(jm, data):
    g = np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))
    codes = g[f"cfs_codes_{data}"].astype(np.int64)
    y = np.unique(g[f"y_{data}"], return_inverse=True)[1]
    xa, neg_c, toff, marg, kept = device_view(codes, y)
    p = codes.shape[1]
    m = finish(jm, neg_c, toff, marg, codes.shape[0], 1, 1.0)
    full = np.zeros((p + 1, p + 1))
    full[np.ix_(kept, kept)] = m
    np.testing.assert_allclose(full[p, :p], g[f"cfs_rcf_{data}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(full[:p, :p], g[f"cfs_rff_{data}"], rtol=0, atol=1e-6)
    # and the float64 oracle to rounding
    o_cf, o_ff = R.su_matrices(codes, y)
    np.testing.assert_allclose(full[p, :p], o_cf, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(full[:p, :p], o_ff, rtol=1e-12, atol=1e-14)


def test_tables_rebuilt_from_reduced_counts_are_exact(jm):
    g = np.load(os.path.join(HERE, "golden", "joint_vectors.npz"))
    x, y = g["X_states"], g["y_states"]
    xa, neg_c, toff, marg, kept = device_view(x, y)
    ptr = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rs = np.random.RandomState(0)
    for _ in range(200):
        c, h = sorted(rs.choice(len(kept), 2, replace=False))
        table = np.zeros((16, 16), np.int64)
        jm.jm_table_host(ptr(neg_c, C.c_int32), C.c_int64(neg_c.shape[1]), ptr(toff, C.c_int32), C.c_int64(c),
                         C.c_int64(h), ptr(marg, C.c_int32), C.c_int64(x.shape[0]), ptr(table, C.c_int64))
        ca = np.unique(xa[:, kept[c]], return_inverse=True)[1]
        cb = np.unique(xa[:, kept[h]], return_inverse=True)[1]
        ref = R.joint_counts(ca, cb)
        assert np.array_equal(table[:ref.shape[0], :ref.shape[1]], ref)
