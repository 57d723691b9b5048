"""Seconds-long GPU check of the joint-count kernels against the oracle, with minimal imports (ctypes
binding loaded from its file: no sklearn / torch).  python tools/check_joint_quick.py [--time]"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("fs_native", os.path.join(ROOT, "fastselect_b200", "_native.py"))
N = importlib.util.module_from_spec(spec)
spec.loader.exec_module(N)
from oracle import ref_oracle as R  # noqa: E402


def codes_of(x):
    return np.stack([np.unique(x[:, f], return_inverse=True)[1] for f in range(x.shape[1])], axis=1).astype(np.int32)


def oracle_full(xa, kind):
    xc = codes_of(xa)
    q = xc.shape[1]
    vec, mat = (R.mi_matrices if kind == 0 else R.su_matrices)(xc[:, :-1], xc[:, -1])
    full = np.zeros((q, q))
    full[:q - 1, :q - 1] = mat
    full[q - 1, :q - 1] = vec
    full[:q - 1, q - 1] = vec
    return full


def open_codes(xa):
    ds = N.Dataset(np.ascontiguousarray(xa), np.zeros(xa.shape[0], np.int32), 1)
    ds.set_features(np.ones(xa.shape[1], np.uint8), np.ones(xa.shape[1], np.float32),
                    N.FS_ARITH_F64 if xa.dtype == np.float64 else N.FS_ARITH_F32)
    return ds


def states(seed, n, p, dtype, spread=1):
    rs = np.random.RandomState(seed)
    x = np.empty((n, p + 1), np.int64)
    for f in range(p + 1):
        x[:, f] = rs.randint(0, 2 + f % 15, n) * spread + (f % 3)
    x[:, 3] = 7
    x[:, 9] = x[:, 2] * 2 + 1
    return x.astype(dtype)


rs = np.random.RandomState(0)
cases = {"geno": rs.randint(0, 3, (300, 65)).astype(np.uint8), "states": states(5, 333, 50, np.uint8),
         "f64": states(7, 200, 33, np.float64, 1 << 30), "wide": rs.randint(0, 3, (1000, 701)).astype(np.uint8)}
t0 = time.perf_counter()
for name, xa in cases.items():
    with open_codes(xa) as ds:
        for kind in (0, 1):
            m = ds.joint_matrix(kind, np.log(2.0))
            np.testing.assert_allclose(m, oracle_full(xa, kind), rtol=1e-11, atol=1e-15, err_msg=f"{name} kind {kind}")
            assert np.array_equal(m, m.T) and not m.diagonal().any(), name
        q = xa.shape[1]
        pairs = rs.randint(0, q, (64, 2))
        pairs = pairs[pairs[:, 0] != pairs[:, 1]]
        t = ds.joint_tables(pairs)
        xc = codes_of(xa)
        for (a, b), tab in zip(pairs, t):
            ref = R.joint_counts(xc[:, a], xc[:, b])
            assert np.array_equal(tab[:ref.shape[0], :ref.shape[1]], ref) and tab.sum() == xa.shape[0], (name, a, b)
    print("ok", name, xa.shape, flush=True)
xa = states(11, 400, 120, np.uint8)
with open_codes(xa) as ds:
    full = ds.joint_matrix(0, 1.0)
    os.environ["FS_B200_JOINT_SLAB_MB"] = "1"
    banded, st = ds.joint_matrix(0, 1.0, want_stats=True)
    parts = [ds.joint_matrix(0, 1.0, pos_begin=lo, pos_end=hi) for lo, hi in ((0, 17), (17, 60), (60, 121))]
    del os.environ["FS_B200_JOINT_SLAB_MB"]
    assert st["n_chunks"] > 1 and np.array_equal(banded, full) and np.array_equal(sum(parts), full)
print(f"ok bands / position ranges; all checks {time.perf_counter() - t0:.2f} s", flush=True)
if "--time" in sys.argv:
    n = p = 10000
    xa = np.empty((n, p + 1), np.uint8)
    xa[:] = np.random.RandomState(1).randint(0, 3, (n, p + 1))
    xa[:, -1] &= 1
    with open_codes(xa) as ds:
        for rep in range(3):
            ds.set_features(np.ones(p + 1, np.uint8), np.ones(p + 1, np.float32), N.FS_ARITH_F32)
            _, st = ds.joint_matrix(0, np.log(2.0), want_stats=True)
            print({k: round(v, 3) for k, v in st.items() if k in ("ms_total", "ms_gather", "ms_dist_tensor", "ms_reduce")}, flush=True)
