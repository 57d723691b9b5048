"""Golden vectors for the joint-count path, made by running the REFERENCE.

Run in the authoring container only (needs /root/reference, numba, sklearn):

    python tests/golden/make_golden_joint.py

Stores, per data set, the inputs and the reference CPU path's outputs:
  * mutual_information.calculate_mi_matrices(backend="cpu") in bits and nats  (mutual_information.py:158-196)
  * mRMR(method=MID / MIQ).top_features_                                      (mRMR.py:66-136)
  * CFS._precompute_correlations_cpu's r_cf / r_ff (float32)                  (CFS.py:81-104)
  * CFS(backend="cpu").selected_indices_ / merit_                             (CFS.py:296-401)
Data sets "mrmr_fixture", "mrmr_dup" and "cfs_fixture" rebuild the fixtures of the reference's own
tests (tests/test_mrmr.py:13-34, :104-131; tests/test_cfs.py:9-56); nothing under /root/reference is copied.
"""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")
from fast_select import mutual_information as ref_mi  # noqa: E402
from fast_select.CFS import CFS, _precompute_correlations_cpu  # noqa: E402
from fast_select.mRMR import mRMR  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def mrmr_fixture():
    from sklearn.datasets import make_classification

    x, y = make_classification(n_samples=100, n_features=20, n_informative=5, n_redundant=2, n_classes=4,
                               random_state=42)
    edges = np.percentile(x, [25, 50, 75], axis=0)
    xd = np.empty_like(x, dtype=np.int64)
    for f in range(x.shape[1]):
        xd[:, f] = np.digitize(x[:, f], bins=edges[:, f])
    return xd, y.astype(np.int64)


def mrmr_dup():
    rng = np.random.default_rng(42)
    n, p = 200, 10
    y = rng.integers(0, 2, n)
    x = rng.integers(0, 3, size=(n, p))
    x[:, 0] = (y + (rng.random(n) < 0.10).astype(int)) % 2
    x[:, 1] = x[:, 0]
    x[:, 9] = (y + (rng.random(n) < 0.05).astype(int)) % 2
    return x.astype(np.int64), y.astype(np.int64)


def genotype(seed=11, n=300, p=64):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 3, (n, p))
    y = ((x[:, 3] == 1) & (x[:, 7] != 2)).astype(np.int64) ^ (rs.random_sample(n) < 0.1)
    x[:, 20] = x[:, 3]                       # duplicate of a relevant column
    return x.astype(np.int64), y.astype(np.int64)


def states(seed=12, n=257, p=50, n_classes=3):
    """2..12 states per column, a constant column, a non-contiguous value set and a copy."""
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, n)
    x = np.empty((n, p), np.int64)
    for f in range(p):
        x[:, f] = rs.randint(0, 2 + f % 11, n)
    x[:, 5] = (y + rs.randint(0, 2, n)) % 4
    x[:, 6] = 3                              # constant
    x[:, 7] = 2 * x[:, 5] + 1                # same partition, other values (1, 3, 5, 7)
    x[:, 8] = y
    return x, y.astype(np.int64)


def cfs_fixture():
    np.random.seed(42)
    n = 200
    y = np.random.randint(0, 2, n)
    f0 = y + np.random.normal(0, 0.1, n)
    f1 = f0 + np.random.normal(0, 0.05, n)
    f2 = y + np.random.normal(0, 0.5, n)
    f2[y == 0] -= 0.5
    f3 = np.random.rand(n) * 10
    f4 = np.full(n, 5.0)
    np.random.randint(0, 40, n)              # the fixture's sixth column (unused: > 32 states)
    return np.vstack([f0, f1, f2, f3, f4]).T, y.astype(np.int64)


def cfs_encode(x, y, n_bins=10, strategy="uniform"):
    """The coding CFS.fit applies before the correlations (CFS.py:319-337), restated."""
    from sklearn.preprocessing import KBinsDiscretizer

    p = x.shape[1]
    enc = np.zeros(x.shape, np.int32)
    n_states = np.zeros(p, np.int32)
    if np.issubdtype(x.dtype, np.floating):
        enc[:] = KBinsDiscretizer(n_bins=n_bins, encode="ordinal", strategy=strategy, subsample=None).fit_transform(x)
        n_states[:] = n_bins
    else:
        for f in range(p):
            u, inv = np.unique(x[:, f], return_inverse=True)
            enc[:, f] = inv
            n_states[f] = len(u)
    uy, y_enc = np.unique(y, return_inverse=True)
    return enc, n_states, y_enc.astype(np.int32), len(uy)


MI_DATA = {"mrmr_fixture": mrmr_fixture, "mrmr_dup": mrmr_dup, "geno": genotype, "states": states}
CFS_DATA = {"cfs_fixture": cfs_fixture, "geno": genotype, "states": states, "mrmr_fixture": mrmr_fixture}


def main():
    arrays, meta = {}, []
    for name, fn in {**MI_DATA, **CFS_DATA}.items():
        x, y = fn()
        arrays[f"X_{name}"] = x
        arrays[f"y_{name}"] = y
    for name in MI_DATA:
        x, y = arrays[f"X_{name}"], arrays[f"y_{name}"]
        # mRMR.fit's coding (mRMR.py:90-92) is what calculate_mi_matrices sees
        u = np.unique(np.concatenate([np.unique(x), np.unique(y)]))
        xe, ye = np.searchsorted(u, x), np.searchsorted(u, y)
        for unit in ("bit", "nat"):
            rel, red = ref_mi.calculate_mi_matrices(xe, ye, backend="cpu", unit=unit)
            arrays[f"mi_rel_{unit}_{name}"] = rel
            arrays[f"mi_red_{unit}_{name}"] = red
        n_sel = min(8, x.shape[1])
        for method in ("MID", "MIQ"):
            est = mRMR(n_features_to_select=n_sel, method=method, backend="cpu").fit(x.copy(), y.copy())
            arrays[f"mrmr_top_{method}_{name}"] = np.asarray(est.top_features_)
        meta.append(dict(kind="mi", data=name, n_select=n_sel))
    for name in CFS_DATA:
        x, y = arrays[f"X_{name}"], arrays[f"y_{name}"]
        enc, n_states, y_enc, ky = cfs_encode(x, y)
        r_cf, r_ff = _precompute_correlations_cpu(enc, y_enc, n_states, ky)
        est = CFS(backend="cpu", n_jobs=1).fit(x.copy(), y.copy())
        arrays[f"cfs_codes_{name}"] = enc
        arrays[f"cfs_rcf_{name}"] = r_cf
        arrays[f"cfs_rff_{name}"] = r_ff
        arrays[f"cfs_sel_{name}"] = np.asarray(est.selected_indices_, np.int64)
        arrays[f"cfs_merit_{name}"] = np.array([est.merit_], np.float64)
        meta.append(dict(kind="cfs", data=name))
    import numba
    import sklearn
    meta.append(dict(kind="versions", numba=numba.__version__, sklearn=sklearn.__version__, numpy=np.__version__))
    np.savez_compressed(os.path.join(HERE, "joint_vectors.npz"), **arrays)
    with open(os.path.join(HERE, "joint_vectors.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"{len(meta) - 1} cases written")


if __name__ == "__main__":
    main()
