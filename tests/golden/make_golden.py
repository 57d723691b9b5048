"""Generate the golden vectors in tests/golden/ by running the REFERENCE itself.

Run in the authoring container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

Each case stores the inputs, the constructor parameters and the reference CPU
backend's ``feature_importances_`` / ``top_features_`` / ``is_discrete_``.  The
fixtures "A" and "B" are the hand-written matrices of the reference's own tests
(tests/test_multisurf.py:19-33, tests/test_surf.py:22-32, tests/test_relieff.py:21-31);
the rest are seeded synthetic data.  Nothing under /root/reference is copied.
"""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, "/root/reference/src")
warnings.filterwarnings("ignore")
from fast_select import MultiSURF, SURF, ReliefF, TuRF  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fixture_a():
    x = np.array([[1.1, 5.0, 10, 3.0], [1.2, 4.0, 10, 3.0], [2.3, 6.0, 10, 3.0], [2.5, 5.5, 10, 3.0],
                  [1.5, 4.5, 20, 3.0], [8.8, 5.0, 20, 3.0], [8.9, 4.0, 20, 3.0], [9.5, 6.0, 20, 3.0],
                  [10.5, 4.5, 20, 3.0], [10.5, 4.5, 10, 3.0]], dtype=np.float32)
    y = np.array([0] * 5 + [1] * 5, dtype=np.int32)
    return x, y


def fixture_b():
    x = np.array([[0.1, 5.0, 10, 3.0], [0.2, 4.0, 10, 3.0], [0.3, 6.0, 10, 3.0],
                  [10.8, 5.0, 20, 3.0], [10.9, 4.0, 20, 3.0], [11.0, 6.0, 20, 3.0]], dtype=np.float32)
    y = np.array([0, 0, 0, 1, 1, 1], dtype=np.int32)
    return x, y


def genotype(seed, n, p, n_classes=2):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 3, (n, p)).astype(np.float64)
    risk = (x[:, 1] == 1) & (x[:, 3] == 1)
    y = rs.randint(0, n_classes, n)
    y[risk] = 1
    return x, y.astype(np.int64)


def gaussian(seed, n, p, n_classes=2):
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, n)
    x = rs.standard_normal((n, p))
    x[:, 0] += 1.5 * y
    x[:, 2] -= 1.0 * y
    return x, y.astype(np.int64)


def mixed(seed, n, p, n_classes=3):
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, n)
    x = np.empty((n, p))
    h = p // 2
    x[:, :h] = rs.randint(0, 3, (n, h))
    x[:, h:] = rs.standard_normal((n, p - h))
    x[:, 0] = (y + rs.randint(0, 2, n)) % 3
    x[:, h] += 1.2 * y
    x[:, p - 1] = 7.0  # constant column
    return x, y.astype(np.int64)


DATA = {
    "A": fixture_a, "B": fixture_b,
    "geno2": lambda: genotype(1, 64, 48, 2),
    "geno3": lambda: genotype(2, 60, 40, 3),
    "gauss2": lambda: gaussian(3, 80, 30, 2),
    "gauss3": lambda: gaussian(4, 70, 25, 3),
    "mixed3": lambda: mixed(5, 72, 36, 3),
    "mixed2_mid": lambda: mixed(6, 300, 200, 2),
    "geno2_mid": lambda: genotype(7, 256, 384, 2),
    # SURF's float32 mean distance is summation-order sensitive here (2 neighbour pairs flip
    # between a sequential / correctly-rounded sum and the order the reference really uses)
    "gauss2_long": lambda: gaussian(2, 1500, 4, 2),
}

CASES = []
for d, dls in (("A", (4, 10)), ("B", (3, 10))):
    for dl in dls:
        for star in (False, True):
            CASES.append((d, "MultiSURF", dict(discrete_limit=dl, use_star=star)))
            CASES.append((d, "SURF", dict(discrete_limit=dl, use_star=star)))
        for k in (1, 2, 3):
            CASES.append((d, "ReliefF", dict(discrete_limit=dl, n_neighbors=k)))
for d in ("geno2", "geno3", "gauss2", "gauss3", "mixed3", "mixed2_mid", "geno2_mid"):
    for star in (False, True):
        CASES.append((d, "MultiSURF", dict(use_star=star)))
        CASES.append((d, "SURF", dict(use_star=star)))
    for k in (1, 5, 10):
        CASES.append((d, "ReliefF", dict(n_neighbors=k)))
for star in (False, True):
    CASES.append(("gauss2_long", "SURF", dict(use_star=star)))
    CASES.append(("gauss2_long", "MultiSURF", dict(use_star=star)))
CASES.append(("gauss2_long", "ReliefF", dict(n_neighbors=10)))
CASES.append(("gauss2", "MultiSURF", dict(discrete_limit=100)))   # n <= limit: every column discrete
CASES.append(("mixed3", "SURF", dict(discrete_limit=2, use_star=True)))

CLS = {"MultiSURF": MultiSURF, "SURF": SURF, "ReliefF": ReliefF}


def main():
    arrays, meta = {}, []
    for name, fn in DATA.items():
        x, y = fn()
        arrays[f"X_{name}"] = x
        arrays[f"y_{name}"] = y
    for idx, (d, algo, params) in enumerate(CASES):
        x, y = arrays[f"X_{d}"], arrays[f"y_{d}"]
        if algo == "ReliefF" and params["n_neighbors"] >= x.shape[0]:
            continue
        est = CLS[algo](n_features_to_select=min(3, x.shape[1]), backend="cpu", n_jobs=1, **params)
        est.fit(x.copy(), y.copy())
        arrays[f"scores_{idx}"] = np.asarray(est.feature_importances_)
        arrays[f"top_{idx}"] = np.asarray(est.top_features_)
        arrays[f"isd_{idx}"] = np.asarray(est.is_discrete_)
        meta.append(dict(idx=idx, data=d, algo=algo, params=params))
    # TuRF through the reference's own estimators
    for d, algo, params in (("geno2_mid", "MultiSURF", {}), ("mixed2_mid", "SURF", {}),
                            ("gauss2", "ReliefF", dict(n_neighbors=5))):
        x, y = arrays[f"X_{d}"], arrays[f"y_{d}"]
        t = TuRF(CLS[algo](n_features_to_select=5, backend="cpu", n_jobs=1, **params),
                 n_features_to_select=5, pct_remove=0.2)
        t.fit(x.copy(), y.copy())
        idx = len(meta) + 1000
        arrays[f"scores_{idx}"] = np.asarray(t.feature_importances_)
        arrays[f"top_{idx}"] = np.asarray(t.top_features_)
        meta.append(dict(idx=idx, data=d, algo="TuRF:" + algo, params=params, turf=dict(n=5, pct=0.2)))
    # README quick-start (README.md:79-89): inputs are regenerated from the seed, only outputs stored
    from sklearn.datasets import make_classification
    import sklearn
    x, y = make_classification(n_samples=500, n_features=1000, n_informative=20, n_redundant=100, random_state=42)
    est = MultiSURF(n_features_to_select=15, backend="cpu", n_jobs=1).fit(x, y)
    arrays["readme_scores"] = est.feature_importances_
    arrays["readme_top"] = est.top_features_
    arrays["readme_xsum"] = np.array([x.sum(), np.abs(x).sum(), float(y.sum())])
    meta.append(dict(idx=-1, data="readme", algo="MultiSURF", params=dict(n_features_to_select=15),
                     sklearn=sklearn.__version__))
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **arrays)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as fh:
        json.dump(meta, fh, indent=1)
    print(f"{len(meta)} cases written")


if __name__ == "__main__":
    main()
