"""Debugging aid: where the time of one end-to-end MultiSURF.fit (host buffers) goes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from sklearn.utils.validation import validate_data  # noqa: E402

import bench  # noqa: E402
import fastselect_b200 as fsb  # noqa: E402
from fastselect_b200 import _native, _relief  # noqa: E402

w = bench.make_workload("c3", 1, "weak")
x = torch.from_numpy(w["x"]).pin_memory().numpy()
y = w["y"]
fsb.MultiSURF(n_features_to_select=10, backend="gpu").fit(x, y)      # warm


def T(label, t0):
    t1 = time.perf_counter()
    print(f"{label:28s} {1e3 * (t1 - t0):8.3f} ms")
    return t1


for rep in range(2):
    print("--- rep", rep)
    est = fsb.MultiSURF(n_features_to_select=10, backend="gpu")
    t00 = t = time.perf_counter()
    xv, yv = validate_data(est, x, y, y_numeric=True, dtype=[np.float32, np.int8, np.uint8], ensure_2d=True)
    t = T("validate_data", t)
    y_enc = np.unique(yv, return_inverse=True)[1].astype(np.int32)
    t = T("np.unique(y)", t)
    ds = _native.Dataset(xv, y_enc, int(y_enc.max()) + 1)
    t = T("Dataset (upload + scan)", t)
    cmin, cmax, cnt = ds.column_stats()
    t = T("column_stats", t)
    ranges = cmax.astype(np.float32) - cmin.astype(np.float32)
    ranges[ranges == 0] = 1
    recip = (1.0 / ranges).astype(np.float32)
    isd = _relief._is_discrete(xv, cnt, 10)
    t = T("ranges / is_discrete", t)
    ds.set_features(isd, recip, _native.FS_ARITH_F32)
    t = T("set_features", t)
    wsum, st = ds.score(_native.FS_MULTISURF, want_stats=True)
    t = T("score", t)
    scores = (wsum / ds.n).astype(np.float32)
    top = np.argsort(scores)[::-1][:10]
    t = T("finish (argsort)", t)
    ds.close()
    t = T("close", t)
    print("total", round(1e3 * (t - t00), 3), "ms; score stats", {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")})
    t0 = time.perf_counter()
    fsb.MultiSURF(n_features_to_select=10, backend="gpu").fit(x, y)
    print("MultiSURF.fit", round(1e3 * (time.perf_counter() - t0), 3), "ms")
