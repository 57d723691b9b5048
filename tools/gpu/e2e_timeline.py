"""Round-2 diagnostic: where do the 12 ms of MultiSURF(backend='gpu').fit on C3 (4000 x 100 000 int8, pinned host
memory) go?  Wall-clock segments of the host path, averaged over a few warm fits."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402  (pinned memory only)

import fastselect_b200 as fsb  # noqa: E402
from fastselect_b200 import _native, _relief  # noqa: E402

n, p = 4000, 100000
rs = np.random.RandomState(0)
y = rs.randint(0, 2, n)
x = rs.randint(0, 3, (n, p)).astype(np.int8)
xp = torch.from_numpy(x).pin_memory().numpy()

seg = {}


def timed(name, fn):
    def wrap(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        seg[name] = seg.get(name, 0.0) + time.perf_counter() - t0
        return r
    return wrap


_relief.validate_data = timed("validate_data", _relief.validate_data)
_relief._narrow_integers = timed("narrow_integers", _relief._narrow_integers)
_relief.open_dataset = timed("open_dataset (upload + scan)", _relief.open_dataset)
_relief._is_discrete = timed("is_discrete", _relief._is_discrete)
_native.Dataset.column_stats = timed("column_stats", _native.Dataset.column_stats)
_native.Dataset.set_features = timed("set_features", _native.Dataset.set_features)
_native.Dataset.score = timed("ds.score", _native.Dataset.score)
_native.Dataset.close = timed("ds.close", _native.Dataset.close)
_relief._ReliefBase._finish = timed("finish (ranking)", _relief._ReliefBase._finish)
np_unique = np.unique
_relief.np.unique = timed("np.unique(y)", np_unique)

for src, label in ((xp, "pinned"), (x, "pageable")):
    for _ in range(2):
        fsb.MultiSURF(n_features_to_select=10, backend="gpu").fit(src, y)
    seg.clear()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        fsb.MultiSURF(n_features_to_select=10, backend="gpu").fit(src, y)
    tot = (time.perf_counter() - t0) / reps
    print(f"{label}: {1e3 * tot:.2f} ms per fit")
    acc = 0.0
    for k, v in seg.items():
        print(f"   {k:32s} {1e3 * v / reps:7.3f} ms")
        acc += v / reps
    print(f"   {'(unaccounted)':32s} {1e3 * (tot - acc):7.3f} ms")
