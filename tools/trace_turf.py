"""Debugging aid: host-side phase timings (FS_B200_TRACE) of a few TuRF-like re-scoring passes."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import fastselect_b200 as fsb  # noqa: E402

n, p = int(sys.argv[1]), int(sys.argv[2])
w = bench.make_c5(n, p)
base = fsb.MultiSURF(n_features_to_select=10, backend="gpu")
sess, _ = base._open_session(w["x"], w["y"])
sess.score(want_stats=True)
print("full pass", {k: round(v, 3) for k, v in sess.last_stats.items() if k.startswith("ms_")}, flush=True)
if os.environ.get("TRACE_HOST"):
    os.environ["FS_B200_TRACE"] = "1"
act = np.arange(p)
rs = np.random.RandomState(0)
for it in range(4):
    act = np.delete(act, rs.choice(len(act), len(act) // 10, replace=False))
    t = time.perf_counter()
    sess.score(act, want_stats=True)
    print("score", len(act), round(1e3 * (time.perf_counter() - t), 3), "ms",
          {k: round(v, 3) for k, v in sess.last_stats.items() if k.startswith("ms_")}, flush=True)
sess.close()
