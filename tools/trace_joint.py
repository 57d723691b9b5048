"""Time the joint-count path (fs_joint_matrix) on one GPU: python tools/trace_joint.py [n] [p] [reps].
Prints the per-phase device times of every repetition (fs_stats) and the end-to-end wall time of
calculate_mi_matrices (upload, column scan, encode, GEMM, finishing kernel, result copy)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastselect_b200 import _mi, _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rs = np.random.RandomState(0)
x = rs.randint(0, 3, (n, p)).astype(np.uint8)
y = rs.randint(0, 2, n).astype(np.uint8)
xa = _mi._stack_for_upload(x, y)
for r in range(reps):
    st = {}
    t0 = time.perf_counter()
    m = _mi.joint_matrix(xa, _native.FS_JOINT_MI, np.log(2.0), stats_out=st)
    wall = time.perf_counter() - t0
    keep = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()
            if k in ("ms_total", "ms_gather", "ms_dist_tensor", "ms_reduce", "launches", "n_chunks", "onehot_k", "ops_dist_tensor")}
    units = float(n) * (p + 1) * p / 2
    print(f"joint n={n} p={p} rep={r} wall {wall * 1e3:.1f} ms  sample*pairs/s (wall) {units / wall:.3e}  {keep}", flush=True)
print("relevance max", float(m[p, :p].max()), "redundancy max", float(m[:p, :p].max()))
