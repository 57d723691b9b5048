#!/bin/bash
# round 2, GPU call 41 (1 GPU): joint-count finishing kernel with one float64 division per cell instead of four: parity, J1 bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call41; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_joint.py -m gpu -x -q > $O/pytest_joint.log 2>&1; echo "rc=$?" >> $O/pytest_joint.log
tail -n 3 $O/pytest_joint.log
timeout 300 python bench.py --workload j1 --steps 5 --warmup 3 > $O/bench_j1.json 2> $O/bench_j1.err; echo "rc=$?" >> $O/bench_j1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_call41/bench_j1.json") if l.startswith("{")][-1])
print("ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, d.get("parity"), {k:round(v.get("frac",0),3) for k,v in d.get("kernels",{}).items()})
PY
