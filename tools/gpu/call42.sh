#!/bin/bash
# round 2, GPU call 42 (1 GPU), final build (after the joint-count change): full GPU suite, smoke, default bench (what the driver runs at round end)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call42; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
( time timeout 1200 python -m pytest tests/ -x -q -m gpu ) > $O/pytest_gpu_all.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_all.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
tail -n 6 $O/pytest_gpu_all.log; tail -n 3 $O/smoke.log; tail -n 4 $O/bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_call42/bench_default.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in (d.get("phases_ms") or {}).items() if v}, d.get("parity"), (d.get("roofline") or {}).get("frac"), d.get("e2e"), d.get("cpu_baseline"))
PY
