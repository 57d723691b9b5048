#!/bin/bash
# round 2, GPU call 19 (1 GPU): what the driver runs at round end -- full GPU suite, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call19; mkdir -p $O
( time python -m pytest tests/ -x -q -m gpu ) > $O/pytest_gpu_all.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_all.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
( time python bench.py --impl reference ) > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
tail -n 6 $O/pytest_gpu_all.log; tail -n 5 $O/smoke.log; tail -n 5 $O/bench_default.err $O/bench_reference.err
python - <<'PY'
import json
for f in ("bench_default","bench_reference"):
    d=json.loads([l for l in open(f"gpurun_out/r02_call19/{f}.json") if l.startswith("{")][-1])
    print(f, {k:d.get(k) for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","gpu_launches")}, d.get("parity"), d.get("roofline",{}).get("frac"), d.get("cpu_baseline"), d.get("e2e"))
PY
