#!/bin/bash
# round 2, GPU call 34 (1 GPU), default build: what the driver runs at round end -- full GPU suite, smoke, default bench,
# reference arm -- plus the C5 line and the ncu capture of the final accumulation kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call34; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
( time timeout 1200 python -m pytest tests/ -x -q -m gpu ) > $O/pytest_gpu_all.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_all.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
( time python bench.py --impl reference ) > $O/bench_reference.json 2> $O/bench_reference.err; echo "rc=$?" >> $O/bench_reference.err
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > $O/bench_c5.json 2> $O/bench_c5.err; echo "rc=$?" >> $O/bench_c5.err
tail -n 6 $O/pytest_gpu_all.log; tail -n 3 $O/smoke.log; tail -n 4 $O/bench_default.err $O/bench_reference.err
python - <<'PY'
import json
for f in ("bench_default","bench_reference","bench_c5"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02_call34/{f}.json") if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in (d.get("phases_ms") or {}).items() if v}, d.get("parity"), (d.get("roofline") or {}).get("frac"), d.get("e2e"))
    except Exception as e: print(f, "failed", e)
PY
C3="python bench.py --steps 1 --warmup 1 --no-parity"
FS_BENCH_SKIP_CPU=1 ncu --set full --clock-control none --import-source on -k regex:tc_accum_merged -s 1 -c 1 -o $O/prof_pairs $C3 > $O/ncu_pairs.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_pairs.ncu-rep --page details > $O/details_pairs.txt 2>/dev/null
FS_BENCH_SKIP_CPU=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-parity > $O/ncu_launches.log 2>&1; echo "launch list rc=$?"
