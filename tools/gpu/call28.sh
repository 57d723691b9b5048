#!/bin/bash
# round 2, GPU call 28 (1 GPU), default build: full GPU suite on the new accumulation kernel (CTA pairs, merged planes),
# C3 / C5 bench lines with parity, ncu of the kernel
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call28; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
( time timeout 1200 python -m pytest tests/ -x -q -m gpu ) > $O/pytest_gpu_all.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_all.log
tail -n 8 $O/pytest_gpu_all.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 > $O/bench_c5.json 2> $O/bench_c5.err; echo "rc=$?" >> $O/bench_c5.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call28/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e"%(d["ms_per_step"],d["value"]), {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, d.get("parity"), {k:round(v["frac"],3) for k,v in d.get("kernels",{}).items()}, d.get("e2e"))
    except Exception as e: print(f, "failed", e)
PY
C3="python bench.py --steps 1 --warmup 1 --no-parity"
FS_BENCH_SKIP_CPU=1 ncu --set full --clock-control none --import-source on -k regex:tc_accum_merged -s 1 -c 1 -o $O/prof_pairs $C3 > $O/ncu_pairs.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_pairs.ncu-rep --page details > $O/details_pairs.txt 2>/dev/null
ls -la $O | tail -8
