#!/bin/bash
# round 2, GPU call 30 (1 GPU), default build: merged kernel adds exact 64-bit column sums with atomics (no float64
# in the kernel, no per-unit partial vectors): diagnostic, parity subset, C3 bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call30; mkdir -p $O
timeout 180 python tools/gpu/diag_merged.py > $O/diag.log 2>&1; echo "rc=$?" >> $O/diag.log
tail -n 3 $O/diag.log
if ! grep -q "^rc=0" $O/diag.log; then echo "diagnostic failed: stopping"; cat $O/diag.log; exit 0; fi
export FS_BENCH_SKIP_CPU=1
for i in 1 2; do timeout 200 python bench.py --steps 10 --warmup 3 > $O/c3_run$i.json 2> $O/c3_run$i.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call30/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e: print(f, "failed", e)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimators.py tests/test_gpu_group.py -m gpu -x -q > $O/pytest_sub.log 2>&1; echo "rc=$?" >> $O/pytest_sub.log
tail -n 3 $O/pytest_sub.log
