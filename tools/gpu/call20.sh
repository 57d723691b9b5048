#!/bin/bash
# round 2, GPU call 20 (1 GPU): merged-plane accumulation epilogue -- diagnostic against the paired kernel,
# parity suites, C3 bench in both modes
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call20; mkdir -p $O
timeout 300 python tools/gpu/diag_merged.py > $O/diag.log 2>&1; echo "rc=$?" >> $O/diag.log
cat $O/diag.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 > $O/$name.json 2> $O/$name.err; }
run c3_merged FS_B200_ACCUM_PAIR=2
run c3_paired FS_B200_ACCUM_PAIR=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call20/c3_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d["parity"])
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 $O/c3_merged.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimators.py tests/test_gpu_group.py -m gpu -x -q > $O/pytest_sub.log 2>&1; echo "rc=$?" >> $O/pytest_sub.log
tail -n 5 $O/pytest_sub.log
