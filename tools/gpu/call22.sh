#!/bin/bash
# round 2, GPU call 22 (1 GPU): merged-plane accumulation kernel as CTA pairs (FS_B200_ACCUM_PAIR=3, the new default):
# diagnostic against the single-CTA variants, C3 bench in modes 3 / 2, pure-mainloop (8) and pure-epilogue (16) experiments
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call22; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 180 python tools/gpu/diag_merged.py > $O/diag.log 2>&1; echo "rc=$?" >> $O/diag.log
cat $O/diag.log
if ! grep -q "^rc=0" $O/diag.log; then echo "diagnostic failed: stopping"; exit 0; fi
export FS_BENCH_SKIP_CPU=1
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 10 --warmup 3 > $O/$name.json 2> $O/$name.err; }
run c3_pairs FS_B200_ACCUM_PAIR=3
run c3_merged FS_B200_ACCUM_PAIR=2
for e in 8 16; do
  FS_B200_ACCUM_PAIR=3 FS_B200_ACCUM_EXP=$e timeout 200 python bench.py --steps 10 --warmup 3 --no-parity > $O/pairs_exp$e.json 2> $O/pairs_exp$e.err
  FS_B200_ACCUM_PAIR=2 FS_B200_ACCUM_EXP=$e timeout 200 python bench.py --steps 10 --warmup 3 --no-parity > $O/merged_exp$e.json 2> $O/merged_exp$e.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call22/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v}, (d.get("parity") or {}).get("ok"))
    except Exception as e: print(f, "failed", e)
PY
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_estimators.py tests/test_gpu_group.py -m gpu -x -q > $O/pytest_sub.log 2>&1; echo "rc=$?" >> $O/pytest_sub.log
tail -n 5 $O/pytest_sub.log
