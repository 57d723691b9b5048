#!/bin/bash
# round 2, GPU call 21 (1 GPU): where does the merged accumulation kernel spend its time?  Experiment builds
# (-DFS_ACCUM_EXPERIMENTS): 1 no codes loads, 2 no constants loads, 3 both, 4 no arithmetic, 7 TMEM loads only,
# 8 epilogue only waits and arrives; then ncu (full set, source) of the product kernel.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call21; mkdir -p $O
export FS_BENCH_SKIP_CPU=1
for e in 0 1 2 3 4 7 8; do
  FS_B200_ACCUM_EXP=$e timeout 200 python bench.py --steps 10 --warmup 3 --no-parity > $O/exp$e.json 2> $O/exp$e.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call21/exp*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v})
    except Exception as e: print(f, "failed", e)
PY
C3="python bench.py --steps 1 --warmup 1 --no-parity"
ncu --set full --clock-control none --import-source on -k regex:tc_accum_merged -s 1 -c 1 -o $O/prof_merged $C3 > $O/ncu_merged.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail
