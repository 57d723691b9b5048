#!/bin/bash
# round 2, GPU call 27 (1 GPU): pairs hand the accumulator back with a non-fencing remote arrive (the cluster-scope release was a MEMBAR.ALL.GPU per warp and item); elect.sync issuers in the distance kernels and the older accumulation kernels
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call27; mkdir -p $O
timeout 180 python tools/gpu/diag_merged.py > $O/diag.log 2>&1; echo "rc=$?" >> $O/diag.log
tail -n 3 $O/diag.log
if ! grep -q "^rc=0" $O/diag.log; then echo "diagnostic failed: stopping"; cat $O/diag.log; exit 0; fi
export FS_BENCH_SKIP_CPU=1
run() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-parity > $O/$name.json 2> $O/$name.err; }
for e in 0 8; do run pairs_exp$e FS_B200_ACCUM_PAIR=3 FS_B200_ACCUM_EXP=$e; done
run single_exp0 FS_B200_ACCUM_PAIR=2 FS_B200_ACCUM_EXP=0
run generic_rows FS_B200_ACCUM_PAIR=0
run paired_old FS_B200_ACCUM_PAIR=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call27/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], "accum %.3f"%d["phases_ms"]["ms_accum_tensor"], "dist %.3f"%d["phases_ms"]["ms_dist_tensor"])
    except Exception as e: print(f, "failed", e)
PY
