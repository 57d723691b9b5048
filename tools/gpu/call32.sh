#!/bin/bash
# round 2, GPU call 32 (1 GPU): L2 blocking of the merged accumulation kernel at n = 20 000 (C5 cut, 20 000 x 100 000):
# band of groups walked together (FS_B200_ACCUM_SPAN) x tiles per group (FS_B200_ACCUM_GROUP_TILES)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call32; mkdir -p $O
export FS_BENCH_SKIP_CPU=1
run() { name=$1; shift; env "$@" timeout 300 python bench.py --workload c5 --features 100000 --steps 1 --warmup 1 --no-parity > $O/$name.json 2> $O/$name.err; }
run default X=1
run span6_gt4 FS_B200_ACCUM_SPAN=6 FS_B200_ACCUM_GROUP_TILES=4
run span4_gt4 FS_B200_ACCUM_SPAN=4 FS_B200_ACCUM_GROUP_TILES=4
run span8_gt2 FS_B200_ACCUM_SPAN=8 FS_B200_ACCUM_GROUP_TILES=2
run span3_gt8 FS_B200_ACCUM_SPAN=3 FS_B200_ACCUM_GROUP_TILES=8
run span12_gt2 FS_B200_ACCUM_SPAN=12 FS_B200_ACCUM_GROUP_TILES=2
run span2_gt4 FS_B200_ACCUM_SPAN=2 FS_B200_ACCUM_GROUP_TILES=4
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call32/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], "accum %.1f"%d["phases_ms"]["ms_accum_tensor"], d.get("top_features", [])[:3])
    except Exception as e: print(f, "failed", e)
PY
