#!/bin/bash
# round 2, GPU call 9 (1 GPU): why did the single-CTA accumulation get slower?  launch path / arrival style / clusters
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call9; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-parity > $O/$name.json 2> $O/$name.err; }
run cg2 FS_B200_TRACE=1
run nocg2_exlaunch_arriveall FS_B200_ACCUM_CG2=0
run nocg2_exlaunch_arrivewarp FS_B200_ACCUM_CG2=0 FS_B200_ACCUM_ARRIVE_ALL=0
run nocg2_legacy_arriveall FS_B200_ACCUM_CG2=0 FS_B200_ACCUM_LEGACY_LAUNCH=1
run nocg2_legacy_arrivewarp FS_B200_ACCUM_CG2=0 FS_B200_ACCUM_LEGACY_LAUNCH=1 FS_B200_ACCUM_ARRIVE_ALL=0
grep -h "resident clusters" $O/cg2.err | sort | uniq -c
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call9/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v})
    except Exception as e: print(f, "failed", e)
PY
