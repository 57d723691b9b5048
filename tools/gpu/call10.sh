#!/bin/bash
# round 2, GPU call 10 (1 GPU): per-item constants staged in shared memory -- tests, then C3 pair / single-CTA
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call10; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-parity > $O/$name.json 2> $O/$name.err; }
run cg2 FS_B200_TRACE=0
run nocg2 FS_B200_ACCUM_CG2=0
run nocg2_nopairepi FS_B200_ACCUM_CG2=0 FS_B200_ACCUM_PAIR=0
run cg2_nopairepi FS_B200_ACCUM_PAIR=0
tail -n 4 $O/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call10/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v})
    except Exception as e: print(f, "failed", e)
PY
