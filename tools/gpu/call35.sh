#!/bin/bash
# round 2, GPU call 35 (1 GPU): staged upload of pageable matrices (host thread pool -> page-locked blocks -> DMA) against
# the driver's own staging (FS_B200_NO_STAGED_UPLOAD=1): end-to-end fit from an ordinary numpy array, C3
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call35; mkdir -p $O
export FS_BENCH_SKIP_CPU=1
timeout 300 python bench.py --steps 5 --warmup 3 > $O/staged.json 2> $O/staged.err
FS_B200_NO_STAGED_UPLOAD=1 timeout 300 python bench.py --steps 5 --warmup 3 > $O/driver.json 2> $O/driver.err
timeout 300 python bench.py --steps 5 --warmup 3 > $O/staged2.json 2> $O/staged2.err
nproc; python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call35/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        e=d["e2e"]; print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], "pinned %.2f ms, pageable %.2f ms per fit"%(1e3*e["seconds_per_fit"],1e3*e["pageable_seconds_per_fit"]), d["parity"]["ok"], e["matches_resident_run"])
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_estimators.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 2
