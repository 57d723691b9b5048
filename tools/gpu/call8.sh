#!/bin/bash
# round 2, GPU call 8 (1 GPU): CTA-pair accumulation kernel -- parity tests, then C3 with and without it
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call8; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-parity > $O/bench_c3_cg2.json 2> $O/bench_c3_cg2.err; echo "rc=$?" >> $O/bench_c3_cg2.err
FS_B200_ACCUM_CG2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-parity > $O/bench_c3_nocg2.json 2> $O/bench_c3_nocg2.err; echo "rc=$?" >> $O/bench_c3_nocg2.err
timeout 900 python -m pytest tests/test_gpu_shapes_more.py -m gpu -x -q -k "c5_full or c3_full or c4_full_shape_production" > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
tail -n 4 $O/pytest_gpu.log $O/pytest_full.log; tail -n 2 $O/*.err
python - <<'PY'
import json
for f in ("bench_c3_cg2","bench_c3_nocg2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02_call8/{f}.json") if l.startswith("{")][-1])
        print(f, "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v})
    except Exception as e: print(f, "failed", e)
PY
