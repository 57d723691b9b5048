#!/bin/bash
# round 2, GPU call 16 (1 GPU): SURF accumulation on the float32 image -- parity, full C4 SURF shapes, bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call16; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_shapes_more.py -m gpu -x -q -k "c4" > $O/pytest_c4.log 2>&1; echo "rc=$?" >> $O/pytest_c4.log
timeout 900 python bench.py --workload c4surf --steps 2 --warmup 2 > $O/bench_c4surf.json 2> $O/bench_c4surf.err; echo "rc=$?" >> $O/bench_c4surf.err
tail -n 4 $O/pytest_gpu.log $O/pytest_c4.log
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_call16/bench_c4surf.json") if l.startswith("{")][-1])
print("ms/step %.1f"%d["ms_per_step"], {k:round(v,2) for k,v in d["phases_ms"].items() if v}, d["parity"])
PY
