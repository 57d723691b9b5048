#!/bin/bash
# round 2, GPU call 13 (1 GPU): accumulation prefetches + fused slice encode -- tests (incl. emulated groups), C3
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call13; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
tail -n 4 $O/pytest_gpu.log; tail -n 2 $O/bench_c3.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_call13/bench_c3.json") if l.startswith("{")][-1])
print("ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d["parity"])
PY
