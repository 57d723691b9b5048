#!/bin/bash
# round 2, GPU call 15 (1 GPU): ring assignment of cross-rank blocks (emulated ranks), joint finishing kernel,
# then the single-GPU lines of every workload
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call15; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -n 4 $O/pytest_gpu.log
for w in c3 j1 c2 c4 c4surf c5; do
  steps=5; [ $w = c5 ] && steps=1; [ $w = c4 ] && steps=2; [ $w = c4surf ] && steps=2; [ $w = c3 ] && steps=10
  timeout 900 python bench.py --workload $w --steps $steps --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err; echo "rc=$?" >> $O/bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call15/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e"%(d["ms_per_step"],d["value"]), {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, d.get("parity"), {k:round(v["frac"],3) for k,v in d.get("kernels",{}).items()})
    except Exception as e: print(f, "failed", e)
PY
