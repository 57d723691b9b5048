#!/bin/bash
# round 2, GPU call 12 (2 GPUs): single-process multi-GPU entry; torchrun weak / strong C3, reduced C5 (TuRF), C4
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call12; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=20
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "single_process" > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c3_2gpu_weak.json 2> $O/bench_c3_2gpu_weak.err; echo "rc=$?" >> $O/bench_c3_2gpu_weak.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong > $O/bench_c3_2gpu_strong.json 2> $O/bench_c3_2gpu_strong.err; echo "rc=$?" >> $O/bench_c3_2gpu_strong.err
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 1 --warmup 1 --workload c5 --n 8000 --p 100000 > $O/bench_c5small_2gpu.json 2> $O/bench_c5small_2gpu.err; echo "rc=$?" >> $O/bench_c5small_2gpu.err
timeout 300 python bench.py --steps 1 --warmup 1 --workload c5 --n 8000 --p 100000 > $O/bench_c5small_1gpu.json 2> $O/bench_c5small_1gpu.err; echo "rc=$?" >> $O/bench_c5small_1gpu.err
timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 1 --workload c4 --scaling strong > $O/bench_c4_2gpu.json 2> $O/bench_c4_2gpu.err; echo "rc=$?" >> $O/bench_c4_2gpu.err
tail -n 6 $O/pytest_multi.log; tail -n 4 $O/*.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call12/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f e2e %.1f ms"%(d["ms_per_step"],1e3*d["e2e"]["seconds_per_fit"]), {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d.get("parity"))
    except Exception as e: print(f, "failed", e)
PY
