#!/bin/bash
# round 2, GPU call 14 (8 GPUs): weak / strong C3, C5 (TuRF, full shape), C4 strong, single-process multi-GPU entry
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call14; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=30
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt; free -g >> $O/topo.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu_weak.json 2> $O/bench_c3_8gpu_weak.err; echo "rc=$?" >> $O/bench_c3_8gpu_weak.err
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > $O/bench_c3_8gpu_strong.json 2> $O/bench_c3_8gpu_strong.err; echo "rc=$?" >> $O/bench_c3_8gpu_strong.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --steps 1 --warmup 1 --workload c5 > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err; echo "rc=$?" >> $O/bench_c5_8gpu.err
timeout 400 $TR --master-port 29524 bench.py --gpus 8 --steps 2 --warmup 1 --workload c4 --scaling strong --no-parity > $O/bench_c4_8gpu.json 2> $O/bench_c4_8gpu.err; echo "rc=$?" >> $O/bench_c4_8gpu.err
timeout 200 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "single_process" > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
tail -n 3 $O/pytest_multi.log; tail -n 3 $O/*.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call14/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e e2e %.1f ms"%(d["ms_per_step"],d["value"],1e3*d["e2e"]["seconds_per_fit"]), {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d.get("parity"))
    except Exception as e: print(f, "failed", e)
PY
