#!/bin/bash
# round 2, GPU call 38 (2 GPUs): shard boundaries on super-blocks of 256 target rows (group_shard_starts): torchrun bench (weak and strong,
# parity object: ranks identical, bitwise equal to one GPU, oracle), single-process multi-GPU entry
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call38; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c3_2gpu_weak.json 2> $O/bench_c3_2gpu_weak.err; echo "rc=$?" >> $O/bench_c3_2gpu_weak.err
timeout 400 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong > $O/bench_c3_2gpu_strong.json 2> $O/bench_c3_2gpu_strong.err; echo "rc=$?" >> $O/bench_c3_2gpu_strong.err
timeout 400 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "single_process or group" > $O/pytest_single_process_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_single_process_2gpu.log
tail -n 3 $O/*.err $O/pytest_single_process_2gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call38/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e"%(d["ms_per_step"],d["value"]), {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, d.get("parity"), d.get("e2e",{}).get("seconds_per_fit"))
    except Exception as e: print(f, "failed", e)
PY
