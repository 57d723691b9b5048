#!/bin/bash
# round 2, GPU call 17 (8 GPUs): ring assignment of cross-rank blocks -- weak / strong C3 and C5 again
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call17; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_c3_8gpu_weak.json 2> $O/bench_c3_8gpu_weak.err; echo "rc=$?" >> $O/bench_c3_8gpu_weak.err
timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > $O/bench_c3_8gpu_strong.json 2> $O/bench_c3_8gpu_strong.err; echo "rc=$?" >> $O/bench_c3_8gpu_strong.err
timeout 600 $TR --master-port 29533 bench.py --gpus 8 --steps 1 --warmup 1 --workload c5 > $O/bench_c5_8gpu.json 2> $O/bench_c5_8gpu.err; echo "rc=$?" >> $O/bench_c5_8gpu.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR4 --master-port 29534 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_c3_4gpu_weak.json 2> $O/bench_c3_4gpu_weak.err; echo "rc=$?" >> $O/bench_c3_4gpu_weak.err
tail -n 2 $O/*.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call17/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e e2e %.1f ms"%(d["ms_per_step"],d["value"],1e3*d["e2e"]["seconds_per_fit"]), {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d.get("parity",{}).get("ok"))
    except Exception as e: print(f, "failed", e)
PY
