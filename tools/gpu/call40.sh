#!/bin/bash
# round 2, GPU call 40 (4 GPUs): final build on four GPUs (torchrun): C3 weak with the parity object
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call40; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_c3_4gpu_weak.json 2> $O/bench_c3_4gpu_weak.err; echo "rc=$?" >> $O/bench_c3_4gpu_weak.err
tail -n 2 $O/bench_c3_4gpu_weak.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call40/bench_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.3e"%(d["ms_per_step"],d["value"]), {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, d.get("parity"), d.get("e2e",{}).get("seconds_per_fit"))
    except Exception as e: print(f, "failed", e)
PY
