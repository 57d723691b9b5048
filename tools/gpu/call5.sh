#!/bin/bash
# round 2, GPU call 5 (2 GPUs): single-process multi-GPU entry, torchrun bench weak / strong at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call5; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "single_process" > $O/pytest_multi.log 2>&1; echo "rc=$?" >> $O/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c3_2gpu_weak.json 2> $O/bench_c3_2gpu_weak.err; echo "rc=$?" >> $O/bench_c3_2gpu_weak.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > $O/bench_c3_2gpu_strong.json 2> $O/bench_c3_2gpu_strong.err; echo "rc=$?" >> $O/bench_c3_2gpu_strong.err
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_c3_1gpu.json 2> $O/bench_c3_1gpu.err; echo "rc=$?" >> $O/bench_c3_1gpu.err
tail -n 5 $O/pytest_multi.log; tail -n 8 $O/*.err; head -c 1500 $O/bench_c3_2gpu_weak.json
