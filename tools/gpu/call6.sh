#!/bin/bash
# round 2, GPU call 6 (1 GPU): tests on the list-caching build, C3 bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call6; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; echo "rc=$?" >> $O/bench_c3.err
tail -n 4 $O/pytest_gpu.log; tail -n 3 $O/bench_c3.err; head -c 600 $O/bench_c3.json
