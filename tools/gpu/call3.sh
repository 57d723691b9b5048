#!/bin/bash
# round 2, GPU call 3: emulated multi-GPU group tests + the regular GPU suite on the new build
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q --durations=8 > $O/pytest_group.log 2>&1; echo "rc=$?" >> $O/pytest_group.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_shapes_more.py --deselect tests/test_gpu_group.py > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -n 25 $O/pytest_group.log; tail -n 5 $O/pytest_gpu.log
