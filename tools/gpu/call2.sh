#!/bin/bash
# round 2, GPU call 2: the new full-shape parity tests (C4 SURF/SURF*, C5 two TuRF passes, production rows)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call2; mkdir -p $O
timeout 2400 python -m pytest tests/test_gpu_shapes_more.py -m gpu -x -q --durations=12 > $O/pytest_shapes_more.log 2>&1; echo "rc=$?" >> $O/pytest_shapes_more.log
tail -n 30 $O/pytest_shapes_more.log
