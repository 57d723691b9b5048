#!/bin/bash
# round 2, GPU call 36 (1 GPU): host timeline of one fit on C3
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call36; mkdir -p $O
timeout 300 python tools/gpu/e2e_timeline.py > $O/timeline.txt 2>&1; echo "rc=$?" >> $O/timeline.txt
cat $O/timeline.txt
