#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call4; mkdir -p $O
timeout 300 python tools/gpu/diag_group.py > $O/diag_group.log 2>&1; echo "rc=$?" >> $O/diag_group.log
cat $O/diag_group.log | tail -40
