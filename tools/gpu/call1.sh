#!/bin/bash
# round 2, GPU call 1: probes never run in round 1, FP4 peak, sanitizer passes, baseline bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
timeout 60 tools/build/tmem_frag_probe > $O/tmem_frag_probe.txt 2>&1; echo "rc=$?" >> $O/tmem_frag_probe.txt
timeout 120 tools/build/tc_dist_cg2_probe > $O/tc_dist_cg2_probe.txt 2>&1; echo "rc=$?" >> $O/tc_dist_cg2_probe.txt
timeout 200 tools/build/fp4_peak 4 > $O/fp4_peak.txt 2>&1; echo "rc=$?" >> $O/fp4_peak.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
SEL='multisurf_rows_match_oracle or surf_rows_match_oracle or general_encode or incremental or symmetric or c3_shape or relieff_rows_match_oracle'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multisurf_rows_match_oracle or general_encode or incremental or symmetric" > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_joint.py -m gpu -x -q -k "tables_are_bit_exact or matrix_matches_oracle" > $O/sanitizer_memcheck_joint.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_joint.log
tail -3 $O/*.log $O/*.txt
