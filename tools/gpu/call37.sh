#!/bin/bash
# round 2, GPU call 37 (1 GPU): the two new GPU tests (staged upload of strided pageable views; merged accumulation kernel over
# more than one launch of tiles), and ncu of the accumulation kernel at n = 20 000 (first TuRF pass of a C5 cut)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call37; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_estimators.py -m gpu -x -q -k "staged or many_tiles" > $O/pytest_new.log 2>&1; echo "rc=$?" >> $O/pytest_new.log
tail -n 12 $O/pytest_new.log
C5="python bench.py --workload c5 --features 100000 --steps 1 --warmup 0 --no-parity"
FS_BENCH_SKIP_CPU=1 timeout 600 ncu --set full --clock-control none -k regex:tc_accum_merged -s 0 -c 1 -o $O/prof_c5cut $C5 > $O/ncu_c5cut.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_c5cut.ncu-rep --page details > $O/details_c5cut.txt 2>/dev/null
grep -E "Duration|SM Frequency|DRAM Throughput|L2 Hit Rate|Issue Slots Busy|TMEM|pipeline|Mem Busy|Max Bandwidth" $O/details_c5cut.txt | head -12
rm -f $O/prof_c5cut.ncu-rep
