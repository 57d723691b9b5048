#!/bin/bash
# round 2, GPU call 43 (1 GPU): ncu of the distance GEMM of the final build (elect.sync issuer)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call43; mkdir -p $O
FS_BENCH_SKIP_CPU=1 timeout 100 ncu --set full --clock-control none -k regex:tc_dist_pair -s 1 -c 1 -o $O/prof_dist python bench.py --steps 1 --warmup 1 --no-parity > $O/ncu_dist.log 2>&1; echo "ncu rc=$?"
ncu -i $O/prof_dist.ncu-rep --page details > $O/details_dist.txt 2>/dev/null
ncu -i $O/prof_dist.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
for a,b in zip(h,v):
    if a in ('gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum'): print(a,b)
" | tee $O/dist_raw_summary.txt
rm -f $O/prof_dist.ncu-rep
