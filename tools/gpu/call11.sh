#!/bin/bash
# round 2, GPU call 11 (1 GPU): ncu launch list + full captures of the C3 kernels (shipped build)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call11; mkdir -p $O
CMD="python bench.py --steps 2 --warmup 1 --no-parity"
$CMD > $O/plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c3.csv $CMD > $O/ncu_launch.log 2>&1; echo "launch list rc=$?"
for k in tc_accum_kernel tc_dist_pair_kernel onehot_encode_v3_kernel select_threshold_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $O/prof_$k $CMD > $O/ncu_$k.log 2>&1; echo "$k rc=$?"
  if [ -f $O/prof_$k.ncu-rep ]; then
    ncu -i $O/prof_$k.ncu-rep --page raw --csv > $O/raw_$k.csv 2>/dev/null
    ncu -i $O/prof_$k.ncu-rep --page details > $O/details_$k.txt 2>/dev/null
    ncu -i $O/prof_$k.ncu-rep --page source --csv > $O/source_$k.csv 2>/dev/null
  fi
done
ls -la $O | head -30; tail -3 $O/ncu_tc_accum_kernel.log
