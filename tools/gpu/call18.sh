#!/bin/bash
# round 2, GPU call 18 (1 GPU): float64-scaling epilogue experiment; ncu of the CUDA-core kernels (C4 reduced, C2)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call18; mkdir -p $O
export FS_B200_BARRIER_TIMEOUT_S=15
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 > $O/$name.json 2> $O/$name.err; }
run c3_int FS_B200_ACCUM_F64=0
run c3_f64 FS_B200_ACCUM_F64=1
FS_B200_ACCUM_F64=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_estimators.py -m gpu -x -q > $O/pytest_f64.log 2>&1; echo "rc=$?" >> $O/pytest_f64.log
timeout 600 python -m pytest tests/test_gpu_estimators.py -m gpu -x -q > $O/pytest_est.log 2>&1; echo "rc=$?" >> $O/pytest_est.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call18/c3_*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d["phases_ms"].items() if v}, d["parity"]["ok"])
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 $O/pytest_f64.log $O/pytest_est.log
# ncu: CUDA-core kernels on a reduced C4 (MultiSURF*, 8000 x 20000 mixed) and on C2 (ReliefF 10000 x 10000)
C4="python bench.py --workload c4 --samples 8000 --features 20000 --steps 1 --warmup 1 --no-parity"
$C4 > $O/plain_c4.log 2>&1 && for k in dist_general_kernel accum_general_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/prof_c4_$k $C4 > $O/ncu_c4_$k.log 2>&1; echo "$k rc=$?"
  [ -f $O/prof_c4_$k.ncu-rep ] && ncu -i $O/prof_c4_$k.ncu-rep --page details > $O/details_c4_$k.txt 2>/dev/null && rm -f $O/prof_c4_$k.ncu-rep
done
C2="python bench.py --workload c2 --steps 1 --warmup 1 --no-parity"
$C2 > $O/plain_c2.log 2>&1 && for k in dist_general_kernel relieff_select_kernel relieff_gather_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/prof_c2_$k $C2 > $O/ncu_c2_$k.log 2>&1; echo "$k rc=$?"
  [ -f $O/prof_c2_$k.ncu-rep ] && ncu -i $O/prof_c2_$k.ncu-rep --page details > $O/details_c2_$k.txt 2>/dev/null && rm -f $O/prof_c2_$k.ncu-rep
done
ls -la $O | tail -20
