#!/bin/bash
# round 2, GPU call 31 (1 GPU): operand ring of the pair kernel, 6 stages (shipped) against 7 (variant library), on C3 and
# on a C5 cut (20 000 x 100 000)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call31; mkdir -p $O
export FS_BENCH_SKIP_CPU=1
L=fastselect_b200/lib/libfastselect_b200.so
cp $L /tmp/lib_default.so
for v in default stages7 default2; do
  [ $v = stages7 ] && cp tools/build/variants/lib_stages7.so $L
  [ $v = default2 ] && cp /tmp/lib_default.so $L
  timeout 200 python bench.py --steps 10 --warmup 3 --no-parity > $O/c3_$v.json 2> $O/c3_$v.err
  timeout 400 python bench.py --workload c5 --features 100000 --steps 1 --warmup 1 --no-parity > $O/c5cut_$v.json 2> $O/c5cut_$v.err
done
cp /tmp/lib_default.so $L
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call31/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v})
    except Exception as e: print(f, "failed", e)
PY
