#!/bin/bash
# round 2, GPU call 24 (1 GPU): merged kernel with the producer / MMA issuer as the LAST two warps, tile table in
# the constant bank, constant TMEM base: diagnostic, product and bare mainloop (8) / bare epilogue (16) in both modes;
# 32 = At prefetch on, 64 = codes prefetch on
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call24; mkdir -p $O
timeout 180 python tools/gpu/diag_merged.py > $O/diag.log 2>&1; echo "rc=$?" >> $O/diag.log
cat $O/diag.log
if ! grep -q "^rc=0" $O/diag.log; then echo "diagnostic failed: stopping"; exit 0; fi
export FS_BENCH_SKIP_CPU=1
run() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-parity > $O/$name.json 2> $O/$name.err; }
for e in 0 8 64; do run pairs_exp$e FS_B200_ACCUM_PAIR=3 FS_B200_ACCUM_EXP=$e; done
for e in 0 8 16 64; do run single_exp$e FS_B200_ACCUM_PAIR=2 FS_B200_ACCUM_EXP=$e; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call24/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], "accum %.3f"%d["phases_ms"]["ms_accum_tensor"])
    except Exception as e: print(f, "failed", e)
PY
