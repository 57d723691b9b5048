#!/bin/bash
# round 2, GPU call 23 (1 GPU): TMEM load shapes (microbenchmark); decomposition of the merged epilogue's own time
# (no operand traffic, no MMAs: 16 + {1 no codes, 2 no constants, 4 no arithmetic, 7 TMEM loads only, 8 nothing});
# effect of the L2 prefetch of the next unit (32 = off) on the product and on the bare mainloop (8 / 40)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call23; mkdir -p $O
timeout 120 tools/build/tmem_ld_bench > $O/tmem_ld_bench.txt 2>&1; echo "rc=$?" >> $O/tmem_ld_bench.txt
cat $O/tmem_ld_bench.txt
export FS_BENCH_SKIP_CPU=1
run() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 10 --warmup 3 --no-parity > $O/$name.json 2> $O/$name.err; }
for e in 16 17 18 20 23 24 0 32; do run single_exp$e FS_B200_ACCUM_PAIR=2 FS_B200_ACCUM_EXP=$e; done
for e in 0 32 8 40; do run pairs_exp$e FS_B200_ACCUM_PAIR=3 FS_B200_ACCUM_EXP=$e; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call23/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], "accum %.3f"%d["phases_ms"]["ms_accum_tensor"])
    except Exception as e: print(f, "failed", e)
PY
