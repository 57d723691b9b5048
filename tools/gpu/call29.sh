#!/bin/bash
# round 2, GPU call 29 (1 GPU): unit order of the merged accumulation kernel at n = 20 000 (C5): blocks-major (0) against
# group-major (1); C3 with both
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02_call29; mkdir -p $O
export FS_BENCH_SKIP_CPU=1
for o in 0 1; do
  FS_B200_ACCUM_ORDER=$o timeout 600 python bench.py --workload c5 --steps 1 --warmup 1 > $O/c5_order$o.json 2> $O/c5_order$o.err
  FS_B200_ACCUM_ORDER=$o timeout 200 python bench.py --steps 10 --warmup 3 --no-parity > $O/c3_order$o.json 2> $O/c3_order$o.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_call29/*.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f"%d["ms_per_step"], {k:round(v,3) for k,v in d.get("phases_ms",{}).items() if v}, (d.get("parity") or {}).get("ok"), d.get("top_features", [])[:4])
    except Exception as e: print(f, "failed", e)
PY
