#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 40 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for o in 1 2 4 8 1 4; do
PHNREC_FRONT_OVERSUB=$o timeout 200 python bench.py $B > $O/r2G_o$o.json 2> $O/r2G_o$o.err; show r2G_o$o
done
PHNREC_VIT_INLINE=1 timeout 200 python bench.py $B > $O/r2G_inline.json 2> $O/r2G_inline.err; show r2G_inline
PHNREC_VIT_INLINE=1 PHNREC_FRONT_OVERSUB=1 timeout 200 python bench.py $B > $O/r2G_inline_o1.json 2> $O/r2G_inline_o1.err; show r2G_inline_o1
