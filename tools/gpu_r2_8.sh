#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for m in device async sync; do timeout 120 python tools/e2e_timeline.py $m 10 $O/r2h_tl_$m.txt 2>&1 | tail -14; done
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), j["clocks"])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 200 python bench.py $B > $O/r2h_side.json 2> $O/r2h_side.err; show r2h_side
PHNREC_VIT_INLINE=1 timeout 200 python bench.py $B > $O/r2h_inline.json 2> $O/r2h_inline.err; show r2h_inline
timeout 200 python bench.py $B > $O/r2h_side2.json 2> $O/r2h_side2.err; show r2h_side2
