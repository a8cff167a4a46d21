#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), j["clocks"])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 200 python bench.py $B > $O/r2n_def.json 2> $O/r2n_def.err; show r2n_def
timeout 200 python bench.py --steps 20 --warmup 3 $B > $O/r2n_20.json 2> $O/r2n_20.err; show r2n_20
timeout 200 python bench.py --steps 100 --warmup 5 $B > $O/r2n_100.json 2> $O/r2n_100.err; show r2n_100
PHNREC_VIT_INLINE=1 timeout 200 python bench.py --steps 20 --warmup 3 $B > $O/r2n_20i.json 2> $O/r2n_20i.err; show r2n_20i
PHNREC_BENCH_NO_SAMPLER=1 timeout 200 python bench.py --steps 20 --warmup 3 $B > $O/r2n_20ns.json 2> $O/r2n_20ns.err; show r2n_20ns
