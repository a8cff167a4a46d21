#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python tools/e2e_timeline.py async 8 $O/r2l_tl_async.txt 2>&1 | tail -6
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), j["clocks"]["samples"])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for i in 1 2 3; do
timeout 200 python bench.py $B > $O/r2l_a_$i.json 2> $O/r2l_a_$i.err; show r2l_a_$i
PHNREC_BENCH_NO_SAMPLER=1 timeout 200 python bench.py $B > $O/r2l_nos_$i.json 2> $O/r2l_nos_$i.err; show r2l_nos_$i
done
PHNREC_VIT_INLINE=1 PHNREC_BENCH_NO_SAMPLER=1 timeout 200 python bench.py $B > $O/r2l_inl.json 2> $O/r2l_inl.err; show r2l_inl
timeout 300 python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > $O/r2l_100.json 2> $O/r2l_100.err; show r2l_100
