#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
L=$PWD/phnrec_b200/lib
timeout 120 python tools/e2e_timeline.py device 8 $O/r2k_tl_device.txt 2>&1 | tail -6
timeout 120 python tools/e2e_timeline.py async 8 $O/r2k_tl_async.txt 2>&1 | tail -6
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for i in 1 2; do
timeout 200 python bench.py $B > $O/r2k_minb32_$i.json 2> $O/r2k_minb32_$i.err; show r2k_minb32_$i
PHNREC_B200_LIB=$L/libphnrec_b200_vit1.so timeout 200 python bench.py $B > $O/r2k_minb1_$i.json 2> $O/r2k_minb1_$i.err; show r2k_minb1_$i
PHNREC_VIT_INLINE=1 timeout 200 python bench.py $B > $O/r2k_inline_$i.json 2> $O/r2k_inline_$i.err; show r2k_inline_$i
done
