#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2j_pytest.log
tail -4 $O/r2j_pytest.log
timeout 120 python tools/e2e_timeline.py async 8 $O/r2j_tl_async.txt 2>&1 | tail -10
timeout 120 python tools/e2e_timeline.py device 8 $O/r2j_tl_device.txt 2>&1 | tail -10
PHNREC_FRONT_OVERSUB=1 timeout 120 python tools/e2e_timeline.py device 8 $O/r2j_tl_device_o1.txt 2>&1 | tail -10
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), j["clocks"])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 200 python bench.py $B > $O/r2j_o4.json 2> $O/r2j_o4.err; show r2j_o4
PHNREC_FRONT_OVERSUB=1 timeout 200 python bench.py $B > $O/r2j_o1.json 2> $O/r2j_o1.err; show r2j_o1
PHNREC_FRONT_OVERSUB=8 timeout 200 python bench.py $B > $O/r2j_o8.json 2> $O/r2j_o8.err; show r2j_o8
PHNREC_VIT_INLINE=1 timeout 200 python bench.py $B > $O/r2j_inline.json 2> $O/r2j_inline.err; show r2j_inline
