#!/bin/bash
# Round 2, GPU call 7 (page-locked upload arena): decoder on its own stream, phn_recognize_async / phn_wait.  Tests + bench A/B (decoder inline vs side stream).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
tail -5 $O/r2g_pytest.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 200 python bench.py $B > $O/r2g_side.json 2> $O/r2g_side.err; show r2g_side
PHNREC_VIT_INLINE=1 timeout 200 python bench.py $B > $O/r2g_inline.json 2> $O/r2g_inline.err; show r2g_inline
timeout 200 python bench.py $B > $O/r2g_side2.json 2> $O/r2g_side2.err; show r2g_side2
timeout 200 python bench.py --config en $B > $O/r2g_en.json 2> $O/r2g_en.err; show r2g_en
