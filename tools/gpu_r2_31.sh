#!/bin/bash
# which role bounds k_wave_tc: producers / epilogue / MMAs switched off one at a time (results are garbage, only K-wave's time matters)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 10 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 0.5"
for d in 0 1 2 4 3 5 6 7; do
PHNREC_WTC_DBG=$d timeout 120 python bench.py $B > $O/r2L_dbg$d.json 2> $O/r2L_dbg$d.err
python - $d <<'PY'
import json,sys
d=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2L_dbg{d}.json")); print("dbg", d, "K-wave ms", [k["ms"] for k in j["roofline"]["kernels"] if k["kernel"]=="K-wave"], "step", round(j["ms_per_step"],3))
except Exception as e: print(d, "ERR", e, open(f"gpurun_out/r2L_dbg{d}.err").read()[-800:])
PY
done
