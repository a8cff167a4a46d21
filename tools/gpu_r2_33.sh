#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 90 python tools/wave_tc_check.py > $O/r2O_check.txt 2>&1; echo "check rc=$?"; tail -3 $O/r2O_check.txt
if ! grep -q "max |dmel|" $O/r2O_check.txt; then echo "CHECK FAILED - stopping"; exit 1; fi
PHNREC_WTC_DBG=8 timeout 120 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 0.1 > $O/r2O_tl.json 2> $O/r2O_tl.err; grep "wtc tile" $O/r2O_tl.err | tail -16
B="--steps 10 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 0.5"
for d in 0 2; do
PHNREC_WTC_DBG=$d timeout 120 python bench.py $B > $O/r2O_dbg$d.json 2> $O/r2O_dbg$d.err
python - $d <<'PY'
import json,sys
d=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2O_dbg{d}.json")); print("dbg", d, "K-wave ms", [k["ms"] for k in j["roofline"]["kernels"] if k["kernel"]=="K-wave"], "step", round(j["ms_per_step"],3), "e2e", round(j["e2e"]["ms_per_step"],3))
except Exception as e: print(d, "ERR", e, open(f"gpurun_out/r2O_dbg{d}.err").read()[-800:])
PY
done
