#!/bin/bash
# deferred decoder beside K-stc (opt-in PHNREC_VIT_DEFER=1): tests in both settings, A/B, timeline
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2Z_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2Z_smoke.txt
if ! grep -q "mode 1 ok" $O/r2Z_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
PHNREC_VIT_DEFER=1 timeout 400 python -m pytest tests/test_gpu_async.py tests/test_gpu_tensor_core.py tests/test_gpu_full_size.py -q -x --timeout 200 > $O/r2Z_pytest_defer.log 2>&1; echo "rc=$?" >> $O/r2Z_pytest_defer.log; tail -3 $O/r2Z_pytest_defer.log
timeout 400 python -m pytest tests/test_gpu_async.py tests/test_gpu_tensor_core.py tests/test_gpu_full_size.py -q -x --timeout 200 > $O/r2Z_pytest.log 2>&1; echo "rc=$?" >> $O/r2Z_pytest.log; tail -3 $O/r2Z_pytest.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
for i in 1 2; do
PHNREC_VIT_DEFER=1 timeout 120 python bench.py $B > $O/r2Z_defer_$i.json 2> $O/r2Z_defer_$i.err; show r2Z_defer_$i
timeout 120 python bench.py $B > $O/r2Z_base_$i.json 2> $O/r2Z_base_$i.err; show r2Z_base_$i
done
PHNREC_VIT_DEFER=1 timeout 100 python tools/e2e_timeline.py device 40 $O/r2Z_tl_device.txt > /dev/null 2>&1; sed -n 30,33p $O/r2Z_tl_device.txt | cut -c1-110
PHNREC_VIT_DEFER=1 timeout 100 python tools/e2e_timeline.py async 40 $O/r2Z_tl_async.txt > /dev/null 2>&1; sed -n 30,33p $O/r2Z_tl_async.txt | cut -c1-150
