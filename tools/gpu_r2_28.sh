#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
L=$PWD/phnrec_b200/lib
timeout 90 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2I_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/r2I_smoke.txt
if ! grep -q "mode 1 ok" $O/r2I_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_full_size.py tests/test_gpu_async.py -q -x --timeout 120 > $O/r2I_pytest.log 2>&1; echo "rc=$?" >> $O/r2I_pytest.log; tail -6 $O/r2I_pytest.log
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
timeout 100 python bench.py $B > $O/r2I_wg_$i.json 2> $O/r2I_wg_$i.err; show r2I_wg_$i
PHNREC_B200_LIB=$L/libphnrec_b200_v5.so timeout 100 python bench.py $B > $O/r2I_v5_$i.json 2> $O/r2I_v5_$i.err; show r2I_v5_$i
done
for c in hu ru en; do
timeout 100 python bench.py --config $c $B > $O/r2I_wg_$c.json 2> $O/r2I_wg_$c.err; show r2I_wg_$c
done
for n in 0 2; do timeout 60 python tools/tc_timeline.py $n > $O/r2I_timeline_$n.txt 2>&1; cat $O/r2I_timeline_$n.txt; done
