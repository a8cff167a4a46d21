#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
L=$PWD/phnrec_b200/lib
timeout 600 python -m pytest tests/test_gpu_tensor_core.py -q -x > $O/r2v_pytest.log 2>&1; echo "rc=$?" >> $O/r2v_pytest.log; tail -3 $O/r2v_pytest.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for i in 1 2 3; do
timeout 200 python bench.py $B > $O/r2v_stc3_$i.json 2> $O/r2v_stc3_$i.err; show r2v_stc3_$i
PHNREC_B200_LIB=$L/libphnrec_b200_stc2.so timeout 200 python bench.py $B > $O/r2v_stc2_$i.json 2> $O/r2v_stc2_$i.err; show r2v_stc2_$i
done
for c in en hu; do timeout 200 python bench.py --config $c $B > $O/r2v_stc3_$c.json 2> $O/r2v_stc3_$c.err; show r2v_stc3_$c; PHNREC_B200_LIB=$L/libphnrec_b200_stc2.so timeout 200 python bench.py --config $c $B > $O/r2v_stc2_$c.json 2> $O/r2v_stc2_$c.err; show r2v_stc2_$c; done
