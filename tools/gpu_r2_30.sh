#!/bin/bash
# first runs of the tensor-core front end (k_wave_tc.cu): correctness gate under a short timeout, then tests and bench A/B
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 90 python tools/wave_tc_check.py > $O/r2K_check.txt 2>&1; echo "check rc=$?"; tail -5 $O/r2K_check.txt
if ! grep -q "max |dmel|" $O/r2K_check.txt; then echo "CHECK FAILED - stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_full_size.py tests/test_gpu_async.py -q -x --timeout 120 > $O/r2K_pytest.log 2>&1; echo "rc=$?" >> $O/r2K_pytest.log; tail -6 $O/r2K_pytest.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
for i in 1 2; do
timeout 120 python bench.py $B > $O/r2K_tc_$i.json 2> $O/r2K_tc_$i.err; show r2K_tc_$i
PHNREC_WAVE_TC=0 timeout 120 python bench.py $B > $O/r2K_fft_$i.json 2> $O/r2K_fft_$i.err; show r2K_fft_$i
done
