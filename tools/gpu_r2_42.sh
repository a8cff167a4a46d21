#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python tools/wave_tc_check.py > $O/r2a_check.txt 2>&1; echo "check rc=$?"; tail -5 $O/r2a_check.txt
if ! grep -q "lin16 labels" $O/r2a_check.txt; then echo "CHECK FAILED - stopping"; exit 1; fi
timeout 100 python tools/wave_lin16_time.py 2>&1 | tail -1
PHNREC_WAVE_TC=0 timeout 100 python tools/wave_lin16_time.py 2>&1 | tail -1
timeout 400 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_async.py tests/test_gpu_full_size.py -q -x --timeout 200 > $O/r2a_pytest.log 2>&1; echo "rc=$?" >> $O/r2a_pytest.log; tail -3 $O/r2a_pytest.log
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > $O/r2a_bench.json 2> $O/r2a_bench.err; python -c "
import json; j=json.load(open('gpurun_out/r2a_bench.json')); print(round(j['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"
