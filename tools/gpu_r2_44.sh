#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_tensor_core.py tests/test_gpu_cli.py tests/test_gpu_async.py -q -x --timeout 200 > $O/r2c_pytest.log 2>&1; echo "rc=$?" >> $O/r2c_pytest.log; tail -3 $O/r2c_pytest.log
for i in 1 2; do timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > $O/r2c_bench_$i.json 2> $O/r2c_bench_$i.err; python -c "
import json; j=json.load(open('gpurun_out/r2c_bench_$i.json')); print(round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"; done
