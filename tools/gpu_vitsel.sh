#!/bin/bash
# decoder form chosen by residency: HU (auto / forced panels), EN sweep (panels / forced direct), CZ unchanged; decoder parity tests
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_core.py -q -x 2>&1 | tail -2
run() { timeout 300 python bench.py --config $1 --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > $O/vs_$2.json 2> $O/vs_$2.err; python -c "
import json; j=json.load(open('gpurun_out/vs_$2.json')); print('$2', round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"; }
run hu hu_auto
PHNREC_VIT_DIRECT=0 run hu hu_panels
run cz cz_auto
run en_sweep sweep_auto
PHNREC_VIT_DIRECT=1 run en_sweep sweep_direct
run ru ru_auto
