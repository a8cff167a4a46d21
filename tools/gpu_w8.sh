#!/bin/bash
# k_wave_tc (8 kHz) variants: front-end tests, then the per-kernel times of two CZ bench runs
cd "$(dirname "$0")/.." || exit 1
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q -x -k "dft_front_end or front_end_mel" 2>&1 | tail -2
for i in 1 2; do timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > gpurun_out/w8_$i.json 2> gpurun_out/w8_$i.err; python -c "
import json; j=json.load(open('gpurun_out/w8_$i.json')); print(round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"; done
timeout 120 python bench.py --config cz_lin16 --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > gpurun_out/w8_l.json 2> gpurun_out/w8_l.err; python -c "
import json; j=json.load(open('gpurun_out/w8_l.json')); print('lin16', round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"
