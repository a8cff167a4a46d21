#!/bin/bash
# k_wave_tc16 variants: correctness on the ragged 16 kHz batch + the 16 kHz test, then the per-kernel times of an EN step
cd "$(dirname "$0")/.." || exit 1
timeout 300 python tools/wave_tc16_check.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -q -x -k "16khz" 2>&1 | tail -2
timeout 300 python bench.py --config en --steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > gpurun_out/w16_en.json 2> gpurun_out/w16_en.err; python -c "
import json; j=json.load(open('gpurun_out/w16_en.json')); print(round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in j['roofline']['kernels']])"
