#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_plp.py tests/test_gpu_trap_systems.py tests/test_gpu_parity.py -q -x > $O/r2w_pytest_trap.log 2>&1; echo "rc=$?" >> $O/r2w_pytest_trap.log
tail -30 $O/r2w_pytest_trap.log
