#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > $O/r2b_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2b_pytest_gpu.log; tail -4 $O/r2b_pytest_gpu.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/wave_tc_check.py > $O/sanitize_memcheck_wave_tc.log 2>&1; echo "memcheck wave_tc_check rc=$?"; tail -4 $O/sanitize_memcheck_wave_tc.log
