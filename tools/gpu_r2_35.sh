#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 90 python tools/wave_tc_check.py > $O/r2Q_check.txt 2>&1; echo "check rc=$?"; tail -2 $O/r2Q_check.txt
if ! grep -q "max |dmel|" $O/r2Q_check.txt; then echo "CHECK FAILED - stopping"; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_tensor_core.py tests/test_gpu_async.py tests/test_gpu_full_size.py -q --timeout 200 > $O/r2Q_pytest.log 2>&1; echo "rc=$?" >> $O/r2Q_pytest.log; tail -3 $O/r2Q_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2Q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2Q_launch_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_wave_tc" -s 1 -c 1 -o $O/r2Q_wave_tc -f python tools/step_once.py cz 3 > $O/r2Q_ncu_wave.log 2>&1; tail -2 $O/r2Q_ncu_wave.log
ls -la $O/r2Q_wave_tc.ncu-rep
