#!/bin/bash
# shipped-state check after the live mode: smoke gate, the whole GPU suite, one default bench line, the reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2J_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/r2J_smoke.txt
if ! grep -q "mode 1 ok" $O/r2J_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 200 python -m pytest tests/test_gpu_cli.py -q -x --timeout 120 -k "live" > $O/r2J_pytest_live.log 2>&1; echo "rc=$?" >> $O/r2J_pytest_live.log; tail -8 $O/r2J_pytest_live.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 200 > $O/r2J_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2J_pytest_gpu.log; tail -6 $O/r2J_pytest_gpu.log
timeout 300 python bench.py > $O/r2J_bench.json 2> $O/r2J_bench.err; echo "bench rc=$?"; cut -c1-600 $O/r2J_bench.json
