#!/bin/bash
# the whole GPU suite on the final build
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/r2j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_pytest.log
