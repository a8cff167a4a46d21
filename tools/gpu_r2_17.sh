#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python tools/cli_bench.py 1000 /tmp/phn_cli > $O/r2q_cli_bench.txt 2> $O/r2q_cli_bench.err; cat $O/r2q_cli_bench.txt; cat $O/r2q_cli_bench.err | tail -30
timeout 900 python tools/cli_bench.py 8000 /tmp/phn_cli8 > $O/r2q_cli_bench8.txt 2> $O/r2q_cli_bench8.err; cat $O/r2q_cli_bench8.txt; cat $O/r2q_cli_bench8.err | tail -16
