#!/bin/bash
# launch list of the final build (default config)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2n_launch_bench.log 2>&1; echo "rc=$?"
wc -l $O/r2n_launches.csv
