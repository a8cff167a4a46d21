#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
PHNREC_TC_PAIR=0 timeout 120 python tools/tc_timeline.py 0 > $O/r2z_timeline_single_0.txt 2>&1; cat $O/r2z_timeline_single_0.txt
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
PHNREC_TC_PAIR=0 timeout 200 python bench.py $B > $O/r2z_single.json 2> $O/r2z_single.err; python -c "
import json; j=json.load(open('gpurun_out/r2z_single.json')); print('single-CTA', j['kernel_ms'])"
