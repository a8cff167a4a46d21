#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
PHNREC_WTC_DBG=8 timeout 200 python tools/step_once.py cz 2 > gpurun_out/w8_tl.txt 2>&1; tail -40 gpurun_out/w8_tl.txt
