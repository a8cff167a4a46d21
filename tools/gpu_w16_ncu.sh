#!/bin/bash
# ncu --set full capture of k_wave_tc16 (third EN step), with source
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wave_tc16 -s 2 -c 1 -f -o gpurun_out/w16 python tools/step_once.py en 3 > gpurun_out/w16_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/w16_ncu.log
ls -la gpurun_out/w16.ncu-rep
