#!/bin/bash
# Round 2, GPU call 4: V6 timeline; ncu --set full of one whole step (V6) and of the V5 MLP for comparison.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > $O/r2d_timeline_v6_$n.txt 2>&1; cat $O/r2d_timeline_v6_$n.txt; done
PHNREC_TC_V5=1 timeout 120 python tools/tc_timeline.py 0 > $O/r2d_timeline_v5_0.txt 2>&1
# steps: synth (1 launch) + per step 7 launches (wave, mean, stc, mlp x3, vit); skip synth + 2 warm steps
timeout 900 ncu --set full --import-source on --clock-control none -s 15 -c 7 -o $O/r2d_step_v6 -f python tools/step_once.py cz 3 > $O/r2d_ncu_v6.log 2>&1; tail -2 $O/r2d_ncu_v6.log
PHNREC_TC_V5=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mlp_tc -s 6 -c 1 -o $O/r2d_mlp_v5 -f python tools/step_once.py cz 3 > $O/r2d_ncu_v5.log 2>&1; tail -2 $O/r2d_ncu_v5.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_wave_pair|k_stc_f2" -s 4 -c 2 -o $O/r2d_front_en -f python tools/step_once.py en 3 > $O/r2d_ncu_en.log 2>&1; tail -2 $O/r2d_ncu_en.log
ls -la $O/*.ncu-rep
