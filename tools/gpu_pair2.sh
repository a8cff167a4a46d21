mkdir -p gpurun_out
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > gpurun_out/timeline_pair_$n.txt 2>&1; cat gpurun_out/timeline_pair_$n.txt; done
