#!/bin/bash
# Round 2, GPU call 5: default = V5 schedule + FFMA2 K-stc.  Tests, bench lines of every config, launch list, ncu --set full
# of one step (kept under the 64 MiB pull limit: no source import for the step capture).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
tail -5 $O/r2e_pytest.log
timeout 400 python bench.py > $O/r2e_bench_cz.json 2> $O/r2e_bench_cz.err; cat $O/r2e_bench_cz.json
for c in hu ru en en_sweep; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > $O/r2e_bench_$c.json 2> $O/r2e_bench_$c.err; cat $O/r2e_bench_$c.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2e_launch_bench.log 2>&1
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > $O/r2e_timeline_$n.txt 2>&1; done
timeout 900 ncu --set full --clock-control none -s 15 -c 7 -o $O/r2e_step -f python tools/step_once.py cz 3 > $O/r2e_ncu_step.log 2>&1; tail -2 $O/r2e_ncu_step.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mlp_tc -s 6 -c 1 -o $O/r2e_mlp_band -f python tools/step_once.py cz 3 > $O/r2e_ncu_mlp.log 2>&1; tail -2 $O/r2e_ncu_mlp.log
ls -la $O/*.ncu-rep; du -sh $O
