# Round-end GPU evidence: tests, bench (both modes), timelines, ncu full capture of one step, ncu launch list.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --mode exact --no-cpu-baseline > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; echo "bench exact rc=$?"
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > gpurun_out/timeline_$n.txt 2>&1; done
# one whole step of the device-resident path under ncu --set full (trial call + 3 warm-up steps skipped: 7 launches each)
timeout 900 ncu --set full --import-source on --clock-control none -s 28 -c 7 -o gpurun_out/r1_step_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_tc.json
