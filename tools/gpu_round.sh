set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > gpurun_out/timeline_$n.txt 2>&1; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mlp_tc -s 12 -c 3 -o gpurun_out/r1_mlp_v4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_v4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_tc.json
cat gpurun_out/timeline_0.txt
