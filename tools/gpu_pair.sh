mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -x -q 2>&1 | tail -6
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_pair.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_pair.json'))
    print("value", j["value"], "ms", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["kernel_ms"])
except Exception as e: print("no bench", e)
PY
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > gpurun_out/timeline_pair_$n.txt 2>&1; cat gpurun_out/timeline_pair_$n.txt; done
