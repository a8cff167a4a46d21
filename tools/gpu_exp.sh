mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensor_core.py -x -q 2>&1 | tail -5
for e in 0 1 3; do
  export PHNREC_TC_EXP=$e
  echo "=== EXP $e"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$e.json 2> gpurun_out/bench_e$e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_e$e.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_e$e.json'))
print("value", j["value"], "ms", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["kernel_ms"])
PY
  for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > gpurun_out/timeline_e${e}_$n.txt 2>&1; cat gpurun_out/timeline_e${e}_$n.txt; done
done
