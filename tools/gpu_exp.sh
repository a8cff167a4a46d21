mkdir -p gpurun_out
for e in 0 1 3; do
  export PHNREC_TC_E2=$e
  echo "=== E2 policy $e"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$e.json 2> gpurun_out/bench_e$e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_e$e.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_e$e.json'))
print("value", j["value"], "ms", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["kernel_ms"])
PY
done
