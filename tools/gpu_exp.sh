mkdir -p gpurun_out
for e in 1 0 2 1; do
  export PHNREC_TC_E2=$e
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$e.json 2> gpurun_out/bench_e$e.err
  python - <<PY
import json
j=json.load(open('gpurun_out/bench_e$e.json'))
print("E2 policy $e: ms", round(j["ms_per_step"],3), "mlp", j["kernel_ms"]["mlp"], "frac", round(j["roofline"]["frac"],4))
PY
done
