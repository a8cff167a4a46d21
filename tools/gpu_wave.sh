# front-end check: tensor-core tests (mel closeness, silence, batch invariance), bench line
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_tensor_core.py -x -q 2>&1 | tail -8
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_q.json'))
print("value", j["value"], "ms", j["ms_per_step"], "e2e ms", j["e2e"]["ms_per_step"], "frac", j["roofline"]["frac"], j["kernel_ms"])
PY
