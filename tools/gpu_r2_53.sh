#!/bin/bash
# 8-GPU bench line of the final build
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2m_bench_cz_8gpu.json 2> $O/r2m_bench_cz_8gpu.err; echo "rc=$?"
python - <<'PY'
import json
try:
    t=[l for l in open("gpurun_out/r2m_bench_cz_8gpu.json") if l.startswith("{")]
    j=json.loads(t[-1]); print(j.get("n_gpus"), j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"), (j.get("e2e") or {}).get("ms_per_step"))
except Exception as e: print("ERR", e, open("gpurun_out/r2m_bench_cz_8gpu.err").read()[-600:])
PY
