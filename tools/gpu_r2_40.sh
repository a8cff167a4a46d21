#!/bin/bash
# 2-GPU box: list mode of the CLI on all GPUs (exact and tensor-core mode), the 2-rank bench line
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -4
timeout 400 python -m pytest tests/test_gpu_cli.py -q -x --timeout 300 -k "list_mode" > $O/r2Y_pytest_2gpu.log 2>&1; echo "rc=$?" >> $O/r2Y_pytest_2gpu.log; tail -4 $O/r2Y_pytest_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2Y_bench_cz_2gpu.json 2> $O/r2Y_bench_cz_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r2Y_bench_cz_2gpu.json")); print("2 GPUs:", round(j["ms_per_step"],3), "ms", round(j["value"]/1e6,3), "M xRT; e2e", round(j["e2e"]["ms_per_step"],3), round(j["e2e"]["value"]/1e6,3))
except Exception as e: print("ERR", e, open("gpurun_out/r2Y_bench_cz_2gpu.err").read()[-1500:])
PY
