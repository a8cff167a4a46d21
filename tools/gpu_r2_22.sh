#!/bin/bash
# 8 GPUs: the scaling bench line the driver will take (N = 8), plus the reference arm under torchrun
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2x_bench_8gpu.json 2> $O/r2x_bench_8gpu.err; cut -c1-900 $O/r2x_bench_8gpu.json; tail -3 $O/r2x_bench_8gpu.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2x_bench_8gpu.json")); print("8 GPUs: value", round(j["value"]/1e6,2), "M xRT,", round(j["ms_per_step"],3), "ms; e2e", round(j["e2e"]["value"]/1e6,2), "M xRT,", round(j["e2e"]["ms_per_step"],3), "ms", j["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/r2x_bench_4gpu.json 2> $O/r2x_bench_4gpu.err
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2x_bench_4gpu.json")); print("4 GPUs: value", round(j["value"]/1e6,2), "M xRT,", round(j["ms_per_step"],3), "ms; e2e", round(j["e2e"]["value"]/1e6,2), "M xRT,", round(j["e2e"]["ms_per_step"],3), "ms")
PY
