#!/bin/bash
# 2-GPU check of the final build: bench under torchrun (both arms), list-mode / multi-device tests
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2g_bench_cz_2gpu.json 2> $O/r2g_bench_cz_2gpu.err; echo "rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2g_bench_ref_2gpu.json 2> $O/r2g_bench_ref_2gpu.err; echo "rc=$?"
timeout 400 python -m pytest tests/test_gpu_cli.py tests/test_gpu_async.py -q -x --timeout 200 > $O/r2g_pytest_2gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2g_pytest_2gpu.log
python - <<'PY'
import json
for n in ("cz_2gpu","ref_2gpu"):
    try:
        t=[l for l in open(f"gpurun_out/r2g_bench_{n}.json") if l.startswith("{")]
        j=json.loads(t[-1]); print(n, j.get("n_gpus"), j.get("value"), j.get("ms_per_step"), (j.get("e2e") or {}).get("value"), j.get("impl"))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2g_bench_{n}.err").read()[-600:])
PY
