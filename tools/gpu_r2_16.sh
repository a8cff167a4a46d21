#!/bin/bash
# 2 GPUs: full test suite (list-mode tests use all visible GPUs), CLI throughput 1 vs 2 GPUs, streaming throughput, 2-GPU bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -q > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2p_pytest.log
tail -5 $O/r2p_pytest.log
timeout 600 python tools/cli_bench.py 1000 /tmp/phn_cli > $O/r2p_cli_bench.txt 2> $O/r2p_cli_bench.err; cat $O/r2p_cli_bench.txt; tail -3 $O/r2p_cli_bench.err
for cfg in "1024 125 4 tc" "1024 500 4 tc" "256 125 4 exact"; do timeout 300 python tools/stream_bench.py $cfg >> $O/r2p_stream_bench.txt 2>> $O/r2p_stream_bench.err; done; cat $O/r2p_stream_bench.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2p_bench_2gpu.json 2> $O/r2p_bench_2gpu.err; cat $O/r2p_bench_2gpu.json | cut -c1-600
