#!/bin/bash
# HU / RU bench lines of the final build
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for c in hu ru; do timeout 400 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $O/r2k_bench_$c.json 2> $O/r2k_bench_$c.err; echo "rc=$?"; done
python - <<'PY'
import json
for n in ("hu","ru"):
    try:
        j=json.load(open(f"gpurun_out/r2k_bench_{n}.json")); print(f"{n:4s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j.get("parity",{}).get("seg_agree"), round(j["roofline"]["frac"],3))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2k_bench_{n}.err").read()[-800:])
PY
