#!/bin/bash
# final evidence of the build with the 16 kHz tensor-core front end: launch list of the EN config, default bench line (with CPU baselines), reference arm, smoke
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2f_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2f_smoke.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_en.csv python bench.py --config en --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2f_launch_bench.log 2>&1; echo "launch list rc=$?"
timeout 400 python bench.py > $O/r2f_bench_cz.json 2> $O/r2f_bench_cz.err; echo "rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2f_bench_reference.json 2> $O/r2f_bench_reference.err; echo "rc=$?"
python - <<'PY'
import json
j=json.load(open("gpurun_out/r2f_bench_cz.json")); print("cz", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j["roofline"]["frac"], j["cpu_baseline"]["value"])
j=json.load(open("gpurun_out/r2f_bench_reference.json")); print("ref", j.get("value"), j.get("unit"), j.get("impl"))
PY
