#!/bin/bash
# final evidence of the round: whole GPU suite, default bench line, the other configs, reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2W_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2W_smoke.txt
if ! grep -q "mode 1 ok" $O/r2W_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > $O/r2W_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2W_pytest_gpu.log; tail -4 $O/r2W_pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 400 python bench.py > $O/r2W_bench_cz.json 2> $O/r2W_bench_cz.err; show r2W_bench_cz
for c in hu ru en; do
timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $O/r2W_bench_$c.json 2> $O/r2W_bench_$c.err; show r2W_bench_$c
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2W_bench_reference.json 2> $O/r2W_bench_reference.err; cut -c1-400 $O/r2W_bench_reference.json
