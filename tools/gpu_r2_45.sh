#!/bin/bash
# final lines: default bench (configs[1]), the same system with 16-bit linear input, smoke
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2d_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2d_smoke.txt
timeout 400 python bench.py > $O/r2d_bench_cz.json 2> $O/r2d_bench_cz.err; echo "rc=$?"
timeout 300 python bench.py --config cz_lin16 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2d_bench_cz_lin16.json 2> $O/r2d_bench_cz_lin16.err; echo "rc=$?"
PHNREC_WAVE_TC=0 timeout 300 python bench.py --config cz_lin16 --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/r2d_bench_cz_lin16_fft.json 2> $O/r2d_bench_cz_lin16_fft.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("cz","cz_lin16","cz_lin16_fft"):
    try:
        j=json.load(open(f"gpurun_out/r2d_bench_{n}.json")); print(f"{n:14s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j.get("parity",{}).get("seg_agree"))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2d_bench_{n}.err").read()[-800:])
PY
