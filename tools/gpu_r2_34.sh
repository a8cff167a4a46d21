#!/bin/bash
# Evidence run with the tensor-core front end: gate, whole GPU suite, default bench line + A/B against the FFT front end,
# the other configs, launch list and an ncu --set full capture of K-wave (k_wave_tc)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 90 python tools/wave_tc_check.py > $O/r2P_check.txt 2>&1; echo "check rc=$?"; tail -2 $O/r2P_check.txt
if ! grep -q "max |dmel|" $O/r2P_check.txt; then echo "CHECK FAILED - stopping"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q --timeout 200 > $O/r2P_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r2P_pytest_gpu.log; tail -5 $O/r2P_pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/{n}.json")); print(f"{n:18s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]])
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1500:])
PY
}
timeout 300 python bench.py > $O/r2P_bench_cz.json 2> $O/r2P_bench_cz.err; show r2P_bench_cz
B="--steps 20 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
for i in 1 2; do
timeout 120 python bench.py $B > $O/r2P_tc_$i.json 2> $O/r2P_tc_$i.err; show r2P_tc_$i
PHNREC_WAVE_TC=0 timeout 120 python bench.py $B > $O/r2P_fft_$i.json 2> $O/r2P_fft_$i.err; show r2P_fft_$i
done
for c in hu ru en; do
timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > $O/r2P_bench_$c.json 2> $O/r2P_bench_$c.err; show r2P_bench_$c
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2P_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2P_launch_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_wave_tc" -s 3 -c 1 -o $O/r2P_wave_tc -f python tools/step_once.py cz 3 > $O/r2P_ncu_wave.log 2>&1; tail -2 $O/r2P_ncu_wave.log
ls -la $O/r2P_wave_tc.ncu-rep
