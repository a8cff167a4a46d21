#!/bin/bash
# Round 2, evidence run of the final build: tests, bench lines, launch list, ncu captures, timelines, sanitizer
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r2F_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2F_pytest.log; tail -4 $O/r2F_pytest.log
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2F_smoke.txt 2>&1; tail -2 $O/r2F_smoke.txt
timeout 400 python bench.py > $O/r2F_bench_cz.json 2> $O/r2F_bench_cz.err; cut -c1-300 $O/r2F_bench_cz.json
for c in hu ru en en_sweep; do timeout 400 python bench.py --config $c --no-cpu-baseline > $O/r2F_bench_$c.json 2> $O/r2F_bench_$c.err; done
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2F_bench_reference.json 2> $O/r2F_bench_reference.err; cut -c1-400 $O/r2F_bench_reference.json
timeout 400 python bench.py --mode exact --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/r2F_bench_exact.json 2> $O/r2F_bench_exact.err; cut -c1-200 $O/r2F_bench_exact.json
python - <<'PY'
import json
for c in ("cz","hu","ru","en","en_sweep","exact"):
    try:
        j=json.load(open(f"gpurun_out/r2F_bench_{c}.json")); print(f"{c:9s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), "frac", j["roofline"].get("frac"), j["roofline"].get("frac_burst"))
    except Exception as e: print(c, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2F_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > $O/r2F_launch_bench.log 2>&1
for n in 0 2; do timeout 120 python tools/tc_timeline.py $n > $O/r2F_timeline_$n.txt 2>&1; done
timeout 900 ncu --set full --clock-control none -k regex:"k_wave_pair|k_sentence_mean|k_stc_f2|k_mlp_tc|k_viterbi" -s 14 -c 7 -o $O/r2F_step -f python tools/step_once.py cz 3 > $O/r2F_ncu_step.log 2>&1; tail -2 $O/r2F_ncu_step.log
bash tools/gpu_sanitize.sh 2>&1 | tail -12
cp gpurun_out/sanitize_memcheck.log $O/r2F_sanitize_memcheck.log; cp gpurun_out/sanitize_memcheck_grid2.log $O/r2F_sanitize_memcheck_grid2.log; cp gpurun_out/sanitize_racecheck_wave.log $O/r2F_sanitize_racecheck_wave.log
ls -la $O/*.ncu-rep; du -sh $O
