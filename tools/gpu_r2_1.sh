#!/bin/bash
# Round 2, GPU call 1: full GPU test suite, tensor-core bound on the benchmarked workload, bench lines of every config,
# A/B of the packed-fp32 epilogue against the scalar one, band-net timeline.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest.log
tail -5 $O/r2_pytest.log
for c in cz:1000 hu:128 ru:128 en:128; do
  timeout 300 python tools/tc_bound.py --config ${c%%:*} --utts ${c##*:} --out $O/r2_tc_bound_${c%%:*}.json > /dev/null 2> $O/r2_tc_bound_${c%%:*}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_tc_bound_*.json")):
    j=json.load(open(f)); print(f, {k:j[k] for k in ("rel_logp_max","rel_logp_p999","rel_logp_mean","frame_argmax_agree","utt_identical","seg_agree","utt_same_phone_sequence","inf_mismatch")})
PY
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_cz.json 2> $O/r2_bench_cz.err; cat $O/r2_bench_cz.json | cut -c1-1500
PHNREC_B200_LIB=$PWD/phnrec_b200/lib/libphnrec_b200_scalar.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/r2_bench_cz_scalar.json 2> $O/r2_bench_cz_scalar.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $O/r2_bench_cz_b.json 2> $O/r2_bench_cz_b.err
for c in hu ru en en_sweep; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        j=json.load(open(f)); print(f, round(j["ms_per_step"],3), "ms", round(j["value"]), "xRT e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"), (j.get("roofline") or {}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 200 python tools/tc_timeline.py > $O/r2_timeline.txt 2>&1; tail -20 $O/r2_timeline.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; cut -c1-600 $O/r2_bench_ref.json
