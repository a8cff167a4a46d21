#!/bin/bash
# Round 2, GPU call 2: diagnostics (cover property, RU outliers), K-stc fp32 form vs mma form, K-mlp epilogue variants.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python tools/diag_r2.py all > $O/r2_diag.txt 2>&1; tail -40 $O/r2_diag.txt
B="--steps 10 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
run() { # name, env...
  local n=$1; shift
  env "$@" timeout 300 python bench.py $B > $O/r2b_$n.json 2> $O/r2b_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2b_{n}.json")); print(f"{n:10s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2b_{n}.err").read()[-500:])
PY
}
L=$PWD/phnrec_b200/lib
run base X=1
run stcmma PHNREC_STC=mma
run ld32 PHNREC_B200_LIB=$L/libphnrec_b200_ld32.so
run rcp4 PHNREC_B200_LIB=$L/libphnrec_b200_rcp4.so
run ps PHNREC_B200_LIB=$L/libphnrec_b200_ps.so
run all PHNREC_B200_LIB=$L/libphnrec_b200_all.so
run base2 X=1
timeout 300 python tools/tc_bound.py --config cz --utts 256 --out $O/r2_tc_bound_cz256_f2.json > /dev/null 2>&1
timeout 300 python tools/tc_bound.py --config en --utts 128 --out $O/r2_tc_bound_en_f2.json > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_tc_bound_*f2.json")):
    j=json.load(open(f)); print(f, {k:j[k] for k in ("rel_logp_max","rel_logp_p999","rel_logp_mean","frame_argmax_agree","utt_identical","seg_agree")})
PY
timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> $O/r2_pytest2.log
tail -15 $O/r2_pytest2.log
