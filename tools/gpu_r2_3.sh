#!/bin/bash
# Round 2, GPU call 3: V6 schedule of the tensor-core MLP (in-place H, two epilogue groups) vs V5; scalar-FMA K-stc; RU outlier fix.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 10 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1"
run() { # name, env...
  local n=$1; shift
  env "$@" timeout 120 python bench.py $B > $O/r2c_$n.json 2> $O/r2c_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r2c_{n}.json")); print(f"{n:10s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), j.get("kernel_ms"))
except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2c_{n}.err").read()[-800:])
PY
}
L=$PWD/phnrec_b200/lib
timeout 120 python -c "
import __graft_entry__ as g
g.smoke()" > $O/r2c_smoke.txt 2>&1; tail -3 $O/r2c_smoke.txt
run v6 X=1
run v5 PHNREC_TC_V5=1
run v6stcs PHNREC_B200_LIB=$L/libphnrec_b200_stcs.so
run v6b X=1
for c in hu ru en; do
  timeout 200 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-parity --profile-seconds 1 > $O/r2c_$c.json 2> $O/r2c_$c.err
  python -c "
import json
j=json.load(open('gpurun_out/r2c_$c.json')); print('$c', round(j['ms_per_step'],3), j['kernel_ms'])"
done
for c in cz:256 ru:128; do
  timeout 300 python tools/tc_bound.py --config ${c%%:*} --utts ${c##*:} --out $O/r2c_tc_bound_${c%%:*}.json > /dev/null 2> $O/r2c_tc_bound_${c%%:*}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_tc_bound_*.json")):
    j=json.load(open(f)); print(f, {k:j[k] for k in ("rel_logp_max","rel_logp_p999","rel_logp_mean","frame_argmax_agree","utt_identical","seg_agree","inf_mismatch")})
PY
timeout 900 python -m pytest tests -m gpu -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
tail -15 $O/r2c_pytest.log
