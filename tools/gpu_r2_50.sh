#!/bin/bash
# k_wave_tc with 8 producer warps for A-law: full GPU suite, default + lin16 bench lines, memcheck / racecheck of the ragged A-law batch
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q -x --timeout 200 > $O/r2i_pytest.log 2>&1; echo "rc=$?" >> $O/r2i_pytest.log; tail -3 $O/r2i_pytest.log
timeout 400 python bench.py > $O/r2i_bench_cz.json 2> $O/r2i_bench_cz.err; echo "rc=$?"
timeout 300 python bench.py --config cz_lin16 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2i_bench_cz_lin16.json 2> $O/r2i_bench_cz_lin16.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("cz","cz_lin16"):
    try:
        j=json.load(open(f"gpurun_out/r2i_bench_{n}.json")); print(f"{n:10s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j.get("parity",{}).get("seg_agree"), j["roofline"]["frac"])
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2i_bench_{n}.err").read()[-800:])
PY
cat > /tmp/san8.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import phnrec_b200 as pb
rec = pb.Recognizer('oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500', device=0)
rec.set_wave_format('alaw')
a = rec.synth_audio(24000, 8, seed=3)
utts = [a[0].tobytes(), a[1].tobytes()[:9001], a[2].tobytes()[:300], a[3].tobytes(), a[4].tobytes()[:16000], a[5].tobytes()[:201], a[6].tobytes()[:5], b"", a[7].tobytes()[:12345]]
rec.set_mlp_mode(pb.MLP_TC_F16)
print('CZ alaw', [len(l) for l in rec.recognize(utts)])
print('CZ alaw single', [len(l) for l in rec.recognize([utts[1]])])
rec.set_wave_format('lin16')
print('CZ lin16', [len(l) for l in rec.recognize(utts)])
rec.close()
PY
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san8.py > $O/sanitize_memcheck_wave_tc.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitize_memcheck_wave_tc.log
timeout 400 compute-sanitizer --tool racecheck --kernel-name kns=k_wave_tc --error-exitcode 9 python /tmp/san8.py > $O/sanitize_racecheck_wave_tc.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" $O/sanitize_racecheck_wave_tc.log; grep "Race reported\|Read access\|SUMMARY" $O/sanitize_racecheck_wave_tc.log | head -8
