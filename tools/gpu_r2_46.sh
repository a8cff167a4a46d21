#!/bin/bash
# 16 kHz tensor-core front end (k_wave_tc16.cu): full GPU suite, EN bench lines with both front ends, default line, memcheck of the 16 kHz batch
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q -x --timeout 200 > $O/r2e_pytest.log 2>&1; echo "rc=$?" >> $O/r2e_pytest.log; tail -3 $O/r2e_pytest.log
timeout 300 python bench.py --config en --steps 20 --warmup 3 --no-cpu-baseline > $O/r2e_bench_en.json 2> $O/r2e_bench_en.err; echo "rc=$?"
PHNREC_WAVE_TC=0 timeout 300 python bench.py --config en --steps 20 --warmup 3 --no-cpu-baseline --no-parity > $O/r2e_bench_en_fft.json 2> $O/r2e_bench_en_fft.err; echo "rc=$?"
timeout 300 python bench.py --config en_sweep --steps 20 --warmup 3 --no-cpu-baseline > $O/r2e_bench_en_sweep.json 2> $O/r2e_bench_en_sweep.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2e_bench_cz.json 2> $O/r2e_bench_cz.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("en","en_fft","en_sweep","cz"):
    try:
        j=json.load(open(f"gpurun_out/r2e_bench_{n}.json")); print(f"{n:10s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j.get("parity",{}).get("seg_agree"))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2e_bench_{n}.err").read()[-800:])
PY
cat > /tmp/san16.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import phnrec_b200 as pb
rec = pb.Recognizer('oracle/_ref/models/PHN_EN_TIMIT_LCRC_N500', device=0)
a = rec.synth_audio(48000, 6, seed=5)
utts = [a[0].tobytes(), a[1].tobytes()[:9001], a[2].tobytes()[:700], a[3].tobytes()[:801], a[4].tobytes()[:3], b"", a[5].tobytes()[:20000]]
rec.set_mlp_mode(pb.MLP_TC_F16)
print('EN16k', [len(l) for l in rec.recognize(utts)])
print('EN16k single', [len(l) for l in rec.recognize([utts[1]])])
rec.close()
PY
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san16.py > $O/sanitize_memcheck_wave_tc16.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitize_memcheck_wave_tc16.log
timeout 400 compute-sanitizer --tool racecheck --kernel-regex kns=k_wave_tc16 --error-exitcode 9 python /tmp/san16.py > $O/sanitize_racecheck_wave_tc16.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" $O/sanitize_racecheck_wave_tc16.log; tail -3 $O/sanitize_racecheck_wave_tc16.log
