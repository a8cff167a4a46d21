#!/bin/bash
# EN bench line with the tensor_work entry of the 16 kHz front end; racecheck of k_wave_tc16 on a small ragged batch
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --config en --steps 20 --warmup 3 --no-cpu-baseline > $O/r2h_bench_en.json 2> $O/r2h_bench_en.err; echo "rc=$?"
python -c "
import json; j=json.load(open('gpurun_out/r2h_bench_en.json')); print(round(j['ms_per_step'],3), round(j['e2e']['ms_per_step'],3), [k for k in j['roofline']['kernels'] if k['kernel']=='K-wave'])"
cat > /tmp/san16.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import phnrec_b200 as pb
rec = pb.Recognizer('oracle/_ref/models/PHN_EN_TIMIT_LCRC_N500', device=0)
a = rec.synth_audio(48000, 4, seed=5)
utts = [a[0].tobytes(), a[1].tobytes()[:9001], a[2].tobytes()[:700], a[3].tobytes()[:20000]]
rec.set_mlp_mode(pb.MLP_TC_F16)
rec.set_force_exact_wave(False) if hasattr(rec, 'set_force_exact_wave') else None
m = rec.mel_tc(utts) if hasattr(rec, 'mel_tc') else None
print('EN16k', [len(l) for l in rec.recognize(utts)])
rec.close()
PY
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=k_wave_tc16 --error-exitcode 9 python /tmp/san16.py > $O/sanitize_racecheck_wave_tc16.log 2>&1; echo "racecheck rc=$?"; tail -6 $O/sanitize_racecheck_wave_tc16.log
