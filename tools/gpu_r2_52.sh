#!/bin/bash
# final build (decoder form chosen by residency): whole GPU suite, HU / default bench lines, memcheck of a 1000-utterance HU batch
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x --timeout 200 > $O/r2l_pytest.log 2>&1; echo "rc=$?" >> $O/r2l_pytest.log; tail -3 $O/r2l_pytest.log
timeout 400 python bench.py --config hu --steps 20 --warmup 3 --no-cpu-baseline > $O/r2l_bench_hu.json 2> $O/r2l_bench_hu.err; echo "rc=$?"
timeout 400 python bench.py > $O/r2l_bench_cz.json 2> $O/r2l_bench_cz.err; echo "rc=$?"
python - <<'PY'
import json
for n in ("hu","cz"):
    try:
        j=json.load(open(f"gpurun_out/r2l_bench_{n}.json")); print(f"{n:4s}", round(j["ms_per_step"],3), "ms e2e", round(j["e2e"]["ms_per_step"],3), [(k["kernel"], k["ms"]) for k in j["roofline"]["kernels"]], j.get("parity",{}).get("seg_agree"), round(j["roofline"]["frac"],3))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r2l_bench_{n}.err").read()[-800:])
PY
cat > /tmp/sanhu.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import phnrec_b200 as pb
rec = pb.Recognizer('oracle/_ref/models/PHN_HU_SPDAT_LCRC_N1500', device=0)
rec.set_wave_format('alaw')
a = rec.synth_audio(4000, 4, seed=3).reshape(-1)
rng = np.random.default_rng(1)
utts = []
for _ in range(1000):
    n = int(rng.integers(150, 900)); o = int(rng.integers(0, a.size - n)); utts.append(a[o:o + n].tobytes())
utts += [b"", a[:5].tobytes()]
rec.set_mlp_mode(pb.MLP_TC_F16)
print('HU 1002 utterances', sum(len(l) for l in rec.recognize(utts)))
rec.close()
PY
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/sanhu.py > $O/sanitize_memcheck_vit_direct.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitize_memcheck_vit_direct.log
