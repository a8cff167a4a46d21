# memcheck of the fused path on a small ragged batch (both MLP modes)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import phnrec_b200 as pb
rec = pb.Recognizer('oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500', device=0)
rec.set_wave_format('alaw')
a = rec.synth_audio(24000, 5, seed=3)
utts = [a[0].tobytes(), a[1].tobytes()[:9001], a[2].tobytes()[:300], a[3].tobytes(), a[4].tobytes()[:16000]]
for mode in (pb.MLP_TC_F16, pb.MLP_EXACT_FP32):
    rec.set_mlp_mode(mode)
    lab = rec.recognize(utts)
    print(mode, [len(l) for l in lab])
rec.close()
# 16 kHz lin16 system (512-point FFT, 23 banks: the two-pass filterbank of the fast front end)
rec = pb.Recognizer('oracle/_ref/models/PHN_EN_TIMIT_LCRC_N500', device=0)
a = rec.synth_audio(48000, 3, seed=5)
utts = [a[0].tobytes(), a[1].tobytes()[:9002], a[2].tobytes()[:700]]
rec.set_mlp_mode(pb.MLP_TC_F16)
print('EN', [len(l) for l in rec.recognize(utts)])
rec.close()
# streaming mode (state kernels, windows, resumable decoder) and two batches in flight
rec = pb.Recognizer('oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500', device=0)
a = open('oracle/_ref/audio/test.raw', 'rb').read()
rec.stream_open(3)
n = 0
for k in range(0, 30000, 3001):
    last = k + 3001 >= 30000
    out = rec.stream_push([0, 2], [a[k:k + 3001], a[40000 + k:40000 + k + 2000]], [last, last])
    n += sum(len(x) for x in out)
print('stream labels', n)
rec.set_mlp_mode(pb.MLP_TC_F16)
print('pipelined', [len(b) for b in rec.recognize_pipelined([[a[:20000], a[:9000]], [a[20000:50000]], [a[:398], a[1000:12000]]])])
rec.close()
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/sanitize_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --kernel-regex kns=k_wave_pair python /tmp/san.py > gpurun_out/sanitize_racecheck_wave.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/sanitize_racecheck_wave.log
# the same batch with one CTA pair for all tiles: several tiles per CTA, so the activation ring of the MLP kernel wraps
PHNREC_TC_GRID=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_memcheck_grid2.log 2>&1; echo "memcheck (one pair) rc=$?"
tail -6 gpurun_out/sanitize_memcheck_grid2.log
# the tensor-core front end: shared-memory hazards between its roles (racecheck sees generic-proxy accesses), and
# memcheck + initcheck of a batch whose last utterances end at the very end of the audio buffer
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 --kernel-regex kns=k_wave_tc python /tmp/san.py > gpurun_out/sanitize_racecheck_wave_tc.log 2>&1; echo "racecheck k_wave_tc rc=$?"
tail -5 gpurun_out/sanitize_racecheck_wave_tc.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/wave_tc_check.py > gpurun_out/sanitize_memcheck_wave_tc.log 2>&1; echo "memcheck wave_tc_check rc=$?"
tail -6 gpurun_out/sanitize_memcheck_wave_tc.log
