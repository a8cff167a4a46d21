"""K-wave time for 8 kHz lin16 input in the tensor-core pipeline (PHNREC_WAVE_TC=0: register FFT, default: DFT on the tensor cores)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb
n_utt, nbytes = 1000, 160000
rec = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500", device=0)
rec.set_wave_format("lin16")
rec.set_mlp_mode(pb.MLP_TC_F16)
boff = np.arange(n_utt + 1, dtype=np.int64) * nbytes
d = rec.device_alloc(n_utt * nbytes)
rec.synth_audio_device(d, nbytes, n_utt, seed=1000)
for _ in range(3):
    rec.recognize_device(d, boff); rec.sync()
rec.set_profiling(True)
acc = {}
for _ in range(20):
    rec.recognize_device(d, boff); rec.sync()
    for k, (ms, n) in rec.last_timing().items():
        acc[k] = acc.get(k, 0.0) + ms / 20
print({k: round(v, 4) for k, v in acc.items()})
