#!/usr/bin/env python3
"""One configuration, N device-resident steps (for ncu captures): python tools/step_once.py [cfg] [steps] [utts]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb
CFG = {"cz": ("PHN_CZ_SPDAT_LCRC_N1500", "alaw", 80000), "en": ("PHN_EN_TIMIT_LCRC_N500", "lin16", 320000)}
cfg = sys.argv[1] if len(sys.argv) > 1 else "cz"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_utt = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
model, fmt, nbytes = CFG[cfg]
rec = pb.Recognizer(ROOT / "oracle/_ref/models" / model, device=0)
rec.set_wave_format(fmt)
rec.set_mlp_mode(pb.MLP_TC_F16)
boff = np.arange(n_utt + 1, dtype=np.int64) * nbytes
d = rec.device_alloc(n_utt * nbytes)
rec.synth_audio_device(d, nbytes, n_utt, seed=1000)
for _ in range(steps):
    rec.recognize_device(d, boff)
    rec.sync()
print("done")
