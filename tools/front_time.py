#!/usr/bin/env python3
"""GPU diagnostic: per-kernel times (phn_last_timing) of the fused tensor-core path for one shipped system.
Usage: python tools/front_time.py [model dir name] [alaw|lin16] [utterances] [seconds per utterance]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "PHN_EN_TIMIT_LCRC_N500"
fmt = sys.argv[2] if len(sys.argv) > 2 else "lin16"
n_utt = int(sys.argv[3]) if len(sys.argv) > 3 else 500
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 10.0
rec = pb.Recognizer(ROOT / "oracle/_ref/models" / name, device=0)
rec.set_wave_format(fmt)
rec.set_mlp_mode(pb.MLP_TC_F16)
nbytes = int(secs * rec.sample_freq) * (1 if fmt == "alaw" else 2)
boff = np.arange(n_utt + 1, dtype=np.int64) * nbytes
d = rec.device_alloc(n_utt * nbytes)
rec.synth_audio_device(d, nbytes, n_utt, seed=3)
rec.set_profiling(True)
best = None
for _ in range(6):
    rec.recognize_device(d, boff)
    rec.sync()
    t = rec.last_timing()
    if best is None or t["wave"][0] < best["wave"][0]:
        best = t
print(name, fmt, n_utt, "utterances x", secs, "s:", {k: round(v[0], 4) for k, v in best.items()})
