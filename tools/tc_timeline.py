#!/usr/bin/env python3
"""GPU diagnostic: timeline (clock64) of one 128-frame tile of the tensor-core MLP kernel, CTA 0, second tile.
Usage: python tools/tc_timeline.py [net 0|1|2]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb  # noqa: E402

net = int(sys.argv[1]) if len(sys.argv) > 1 else 0
rec = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500", device=0)
rec.set_wave_format("alaw")
rec.set_mlp_mode(pb.MLP_TC_F16)
n_utt = 200
boff = np.arange(n_utt + 1, dtype=np.int64) * 80000
d = rec.device_alloc(n_utt * 80000)
rec.synth_audio_device(d, 80000, n_utt, seed=3)
for _ in range(2):
    rec.recognize_device(d, boff)
rec.sync()
L = rec._L
L.phn_debug_tc_timeline(rec._h, net, None)
rec.recognize_device(d, boff)
rec.sync()
out = np.zeros(256, dtype=np.int64)
L.phn_debug_tc_timeline(rec._h, -1, out.ctypes.data)
t = out.reshape(16, 16)
t0 = t[t > 0].min()
names = {2: "L1.beg", 3: "L1.ready", 4: "L1.issued", 5: "L2.beg", 6: "L2.ready", 7: "L2.issued", 8: "e1.wait", 9: "e1.D1", 10: "e1.math", 11: "e1.Hfree", 12: "e1.pub"}
print("chunk " + " ".join(f"{names[k]:>9s}" for k in sorted(names)))
for c in range(12):
    print(f"{c:5d} " + " ".join(f"{(t[c, k] - t0) if t[c, k] else -1:9d}" for k in sorted(names)))
e2 = ["A.begin", "A.D2seen", "A.end", "B.begin", "B.end", "C.begin", "C.end"]   # (V6 stamps 1..6 only)
print("e2 stages of that tile (they run during the next tile):", ", ".join(f"{n} {t[15, i] - t0}" for i, n in enumerate(e2) if t[15, i]))
