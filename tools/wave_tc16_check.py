"""K-wave of the 16 kHz systems on the tensor cores (k_wave_tc16.cu) against the exact front end: ragged lin16 batch."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import model_dir
import phnrec_b200 as pb

r = pb.Recognizer(model_dir("PHN_EN_TIMIT_LCRC_N500"), device=0)
r.set_wave_format("lin16")
rng = np.random.default_rng(5)
a = r.synth_audio(320000, 24, seed=13).copy()
a[2, 100000:200001] = 0
a[6, :9001] = 0
lens = [320000, 319999, 803, 802, 801, 800, 799, 401, 21, 1, 0, 64691, 256000, 1667] + [int(x) for x in rng.integers(500, 320000, 10)]
utts = [a[i].tobytes()[:n] for i, n in enumerate(lens)]
exact = np.concatenate(r.mel(utts))
r.set_mlp_mode(pb.MLP_TC_F16)
lab = r.recognize(utts)
fast = r.fetch_mel(exact.shape[0])
d = np.abs(fast - exact)
silent = (exact == 0.0).all(axis=1)
print("en16k frames", exact.shape[0], "finite", bool(np.isfinite(fast).all()), "max |dmel|", float(d.max()), "at", np.unravel_index(d.argmax(), d.shape),
      "rows >1e-4:", int((d.max(axis=1) > 1e-4).sum()), "silent rows", int(silent.sum()), "kept exactly 0:", bool((fast[silent] == 0.0).all()))
print("en16k labels", sum(len(x) for x in lab))
