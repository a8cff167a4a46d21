"""K-wave on the tensor cores (k_wave_tc.cu) against the exact front end: ragged A-law batch, max |mel difference|."""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import model_dir
import phnrec_b200 as pb

r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
r.set_wave_format("alaw")
rng = np.random.default_rng(3)
a = r.synth_audio(80000, 40, seed=11)
lens = [80000, 79999, 201, 200, 199, 5, 0, 12345, 64000, 333] + [int(x) for x in rng.integers(150, 80000, 30)]
utts = [a[i].tobytes()[:n] for i, n in enumerate(lens)]
exact = np.concatenate(r.mel(utts))
r.set_mlp_mode(pb.MLP_TC_F16)
lab = r.recognize(utts)
fast = r.fetch_mel(exact.shape[0])
d = np.abs(fast - exact)
print("frames", exact.shape[0], "finite", bool(np.isfinite(fast).all()), "max |dmel|", float(d.max()), "at", np.unravel_index(d.argmax(), d.shape),
      "rows >1e-4:", int((d.max(axis=1) > 1e-4).sum()))
print("labels", sum(len(x) for x in lab))

# lin16 (the CLI's default wave format): two parts per tile (high / low bytes); odd byte counts, odd byte offsets, digital silence
r.set_mlp_mode(pb.MLP_EXACT_FP32)
r.set_wave_format("lin16")
a = r.synth_audio(160000, 24, seed=12).copy()
a[3, 40000:90001] = 0
a[5, :7001] = 0
lens = [160000, 159999, 403, 402, 401, 400, 399, 11, 1, 0, 24691, 128000, 667] + [int(x) for x in rng.integers(300, 160000, 11)]
utts = [a[i].tobytes()[:n] for i, n in enumerate(lens)]
exact = np.concatenate(r.mel(utts))
r.set_mlp_mode(pb.MLP_TC_F16)
lab = r.recognize(utts)
fast = r.fetch_mel(exact.shape[0])
d = np.abs(fast - exact)
silent = (exact == 0.0).all(axis=1)
print("lin16 frames", exact.shape[0], "finite", bool(np.isfinite(fast).all()), "max |dmel|", float(d.max()), "at", np.unravel_index(d.argmax(), d.shape),
      "rows >1e-4:", int((d.max(axis=1) > 1e-4).sum()), "silent rows", int(silent.sum()), "kept exactly 0:", bool((fast[silent] == 0.0).all()))
print("lin16 labels", sum(len(x) for x in lab))
