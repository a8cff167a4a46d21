#!/usr/bin/env python3
"""GPU diagnostic: tensor-core MLP mode against the exact fp32 mode (same library, same inputs).
Prints the deviation statistics quoted in DESIGN.md.  Usage: python tools/tc_check.py [model ...]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb  # noqa: E402

MODELS = sys.argv[1:] or ["PHN_CZ_SPDAT_LCRC_N1500", "PHN_EN_TIMIT_LCRC_N500", "PHN_HU_SPDAT_LCRC_N1500", "PHN_RU_SPDAT_LCRC_N1500"]
audio = (ROOT / "oracle/_ref/audio/test.raw").read_bytes()

for name in MODELS:
    rec = pb.Recognizer(ROOT / "oracle/_ref/models" / name, device=0)
    utts = [audio, audio[:50000], audio[20000:90000]]
    if rec.sample_freq == 8000:
        rec.set_wave_format("alaw")
        syn = rec.synth_audio(80000, 16, seed=7)
        rec.set_wave_format("lin16")
    mels = rec.mel(utts)
    rec.set_mlp_mode(pb.MLP_EXACT_FP32)
    pe = rec.posteriors(mels)
    le = rec.decode(pe)
    rec.set_mlp_mode(pb.MLP_TC_F16)
    t0 = time.time()
    pt = rec.posteriors(mels)
    print(f"{name}: tc posteriors call took {time.time() - t0:.3f}s", flush=True)
    lt = rec.decode(pt)
    P3 = rec.n_phonemes * 3
    for u, (a, b) in enumerate(zip(pe, pt)):
        a = a[:, :P3].astype(np.float64); b = b[:, :P3].astype(np.float64)
        nan = int(np.isnan(b).sum())
        la, lb = np.log(np.maximum(a, 1e-45)), np.log(np.maximum(b, 1e-45))
        m = np.abs(la - lb) / np.maximum(1.0, np.abs(la))
        agree = float((a.argmax(1) == b.argmax(1)).mean())
        lab_same = [(int(x["start"]), int(x["end"]), int(x["phn"])) for x in le[u]] == [(int(x["start"]), int(x["end"]), int(x["phn"])) for x in lt[u]]
        print(f"  utt{u} T={a.shape[0]} nan={nan} max|dp|={np.abs(a - b).max():.3e} rel-logp max={m.max():.3e} p99.9={np.quantile(m, 0.999):.3e} "
              f"mean={m.mean():.3e} max|dlogp|={np.abs(la - lb).max():.3e} argmax-agree={agree:.4f} labels-identical={lab_same} "
              f"rowsum={b.sum(1).min():.4f}..{b.sum(1).max():.4f}", flush=True)
    rec.close()
