#!/usr/bin/env python3
"""GPU diagnostic: tensor-core mode on growing batches (several tiles per CTA), compared with the exact mode."""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "PHN_CZ_SPDAT_LCRC_N1500"
sizes = [int(x) for x in sys.argv[2:]] or [256, 1000, 19000, 40000]
rec = pb.Recognizer(ROOT / "oracle/_ref/models" / name, device=0)
audio = (ROOT / "oracle/_ref/audio/test.raw").read_bytes()
mel = rec.mel([audio])[0]
for T in sizes:
    m = np.concatenate([mel] * (T // mel.shape[0] + 1))[:T]
    rec.set_mlp_mode(pb.MLP_EXACT_FP32)
    pe = rec.posteriors([m])[0]
    rec.set_mlp_mode(pb.MLP_TC_F16)
    print(f"T={T} grid={os.environ.get('PHNREC_TC_GRID', 'auto')} launching tc ...", flush=True)
    t0 = time.time()
    pt = rec.posteriors([m])[0]
    dt = time.time() - t0
    la, lb = np.log(np.maximum(pe, 1e-45)), np.log(np.maximum(pt, 1e-45))
    e = np.abs(la - lb) / np.maximum(1.0, np.abs(la))
    print(f"   done in {dt:.3f}s  rel-logp max={e.max():.3e} nan={int(np.isnan(pt).sum())}", flush=True)
