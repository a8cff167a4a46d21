#!/usr/bin/env python3
"""Where the time of back-to-back batches goes (PHNREC_TIMELINE): python tools/e2e_timeline.py device|async|sync [steps] [out]
device: phn_recognize_device on resident audio; async: phn_recognize_async + phn_wait (two in flight); sync: phn_recognize."""
import os
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
mode = sys.argv[1] if len(sys.argv) > 1 else "async"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = sys.argv[3] if len(sys.argv) > 3 else f"gpurun_out/e2e_timeline_{mode}.txt"
os.environ["PHNREC_TIMELINE"] = out
if os.path.exists(out):
    os.remove(out)
import phnrec_b200 as pb
n_utt, nbytes = 1000, 80000
rec = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500", device=0)
rec.set_wave_format("alaw")
rec.set_mlp_mode(pb.MLP_TC_F16)
boff = np.arange(n_utt + 1, dtype=np.int64) * nbytes
d = rec.device_alloc(n_utt * nbytes)
rec.synth_audio_device(d, nbytes, n_utt, seed=1000)
h = rec._L.phn_host_alloc_pinned(n_utt * nbytes)
rec.memcpy_d2h(h, d, n_utt * nbytes)
frames = n_utt * rec.num_frames(nbytes)
labels = np.zeros(frames + 48 * n_utt, dtype=pb.LABEL_DTYPE)
loff = np.zeros(n_utt + 1, dtype=np.int64)
flush = None
if os.environ.get("TL_FLUSH"):
    import torch
    stream = torch.cuda.ExternalStream(rec._L.phn_stream(rec._h), device=torch.device("cuda", 0))
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda:0")
t0 = time.perf_counter()
for k in range(steps):
    if flush is not None:
        with torch.cuda.stream(stream):
            flush.zero_()
    if mode == "device":
        rec.recognize_device(d, boff)
    elif mode == "sync":
        rec.recognize_raw(h, boff, labels, loff)
    else:
        rec.recognize_async_raw(h, boff)
        if rec.pending() == 2:
            rec.wait_raw(labels, loff)
while rec.pending():
    rec.wait_raw(labels, loff)
rec.sync()
print(f"{mode}: {(time.perf_counter() - t0) * 1e3 / steps:.3f} ms per step (host clock, {steps} steps incl. the first)")
rec.close()
txt = open(out).read()
rows = [l.split("|") for l in txt.splitlines() if l and l[0] != "#"]
go = [float(r[1].split()[1]) for r in rows]
print("wave-go deltas:", " ".join(f"{b - a:.2f}" for a, b in zip(go, go[1:])))
if steps <= 12:
    print(txt)
