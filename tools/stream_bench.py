#!/usr/bin/env python3
"""Throughput of the streaming mode: N concurrent streams, one block of B ms per stream per push (phn_stream_push).
    python tools/stream_bench.py [streams] [block_ms] [seconds_of_audio_per_stream] [exact|tc]
Prints one JSON line: audio-seconds per second over all streams, pushes per second, ms per push."""
import json
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
block_ms = int(sys.argv[2]) if len(sys.argv) > 2 else 125
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
mode = sys.argv[4] if len(sys.argv) > 4 else "tc"
rec = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500", device=0)
rec.set_wave_format("alaw")
rec.set_mlp_mode(pb.MLP_TC_F16 if mode == "tc" else pb.MLP_EXACT_FP32)
nbytes = int(secs * 8000)
audio = rec.synth_audio(nbytes, n, seed=3)
blk = block_ms * 8
rec.stream_open(n)
sids = list(range(n))
pushes = (nbytes + blk - 1) // blk
labels = 0
# warm-up on a separate short run
for k in range(3):
    rec.stream_push(sids, [audio[i, :blk].tobytes() for i in range(n)], [k == 2] * n)
t0 = time.perf_counter()
for k in range(pushes):
    blocks = [audio[i, k * blk:(k + 1) * blk] for i in range(n)]
    out = rec.stream_push(sids, blocks, [k == pushes - 1] * n)
    labels += sum(len(x) for x in out)
dt = time.perf_counter() - t0
print(json.dumps({"streams": n, "block_ms": block_ms, "mode": mode, "audio_s_per_stream": secs, "pushes": pushes, "ms_per_push": round(dt / pushes * 1e3, 3),
                  "xRT_all_streams": round(n * secs / dt, 1), "labels": labels,
                  "note": "host loop in Python (block slicing + ctypes call included); every push = H2D of the blocks, all kernels, labels D2H"}))
rec.close()
