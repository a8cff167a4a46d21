#!/usr/bin/env python3
"""CLI-level throughput ("run the CLI itself on the 1000-utterance set"): writes N synthetic 10 s A-law files,
runs `phnrec -c CZ -l list -m out.mlf -w alaw` in the tensor-core mode on 1 and on all visible GPUs, prints one JSON line each.
    python tools/cli_bench.py [n_files] [dir]"""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tools.synth_host import synth_audio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
base = Path(sys.argv[2]) if len(sys.argv) > 2 else Path(tempfile.mkdtemp(prefix="phn_cli_"))
base.mkdir(parents=True, exist_ok=True)
model = ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500"
a = synth_audio(80000, n, seed=1000, fmt="alaw", fs=8000)
names = []
for i in range(n):
    f = base / f"u{i:05d}.raw"
    f.write_bytes(a[i].tobytes())
    names.append(str(f))
lst = base / "list.scp"
lst.write_text("\n".join(names) + "\n")
import ctypes
ndev = 1
try:
    import phnrec_b200 as pb
    ndev = int(pb.api.load_library().phn_device_count())
except Exception:
    pass
outs = {}
for tag, env in [("1gpu", {"PHNREC_DEVICES": "0"})] + ([("all", {"PHNREC_DEVICES": "all"})] if ndev > 1 else []):
    for rep in range(2):   # second run: files in the page cache, GPU warm
        mlf = base / f"out_{tag}.mlf"
        t0 = time.perf_counter()
        r = subprocess.run([str(ROOT / "phnrec_b200/bin/phnrec"), "-c", str(model), "-l", str(lst), "-m", str(mlf), "-w", "alaw"],
                           env={**os.environ, "PHNREC_MLP": "tc", "PHNREC_CLI_TIMING": "1", **env}, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
    outs[tag] = mlf.read_bytes()
    sys.stderr.write(f"--- {tag}\n{r.stderr}")
    print(json.dumps({"cli": "phnrec -l list -m out.mlf (PHNREC_MLP=tc)", "devices": env["PHNREC_DEVICES"], "n_gpus_visible": ndev, "files": n,
                      "audio_s": n * 10.0, "wall_s": round(dt, 3), "xRT": round(n * 10.0 / dt, 1),
                      "note": "whole process: CUDA context + model load, reading the files, recognition, writing the MLF"}))
if "all" in outs:
    print(json.dumps({"mlf_identical_1gpu_vs_all": outs["all"] == outs["1gpu"]}))
