#!/usr/bin/env python3
"""bench.py — audio-seconds per second (xRT) of the PhnRec recognition hot path on B200.

Workload (BASELINE.json configs[1], per GPU): PHN_CZ_SPDAT_LCRC_N1500, 8 kHz A-law, 1000 synthetic
10 s utterances.  A step is one pass of the whole path (wave -> mel -> STC -> 3 MLPs -> Viterbi ->
labels) over that batch.  N > 1: one process per GPU (torchrun), every rank runs its own 1000
utterances, no data-path collective (utterances are independent) -> weak scaling.

  value : inputs resident in HBM when the timed region starts, labels left on the device
  e2e   : the reference-facing C-ABI call phn_recognize() with HOST buffers (pinned audio in,
          label arrays out), H2D and D2H copies inside the timed region
  --impl reference : the reference's own CPU implementation (oracle/_ref/phnrec_ref, the
          reference sources compiled by oracle/Makefile) on all host cores, bounded sample.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL = "PHN_CZ_SPDAT_LCRC_N1500"
FS = 8000
UTT_SECONDS = 10.0
BYTES_PER_UTT = 80000           # 10 s of 8 kHz A-law
FRAMES_PER_UTT = 998            # (80000 - 200) / 80 + 1, srec.cpp:945
FLOP_PER_FRAME = 2 * 1530000    # 3 MLPs, from the .nbin header sizes


def model_dir() -> Path:
    p = ROOT / "oracle" / "_ref" / "models" / MODEL
    if (p / "weights" / "merger.nbin").exists():
        return p
    return synth_model_dir()


def synth_model_dir() -> Path:
    """Random-init weights of the CZ N1500 architecture (used only when the staged model data is absent)."""
    d = Path(tempfile.gettempdir()) / "phnrec_b200_synth_model"
    if (d / "weights" / "merger.nbin").exists():
        return d
    for s in ("weights", "windows", "dicts"):
        (d / s).mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(0)

    def nbin(path, nin, nhid, nout):
        up4 = lambda n: (n + 3) // 4 * 4
        w1 = np.zeros((up4(nhid), up4(nin)), np.float32); w1[:nhid, :nin] = rng.standard_normal((nhid, nin)) / np.sqrt(nin)
        w2 = np.zeros((up4(nout), up4(nhid)), np.float32); w2[:nout, :nhid] = rng.standard_normal((nout, nhid)) / np.sqrt(nhid)
        b1 = np.zeros(up4(nhid), np.float32); b2 = np.zeros(up4(nout), np.float32)
        mean = np.zeros(up4(nin), np.float32); dev = np.ones(up4(nin), np.float32)
        with open(path, "wb") as f:
            f.write(np.array([2, nin, nhid, nout], np.int32).tobytes())
            for a in (w1, w2, b1, b2, mean, dev):
                f.write(a.tobytes())
    nbin(d / "weights" / "band0.nbin", 165, 1500, 138)
    nbin(d / "weights" / "band1.nbin", 165, 1500, 138)
    nbin(d / "weights" / "merger.nbin", 276, 1500, 138)
    ham = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(31) / 30)
    (d / "windows" / "band0.window").write_text(" ".join("%f" % v for v in ham[:16]) + "\n")
    (d / "windows" / "band1.window").write_text(" ".join("%f" % v for v in ham[15:]) + "\n")
    (d / "dicts" / "phonemes").write_text("".join("p%d\n" % i for i in range(45)))
    (d / "config").write_text(
        "[source]\nformat=lin16\nsample_freq=8000\n\n[posteriors]\nsystem=LCRC\nlength=31\nadd_c0=true\nhamming=false\n"
        "bunch_size=5\nsoftening_func=none 0 0 0\n\n[params]\nkind=fbanks\n\n[melbanks]\nnbanks=15\nlower_freq=64\n"
        "higher_freq=4000\nvector_size=200\nvector_step=80\n\n[decoder]\ntype=phndec\nnum_states_per_phn=3\n"
        "softening_func=log 0 0 0\nwpenalty=-4.6875\ntime_pruning=40\n\n[offlinenorm]\nsent_mean_norm=true\n\n"
        "[dicts]\nphoneme_list=$C/dicts/phonemes\n")
    return d


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_traffic(frames):
    """DRAM bytes of the MLP launches of one step from the committed ncu --set full capture (same workload), or None."""
    p = ROOT / "profiles" / "r1_mlp_traffic.json"
    try:
        j = json.loads(p.read_text())
        if int(j.get("frames", -1)) == int(frames):
            return int(j["k_mlp_per_step"])
    except Exception:
        pass
    return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"tflops": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0))), "hbm_gbs": float(j["hbm_gbs"]),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


# ------------------------------------------------------------------------------------------ reference CPU arm
def cpu_reference_run(audio: np.ndarray, mdir: Path, cores: int):
    """The reference's own implementation on `cores` host cores: the list is split into `cores` slices,
    one phnrec process each (the reference is single threaded), wall clock first start -> last exit.
    Uses oracle/_ref/phnrec_ref (kind "reference") when built, else the C restatement (kind "port")."""
    from oracle import oracle as orc  # checker / baseline only
    n_utt = audio.shape[0]
    secs = n_utt * UTT_SECONDS
    if orc.have_ref():
        td = Path(tempfile.mkdtemp(prefix="phn_cpu_"))
        try:
            lists = []
            for k in range(cores):
                idx = list(range(k, n_utt, cores))
                if not idx:
                    continue
                lines = []
                for u in idx:
                    f = td / f"u{u}.raw"
                    f.write_bytes(audio[u].tobytes())
                    lines.append(str(f))
                (td / f"l{k}.scp").write_text("\n".join(lines) + "\n")
                lists.append(k)
            t0 = time.perf_counter()
            procs = [subprocess.Popen([str(orc.REF_BIN), "-c", str(mdir), "-w", "alaw", "-l", str(td / f"l{k}.scp"),
                                       "-m", str(td / f"o{k}.mlf")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                     for k in lists]
            for p in procs:
                p.wait()
            dt = time.perf_counter() - t0
        finally:
            shutil.rmtree(td, ignore_errors=True)
        return secs / dt, "reference", dt
    om = orc.Model(mdir)
    t0 = time.perf_counter()
    for u in range(n_utt):
        om.recognize(audio[u].tobytes(), fmt="alaw")
    dt = time.perf_counter() - t0
    return secs / dt, "port", dt


def host_synth_audio(n_utt: int, seed: int = 1) -> np.ndarray:
    """Host-generated A-law bytes for the CPU arm when no GPU generated them (speech-like byte statistics)."""
    rng = np.random.default_rng(seed)
    t = np.arange(BYTES_PER_UTT) / FS
    out = np.zeros((n_utt, BYTES_PER_UTT), np.uint8)
    for u in range(n_utt):
        f = rng.uniform(200, 3000, size=3)
        x = sum(np.sin(2 * np.pi * fi * t) for fi in f) * 2000 * (0.5 - 0.5 * np.cos(2 * np.pi * 4 * t)) + rng.normal(0, 300, t.size)
        x[int(rng.uniform(0, 8) * FS):][:FS] = 0
        pcm = np.clip(x, -32768, 32767).astype(np.int16) >> 3
        sign = np.where(pcm >= 0, 0xD5, 0x55)
        mag = np.where(pcm >= 0, pcm, -pcm - 1).astype(np.int32)
        seg = np.zeros_like(mag)
        for s in range(8):
            seg = np.where(mag > ((0x20 << s) - 1), s + 1, seg)
        seg = np.minimum(seg, 7)
        aval = (seg << 4) | np.where(seg < 2, (mag >> 1) & 0xF, (mag >> np.maximum(seg, 1)) & 0xF)
        out[u] = (aval ^ sign).astype(np.uint8)
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    mdir = model_dir()
    per_step = max(cores, 8) * args.ref_utts_per_core
    audio = host_synth_audio(min(per_step, 64))
    audio = np.concatenate([audio] * ((per_step + audio.shape[0] - 1) // audio.shape[0]))[:per_step]
    for _ in range(args.warmup):
        cpu_reference_run(audio[:cores], mdir, cores)
    t_tot, kind = 0.0, "reference"
    for _ in range(args.steps):
        v, kind, dt = cpu_reference_run(audio, mdir, cores)
        t_tot += dt
    value = per_step * UTT_SECONDS * args.steps / t_tot
    line = {"impl": "reference", "metric": "audio-sec/sec", "value": value, "unit": "xRT", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL}, 8 kHz alaw, {per_step} synthetic 10 s utterances per step (bounded sample)"},
            "cpu_baseline": {"value": value, "unit": "xRT", "cores": cores, "kind": kind,
                             "sample": f"{per_step} utterances x 10 s per step, {cores} phnrec processes"},
            "e2e": {"value": value, "unit": "xRT", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
_OUT = sys.stdout


def _one_line_stdout():
    """stdout carries exactly ONE line, the JSON result: everything else a library may print there (NCCL's version
    banner under torchrun, for instance) goes to stderr."""
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def main():
    _one_line_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "tc", "exact"], help="MLP mode: tcgen05 fp16 or exact fp32")
    ap.add_argument("--utts", type=int, default=1000, help="utterances per GPU per step")
    ap.add_argument("--ref-utts-per-core", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import phnrec_b200 as pb

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    mdir = model_dir()
    rec = pb.Recognizer(mdir, device=local)
    rec.set_wave_format("alaw")
    mode = args.mode
    if mode in ("auto", "tc"):
        try:
            rec.set_mlp_mode(pb.MLP_TC_F16)
            a = rec.synth_audio(BYTES_PER_UTT, 2, seed=5)
            rec.recognize([a[0].tobytes(), a[1].tobytes()])
            mode = "tc"
        except pb.PhnRecError as e:
            if args.mode == "tc":
                raise
            mode = "exact"
    if mode == "exact":
        rec.set_mlp_mode(pb.MLP_EXACT_FP32)

    n_utt = args.utts
    total_bytes = n_utt * BYTES_PER_UTT
    frames = n_utt * rec.num_frames(BYTES_PER_UTT)
    byte_off = (np.arange(n_utt + 1, dtype=np.int64) * BYTES_PER_UTT)
    d_audio = rec.device_alloc(total_bytes)
    rec.synth_audio_device(d_audio, BYTES_PER_UTT, n_utt, seed=1000 + rank)
    rec.sync()

    stream = torch.cuda.ExternalStream(rec._L.phn_stream(rec._h), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        rec.sync()
        barrier()
        evs = []
        for _ in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()                          # evict L2 between timed iterations (not timed)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                step_fn()
                e1.record(stream)
            evs.append((e0, e1))
        rec.sync()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident inputs, labels stay on the device
    def step_device():
        rec.recognize_device(d_audio, byte_off)

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed_loop(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    launches_per_step = sum(n for _, n in rec.last_timing().values())

    # ---- e2e: host buffers through phn_recognize (pinned audio -> H2D -> kernels -> labels D2H)
    import ctypes
    h_audio = rec._L.phn_host_alloc_pinned(total_bytes)
    rec.memcpy_d2h(h_audio, d_audio, total_bytes)
    cap = frames + 48 * n_utt
    labels = np.zeros(cap, dtype=pb.LABEL_DTYPE)
    loff = np.zeros(n_utt + 1, dtype=np.int64)
    nlab = [0]

    def step_host():
        nlab[0] = rec.recognize_raw(h_audio, byte_off, labels, loff)

    ms_e2e = timed_loop(step_host, args.steps, 2)
    d2h_bytes = nlab[0] * 16 + 4 * n_utt
    h2d_bytes = total_bytes + 3 * 8 * (n_utt + 1) + 4

    # ---- per-kernel-family device time (CUDA events on the launching stream, separate pass, not part of `value`)
    rec.set_profiling(True)
    fam = {}
    reps = 2
    for _ in range(reps):
        rec.recognize_device(d_audio, byte_off)
        rec.sync()
        for k, (ms, n) in rec.last_timing().items():
            a = fam.setdefault(k, [0.0, 0])
            a[0] += ms / reps
            a[1] = n
    rec.set_profiling(False)
    peaks = measured_peaks()
    mlp_ms = fam["mlp"][0]
    achieved_tflops = frames * FLOP_PER_FRAME / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
    fam_total = sum(v[0] for v in fam.values()) or 1.0

    audio_seconds = n_utt * UTT_SECONDS * world * args.steps
    value = audio_seconds / (ms_dev * 1e-3)
    e2e = audio_seconds / (ms_e2e * 1e-3)

    if rank == 0:
        line = {
            "metric": "audio-sec/sec", "value": value, "unit": "xRT", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if mode == "tc" else "f32", "data": "synthetic",
            "config": {"workload": f"{MODEL}, 8 kHz alaw, {n_utt} synthetic 10 s utterances per GPU per step "
                                   f"({frames} frames)", "mlp_mode": "tcgen05 fp16 operands, fp32 accumulate" if mode == "tc"
                                   else "exact fp32 (CUDA cores, reference summation order)",
                       "l2": "256 MB buffer written between timed iterations", "model_data": str(mdir.relative_to(ROOT)) if str(mdir).startswith(str(ROOT)) else "random-init CZ N1500 architecture",
                       "parallelism": f"{world} x independent utterance shards, no collective"},
            "e2e": {"value": e2e, "unit": "xRT", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "K-mlp (3 MLPs per frame)", "achieved": achieved_tflops, "peak": peaks["tflops"],
                         "unit": "TFLOP/s", "frac": achieved_tflops / peaks["tflops"], "traffic": measured_traffic(frames) if mode == "tc" else None,
                         "traffic_note": "DRAM bytes of the 3 MLP launches of one step (ncu --set full, profiles/r1_mlp_traffic.json); "
                                         "the bound is the tensor pipe, HBM traffic is ~10 % of peak",
                         "algorithmic": f"{FLOP_PER_FRAME} FLOP/frame x {frames} frames", "peak_source": peaks["source"],
                         "kernel_ms_per_step": mlp_ms, "launches_per_step": fam["mlp"][1]},
            "kernel_ms": {k: round(v[0], 4) for k, v in fam.items()},
            "kernel_share": {k: round(v[0] / fam_total, 4) for k, v in fam.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = max(cores, 8) * args.ref_utts_per_core
            n_gen = min(n_cpu, n_utt)
            sample = np.zeros((n_gen, BYTES_PER_UTT), dtype=np.uint8)
            rec.memcpy_d2h(sample.ctypes.data, d_audio, n_gen * BYTES_PER_UTT)
            sample = np.concatenate([sample] * ((n_cpu + n_gen - 1) // n_gen))[:n_cpu]
            v, kind, dt = cpu_reference_run(sample, mdir, cores)
            line["cpu_baseline"] = {"value": v, "unit": "xRT", "cores": cores, "kind": kind,
                                    "sample": f"{n_cpu} of the same synthetic utterances ({n_cpu * 10} s audio), {cores} single-threaded "
                                              f"phnrec processes, {dt:.1f} s wall"}
        print(json.dumps(line), file=_OUT, flush=True)

    rec._L.phn_host_free_pinned(h_audio)
    rec.device_free(d_audio)
    rec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
