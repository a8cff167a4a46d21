#!/usr/bin/env python3
"""bench.py — audio-seconds per second (xRT) of the PhnRec recognition hot path on B200.

Default workload = BASELINE.json configs[1], per GPU: PHN_CZ_SPDAT_LCRC_N1500, 8 kHz A-law, 1000 synthetic 10 s
utterances.  A step is one pass of the whole path (wave -> mel -> STC -> 3 MLPs -> Viterbi -> labels) over that batch.
N > 1: one process per GPU (torchrun), every rank runs its own 1000 utterances, no data-path collective (utterances are
independent) -> weak scaling.

  value : inputs resident in HBM when the timed region starts, labels left on the device
  e2e   : the reference-facing C-ABI call (phn_recognize / phn_decode) with HOST buffers (pinned audio in, label arrays
          out), H2D and D2H copies inside the timed region
  --config cz|hu|ru|en : the other shipped systems (BASELINE configs[2], [3]) on the same kind of batch
  --config cz_lin16    : configs[1]'s system fed 16-bit linear samples (the CLI's default wave format)
  --config en_sweep    : BASELINE configs[3], the 14-penalty decode from saved posteriors (srec.cpp:1080-1104 with -p)
  --impl reference     : the reference's own CPU implementation (oracle/_ref/phnrec_ref*, the reference sources compiled
                         by oracle/Makefile) on all host cores, bounded sample of the same utterances.
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

UTT_SECONDS = 10.0
FRAMES_PER_UTT = 998            # (80000 - 200) / 80 + 1 = (160000 - 400) / 160 + 1, srec.cpp:945
SWEEP_PENALTIES = [-6.0 + 0.5 * i for i in range(13)] + [None]   # None = the config's own penalty

CONFIGS = {  # --config -> workload
    "cz": dict(model="PHN_CZ_SPDAT_LCRC_N1500", fmt="alaw", fs=8000, bytes_per_utt=80000, baseline="configs[1]"),
    "cz_lin16": dict(model="PHN_CZ_SPDAT_LCRC_N1500", fmt="lin16", fs=8000, bytes_per_utt=160000, baseline="configs[1] with 16-bit linear input"),
    "hu": dict(model="PHN_HU_SPDAT_LCRC_N1500", fmt="alaw", fs=8000, bytes_per_utt=80000, baseline="configs[2]"),
    "ru": dict(model="PHN_RU_SPDAT_LCRC_N1500", fmt="alaw", fs=8000, bytes_per_utt=80000, baseline="configs[2]"),
    "en": dict(model="PHN_EN_TIMIT_LCRC_N500", fmt="lin16", fs=16000, bytes_per_utt=320000, baseline="configs[3]"),
    "en_sweep": dict(model="PHN_EN_TIMIT_LCRC_N500", fmt="lin16", fs=16000, bytes_per_utt=320000, baseline="configs[3]"),
}


def model_dir(name: str, allow_random: bool) -> Path:
    p = ROOT / "oracle" / "_ref" / "models" / name
    if (p / "weights" / "merger.nbin").exists():
        return p
    if not allow_random or name != "PHN_CZ_SPDAT_LCRC_N1500":
        raise SystemExit(f"bench.py: the staged model data {p} is missing (run `make -C oracle ref` where /root/reference exists); "
                         "refusing to measure on substitute weights (--allow-random-weights overrides, config cz only)")
    return synth_model_dir()


def synth_model_dir() -> Path:
    """Random-init weights of the CZ N1500 architecture (only with --allow-random-weights; the line then says so)."""
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


def nbin_dims(mdir: Path):
    """-> [(nin, nhid, nout)] x 3 from the .nbin headers (nn.cpp:464-531)."""
    out = []
    for n in ("band0", "band1", "merger"):
        h = np.fromfile(mdir / "weights" / f"{n}.nbin", dtype=np.int32, count=4)
        out.append((int(h[1]), int(h[2]), int(h[3])))
    return out


def workload_text(cfg_name: str, cfg: dict, n_utt: int) -> str:
    rate = "8 kHz" if cfg["fs"] == 8000 else "16 kHz"
    txt = f"{cfg['model']}, {rate} {cfg['fmt']}, {n_utt} synthetic 10 s utterances per GPU per step ({n_utt * FRAMES_PER_UTT} frames)"
    if cfg_name == "en_sweep":
        txt += f"; step = decoder under {len(SWEEP_PENALTIES)} insertion penalties from saved posteriors"
    return txt


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        # NVML in-process (the library nvidia-smi itself reads): a query costs microseconds.  Spawning nvidia-smi every
        # 100 ms stalls the driver for tens of milliseconds per call - invisible to per-kernel events, but inside a span
        # timed end to end.  nvidia-smi remains the fallback when the NVML binding is missing.
        h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            h = None
        while not self._stop_evt.is_set():
            try:
                if h is not None:
                    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
                    rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                    act = lambda bit: "Active" if rs & bit else "Not Active"
                    self.rows.append([str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4), f"{pw:.2f}"])
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05 if h is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)

        def num(r, i):
            try:
                return float(r[i])
            except Exception:
                return None
        sm = [v for v in (num(r, 0) for r in self.rows) if v is not None]
        mx = [v for v in (num(r, 1) for r in self.rows) if v is not None]
        pw = [v for v in (num(r, 6) for r in self.rows if len(r) > 6) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def measured_traffic(frames, mode, config="cz"):
    """DRAM bytes of the MLP launches of one step from the committed ncu --set full capture (same workload), or None."""
    for name in ("r2_mlp_traffic.json", "r1_mlp_traffic.json"):
        try:
            j = json.loads((ROOT / "profiles" / name).read_text())
            if int(j.get("frames", -1)) == int(frames) and mode == "tc" and config == j.get("config", "cz"):
                return int(j["k_mlp_per_step"]), name
        except Exception:
            pass
    return None, None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return {"tflops_sustained": float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1400.0))),
                "tflops_burst": float(j.get("bf16_tflops", j.get("bf16_tflops_sustained", 1590.0))),
                "hbm_gbs": float(j["hbm_gbs"]), "source": "MEASURED_PEAKS.json"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback of B200_PROFILING.md"}


# ------------------------------------------------------------------------------------------ reference CPU arm
def ref_binaries():
    """-> {"noblas": path | None, "blas": path | None}: the reference's own sources compiled by oracle/Makefile."""
    d = ROOT / "oracle" / "_ref"
    out = {}
    for k, n in (("noblas", "phnrec_ref"), ("blas", "phnrec_ref_blas")):
        p = d / n
        out[k] = p if p.exists() and os.access(p, os.X_OK) else None
    return out


def cpu_reference_run(audio: np.ndarray, mdir: Path, cores: int, fmt: str, binary):
    """The reference's own implementation on `cores` host cores: the list is split into `cores` slices, one phnrec process
    each (the reference is single threaded), wall clock first start -> last exit.  binary = a compiled reference
    (kind "reference") or None -> the C restatement (kind "port")."""
    n_utt = audio.shape[0]
    secs = n_utt * UTT_SECONDS
    if binary is not None:
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
            env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
            t0 = time.perf_counter()
            procs = [subprocess.Popen([str(binary), "-c", str(mdir), "-w", fmt, "-l", str(td / f"l{k}.scp"),
                                       "-m", str(td / f"o{k}.mlf")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
                     for k in lists]
            rcs = [p.wait() for p in procs]
            dt = time.perf_counter() - t0
            if any(rcs):
                raise RuntimeError(f"reference binary {binary} failed: {rcs}")
        finally:
            shutil.rmtree(td, ignore_errors=True)
        return secs / dt, "reference", dt
    from oracle import oracle as orc  # checker / baseline only
    om = orc.Model(mdir)
    t0 = time.perf_counter()
    for u in range(n_utt):
        om.recognize(audio[u].tobytes(), fmt=fmt)
    dt = time.perf_counter() - t0
    return secs / dt, "port", dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation, all host cores, on a bounded sample of the GPU arm's own
    utterances (tools/synth_host.py reproduces the device generator's bytes; the CUDA library is not loaded here)."""
    if rank != 0:
        return
    from tools.synth_host import synth_audio
    cfg = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    mdir = model_dir(cfg["model"], args.allow_random_weights)
    per_step = max(cores, 8) * args.ref_utts_per_core
    audio = synth_audio(cfg["bytes_per_utt"], min(per_step, args.utts), seed=1000, fmt=cfg["fmt"], fs=cfg["fs"])
    audio = np.concatenate([audio] * ((per_step + audio.shape[0] - 1) // audio.shape[0]))[:per_step]
    bins = ref_binaries()
    if args.config == "en_sweep":
        raise SystemExit("bench.py: --impl reference is defined for the audio -> labels configs (cz, hu, ru, en)")
    # The headline of this arm is the FASTER of the two flavours the reference offers, measured here on the full sample:
    # its BLAS build (north_star: "ATLAS-BLAS CPU build"; ATLAS is not installable offline, OpenBLAS from the image stands
    # in) or the plain-loops build.  The other flavour's single pass is reported beside it.
    probe = {}
    for fl in ("blas", "noblas"):
        if bins[fl] is not None:
            probe[fl] = cpu_reference_run(audio, mdir, cores, cfg["fmt"], bins[fl])
    flavour = max(probe, key=lambda k: probe[k][0]) if probe else "noblas"
    binary = bins.get(flavour)
    for _ in range(max(args.warmup - len(probe), 0)):
        cpu_reference_run(audio[:cores], mdir, cores, cfg["fmt"], binary)
    t_tot, kind = 0.0, "reference"
    for _ in range(args.steps):
        v, kind, dt = cpu_reference_run(audio, mdir, cores, cfg["fmt"], binary)
        t_tot += dt
    value = per_step * UTT_SECONDS * args.steps / t_tot
    other = None
    names = {"blas": "USE_BLAS build, sgemv per frame (nn.cpp:760) on OpenBLAS 0.3.15, 1 thread per process: substitute for ATLAS, "
                     "which is not installable offline", "noblas": "no-BLAS loops (nn.cpp:771-793)"}
    for fl, (v2, _, dt2) in probe.items():
        if fl != flavour:
            other = {"value": v2, "unit": "xRT", "cores": cores, "kind": "reference", "flavour": names[fl],
                     "sample": f"{per_step} utterances x 10 s, one pass, {dt2:.1f} s wall"}
    flav_txt = names[flavour]
    line = {"impl": "reference", "metric": "audio-sec/sec", "value": value, "unit": "xRT", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generator and seed as the GPU arm; tools/synth_host.py)",
            "config": {"workload": workload_text(args.config, cfg, args.utts)},
            "cpu_baseline": {"value": value, "unit": "xRT", "cores": cores, "kind": kind, "flavour": flav_txt,
                             "sample": f"the first {min(per_step, args.utts)} utterances of that workload ({per_step} x 10 s per step), "
                                       f"{cores} single-threaded phnrec processes over disjoint list slices"},
            "e2e": {"value": value, "unit": "xRT", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if other:
        line["cpu_baseline_other_flavour"] = other
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
    ap.add_argument("--config", default="cz", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="tc", choices=["tc", "exact"], help="MLP mode: tcgen05 fp16 or exact fp32")
    ap.add_argument("--utts", type=int, default=1000, help="utterances per GPU per step")
    ap.add_argument("--ref-utts-per-core", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--profile-seconds", type=float, default=2.0, help="length of the per-kernel profiling pass")
    ap.add_argument("--allow-random-weights", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    sweep = args.config == "en_sweep"

    import torch
    import torch.distributed as dist
    import phnrec_b200 as pb

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    mdir = model_dir(cfg["model"], args.allow_random_weights)
    rec = pb.Recognizer(mdir, device=local)
    rec.set_wave_format(cfg["fmt"])
    mode = args.mode
    rec.set_mlp_mode(pb.MLP_TC_F16 if mode == "tc" else pb.MLP_EXACT_FP32)   # (a missing kernel image raises: no fallback)

    n_utt = args.utts
    bpu = cfg["bytes_per_utt"]
    total_bytes = n_utt * bpu
    fpu = rec.num_frames(bpu)
    frames = n_utt * fpu
    byte_off = (np.arange(n_utt + 1, dtype=np.int64) * bpu)
    frame_off = (np.arange(n_utt + 1, dtype=np.int64) * fpu)
    d_audio = rec.device_alloc(total_bytes)
    rec.synth_audio_device(d_audio, bpu, n_utt, seed=1000 + rank)
    rec.sync()

    stream = torch.cuda.ExternalStream(rec._L.phn_stream(rec._h), device=dev)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(step_fn, steps, warmup, drain_fn=None):
        """One CUDA-event pair around the WHOLE run of `steps` steps (L2 flushes included in the timed span): the decoder of
        step k runs on the context's decoder stream under the front end of step k+1, so per-step event pairs on the main
        stream would miss it.  rec.sync() joins both streams before the closing event is recorded."""
        for _ in range(warmup):
            with torch.cuda.stream(stream):
                flush.zero_()                          # (also loads torch's fill kernel before the timed span: lazy module
            step_fn()                                  #  loading inside it costs hundreds of milliseconds)
        if drain_fn:
            drain_fn()
        rec.sync()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                flush.zero_()                          # evict L2 between iterations (inside the timed span)
                step_fn()
            if drain_fn:
                drain_fn()
            rec.sync()
            e1.record(stream)
        rec.sync()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import ctypes  # noqa: F401
    h_audio = rec._L.phn_host_alloc_pinned(total_bytes)
    rec.memcpy_d2h(h_audio, d_audio, total_bytes)
    pens = np.array([rec.wpenalty if p is None else p for p in SWEEP_PENALTIES], dtype=np.float32)
    n_pass = len(pens) if sweep else 1

    if sweep:
        # posteriors once (tensor-core or exact nets), resident in HBM; the step is the decoder under 14 penalties
        audio_np = np.ctypeslib.as_array((ctypes.c_uint8 * total_bytes).from_address(h_audio))
        utts = [audio_np[i * bpu:(i + 1) * bpu] for i in range(n_utt)]
        posts = np.concatenate(rec.posteriors(rec.mel(utts)))   # [frames, n_outputs]: host copy for the e2e leg; phn_posteriors
                                                                # also leaves them resident in HBM for phn_decode_device

        def step_device():
            rec.decode_device(pens)
    else:
        def step_device():
            rec.recognize_device(d_audio, byte_off)

    sampler = ClockSampler(local)
    if not os.environ.get("PHNREC_BENCH_NO_SAMPLER"):    # (diagnostic: how much the sampling itself costs)
        sampler.start()
    ms_dev = timed_loop(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler.is_alive() or sampler.rows else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "power_w_max": None}
    launches_per_step = sum(n for _, n in rec.last_timing().values())

    # ---- e2e: host buffers through the C ABI (pinned audio -> H2D -> kernels -> labels D2H)
    cap = (frames + 48 * n_utt) * n_pass
    labels = np.zeros(cap, dtype=pb.LABEL_DTYPE)
    loff = np.zeros(n_utt * n_pass + 1, dtype=np.int64)
    nlab = [0]
    if sweep:
        def step_host():
            rec._ck(rec._L.phn_decode(rec._h, posts.reshape(-1), frame_off, n_utt, pens.ctypes.data, len(pens), labels.ctypes.data, cap, loff))
            nlab[0] = int(loff[-1])
        h2d_bytes = posts.nbytes + 2 * 8 * (n_utt + 1) + 4 * len(pens)
        drain_host = None
    else:
        # the call a list-mode user makes (the CLI does the same): phn_recognize_async for batch k+1 while batch k is on the
        # device, phn_wait for the labels of batch k - every step copies its audio up and its labels down inside the span
        def step_host():
            rec.recognize_async_raw(h_audio, byte_off)
            if rec.pending() == 2:
                nlab[0] = rec.wait_raw(labels, loff)

        def drain_host():
            while rec.pending():
                nlab[0] = rec.wait_raw(labels, loff)
        h2d_bytes = total_bytes + 3 * 8 * (n_utt + 1) + 4
    ms_e2e = timed_loop(step_host, args.steps, 3, drain_host)
    d2h_bytes = nlab[0] * 16 + 4 * n_utt * n_pass

    # ---- per-kernel-family device time: CUDA events on the launching stream, a separate pass of >= profile-seconds so that
    # the GPU is in its sustained state (clocks under the power cap), not part of `value`
    rec.set_profiling(True)
    fam = {}
    reps = max(2, int(np.ceil(args.profile_seconds * 1000.0 / max(ms_dev / args.steps, 1e-3))))
    psampler = ClockSampler(local)
    psampler.start()
    t_prof0 = time.perf_counter()
    for _ in range(reps):
        step_device()
        rec.sync()
        for k, (ms, n) in rec.last_timing().items():
            a = fam.setdefault(k, [0.0, 0])
            a[0] += ms / reps
            a[1] = n
    t_prof = time.perf_counter() - t_prof0
    pclocks = psampler.stop()
    rec.set_profiling(False)
    peaks = measured_peaks()

    dims = nbin_dims(mdir)
    flop_per_frame = 2 * sum(i * h + h * o for i, h, o in dims)          # 3 MLPs, from the .nbin header sizes
    bps = 2 if cfg["fmt"] == "lin16" else 1
    step_samples = 80 if cfg["fs"] == 8000 else 160
    nb, nin, P3 = rec.nbanks, dims[0][0], 3 * rec.n_phonemes
    alg = {  # algorithmic work per frame: bytes for the streaming kernels, FLOP for the nets
        "wave": ("hbm", step_samples * bps + 4 * nb, "audio in + log-mel out"),
        "mean": ("hbm", 4 * nb, "log-mel in"),
        "stc": ("hbm", 4 * nb + (2 * nin * 2 if mode == "tc" else 2 * nin * 4), "log-mel in + 2 x band-net inputs out (fp16 | fp32)"),
        "mlp": ("tensor" if mode == "tc" else "fp32", flop_per_frame, "3 MLPs"),
        "vit": ("hbm", 4 * P3 * n_pass, "ln p in (per penalty pass)"),
    }
    kernels = []
    for k, (ms, nl) in fam.items():
        if ms <= 0:
            continue
        bound, per_frame, what = alg[k]
        if bound == "hbm":
            ach = frames * per_frame / (ms * 1e-3) / 1e9
            kernels.append({"kernel": f"K-{k}", "bound": "hbm", "algorithmic_per_frame": per_frame, "what": what, "ms": round(ms, 4),
                            "launches": nl, "achieved": round(ach, 1), "unit": "GB/s", "peak": peaks["hbm_gbs"], "frac": round(ach / peaks["hbm_gbs"], 4)})
            if k == "wave" and mode == "tc" and cfg["fs"] == 16000 and os.environ.get("PHNREC_WAVE_TC", "1") != "0":
                # k_wave_tc16.cu: 2 bin passes x 2 window halves x (fp16(sample) against W_hi, W_lo; the rounding error against W_hi), 208 x 256 each
                gf = 2 * 208 * 256 * 2 * 2 * 3
                kernels[-1]["tensor_work"] = {"flop_per_frame": gf, "tflops": round(frames * gf / (ms * 1e-3) / 1e12, 1),
                                              "of_sustained_peak": round(frames * gf / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"], 3),
                                              "note": "DFT as a tcgen05 GEMM; what the kernel executes, not algorithmic work"}
            if k == "wave" and mode == "tc" and cfg["fs"] == 8000 and os.environ.get("PHNREC_WAVE_TC", "1") != "0":
                # k_wave_tc.cu: the windowed DFT runs as a GEMM on the tensor cores (208 x 256 per frame, matrix as fp16 hi + lo); the
                # algorithmic figure above stays audio in + mel out - the GEMM's FLOPs are the kernel's own choice, reported for what they are
                gf = 2 * 208 * 256 * (2 if cfg["fmt"] == "alaw" else 3)   # (lin16: fp16(sample) against W_hi, W_lo; the rounding error against W_hi)
                kernels[-1]["tensor_work"] = {"flop_per_frame": gf, "tflops": round(frames * gf / (ms * 1e-3) / 1e12, 1),
                                              "of_sustained_peak": round(frames * gf / (ms * 1e-3) / 1e12 / peaks["tflops_sustained"], 3),
                                              "note": "DFT as a tcgen05 GEMM; what the kernel executes, not algorithmic work"}
        else:
            ach = frames * per_frame / (ms * 1e-3) / 1e12
            pk = peaks["tflops_sustained"] if bound == "tensor" else 74.4
            kernels.append({"kernel": f"K-{k}", "bound": bound, "algorithmic_per_frame": per_frame, "what": what, "ms": round(ms, 4),
                            "launches": nl, "achieved": round(ach, 1), "unit": "TFLOP/s", "peak": pk, "frac": round(ach / pk, 4)})
    mlp_ms = fam.get("mlp", [0.0, 0])[0]
    fam_total = sum(v[0] for v in fam.values()) or 1.0

    audio_seconds = n_utt * UTT_SECONDS * world * args.steps * n_pass
    value = audio_seconds / (ms_dev * 1e-3)
    e2e = audio_seconds / (ms_e2e * 1e-3)

    if rank == 0:
        workload = workload_text(args.config, cfg, n_utt)
        line = {
            "metric": "audio-sec/sec", "value": value, "unit": "xRT", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16" if mode == "tc" else "f32", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": cfg["baseline"],
                       "mlp_mode": "tcgen05 fp16 operands, fp32 accumulate" if mode == "tc" else "exact fp32 (CUDA cores, reference summation order)",
                       "l2": "192 MB buffer (> 126 MB L2) written between iterations, inside the timed span; a step itself moves ~2 GB",
                       "timing": "one CUDA-event pair around all timed steps (the decoder of step k overlaps the front end of step k+1 on a second stream); both streams joined before the closing event",
                       "model_data": str(mdir.relative_to(ROOT)) if str(mdir).startswith(str(ROOT)) else "RANDOM-INIT CZ N1500 architecture (--allow-random-weights)",
                       "parallelism": f"{world} x independent utterance shards, no collective"},
            "e2e": {"value": e2e, "unit": "xRT", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "phn_decode (host posteriors in, labels out)" if sweep else
                           "phn_recognize_async + phn_wait, two batches in flight (pinned host audio in, host labels out, every step)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        if sweep:
            line["config"]["value_counts"] = f"audio seconds x {n_pass} decoder passes"
        if mlp_ms > 0:
            ach = frames * flop_per_frame / (mlp_ms * 1e-3) / 1e12
            traffic, tfile = measured_traffic(frames, mode, args.config)
            pk = peaks["tflops_sustained"] if mode == "tc" else 74.4
            line["roofline"] = {
                "bound": "tensor" if mode == "tc" else "fp32", "kernel": "K-mlp (3 MLPs per frame)", "achieved": ach, "peak": pk,
                "unit": "TFLOP/s", "frac": ach / pk, "traffic": traffic,
                "frac_burst": ach / peaks["tflops_burst"] if mode == "tc" else None,
                "peak_burst": peaks["tflops_burst"] if mode == "tc" else None,
                "peak_applies": f"sustained: kernel times are CUDA-event averages over a profiling pass of {reps} steps spanning {t_prof:.1f} s "
                                f"(SM clock median {pclocks['sm_mhz']} MHz, reasons {pclocks['reasons']}); frac_burst = the same achieved rate "
                                "over the burst peak, for comparison with a kernel timed alone",
                "traffic_note": (f"DRAM bytes of the MLP launches of one step (ncu --set full, profiles/{tfile})" if tfile else None),
                "algorithmic": f"{flop_per_frame} FLOP/frame x {frames} frames", "peak_source": peaks["source"],
                "kernel_ms_per_step": mlp_ms, "launches_per_step": fam["mlp"][1],
                "kernels": kernels}
        else:   # decoder-only step (en_sweep): the dominant kernel is K-vit, latency bound; reported against HBM like the others
            kv = next((k for k in kernels if k["kernel"] == "K-vit"), None)
            line["roofline"] = {"bound": "hbm", "kernel": "K-vit (K-log + token passing, per penalty pass)", "achieved": kv["achieved"] if kv else None,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kv["frac"] if kv else None, "traffic": None,
                                "peak_source": peaks["source"], "kernels": kernels}
        line["kernel_ms"] = {k: round(v[0], 4) for k, v in fam.items()}
        line["kernel_share"] = {k: round(v[0] / fam_total, 4) for k, v in fam.items()}
        if mode == "tc" and not sweep and not args.no_parity:
            # the precision of the benchmarked mode on the benchmarked audio, outside the timed region: the first 256
            # utterances through both modes (tools/tc_bound.py has the full-set figures under profiles/)
            from tools.tc_bound import measure
            n_par = min(256, n_utt)
            audio_np = np.ctypeslib.as_array((ctypes.c_uint8 * (n_par * bpu)).from_address(h_audio))
            pr = measure(rec, pb, [audio_np[i * bpu:(i + 1) * bpu].tobytes() for i in range(n_par)], chunk=128)
            rec.set_mlp_mode(pb.MLP_TC_F16)
            line["parity"] = {"vs": "exact fp32 mode of this library (bit-identical to the reference binary)", "utterances": n_par,
                              "seg_agree": round(pr["seg_agree"], 5), "utt_identical": round(pr["utt_identical"], 4),
                              "frame_argmax_agree": round(pr["frame_argmax_agree"], 5),
                              "logp_rel_max": pr["rel_logp_max"], "logp_rel_p999": pr["rel_logp_p999"], "logp_rel_mean": pr["rel_logp_mean"],
                              "measure": "|ln p_tc - ln p_exact| / max(1, |ln p_exact|) on the 3P decoder columns, every frame"}
        if world == 1 and not args.no_cpu_baseline and not sweep:
            from tools.synth_host import synth_audio
            cores = os.cpu_count() or 1
            n_cpu = max(cores, 8) * args.ref_utts_per_core
            n_gen = min(n_cpu, n_utt)
            sample = np.zeros((n_gen, bpu), dtype=np.uint8)
            rec.memcpy_d2h(sample.ctypes.data, d_audio, n_gen * bpu)
            host = synth_audio(bpu, min(n_gen, 4), seed=1000 + rank, fmt=cfg["fmt"], fs=cfg["fs"])
            same = bool(np.array_equal(host, sample[:host.shape[0]]))
            sample = np.concatenate([sample] * ((n_cpu + n_gen - 1) // n_gen))[:n_cpu]
            bins = ref_binaries()
            v, kind, dt = cpu_reference_run(sample, mdir, cores, cfg["fmt"], bins["noblas"])
            line["cpu_baseline"] = {"value": v, "unit": "xRT", "cores": cores, "kind": kind, "flavour": "no-BLAS loops (nn.cpp:771-793)",
                                    "sample": f"{n_cpu} of the same synthetic utterances ({n_cpu * 10} s audio), {cores} single-threaded "
                                              f"phnrec processes, {dt:.1f} s wall; host generator reproduces the device bytes: {same}"}
            if bins["blas"] is not None:
                v2, kind2, dt2 = cpu_reference_run(sample, mdir, cores, cfg["fmt"], bins["blas"])
                line["cpu_baseline_blas"] = {"value": v2, "unit": "xRT", "cores": cores, "kind": kind2,
                                             "flavour": "USE_BLAS build, sgemv per frame (nn.cpp:760) on OpenBLAS 0.3.15, 1 thread per process: "
                                                        "substitute for ATLAS (not installable offline)",
                                             "sample": f"same {n_cpu} utterances, {dt2:.1f} s wall"}
        print(json.dumps(line), file=_OUT, flush=True)

    rec._L.phn_host_free_pinned(h_audio)
    rec.device_free(d_audio)
    rec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
