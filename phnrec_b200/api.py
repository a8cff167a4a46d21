"""ctypes binding of include/phnrec_b200.h.

`Recognizer` mirrors the reference's SpeechRec object for the offline hot path
(srec.h:161-199): Init(config dir) -> ProcessOffline stages, with SetWPenalty /
SetWaveFormat knobs.  Every method is a direct call into the CUDA library; when
the library is missing or no B200 is visible the call raises — there is no
Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
WAVE_LIN16, WAVE_ALAW = 0, 1
MLP_EXACT_FP32, MLP_TC_F16 = 0, 1
K_FAMILIES = ("wave", "mean", "stc", "mlp", "vit")
LABEL_DTYPE = np.dtype([("phn", np.int32), ("start", np.int32), ("end", np.int32), ("like", np.float32)])


class PhnRecError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[phnrec_b200 error {code}] {msg.strip()}")
        self.code = code
        self.message = msg


class _Info(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("sample_freq", "wave_format", "nbanks", "vector_size", "vector_step", "fft_size", "n_phonemes",
                 "n_states", "n_outputs", "band_inputs", "merger_inputs", "hidden", "sent_mean_norm", "time_pruning",
                 "mlp_mode", "device")] + [("wpenalty", C.c_float), ("n_params", C.c_int32)]


def lib_path() -> Path:
    # PHNREC_B200_LIB: kernel-development aid (A/B runs of two builds on the same GPU box)
    return Path(os.environ["PHNREC_B200_LIB"]) if os.environ.get("PHNREC_B200_LIB") else PKG / "lib" / "libphnrec_b200.so"


def build(verbose: bool = False) -> None:
    """Compile the CUDA library + CLI for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(PKG / "csrc"), "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("phnrec_b200: build failed")


_lib = None
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")

# every symbol include/phnrec_b200.h declares: (name, restype, argtypes)
_SYMS = [
    ("phn_create", C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    ("phn_destroy", None, [C.c_void_p]),
    ("phn_last_error", C.c_char_p, [C.c_void_p]),
    ("phn_get_info", C.c_int, [C.c_void_p, C.POINTER(_Info)]),
    ("phn_phoneme", C.c_char_p, [C.c_void_p, C.c_int]),
    ("phn_config_get", C.c_char_p, [C.c_void_p, C.c_char_p, C.c_char_p]),
    ("phn_set_penalty", C.c_int, [C.c_void_p, C.c_float]),
    ("phn_set_wave_format", C.c_int, [C.c_void_p, C.c_int]),
    ("phn_set_mlp_mode", C.c_int, [C.c_void_p, C.c_int]),
    ("phn_num_frames", C.c_int64, [C.c_void_p, C.c_int64]),
    ("phn_label_capacity", C.c_int64, [C.c_void_p, _i64p, C.c_int]),
    ("phn_mel", C.c_int, [C.c_void_p, C.c_void_p, _i64p, C.c_int, C.c_void_p, _i64p]),
    ("phn_posteriors", C.c_int, [C.c_void_p, _f32p, _i64p, C.c_int, _f32p]),
    ("phn_decode", C.c_int, [C.c_void_p, _f32p, _i64p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, _i64p]),
    ("phn_recognize", C.c_int, [C.c_void_p, C.c_void_p, _i64p, C.c_int, C.c_void_p, C.c_int64, _i64p, C.c_void_p]),
    ("phn_recognize_async", C.c_int, [C.c_void_p, C.c_void_p, _i64p, C.c_int]),
    ("phn_wait", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _i64p, C.c_void_p]),
    ("phn_pending", C.c_int, [C.c_void_p]),
    ("phn_stream_open", C.c_int, [C.c_void_p, C.c_int]),
    ("phn_stream_count", C.c_int, [C.c_void_p]),
    ("phn_stream_reset", C.c_int, [C.c_void_p, C.c_int]),
    ("phn_stream_push", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, _i64p, C.c_void_p, C.c_void_p, C.c_int64, _i64p]),
    ("phn_recognize_device", C.c_int, [C.c_void_p, C.c_void_p, _i64p, C.c_int]),
    ("phn_decode_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    ("phn_sync", C.c_int, [C.c_void_p]),
    ("phn_fetch_labels", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("phn_convert_weights", C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p]),
    ("phn_fetch_mel", C.c_int, [C.c_void_p, _f32p]),
    ("phn_fetch_posteriors", C.c_int, [C.c_void_p, _f32p]),
    ("phn_fetch_logp", C.c_int, [C.c_void_p, _f32p]),
    ("phn_stream", C.c_void_p, [C.c_void_p]),
    ("phn_device_alloc", C.c_void_p, [C.c_void_p, C.c_int64]),
    ("phn_device_free", None, [C.c_void_p, C.c_void_p]),
    ("phn_host_alloc_pinned", C.c_void_p, [C.c_int64]),
    ("phn_host_free_pinned", None, [C.c_void_p]),
    ("phn_memcpy_h2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("phn_memcpy_d2h", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("phn_synth_audio_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64]),
    ("phn_set_profiling", C.c_int, [C.c_void_p, C.c_int]),
    ("phn_last_timing", C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    ("phn_online_norm", C.c_int, [C.c_void_p, _f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("phn_debug_tc_timeline", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("phn_debug_logf", C.c_int, [C.c_void_p, C.c_uint32, C.c_int64, _f32p]),
    ("phn_version", C.c_char_p, []),
    ("phn_device_count", C.c_int, []),
]


def load_library() -> C.CDLL:
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not p.exists():
            raise PhnRecError(40, f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                                  "phnrec_b200 has no CPU path")
        L = C.CDLL(str(p))
        for name, res, args in _SYMS:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _offsets(lengths) -> np.ndarray:
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lengths, dtype=np.int64), out=off[1:])
    return off


class Recognizer:
    """One context per GPU (SpeechRec analogue).  Not thread-safe; contexts are independent."""

    def __init__(self, cfg_dir, device: int = 0, mlp_mode: int = MLP_EXACT_FP32):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.phn_create(str(cfg_dir).encode(), device, C.byref(h))
        if rc:
            raise PhnRecError(rc, (self._L.phn_last_error(None) or b"").decode())
        self._h = h
        info = _Info()
        self._L.phn_get_info(h, C.byref(info))
        for n, _ in _Info._fields_:
            setattr(self, n, getattr(info, n))
        self.phonemes = [self._L.phn_phoneme(h, i).decode() for i in range(self.n_phonemes)]
        self.set_mlp_mode(mlp_mode)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._L.phn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc:
            raise PhnRecError(rc, (self._L.phn_last_error(self._h) or b"").decode())

    def config(self, section: str, variable: str):
        v = self._L.phn_config_get(self._h, section.encode(), variable.encode())
        return None if v is None else v.decode()

    # -- knobs (Decoder::SetWPenalty, SpeechRec::SetWaveFormat)
    def set_penalty(self, wp: float):
        self._ck(self._L.phn_set_penalty(self._h, float(wp)))
        self.wpenalty = float(wp)

    def set_wave_format(self, fmt):
        f = {"lin16": WAVE_LIN16, "alaw": WAVE_ALAW}.get(fmt, fmt)
        self._ck(self._L.phn_set_wave_format(self._h, int(f)))
        self.wave_format = int(f)

    def set_mlp_mode(self, mode: int):
        self._ck(self._L.phn_set_mlp_mode(self._h, int(mode)))
        self.mlp_mode = int(mode)

    def set_profiling(self, on: bool):
        self._ck(self._L.phn_set_profiling(self._h, int(on)))

    def last_timing(self):
        ms = (C.c_float * len(K_FAMILIES))()
        n = (C.c_int64 * len(K_FAMILIES))()
        self._ck(self._L.phn_last_timing(self._h, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(K_FAMILIES)}

    def num_frames(self, nbytes: int) -> int:
        return int(self._L.phn_num_frames(self._h, int(nbytes)))

    # -- helpers
    @staticmethod
    def _concat_audio(utts):
        bufs = [np.frombuffer(u, dtype=np.uint8) if isinstance(u, (bytes, bytearray, memoryview))
                else np.ascontiguousarray(u).view(np.uint8).reshape(-1) for u in utts]
        off = _offsets([b.size for b in bufs])
        audio = np.concatenate(bufs) if bufs else np.zeros(0, dtype=np.uint8)
        return np.ascontiguousarray(audio), off

    def _split_labels(self, labels, off):
        return [labels[off[i]:off[i + 1]].copy() for i in range(len(off) - 1)]

    # -- stages on host buffers
    def mel(self, utts):
        """list of audio byte strings -> list of un-normalised log-mel [T_u, nbanks] (what -t par saves)."""
        audio, boff = self._concat_audio(utts)
        n = len(utts)
        foff = np.zeros(n + 1, dtype=np.int64)
        self._ck(self._L.phn_mel(self._h, audio.ctypes.data, boff, n, None, foff))
        out = np.zeros((int(foff[-1]), self.n_params), dtype=np.float32)   # nbanks, or the PLP coefficients (params/kind = plp)
        self._ck(self._L.phn_mel(self._h, audio.ctypes.data, boff, n, out.ctypes.data, foff))
        return [out[foff[i]:foff[i + 1]] for i in range(n)]

    def posteriors(self, mels):
        """list of mel [T_u, nbanks] -> list of linear posteriors [T_u, n_outputs] (what -t post saves)."""
        mels = [np.ascontiguousarray(m, dtype=np.float32) for m in mels]
        foff = _offsets([m.shape[0] for m in mels])
        mel = np.ascontiguousarray(np.concatenate(mels, axis=0)) if mels else np.zeros((0, self.nbanks), np.float32)
        out = np.zeros((int(foff[-1]), self.n_outputs), dtype=np.float32)
        self._ck(self._L.phn_posteriors(self._h, mel.reshape(-1), foff, len(mels), out.reshape(-1)))
        return [out[foff[i]:foff[i + 1]] for i in range(len(mels))]

    def decode(self, posts, penalties=None):
        """list of posteriors -> list (per penalty: list) of label arrays.  penalties=None: context penalty."""
        posts = [np.ascontiguousarray(p, dtype=np.float32) for p in posts]
        n = len(posts)
        foff = _offsets([p.shape[0] for p in posts])
        post = np.ascontiguousarray(np.concatenate(posts, axis=0)) if posts else np.zeros((0, self.n_outputs), np.float32)
        pen = None if penalties is None else np.ascontiguousarray(penalties, dtype=np.float32)
        npen = 1 if pen is None else int(pen.size)
        cap = int(self._L.phn_label_capacity(self._h, foff, n)) * npen
        labels = np.zeros(cap, dtype=LABEL_DTYPE)
        loff = np.zeros(n * npen + 1, dtype=np.int64)
        self._ck(self._L.phn_decode(self._h, post.reshape(-1), foff, n, None if pen is None else pen.ctypes.data, npen,
                                    labels.ctypes.data, cap, loff))
        per = self._split_labels(labels, loff)
        return per if pen is None else [per[k * n:(k + 1) * n] for k in range(npen)]

    def recognize(self, utts):
        """list of audio byte strings -> list of label arrays (audio -> labels on the GPU, one call)."""
        audio, boff = self._concat_audio(utts)
        n = len(utts)
        foff = np.zeros(n + 1, dtype=np.int64)
        self._ck(self._L.phn_mel(self._h, audio.ctypes.data, boff, n, None, foff))
        cap = int(self._L.phn_label_capacity(self._h, foff, n))
        labels = np.zeros(cap, dtype=LABEL_DTYPE)
        loff = np.zeros(n + 1, dtype=np.int64)
        self._ck(self._L.phn_recognize(self._h, audio.ctypes.data, boff, n, labels.ctypes.data, cap, loff, None))
        return self._split_labels(labels, loff)

    # -- raw entry points for bench.py (pre-concatenated buffers, no Python-side copies)
    def recognize_raw(self, audio_ptr: int, byte_off: np.ndarray, labels: np.ndarray, label_off: np.ndarray) -> int:
        self._ck(self._L.phn_recognize(self._h, audio_ptr, byte_off, len(byte_off) - 1, labels.ctypes.data, labels.size,
                                       label_off, None))
        return int(label_off[-1])

    def recognize_async_raw(self, audio_ptr: int, byte_off: np.ndarray):
        """Enqueue a batch (phn_recognize_async); the audio must stay valid until the matching wait_raw()."""
        self._ck(self._L.phn_recognize_async(self._h, audio_ptr, byte_off, len(byte_off) - 1))

    def wait_raw(self, labels: np.ndarray, label_off: np.ndarray) -> int:
        """Labels of the oldest batch in flight (phn_wait)."""
        self._ck(self._L.phn_wait(self._h, labels.ctypes.data, labels.size, label_off, None))
        return int(label_off[-1])

    def pending(self) -> int:
        return int(self._L.phn_pending(self._h))

    def recognize_pipelined(self, batches):
        """list of batches (each a list of audio byte strings) -> list of lists of label arrays, two batches in flight."""
        out, keep = [], []
        for utts in batches:
            audio, boff = self._concat_audio(utts)
            n = len(utts)
            foff = np.zeros(n + 1, dtype=np.int64)
            self._ck(self._L.phn_mel(self._h, audio.ctypes.data, boff, n, None, foff))
            keep.append((audio, boff, n, int(self._L.phn_label_capacity(self._h, foff, n))))
        done = 0
        for k, (audio, boff, n, cap) in enumerate(keep):
            self.recognize_async_raw(audio.ctypes.data, boff)
            if self.pending() == 2 or k == len(keep) - 1:
                while self.pending() and (self.pending() == 2 or k == len(keep) - 1):
                    _, _, nd, capd = keep[done]
                    labels = np.zeros(capd, dtype=LABEL_DTYPE)
                    loff = np.zeros(nd + 1, dtype=np.int64)
                    self.wait_raw(labels, loff)
                    out.append(self._split_labels(labels, loff))
                    done += 1
        return out

    # -- streaming (online) mode: SpeechRec::ProcessOnline for many concurrent streams
    def stream_open(self, n_streams: int):
        self._ck(self._L.phn_stream_open(self._h, int(n_streams)))

    def stream_reset(self, sid: int):
        self._ck(self._L.phn_stream_reset(self._h, int(sid)))

    def stream_push(self, sids, blocks, last=None):
        """One block of audio bytes per stream in `sids`; returns the labels that became final, one array per pushed stream."""
        n = len(sids)
        audio, boff = self._concat_audio(blocks)
        sid = np.ascontiguousarray(sids, dtype=np.int32)
        lst = None if last is None else np.ascontiguousarray([1 if x else 0 for x in last], dtype=np.int32)
        cap = int(sum(len(b) for b in blocks) // 20 + 64 * n + 64)
        labels = np.zeros(cap, dtype=LABEL_DTYPE)
        loff = np.zeros(n + 1, dtype=np.int64)
        self._ck(self._L.phn_stream_push(self._h, sid.ctypes.data, n, audio.ctypes.data, boff, None if lst is None else lst.ctypes.data,
                                         labels.ctypes.data, cap, loff))
        return self._split_labels(labels, loff)

    def recognize_device(self, d_audio: int, byte_off: np.ndarray):
        self._ck(self._L.phn_recognize_device(self._h, d_audio, byte_off, len(byte_off) - 1))

    def decode_device(self, penalties=None):
        """Penalty sweep from the posteriors resident in the context (after posteriors()); fetch with fetch_labels()."""
        pen = None if penalties is None else np.ascontiguousarray(penalties, dtype=np.float32)
        self._pen_keepalive = pen
        self._ck(self._L.phn_decode_device(self._h, None if pen is None else pen.ctypes.data, 1 if pen is None else int(pen.size)))

    def sync(self):
        self._ck(self._L.phn_sync(self._h))

    def fetch_labels(self, n_seg: int, cap: int):
        labels = np.zeros(cap, dtype=LABEL_DTYPE)
        loff = np.zeros(n_seg + 1, dtype=np.int64)
        self._ck(self._L.phn_fetch_labels(self._h, labels.ctypes.data, cap, loff.ctypes.data))
        return self._split_labels(labels, loff)

    def fetch_mel(self, total_frames: int):
        out = np.zeros((total_frames, self.n_params), dtype=np.float32)
        self._ck(self._L.phn_fetch_mel(self._h, out.reshape(-1)))
        return out

    def fetch_posteriors(self, total_frames: int):
        out = np.zeros((total_frames, self.n_outputs), dtype=np.float32)
        self._ck(self._L.phn_fetch_posteriors(self._h, out.reshape(-1)))
        return out

    def fetch_logp(self, total_frames: int):
        """ln p as the decoder of the last call consumed it: [total_frames, 3 * n_phonemes]."""
        out = np.zeros((total_frames, 3 * self.n_phonemes), dtype=np.float32)
        self._ck(self._L.phn_fetch_logp(self._h, out.reshape(-1)))
        return out

    def device_alloc(self, nbytes: int) -> int:
        p = self._L.phn_device_alloc(self._h, int(nbytes))
        if not p:
            raise PhnRecError(33, f"device allocation of {nbytes} bytes failed")
        return int(p)

    def device_free(self, p: int):
        self._L.phn_device_free(self._h, C.c_void_p(p))

    def synth_audio_device(self, d_audio: int, bytes_per_utt: int, n_utt: int, seed: int = 1):
        self._ck(self._L.phn_synth_audio_device(self._h, d_audio, int(bytes_per_utt), int(n_utt), int(seed)))

    def memcpy_h2d(self, dst: int, src_ptr: int, nbytes: int):
        self._ck(self._L.phn_memcpy_h2d(self._h, dst, src_ptr, int(nbytes)))

    def memcpy_d2h(self, dst_ptr: int, src: int, nbytes: int):
        self._ck(self._L.phn_memcpy_d2h(self._h, dst_ptr, src, int(nbytes)))

    def synth_audio(self, bytes_per_utt: int, n_utt: int, seed: int = 1) -> np.ndarray:
        """Synthetic audio generated on the device, returned as a host uint8 array [n_utt, bytes_per_utt]."""
        d = self.device_alloc(bytes_per_utt * n_utt)
        try:
            self.synth_audio_device(d, bytes_per_utt, n_utt, seed)
            out = np.zeros(bytes_per_utt * n_utt, dtype=np.uint8)
            self.memcpy_d2h(out.ctypes.data, d, out.nbytes)
        finally:
            self.device_free(d)
        return out.reshape(n_utt, bytes_per_utt)

    def debug_logf(self, first_bits: int, n: int) -> np.ndarray:
        """The device's logf on the n consecutive float bit patterns from first_bits (verification aid)."""
        out = np.zeros(n, dtype=np.float32)
        self._ck(self._L.phn_debug_logf(self._h, int(first_bits), int(n), out))
        return out

    def online_norm(self, x: np.ndarray, interval: int, mean_norm: bool, var_norm: bool) -> np.ndarray:
        y = np.array(x, dtype=np.float32, copy=True, order="C")
        self._ck(self._L.phn_online_norm(self._h, y.reshape(-1), y.shape[0], y.shape[1], interval, int(mean_norm),
                                         int(var_norm)))
        return y


# ---------------------------------------------------------------------------------------------
# text / file formats next to the path (host side)
# ---------------------------------------------------------------------------------------------
def format_rec(labels, phonemes) -> str:
    """.rec text as PhnDec prints it (phndec.cpp:230,292): "%d00000 %d00000 %s %f"."""
    return "".join("%d00000 %d00000 %s %f\n" % (int(l["start"]), int(l["end"]), phonemes[int(l["phn"])], float(l["like"]))
                   for l in labels)


def format_mlf_entry(name: str, labels, phonemes) -> str:
    """One MLF entry as SpeechRec::ProcessFile + OnWordMLF print it (srec.cpp:137-161,1156,1180)."""
    out = ['"%s"\n' % name]
    for l in labels:
        s, e = int(l["start"]), int(l["end"])
        st = "0" if s == 0 else "%u00000" % s
        en = "0" if e == 0 else "%u00000" % e
        out.append("%s %s %s %f\n" % (st, en, phonemes[int(l["phn"])], float(l["like"])))
    out.append(".\n")
    return "".join(out)


def read_htk(path) -> np.ndarray:
    """Mat::loadHTK (matrix.h:2540-2573): 12-byte big-endian header + big-endian float32 rows."""
    b = Path(path).read_bytes()
    n, _period, size, _kind = struct.unpack(">iihh", b[:12])
    cols = size // 4
    return np.frombuffer(b, dtype=">f4", count=n * cols, offset=12).astype(np.float32).reshape(n, cols)


def write_htk(path, m: np.ndarray) -> None:
    """Mat::saveHTK (matrix.h:2506-2538): sampPeriod 100000, parmKind 6 (header defaults matrix.h:412-423)."""
    m = np.ascontiguousarray(m, dtype=np.float32)
    Path(path).write_bytes(struct.pack(">iihh", m.shape[0], 100000, m.shape[1] * 4, 6) + m.astype(">f4").tobytes())
