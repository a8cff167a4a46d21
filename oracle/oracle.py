"""ctypes front end of the CPU oracle (oracle/phn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(phnrec_b200/) never imports this module.

Also holds the helpers that drive the REAL reference binary built by
oracle/Makefile (oracle/_ref/phnrec_ref) and read/write its HTK files
(matrix.h:2506-2573).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libphn_oracle.so"
REF_DIR = HERE / "_ref"
REF_BIN = REF_DIR / "phnrec_ref"
REF_MODELS = REF_DIR / "models"
REF_AUDIO = REF_DIR / "audio"

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

LABEL_DTYPE = np.dtype([("phn", np.int32), ("start", np.int32), ("end", np.int32), ("like", np.float32)])


def build(force: bool = False) -> None:
    """Compile the C restatement (and the reference binary when /root/reference exists)."""
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "phn_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "restatement"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    L.orc_alaw_table.argtypes = [_i16p]
    L.orc_wave_to_float.argtypes = [C.c_int, vp, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_wave_to_float.restype = C.c_int
    L.orc_mel_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
    L.orc_mel_create.restype = vp
    L.orc_mel_destroy.argtypes = [vp]
    L.orc_mel_set_plp.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]
    L.orc_mel_nparams.argtypes = [vp]
    L.orc_mel_fft_size.argtypes = [vp]
    L.orc_mel_tables.argtypes = [vp, _f32p, _f32p, _i16p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.orc_num_frames.argtypes = [C.c_int] * 3
    L.orc_mel_compute.argtypes = [vp, _f32p, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_mel_compute.restype = C.c_int
    L.orc_sentence_mean_norm.argtypes = [_f32p, C.c_int, C.c_int]
    L.orc_online_norm.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.orc_stc.argtypes = [_f32p, C.c_int, C.c_int, _f32p, C.c_int, _f32p, _f32p]
    L.orc_dct_table.argtypes = [_f32p]
    L.orc_stc_fifo.argtypes = [_f32p, C.c_int, C.c_int, _f32p]
    L.orc_nn_load.argtypes = [C.c_char_p]
    L.orc_nn_load.restype = vp
    L.orc_nn_destroy.argtypes = [vp]
    L.orc_nn_dims.argtypes = [vp, _i32p]
    L.orc_fexp.argtypes = [C.c_double]
    L.orc_fexp.restype = C.c_double
    L.orc_fsigmoid.argtypes = [C.c_float]
    L.orc_fsigmoid.restype = C.c_float
    L.orc_fsoftmax.argtypes = [C.c_int, _f32p]
    L.orc_nn_forward.argtypes = [vp, _f32p, _f32p, C.c_int]
    L.orc_nn_hidden.argtypes = [vp, _f32p, _f32p, C.c_int]
    L.orc_posteriors_from_normed_mel.argtypes = [_f32p, C.c_int, C.c_int, _f32p, vp, vp, vp, _f32p]
    L.orc_log_inplace.argtypes = [_f32p, C.c_size_t]
    L.orc_logf_port.argtypes = [C.c_float]
    L.orc_logf_port.restype = C.c_float
    L.orc_logf_port_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.orc_logf_port_mismatches.restype = C.c_uint64
    L.orc_decode.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, vp, C.c_int]
    L.orc_decode.restype = C.c_int
    L.orc_model_load.argtypes = [C.c_char_p]
    L.orc_model_load.restype = vp
    L.orc_model_destroy.argtypes = [vp]
    L.orc_model_info.argtypes = [vp, _i32p, C.POINTER(C.c_float)]
    L.orc_model_phoneme.argtypes = [vp, C.c_int]
    L.orc_model_phoneme.restype = C.c_char_p
    L.orc_model_net.argtypes = [vp, C.c_int]
    L.orc_model_net.restype = vp
    L.orc_model_windows.argtypes = [vp]
    L.orc_model_windows.restype = C.POINTER(C.c_float)
    L.orc_model_mel.argtypes = [vp]
    L.orc_model_mel.restype = vp
    L.orc_model_num_frames.argtypes = [vp, C.c_int, C.c_int]
    L.orc_model_nparams.argtypes = [vp]
    L.orc_model_mel_from_audio.argtypes = [vp, vp, C.c_int, C.c_int, _f32p]
    L.orc_model_mel_from_audio.restype = C.c_int
    L.orc_model_posteriors.argtypes = [vp, _f32p, C.c_int, _f32p]
    L.orc_model_decode.argtypes = [vp, _f32p, C.c_int, C.c_float, vp, C.c_int]
    L.orc_model_decode.restype = C.c_int
    L.orc_model_recognize.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, vp, C.c_int]
    L.orc_model_recognize.restype = C.c_int
    L.orc_model_recognize_online.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.orc_model_recognize_online.restype = C.c_int
    _lib = L
    return L


FMT = {"lin16": 0, "alaw": 1}


def _buf(b):
    a = np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b.view(np.uint8).reshape(-1)
    a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(C.c_void_p), int(a.size)


def decode(logpost: np.ndarray, P: int, wp: float, S: int = 3, hist: int = 40) -> np.ndarray:
    """PhnDec over log-posteriors [T, ncols] -> structured label array."""
    lp = np.ascontiguousarray(logpost, dtype=np.float32)
    T, nc = lp.shape
    cap = T + 64
    out = np.zeros(cap, dtype=LABEL_DTYPE)
    n = lib().orc_decode(lp, T, nc, P, S, hist, C.c_float(wp), out.ctypes.data_as(C.c_void_p), cap)
    return out[:n].copy()


def logf(x: np.ndarray) -> np.ndarray:
    """glibc logf (the reference's SoftLog), elementwise."""
    y = np.array(x, dtype=np.float32, copy=True, order="C")
    lib().orc_log_inplace(y.reshape(-1), y.size)
    return y


class Model:
    """The reference system loaded from a PHN_* directory (oracle side)."""

    def __init__(self, cfg_dir):
        self.dir = str(cfg_dir)
        self.h = lib().orc_model_load(self.dir.encode())
        if not self.h:
            raise FileNotFoundError(f"oracle: cannot load model dir {cfg_dir}")
        info = np.zeros(12, dtype=np.int32)
        wp = C.c_float()
        lib().orc_model_info(self.h, info, C.byref(wp))
        (self.fs, self.nbanks, self.vs, self.step, self.fmt, self.sent_mean_norm, self.hist, self.S, self.P,
         self.nout, self.nin_band, self.nhid) = (int(v) for v in info)
        self.wpenalty = float(wp.value)
        self.phonemes = [lib().orc_model_phoneme(self.h, i).decode() for i in range(self.P)]

    def close(self):
        if self.h:
            lib().orc_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def net(self, which: int):
        return lib().orc_model_net(self.h, which)

    def net_dims(self, which: int):
        d = np.zeros(3, dtype=np.int32)
        lib().orc_nn_dims(self.net(which), d)
        return tuple(int(v) for v in d)

    def windows(self) -> np.ndarray:
        p = lib().orc_model_windows(self.h)
        return np.ctypeslib.as_array(p, shape=(32,)).copy()

    def num_frames(self, nbytes: int, fmt=None) -> int:
        f = -1 if fmt is None else FMT[fmt]
        return lib().orc_model_num_frames(self.h, nbytes, f)

    def mel(self, audio, fmt=None) -> np.ndarray:
        a, p, n = _buf(audio)
        f = -1 if fmt is None else FMT[fmt]
        T = lib().orc_model_num_frames(self.h, n, f)
        out = np.zeros((T, lib().orc_model_nparams(self.h)), dtype=np.float32)   # nbanks, or the PLP coefficients (params/kind = plp)
        lib().orc_model_mel_from_audio(self.h, p, n, f, out)
        return out

    def stc(self, mel_normed: np.ndarray):
        m = np.ascontiguousarray(mel_normed, dtype=np.float32)
        T = m.shape[0]
        ncoef = self.nin_band // self.nbanks
        XL = np.zeros((T, self.nin_band), dtype=np.float32)
        XR = np.zeros((T, self.nin_band), dtype=np.float32)
        lib().orc_stc(m, T, self.nbanks, self.windows(), ncoef, XL, XR)
        return XL, XR

    def nn_forward(self, which: int, x: np.ndarray) -> np.ndarray:
        nin, _, nout = self.net_dims(which)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, nin)
        out = np.zeros((x.shape[0], nout), dtype=np.float32)
        lib().orc_nn_forward(self.net(which), x, out, x.shape[0])
        return out

    def nn_hidden(self, which: int, x: np.ndarray) -> np.ndarray:
        nin, nhid, _ = self.net_dims(which)
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, nin)
        out = np.zeros((x.shape[0], nhid), dtype=np.float32)
        lib().orc_nn_hidden(self.net(which), x, out, x.shape[0])
        return out

    def sentence_norm(self, mel: np.ndarray) -> np.ndarray:
        m = np.array(mel, dtype=np.float32, copy=True, order="C")
        if self.sent_mean_norm:
            lib().orc_sentence_mean_norm(m, m.shape[0], m.shape[1])
        return m

    def posteriors(self, mel: np.ndarray) -> np.ndarray:
        m = np.ascontiguousarray(mel, dtype=np.float32)
        out = np.zeros((m.shape[0], self.nout), dtype=np.float32)
        lib().orc_model_posteriors(self.h, m, m.shape[0], out)
        return out

    def decode(self, post: np.ndarray, wp=None) -> np.ndarray:
        p = np.ascontiguousarray(post, dtype=np.float32)
        T = p.shape[0]
        cap = T + 64
        out = np.zeros(cap, dtype=LABEL_DTYPE)
        w = self.wpenalty if wp is None else wp
        n = lib().orc_model_decode(self.h, p, T, C.c_float(w), out.ctypes.data_as(C.c_void_p), cap)
        return out[:n].copy()

    def recognize(self, audio, fmt=None, wp=None) -> np.ndarray:
        a, p, n = _buf(audio)
        f = -1 if fmt is None else FMT[fmt]
        T = lib().orc_model_num_frames(self.h, n, f)
        cap = T + 64
        out = np.zeros(cap, dtype=LABEL_DTYPE)
        w = self.wpenalty if wp is None else wp
        k = lib().orc_model_recognize(self.h, p, n, f, C.c_float(w), out.ctypes.data_as(C.c_void_p), cap)
        return out[:k].copy()


    def recognize_online(self, audio, fmt=None, wp=None, interval=-1, mean_norm=-1, var_norm=-1) -> np.ndarray:
        """The online path (SpeechRec::ProcessOnline, srec.cpp:793-927) over the whole signal; -1 = the config's [onlinenorm]."""
        a, p, n = _buf(audio)
        f = -1 if fmt is None else FMT[fmt]
        cap = n // 50 + 128
        out = np.zeros(cap, dtype=LABEL_DTYPE)
        w = self.wpenalty if wp is None else wp
        k = lib().orc_model_recognize_online(self.h, p, n, f, C.c_float(w), int(interval), int(mean_norm), int(var_norm),
                                             out.ctypes.data_as(C.c_void_p), cap)
        return out[:k].copy()


# ----------------------------------------------------------------------------
# text formats (phndec.cpp:230,292 / srec.cpp:137-161)
# ----------------------------------------------------------------------------
def format_rec(labels: np.ndarray, phonemes) -> str:
    """`.rec` text exactly as PhnDec prints it: "%d00000 %d00000 %s %f\\n"."""
    return "".join("%d00000 %d00000 %s %f\n" % (int(l["start"]), int(l["end"]), phonemes[int(l["phn"])], float(l["like"]))
                   for l in labels)


def format_vad(labels: np.ndarray, phonemes) -> str:
    """`.rec` text of the fork's `vadalize` tool (phndecalize.cpp:227-239, 299-314): "%.2f %.2f speech" in seconds
    (float arithmetic) for every segment whose phoneme is not pau / int / spk."""
    out = []
    for l in labels:
        if phonemes[int(l["phn"])] in ("pau", "int", "spk"):
            continue
        s, e = np.float32(int(l["start"])) / np.float32(100), np.float32(int(l["end"])) / np.float32(100)
        out.append("%.2f %.2f speech\n" % (float(s), float(e)))
    return "".join(out)


def parse_rec(text: str):
    """-> list of (start_frame, end_frame, phoneme, score) from .rec / MLF body lines."""
    out = []
    for ln in text.splitlines():
        p = ln.split()
        if len(p) != 4:
            continue
        out.append((int(p[0]) // 100000, int(p[1]) // 100000, p[2], float(p[3])))
    return out


# ----------------------------------------------------------------------------
# HTK parameter files (matrix.h:2506-2573): 12-byte big-endian header + BE f32
# ----------------------------------------------------------------------------
def read_htk(path) -> np.ndarray:
    b = Path(path).read_bytes()
    n, _period, size, _kind = struct.unpack(">iihh", b[:12])
    cols = size // 4
    return np.frombuffer(b, dtype=">f4", count=n * cols, offset=12).astype(np.float32).reshape(n, cols)


def write_htk(path, m: np.ndarray) -> None:
    m = np.ascontiguousarray(m, dtype=np.float32)
    hdr = struct.pack(">iihh", m.shape[0], 100000, m.shape[1] * 4, 6)
    Path(path).write_bytes(hdr + m.astype(">f4").tobytes())


# ----------------------------------------------------------------------------
# the real reference binary
# ----------------------------------------------------------------------------
def have_ref() -> bool:
    return REF_BIN.exists() and os.access(REF_BIN, os.X_OK) and REF_MODELS.is_dir()


def run_ref(args, check=True, cwd=None) -> subprocess.CompletedProcess:
    """Run oracle/_ref/phnrec_ref with the given CLI arguments."""
    return subprocess.run([str(REF_BIN)] + [str(a) for a in args], check=check, capture_output=True, text=True, cwd=cwd)
