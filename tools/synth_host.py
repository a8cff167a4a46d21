"""Host (numpy) restatement of the device synthetic-audio generator phnrec_b200/csrc/k_synth.cu.

Bench / test infrastructure: bench.py's CPU reference arm recognises the SAME utterances the GPU arm does without loading
the CUDA library.  Every operation below is the single correctly rounded fp32 / uint64 operation of the kernel, in the
same order, so the bytes are identical (checked on the GPU box by tests/test_gpu_parity.py::test_synthetic_audio_host_port).
"""
import numpy as np

_F = np.float32
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = (z + _GOLD).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(30))) * _M1).astype(np.uint64)
        z = ((z ^ (z >> np.uint64(27))) * _M2).astype(np.uint64)
        return z ^ (z >> np.uint64(31))


def _u01(h):
    return (h >> np.uint64(40)).astype(_F) * _F(1.0 / 16777216.0)


def _sin_turns(y):
    fr = (y - np.floor(y)).astype(_F)
    z = (_F(2.0) * fr - _F(1.0)).astype(_F)
    p = ((_F(4.0) * z).astype(_F) * (_F(1.0) - np.abs(z)).astype(_F)).astype(_F)
    return -(p * (_F(0.775) + (_F(0.225) * np.abs(p)).astype(_F)).astype(_F)).astype(_F)


def _lin2alaw(v):
    pcm = v >> 3
    neg = pcm < 0
    mask = np.where(neg, 0x55, 0xD5)
    pcm = np.where(neg, -pcm - 1, pcm)
    seg = np.zeros_like(pcm)
    for s in range(8):
        seg = np.where(pcm > ((0x20 << s) - 1), s + 1, seg)
    sat = seg >= 8
    segc = np.minimum(seg, 7)
    aval = (segc << 4) | np.where(segc < 2, (pcm >> 1) & 0xF, (pcm >> np.maximum(segc, 1)) & 0xF)
    return np.where(sat, 0x7F ^ mask, aval ^ mask).astype(np.uint8)


def synth_audio(bytes_per_utt: int, n_utt: int, seed: int, fmt: str, fs: int, first_utt: int = 0) -> np.ndarray:
    """-> uint8 [n_utt, bytes_per_utt]: utterances first_utt .. first_utt + n_utt - 1 of phn_synth_audio_device(seed)."""
    bps = 2 if fmt == "lin16" else 1
    n_per = bytes_per_utt // bps
    out = np.zeros((n_utt, bytes_per_utt), dtype=np.uint8)
    n = np.arange(n_per, dtype=np.uint64)
    t = (n.astype(_F) / _F(fs)).astype(_F)
    with np.errstate(over="ignore"):
        for k in range(n_utt):
            u = first_utt + k
            us = _mix64(np.uint64(seed) ^ (_GOLD * np.uint64(u + 1)))
            dur = _F(0.5) + _u01(_mix64(us ^ np.uint64(1)))
            total = _F(n_per) / _F(fs)
            s0 = _u01(_mix64(us ^ np.uint64(2))) * max(_F(total - dur), _F(0.0))
            silent = (t >= s0) & (t < _F(s0 + dur))
            seg = (n // np.uint64(fs // 10)).astype(np.uint64)
            hs = _mix64(us ^ (np.uint64(0x1000) + seg))
            f1 = (_F(200.0) + (_F(700.0) * _u01(hs)).astype(_F)).astype(_F)
            f2 = (_F(900.0) + (_F(1500.0) * _u01(_mix64(hs ^ np.uint64(11)))).astype(_F)).astype(_F)
            f3 = (_F(2400.0) + (_F(1000.0) * _u01(_mix64(hs ^ np.uint64(23)))).astype(_F)).astype(_F)
            ph = _u01(_mix64(us ^ np.uint64(3)))
            env = (_F(0.5) * (_F(1.0) - _sin_turns((((_F(4.0) * t).astype(_F) + ph).astype(_F) + _F(0.25)).astype(_F))).astype(_F)).astype(_F)
            voiced = ((_sin_turns((f1 * t).astype(_F)) + (_F(0.6) * _sin_turns((f2 * t).astype(_F))).astype(_F)).astype(_F)
                      + (_F(0.3) * _sin_turns((f3 * t).astype(_F))).astype(_F)).astype(_F)
            noise = ((_F(2.0) * _u01(_mix64(us ^ (np.uint64(0x5000000) + n)))).astype(_F) - _F(1.0)).astype(_F)
            x = (((_F(3500.0) * env).astype(_F) * voiced).astype(_F) + (_F(400.0) * noise).astype(_F)).astype(_F)
            x = np.where(silent, _F(0.0), x)
            v = np.clip(np.rint(x).astype(np.int64), -32768, 32767).astype(np.int32)
            if fmt == "lin16":
                out[k] = v.astype("<i2").view(np.uint8)
            else:
                out[k] = _lin2alaw(v)
    return out
