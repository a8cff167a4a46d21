#!/usr/bin/env python3
"""GPU diagnostics of round 2 (scratch): (1) which utterances of the 1000-utterance property test break the cover property, in
which mode; (2) the worst ln p deviations of the RU system."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import phnrec_b200 as pb

def tiles(lab, T):
    if len(lab) == 0:
        return "empty"
    s, e = lab["start"].astype(np.int64), lab["end"].astype(np.int64)
    if s[0] != 0: return f"first start {s[0]}"
    if e[-1] != T: return f"last end {e[-1]}"
    if not (s[1:] >= e[:-1]).all(): return "overlap"
    if not (e > s).all(): return "empty segment"
    return ""

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cover"):
    r = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_CZ_SPDAT_LCRC_N1500", device=0)
    r.set_wave_format("alaw")
    a = r.synth_audio(80000, 1000, seed=2024)
    utts = [a[i].tobytes() for i in range(1000)]
    for mode, name in ((pb.MLP_TC_F16, "tc"), (pb.MLP_EXACT_FP32, "exact")):
        r.set_mlp_mode(mode)
        lab = r.recognize(utts)
        bad = [(i, tiles(l, 998)) for i, l in enumerate(lab) if tiles(l, 998)]
        print(name, "violations:", bad[:10], len(bad))
        for i, why in bad[:2]:
            print(pb.format_rec(lab[i], r.phonemes)[:600])
            np.save(ROOT / "gpurun_out" / f"diag_cover_utt{i}.npy", a[i])
    r.close()
if which in ("all", "ru"):
    r = pb.Recognizer(ROOT / "oracle/_ref/models/PHN_RU_SPDAT_LCRC_N1500", device=0)
    r.set_wave_format("alaw")
    a = r.synth_audio(80000, 128, seed=1000)
    utts = [a[i].tobytes() for i in range(128)]
    F = 128 * 998
    r.set_mlp_mode(pb.MLP_EXACT_FP32); r.recognize(utts); le = r.fetch_logp(F)
    r.set_mlp_mode(pb.MLP_TC_F16); r.recognize(utts); lt = r.fetch_logp(F)
    m = np.abs(lt.astype(np.float64) - le) / np.maximum(1.0, np.abs(le))
    m[~np.isfinite(m)] = 0
    idx = np.argsort(m.ravel())[::-1][:12]
    for k in idx:
        f, c = divmod(int(k), m.shape[1])
        print(f"frame {f} (utt {f // 998} t {f % 998}) col {c}: exact {le[f, c]:.5f} tc {lt[f, c]:.5f} m {m[f, c]:.4f}; row max exact {le[f].max():.4f} tc {lt[f].max():.4f}")
    fr = np.unique([int(k) // m.shape[1] for k in idx])
    print("frames with large deviations:", fr[:20], "count of values with m > 0.1:", int((m > 0.1).sum()))
    # the staged path on the same audio (exact front end + tensor-core nets): does it show the same outliers?
    mels = r.mel(utts)
    pt = np.concatenate(r.posteriors(mels))
    r.set_mlp_mode(pb.MLP_EXACT_FP32)
    pe = np.concatenate(r.posteriors(mels))
    P3 = 3 * r.n_phonemes
    ms = np.abs(np.log(np.maximum(pt[:, :P3], 1e-45).astype(np.float64)) - np.log(np.maximum(pe[:, :P3], 1e-45))) / np.maximum(1.0, np.abs(np.log(np.maximum(pe[:, :P3], 1e-45))))
    print("staged path: max", ms.max(), "at frame", int(ms.max(1).argmax()), "values > 0.1:", int((ms > 0.1).sum()))
    r.close()
