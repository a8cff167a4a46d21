"""The tensor-core mode's precision bound excludes "outlier frames" of the RU system's synthetic set: frames whose
ln p deviates by more than 0.25 on the stated measure (tools/tc_bound.py -> profiles/r2_tc_bound_ru.json).  This CPU test pins
WHY they are excluded: in every such frame a band net's hidden pre-activation lies beyond +-710.5, where the reference's
bit-trick exponential (fexp.h:14-21: (int)(2^20/ln2 * y) + 1072632447) overflows int32 - undefined behaviour in the reference
(x86 wraps: a saturated unit flips or becomes NaN).  The exact mode reproduces the x86 bits; the tensor-core mode saturates."""
import json
import sys

import numpy as np

from conftest import ROOT, model_dir

sys.path.insert(0, str(ROOT / "tools"))

OVERFLOW = (2 ** 31 - 1072632447) / (1048576 / 0.69314718055994530942)   # 710.5


def read_band_nbin(path):
    b = open(path, "rb").read()
    _, nin, nhid, nout = (int(x) for x in np.frombuffer(b[:16], dtype=np.int32))
    a4 = lambda n: (n + 3) // 4 * 4
    f = np.frombuffer(b[16:], dtype=np.float32)
    o = 0
    w1 = f[o:o + a4(nhid) * a4(nin)].reshape(a4(nhid), a4(nin)); o += w1.size
    o += a4(nout) * a4(nhid)
    b1 = f[o:o + a4(nhid)]; o += a4(nhid) + a4(nout)
    mean = f[o:o + a4(nin)]; o += a4(nin)
    dev = f[o:o + a4(nin)]
    return w1[:nhid, :nin].astype(np.float64), b1[:nhid].astype(np.float64), mean[:nin].astype(np.float64), dev[:nin].astype(np.float64)


def test_ru_outlier_frames_lie_in_the_references_overflow_zone(orc):
    import synth_host
    prof = json.loads((ROOT / "profiles" / "r2_tc_bound_ru.json").read_text())
    frames = prof["outlier_frames"]
    assert prof["n_outlier_frames"] == len(frames) and len(frames) >= 1
    md = model_dir("PHN_RU_SPDAT_LCRC_N1500")
    m = orc.Model(md)
    nets = [read_band_nbin(md / "weights" / f"band{i}.nbin") for i in range(2)]
    per_utt = 998                                        # frames of a 10 s utterance (tools/tc_bound.py: 80000 A-law bytes each)
    picked = frames[::4] + [frames[-1]]                  # every fourth (each costs one utterance through the CPU front end)
    for f in picked:
        u, t = divmod(f, per_utt)
        a = synth_host.synth_audio(80000, 1, prof["seed"], "alaw", 8000, first_utt=u)[0].tobytes()
        xs = m.stc(m.sentence_norm(m.mel(a, fmt="alaw")))
        worst = 0.0
        for (w1, b1, mean, dev), x in zip(nets, xs):
            pre = w1 @ ((x[t].astype(np.float64) - mean) * dev) + b1
            worst = max(worst, float(np.abs(pre).max()))
        assert worst > OVERFLOW, (f, worst)
