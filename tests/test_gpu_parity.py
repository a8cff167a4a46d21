"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs
and against the committed reference-built fixtures.  Exact mode bar: BIT-exact mel, posteriors,
labels, boundaries and scores.  Run on the B200 box:  python -m pytest tests -m gpu"""
import numpy as np
import pytest

from conftest import ALL_MODELS, audio_bytes, model_dir, ref_run

import phnrec_b200 as pb

pytestmark = pytest.mark.gpu

RUNS = [("PHN_CZ_SPDAT_LCRC_N1500", "test.raw"), ("PHN_CZ_SPDAT_LCRC_N1500", "8580.wav"),
        ("PHN_EN_TIMIT_LCRC_N500", "test.raw"), ("PHN_HU_SPDAT_LCRC_N1500", "test.raw"),
        ("PHN_RU_SPDAT_LCRC_N1500", "test.raw"), ("PHN_ES", "8580.wav"), ("PHN_ES", "es.wav")]


@pytest.fixture(scope="module")
def recs():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = pb.Recognizer(model_dir(name), device=0)
        return cache[name]
    yield get
    for r in cache.values():
        r.close()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(got, want, what):
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    bad = bits(got) != bits(want)
    if bad.any():
        d = np.abs(got.astype(np.float64) - want.astype(np.float64))
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.size} values differ bitwise, max abs diff {d.max():.3e}")


def labels_equal(a, b):
    return (len(a) == len(b) and np.array_equal(a["phn"], b["phn"]) and np.array_equal(a["start"], b["start"])
            and np.array_equal(a["end"], b["end"]) and np.array_equal(bits(a["like"]), bits(b["like"])))


# ----------------------------------------------------------------------------- fixtures from the reference build
@pytest.mark.parametrize("model,audio", RUNS)
def test_reference_fixtures_end_to_end(recs, model, audio):
    """audio -> mel -> posteriors -> .rec text, against what the reference's own binary wrote."""
    r = recs(model)
    ref = ref_run(model, audio)
    a = audio_bytes(audio)
    mel = r.mel([a])[0]
    assert_bits_equal(mel, ref["mel"], "mel vs reference -t par")
    post = r.posteriors([mel])[0]
    assert_bits_equal(post[ref["post_rows"]], ref["post"], "posteriors vs reference -t post")
    assert pb.format_rec(r.recognize([a])[0], r.phonemes) == str(ref["rec"])
    assert pb.format_rec(r.decode([post], penalties=[-1.5])[0][0], r.phonemes) == str(ref["rec_p15"])


def test_shipped_golden_labels(recs, ref_labels):
    for g, info in ref_labels.items():
        r = recs(info["model"])
        lab = r.recognize([audio_bytes(info["audio"])])[0]
        got = [(int(l["start"]), int(l["end"]), r.phonemes[int(l["phn"])]) for l in lab]
        assert got == [(s, e, p) for s, e, p, _ in info["labels"]], g


def test_mlf_text_matches_reference_mlf(recs, ref_labels):
    info = ref_labels["test/test"]
    r = recs(info["model"])
    lab = r.recognize([audio_bytes(info["audio"])])[0]
    got = "#!MLF!#\n" + pb.format_mlf_entry("8580.rec", lab, r.phonemes)
    want = info["text"]
    # another build wrote the golden: identical up to the last printed digits of the scores
    gl, wl = got.splitlines(), want.splitlines()
    assert len(gl) == len(wl)
    for a, b in zip(gl, wl):
        assert a.split()[:3] == b.split()[:3]


# ----------------------------------------------------------------------------- against the oracle, stage by stage
def ragged_audio(rng, fs_bytes, lengths):
    return [rng.integers(-3000, 3000, size=n, dtype=np.int16).tobytes()[:fs_bytes * n] for n in lengths]


@pytest.mark.parametrize("model", ["PHN_CZ_SPDAT_LCRC_N1500", "PHN_EN_TIMIT_LCRC_N500"])
def test_ragged_batch_every_stage_bit_exact(recs, oracle_models, model):
    r, o = recs(model), oracle_models(model)
    rng = np.random.default_rng(5)
    vs, st = r.vector_size, r.vector_step
    # edge cases: empty, shorter than one window, exactly one window, one sample short of 2 frames,
    # fewer frames than the 15-frame context, around the 41-frame decoder horizon, and long ones
    nsamp = [0, 1, vs - 1, vs, vs + st - 1, vs + st, vs + 5 * st, vs + 13 * st, vs + 14 * st, vs + 39 * st,
             vs + 40 * st, vs + 41 * st, vs + 200 * st + 7, vs + 997 * st]
    base = np.frombuffer(audio_bytes("test.raw"), dtype=np.int16)
    utts = []
    for i, n in enumerate(nsamp):
        seg = base[(i * 997) % 20000:][:n].copy()
        if i % 3 == 0 and n > 50:
            seg[n // 3: n // 2] = 0          # digital silence: the sLn zero guard
        utts.append(seg.tobytes())
    mels = r.mel(utts)
    for u, m in zip(utts, mels):
        assert_bits_equal(m, o.mel(u), f"mel, {len(u)} bytes")
    posts = r.posteriors(mels)
    for m, p in zip(mels, posts):
        assert_bits_equal(p, o.posteriors(m), f"posteriors, {m.shape[0]} frames")
    labs = r.decode(posts)
    for p, l in zip(posts, labs):
        assert labels_equal(l, o.decode(p)), f"labels, {p.shape[0]} frames"
    for u, l in zip(utts, r.recognize(utts)):
        assert labels_equal(l, o.recognize(u)), f"recognize, {len(u)} bytes"


def test_alaw_input_path(recs, oracle_models, orc):
    r, o = recs("PHN_CZ_SPDAT_LCRC_N1500"), oracle_models("PHN_CZ_SPDAT_LCRC_N1500")
    rng = np.random.default_rng(11)
    utts = [rng.integers(0, 256, size=n, dtype=np.uint8).tobytes() for n in (80000, 12345, 199, 200, 201)]
    utts.append(bytes(range(256)) * 40)     # every A-law code
    r.set_wave_format("alaw")
    try:
        for u, m in zip(utts, r.mel(utts)):
            assert_bits_equal(m, o.mel(u, fmt="alaw"), f"alaw mel, {len(u)} bytes")
        lin = np.frombuffer(audio_bytes("test.raw"), dtype=np.int16)
        t = np.zeros(256, dtype=np.int16)
        orc.lib().orc_alaw_table(t)
        inv = {int(v) * 8: i for i, v in enumerate(t)}
        enc = np.array([inv[int(s)] for s in lin], dtype=np.uint8).tobytes()
        lab = r.recognize([enc])[0]
    finally:
        r.set_wave_format("lin16")
    assert pb.format_rec(lab, r.phonemes) == str(ref_run("PHN_CZ_SPDAT_LCRC_N1500", "test.raw")["rec"])


@pytest.mark.parametrize("model", ALL_MODELS)
def test_decoder_bit_exact_on_random_posteriors(recs, oracle_models, model):
    """Flat random posteriors drive the decoder through its odd corners (skipped commits, ties,
    utterances shorter than the 41-frame horizon, zeros -> log = -inf)."""
    r, o = recs(model), oracle_models(model)
    rng = np.random.default_rng(3)
    posts = []
    for T in (1, 2, 3, 11, 39, 40, 41, 42, 43, 100, 333, 1200):
        p = rng.random((T, r.n_outputs)).astype(np.float32) ** 8 + 1e-6
        p /= p.sum(1, keepdims=True)
        if T > 50:
            p[T // 2, :] = 1.0 / r.n_outputs          # exact ties across all states
            p[T // 3, 5] = 0.0                        # logf(0) = -inf
        posts.append(p)
    for wp in (None, 0.0, -10.0):
        if wp is not None:
            r.set_penalty(wp)
        labs = r.decode(posts)
        for p, l in zip(posts, labs):
            assert labels_equal(l, o.decode(p, wp=wp)), (model, p.shape[0], wp)
    r.set_penalty(o.wpenalty)


def test_penalty_sweep_from_saved_posteriors(recs, oracle_models):
    """BASELINE config 4: decode saved posteriors under 14 penalties in one call (-s post -p P)."""
    model = "PHN_EN_TIMIT_LCRC_N500"
    r, o = recs(model), oracle_models(model)
    ref = ref_run(model, "test.raw")
    post = np.ascontiguousarray(ref["post"])
    pens = [-6.0 + 0.5 * i for i in range(13)] + [o.wpenalty]
    out = r.decode([post, post[:100]], penalties=pens)
    for k, wp in enumerate(pens):
        assert labels_equal(out[k][0], o.decode(post, wp=wp)), wp
        assert labels_equal(out[k][1], o.decode(post[:100], wp=wp)), wp
    assert pb.format_rec(out[-1][0], r.phonemes) == str(ref["rec"])


def test_synthetic_alaw_batch_matches_oracle(recs, oracle_models):
    """BASELINE config 2 in miniature: synthetic 8 kHz A-law utterances generated on the device."""
    r, o = recs("PHN_CZ_SPDAT_LCRC_N1500"), oracle_models("PHN_CZ_SPDAT_LCRC_N1500")
    r.set_wave_format("alaw")
    try:
        audio = r.synth_audio(80000, 6, seed=1234)
        assert audio.shape == (6, 80000) and len(np.unique(audio)) > 100
        assert r.num_frames(80000) == 998
        utts = [audio[i].tobytes() for i in range(6)]
        labs = r.recognize(utts)
        for u, l in zip(utts, labs):
            want = o.recognize(u, fmt="alaw")
            assert labels_equal(l, want)
            assert len(l) > 20
        again = r.synth_audio(80000, 6, seed=1234)
        assert np.array_equal(audio, again)
    finally:
        r.set_wave_format("lin16")


@pytest.mark.parametrize("model,fmt,nbytes", [("PHN_CZ_SPDAT_LCRC_N1500", "alaw", 80000), ("PHN_EN_TIMIT_LCRC_N500", "lin16", 320000)])
def test_synthetic_audio_host_port(model, fmt, nbytes):
    """tools/synth_host.py (what bench.py's CPU reference arm recognises) reproduces the device generator byte for byte, so
    both arms of the benchmark see the same utterances."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, str(ROOT))
    from tools.synth_host import synth_audio
    r = pb.Recognizer(model_dir(model), device=0)
    try:
        r.set_wave_format(fmt)
        dev = r.synth_audio(nbytes, 5, seed=1000)
        host = synth_audio(nbytes, 5, seed=1000, fmt=fmt, fs=r.sample_freq)
        assert np.array_equal(dev, host)
        assert np.array_equal(r.synth_audio(nbytes, 7, seed=77)[4:], synth_audio(nbytes, 3, seed=77, fmt=fmt, fs=r.sample_freq, first_utt=4))
    finally:
        r.close()


def test_exact_mode_inside_the_bit_trick_exponentials_overflow_zone(recs, oracle_models):
    """fexp.h:14-21 adds an int32 constant to (int)(2^20/ln2 * y): for |y| > 710.5 the sum overflows (undefined behaviour in
    the reference; x86 wraps, the hidden activation becomes NaN and the net's soft-max degenerates to the uniform
    distribution).  Utterance 99 of the RU synthetic set (seed 1000) has such a frame (783: a band-1 pre-activation of
    822).  The exact mode must still give what the x86 reference gives: posteriors of that neighbourhood bit for bit, and
    the same labels for the whole utterance."""
    import sys
    from conftest import ROOT
    sys.path.insert(0, str(ROOT))
    from tools.synth_host import synth_audio
    r, o = recs("PHN_RU_SPDAT_LCRC_N1500"), oracle_models("PHN_RU_SPDAT_LCRC_N1500")
    a = synth_audio(80000, 1, seed=1000, fmt="alaw", fs=8000, first_utt=99)[0].tobytes()
    r.set_wave_format("alaw")
    try:
        mel = r.mel([a])[0]
        want_mel = o.mel(a, fmt="alaw")
        assert_bits_equal(mel, want_mel, "mel")
        post = r.posteriors([mel])[0]
        want = o.posteriors(want_mel)
        assert np.isfinite(want[783]).all() and np.ptp(want[783, :10]) < 0.2      # (the degenerate frame is finite)
        assert_bits_equal(post[760:800], want[760:800], "posteriors around the overflow frame")
        assert labels_equal(r.recognize([a])[0], o.recognize(a, fmt="alaw"))
    finally:
        r.set_wave_format("lin16")


def test_batch_equals_singletons_and_is_order_independent(recs):
    """Size-independent property: utterances are independent, so batching must not change results."""
    r = recs("PHN_CZ_SPDAT_LCRC_N1500")
    base = audio_bytes("test.raw")
    utts = [base[:30000], base[30000:90000], base[:2000], base]
    together = r.recognize(utts)
    rev = r.recognize(utts[::-1])[::-1]
    for u, a, b in zip(utts, together, rev):
        c = r.recognize([u])[0]
        assert labels_equal(a, c) and labels_equal(b, c)


def test_optional_front_end_arithmetic_equals_reference_binary(tmp_path, oracle_models):
    """dc_shift / scale (srec.cpp:780-788), z_mean_source / preem_coef (melbanks.cpp:111-149), framenorm shift / min_floor
    (srec.cpp:1594-1620): the CUDA path on an edited copy of the model directory against what the reference binary made of
    the same directory (tests/golden/ref_front_variants.npz) - mel bit-identical, .rec text identical; and the fast
    (tensor-core pipeline) front end stays within its stated bound of the exact one under the same switches."""
    from conftest import front_end_variants, variant_model_dir
    for name, model, audio, nbytes, edits, mel, rec in front_end_variants():
        a = audio_bytes(audio)[:nbytes]
        r = pb.Recognizer(variant_model_dir(tmp_path / name, model, edits), device=0)
        try:
            got = r.mel([a])[0]
            assert_bits_equal(got, mel, f"mel, variant {name}")
            assert pb.format_rec(r.recognize([a])[0], r.phonemes) == rec, name
            r.set_mlp_mode(pb.MLP_TC_F16)
            r.recognize([a])
            fast = r.fetch_mel(mel.shape[0])
            assert np.abs(fast - mel).max() <= 1e-3, (name, np.abs(fast - mel).max())
        finally:
            r.close()


def test_batch_arguments_are_validated_before_use(recs):
    """A negative utterance count, offsets that do not start at 0 or decrease, a null audio pointer: PHN_ERR_ARG, nothing is
    read through them (capi.cu: plan_audio runs before any buffer is sized or dereferenced)."""
    r = recs("PHN_CZ_SPDAT_LCRC_N1500")
    a = np.frombuffer(audio_bytes("test.raw"), dtype=np.uint8)[:20000].copy()
    labels = np.zeros(400, dtype=pb.LABEL_DTYPE)
    loff = np.zeros(3, dtype=np.int64)
    for boff, n in ((np.array([0, 20000], np.int64), -1), (np.array([4, 20000], np.int64), 1),
                    (np.array([0, 20000, 10000], np.int64), 2)):
        assert r._L.phn_recognize(r._h, a.ctypes.data, boff, n, labels.ctypes.data, 400, loff, None) == 31
    assert r._L.phn_recognize(r._h, None, np.array([0, 20000], np.int64), 1, labels.ctypes.data, 400, loff, None) == 31
    # and the context is still usable
    assert len(r.recognize([a.tobytes()])[0]) > 3


def test_device_logf_equals_glibc_on_every_float_up_to_one(recs, orc):
    """the device logf (glibc port, device_math.cuh) against the host's libm logf on EVERY float in (0, 1] -
    all 1 065 353 216 bit patterns 0x00000001 .. 0x3F800000, subnormals included - plus zero, one-and-above samples,
    infinity and NaN.  Bit-exact: the decoder's scores are sums of these values."""
    r = recs("PHN_CZ_SPDAT_LCRC_N1500")
    first, last = 0x00000001, 0x3F800000
    step = 1 << 26
    bad = 0
    for lo in range(first, last + 1, step):
        n = min(step, last + 1 - lo)
        got = r.debug_logf(lo, n)
        want = np.arange(lo, lo + n, dtype=np.uint32).view(np.float32).copy()
        orc.lib().orc_log_inplace(want, want.size)
        bad += int((got.view(np.uint32) != want.view(np.uint32)).sum())
    assert bad == 0
    for lo in (0x00000000, 0x3F800001, 0x40000000, 0x7F7FFFF0, 0x7F800000):     # 0, just above 1, 2.0, near FLT_MAX, +inf
        got = r.debug_logf(lo, 16)
        want = np.arange(lo, lo + 16, dtype=np.uint32).view(np.float32).copy()
        orc.lib().orc_log_inplace(want, want.size)
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan)
        assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])


def test_online_norm_equals_reference_object(recs, orc):
    """Row N2: phn_online_norm against the outputs of the reference's own Normalization object (fixture made by
    oracle/_ref/online_ref, tests/golden/ref_online_norm.npz) and against the oracle; var_norm without mean_norm is the
    reference's assert (norm.cpp:150-155) -> PHN_ERR_ARG."""
    from test_oracle_pinned import online_norm_cases
    r = recs("PHN_CZ_SPDAT_LCRC_N1500")
    for cols, interval, mean, var, x, want in online_norm_cases():
        got = r.online_norm(x, interval, bool(mean), bool(var))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (cols, interval, mean, var)
        chk = x.copy()
        orc.lib().orc_online_norm(chk, chk.shape[0], cols, interval, mean, var)
        assert np.array_equal(got.view(np.uint32), chk.view(np.uint32))
    with pytest.raises(pb.PhnRecError) as e:
        r.online_norm(np.ones((10, 15), np.float32), 5, False, True)
    assert e.value.code == 31


def test_capacity_error_reports_needed_size(recs):
    import ctypes as C
    r = recs("PHN_CZ_SPDAT_LCRC_N1500")
    a = np.frombuffer(audio_bytes("test.raw"), dtype=np.uint8)
    boff = np.array([0, a.size], dtype=np.int64)
    labels = np.zeros(2, dtype=pb.LABEL_DTYPE)
    loff = np.zeros(2, dtype=np.int64)
    rc = r._L.phn_recognize(r._h, a.ctypes.data, boff, 1, labels.ctypes.data, 2, loff, None)
    assert rc == 32 and loff[1] == 50


def test_exact_several_mlp_passes_equal_reference(monkeypatch):
    """The exact mode with forced small passes (PHNREC_PASS_FRAMES): still the reference's bits."""
    monkeypatch.setenv("PHNREC_PASS_FRAMES", "256")
    r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
    try:
        ref = ref_run("PHN_CZ_SPDAT_LCRC_N1500", "test.raw")
        lab = r.recognize([audio_bytes("test.raw")])[0]
        assert pb.format_rec(lab, r.phonemes) == str(ref["rec"])
    finally:
        r.close()


@pytest.mark.parametrize("mode", [pb.MLP_EXACT_FP32, pb.MLP_TC_F16])
def test_degenerate_batches(oracle_models, mode):
    """Edge cases of the batch interface: an empty batch, empty utterances (0 bytes: the reference still makes ONE frame,
    srec.cpp:945, and decodes it), an utterance shorter than a window, all in one call."""
    r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
    try:
        r.set_mlp_mode(mode)
        assert r.recognize([]) == []
        a = audio_bytes("test.raw")
        utts = [b"", a[:100], a[:398], b"", a[:4000]]
        got = r.recognize(utts)
        assert len(got) == 5
        om = oracle_models("PHN_CZ_SPDAT_LCRC_N1500")
        for u, g in zip(utts, got):
            want = om.recognize(u)
            assert [(int(x["start"]), int(x["end"])) for x in g] == [(int(x["start"]), int(x["end"])) for x in want]
            if mode == pb.MLP_EXACT_FP32:
                assert pb.format_rec(g, r.phonemes) == pb.format_rec(want, om.phonemes)
    finally:
        r.close()
