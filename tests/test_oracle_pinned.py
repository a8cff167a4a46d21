"""Pins the CPU oracle (oracle/phn_oracle.c) before anything else trusts it.

(a) against the reference's own shipped golden label files (tests/golden/ref_labels.json);
(b) bit-for-bit against outputs of the reference's own sources compiled by oracle/Makefile
    (tests/golden/ref_run_*.npz, produced by tests/golden/make_golden.py);
(c) live against oracle/_ref/phnrec_ref when it is present.
"""
import numpy as np
import pytest

from conftest import ALL_MODELS, audio_bytes, model_dir, ref_run

RUNS = [("PHN_CZ_SPDAT_LCRC_N1500", "test.raw"), ("PHN_CZ_SPDAT_LCRC_N1500", "8580.wav"),
        ("PHN_EN_TIMIT_LCRC_N500", "test.raw"), ("PHN_HU_SPDAT_LCRC_N1500", "test.raw"),
        ("PHN_RU_SPDAT_LCRC_N1500", "test.raw"), ("PHN_ES", "8580.wav"), ("PHN_ES", "es.wav")]


def test_alaw_table_matches_g711(orc):
    t = np.zeros(256, dtype=np.int16)
    orc.lib().orc_alaw_table(t)
    # spot values of ALawTableD5 (alaw.cpp:14-48): table = G.711 expansion / 8
    assert t[0x55 ^ 0x00] == -1 or abs(int(t[0x55])) == 1
    assert int(t.max()) == 4032 and int(t.min()) == -4032
    assert len(set(t.tolist())) == 256 - 0  # all distinct (sign x 128 magnitudes)


@pytest.mark.parametrize("golden", ["test_en.rec", "test.rec.org", "test_hu.rec", "test_ru.rec", "test.rec",
                                    "test/8580.rec", "test/test", "es.rec"])
def test_shipped_golden_labels(orc, oracle_models, ref_labels, golden):
    g = ref_labels[golden]
    m = oracle_models(g["model"])
    lab = m.recognize(audio_bytes(g["audio"]))
    got = [(int(l["start"]), int(l["end"]), m.phonemes[int(l["phn"])]) for l in lab]
    want = [(s, e, p) for s, e, p, _ in g["labels"]]
    assert got == want  # labels + boundaries exact
    sc = np.array([float(l["like"]) for l in lab])
    ws = np.array([x[3] for x in g["labels"]])
    # goldens came from another build of the reference: scores agree to ~4e-5 relative
    assert np.max(np.abs(sc - ws) / np.maximum(1.0, np.abs(ws))) < 2e-4


@pytest.mark.parametrize("model,audio", RUNS)
def test_bit_exact_vs_reference_build(orc, oracle_models, model, audio):
    r = ref_run(model, audio)
    m = oracle_models(model)
    a = audio_bytes(audio)
    mel = m.mel(a)
    assert mel.shape == r["mel"].shape
    assert np.array_equal(mel.view(np.uint32), r["mel"].view(np.uint32))
    post = m.posteriors(mel)
    rows = r["post_rows"]
    assert np.array_equal(post[rows].view(np.uint32), r["post"].view(np.uint32))
    assert orc.format_rec(m.recognize(a), m.phonemes) == str(r["rec"])
    assert orc.format_rec(m.decode(post, wp=-1.5), m.phonemes) == str(r["rec_p15"])


def test_live_reference_binary(orc, oracle_models, tmp_path):
    if not orc.have_ref():
        pytest.skip("oracle/_ref/phnrec_ref not built")
    m = oracle_models("PHN_CZ_SPDAT_LCRC_N1500")
    a = audio_bytes("test.raw")[:40000]
    (tmp_path / "a.raw").write_bytes(a)
    orc.run_ref(["-c", model_dir("PHN_CZ_SPDAT_LCRC_N1500"), "-i", tmp_path / "a.raw", "-o", tmp_path / "a.rec"])
    assert (tmp_path / "a.rec").read_text() == orc.format_rec(m.recognize(a), m.phonemes)


def test_alaw_path_equals_lin16(orc, oracle_models):
    # test.raw is exactly A-law-decoded audio: re-encode losslessly, decode with -w alaw
    m = oracle_models("PHN_CZ_SPDAT_LCRC_N1500")
    raw = np.frombuffer(audio_bytes("test.raw"), dtype=np.int16)
    t = np.zeros(256, dtype=np.int16)
    orc.lib().orc_alaw_table(t)
    inv = {int(v) * 8: i for i, v in enumerate(t)}
    enc = np.array([inv[int(s)] for s in raw], dtype=np.uint8)
    assert np.array_equal(m.mel(enc.tobytes(), fmt="alaw").view(np.uint32), m.mel(raw.tobytes()).view(np.uint32))


def test_stc_clamp_equals_fifo(orc):
    rng = np.random.default_rng(0)
    for T in (1, 3, 14, 15, 16, 31, 77):
        mel = rng.standard_normal((T, 15)).astype(np.float32)
        ctx = np.zeros((T, 15, 31), dtype=np.float32)
        orc.lib().orc_stc_fifo(mel, T, 15, ctx)
        idx = np.clip(np.arange(T)[:, None] - 15 + np.arange(31)[None, :], 0, T - 1)
        want = mel[idx].transpose(0, 2, 1)
        assert np.array_equal(ctx, want)


def test_optional_front_end_arithmetic_equals_reference_binary(orc, tmp_path):
    """Rows W2 / M2 / M3 with the switches no shipped config sets: dc_shift, scale (srec.cpp:780-788), z_mean_source,
    preem_coef (melbanks.cpp:111-149), framenorm shift / min_floor (srec.cpp:1594-1620).  The oracle on an edited copy of
    the model directory reproduces, bit for bit, what the reference binary made of the same directory (fixture)."""
    from conftest import front_end_variants, variant_model_dir, ref_run
    n = 0
    for name, model, audio, nbytes, edits, mel, rec in front_end_variants():
        a = audio_bytes(audio)[:nbytes]
        m = orc.Model(variant_model_dir(tmp_path / name, model, edits))
        got = m.mel(a)
        assert np.array_equal(got.view(np.uint32), mel.view(np.uint32)), name
        assert orc.format_rec(m.recognize(a), m.phonemes) == rec, name
        base = np.asarray(ref_run(model, audio)["mel"])[:mel.shape[0]]
        assert not np.array_equal(base, mel), name      # the switch really changed the features
        m.close()
        n += 1
    assert n == 6


def online_norm_cases():
    """(cols, interval, mean, var, input, reference output) from tests/golden/ref_online_norm.npz: what the reference's own
    Normalization object (norm.cpp, driven by oracle/_ref/online_ref) made of each input."""
    from conftest import GOLDEN
    z = np.load(GOLDEN / "ref_online_norm.npz")
    for k in z.files:
        if k.startswith("y"):
            cols, interval, mean, var = (int(v) for v in k[1:].split("_"))
            yield cols, interval, mean, var, z[f"x{cols}"], z[k]


def test_online_norm_restatement_equals_reference_object(orc):
    """Row N2 pinned: orc_online_norm == Normalization::ProcessFrame (norm.cpp:216-234) bit for bit, including the frame
    that completes the estimate (normalised already), estimates that never complete, and NaN/inf from a zero variance."""
    n = 0
    for cols, interval, mean, var, x, want in online_norm_cases():
        got = x.copy()
        orc.lib().orc_online_norm(got, got.shape[0], cols, interval, mean, var)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (cols, interval, mean, var)
        n += 1
    assert n >= 16


def test_logf_port_equals_glibc(orc):
    # the device logf is a port of orc_logf_port; here: port == libm over a strided sweep of (0, 1]
    bad = orc.lib().orc_logf_port_mismatches(0x00000001, 0x3F800001, 997)
    assert bad == 0


@pytest.mark.parametrize("T", [1, 2, 11, 39, 40, 41, 42, 100])
def test_decoder_short_inputs_run(orc, T):
    rng = np.random.default_rng(T)
    p = rng.random((T, 138)).astype(np.float32) + 1e-3
    p /= p.sum(1, keepdims=True)
    lab = orc.decode(orc.logf(p), 45, -4.6875)
    assert len(lab) >= 1
    assert int(lab[-1]["end"]) == T
    # NB: on flat random posteriors the partial traceback (phndec.cpp:191-234) may skip a
    # commit, so labels are ordered but not necessarily contiguous - reference behaviour.
    assert all(int(a["end"]) <= int(b["start"]) for a, b in zip(lab[:-1], lab[1:]))


def test_all_models_load(oracle_models):
    dims = {"PHN_CZ_SPDAT_LCRC_N1500": (165, 1500, 138), "PHN_HU_SPDAT_LCRC_N1500": (165, 1500, 186),
            "PHN_RU_SPDAT_LCRC_N1500": (165, 1400, 159), "PHN_EN_TIMIT_LCRC_N500": (253, 500, 120),
            "PHN_ES": (165, 1500, 186)}
    for name in ALL_MODELS:
        m = oracle_models(name)
        assert m.net_dims(0) == dims[name] and m.net_dims(1) == dims[name]
        assert m.net_dims(2) == (2 * dims[name][2], dims[name][1], dims[name][2])


@pytest.mark.parametrize("key", ["PHN_CZ_SPDAT_LCRC_N1500/test.raw", "PHN_EN_TIMIT_LCRC_N500/test.raw", "PHN_ES/es.wav"])
def test_vadalize_output_matches_reference_tool(orc, oracle_models, key):
    """the fork's `vadalize` tool (vadalize.cpp + phndecalize.cpp, built as oracle/_ref/vadalize_ref):
    "start end speech" lines for every non-{pau,int,spk} segment.  The oracle's formatter on the oracle's labels must
    reproduce the reference tool's output byte for byte (fixture tests/golden/ref_vad.json)."""
    import json
    from conftest import GOLDEN, audio_bytes
    want = json.loads((GOLDEN / "ref_vad.json").read_text())[key]
    model, audio = key.split("/")
    om = oracle_models(model)
    assert orc.format_vad(om.recognize(audio_bytes(audio)), om.phonemes) == want


def online_stream_cases():
    """tests/golden/ref_online_stream.json: the reference's own ProcessOnline / ProcessTail fed block by block by
    oracle/_ref/online_ref (generated by tests/golden/make_golden.py online)."""
    import json
    from conftest import GOLDEN
    return json.loads((GOLDEN / "ref_online_stream.json").read_text())


def test_online_path_restatement_equals_reference_objects(orc, tmp_path):
    """§8(f) rank 1 pinned: the whole-signal restatement of the online path (orc_model_recognize_online: streaming frame count,
    FrameBasedNormalization, the live normaliser, clamped context, decoder) prints what the reference's streaming objects
    print when they are fed in blocks - every block size, penalty and [onlinenorm] setting of the fixture, labels AND scores."""
    from conftest import audio_bytes, online_case_model_dir
    n = 0
    for c in online_stream_cases():
        m = orc.Model(online_case_model_dir(tmp_path, c))
        a = audio_bytes(c["audio"])[:c["nbytes"]]
        got = m.recognize_online(a, fmt=c["fmt"], wp=c["penalty"])
        assert orc.format_rec(got, m.phonemes) == c["rec"], c["name"]
        m.close()
        n += 1
    assert n >= 18


def test_other_trap_systems_restatement_equals_reference_binary(orc, tmp_path):
    """SURVEY section 8(f) rank 4 pinned: posteriors/system = 1BT, 3BT (as the reference executes it), 1BT_DCT - with and without
    the Hamming window / C0, trap lengths 21, 31, 51, 15 and 23 banks - on synthetic model directories (random nets,
    tests/conftest.py:synthetic_trap_model): the restatement's posteriors and labels are the reference binary's, bit for bit
    (tests/golden/ref_trap_systems.npz, written by oracle/_ref/phnrec_ref)."""
    from conftest import GOLDEN, TRAP_CASES, audio_bytes, synthetic_trap_model
    z = np.load(GOLDEN / "ref_trap_systems.npz")
    for name, system, seed, opt in TRAP_CASES:
        m = orc.Model(synthetic_trap_model(tmp_path / name, system, seed, **opt))
        a = audio_bytes("test.raw")[:int(z[f"nbytes_{name}"])]
        post = m.posteriors(m.mel(a))
        want = z[f"post_{name}"]
        assert post.shape == want.shape, name
        assert np.array_equal(post.view(np.uint32), want.view(np.uint32)), name
        assert orc.format_rec(m.recognize(a), m.phonemes) == str(z[f"rec_{name}"]), name
        m.close()


def test_plp_restatement_equals_reference_class(orc, tmp_path):
    """params/kind = plp (plp.cpp:38-165, dspc.cpp:275-335): the restatement's coefficients are those of the reference's own
    PLPCoefs class (tests/golden/ref_plp.npz, written by oracle/_ref/online_ref plp), bit for bit - default settings, order 8 with
    C0 and no lifter, 23 banks at 16 kHz with pre-emphasis and mean removal."""
    import sys
    from conftest import GOLDEN, ROOT, audio_bytes, variant_model_dir
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    from make_golden import PLP_CASES
    z = np.load(GOLDEN / "ref_plp.npz")
    for name, model, edits, nbytes in PLP_CASES:
        m = orc.Model(variant_model_dir(tmp_path / name, model, edits))
        got = m.mel(audio_bytes("test.raw")[:nbytes])
        assert got.shape == z[name].shape, name
        assert np.array_equal(got.view(np.uint32), z[name].view(np.uint32)), name
        m.close()
