"""GPU tests of the streaming (online) mode: phn_stream_open / phn_stream_push (k_stream.cu) against
  * the reference's own streaming objects fed block by block (tests/golden/ref_online_stream.json, written by
    oracle/_ref/online_ref = SpeechRec::ProcessOnline / ProcessTail, srec.cpp:793-927) - labels, boundaries and scores exact;
  * the oracle's whole-signal restatement of that path (pinned to the same fixture by tests/test_oracle_pinned.py) for many
    interleaved streams with ragged block sizes;
  * the offline path, where the reference's two paths give the same result (no sentence normalisation, >= 15 frames)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, audio_bytes, model_dir, online_case_model_dir, variant_model_dir
import phnrec_b200 as pb

pytestmark = pytest.mark.gpu

CASES = json.loads((GOLDEN / "ref_online_stream.json").read_text())


def stream_one(r, audio: bytes, block: int, sid: int = 0):
    out = []
    pos = 0
    while True:
        blk = audio[pos:pos + block]
        pos += len(blk)
        last = pos >= len(audio)
        out.append(r.stream_push([sid], [blk], [last])[0])
        if last:
            break
    return np.concatenate(out) if out else np.zeros(0, dtype=pb.LABEL_DTYPE)


@pytest.mark.parametrize("case", [c for c in CASES if "bunch4" not in c["name"]], ids=lambda c: c["name"])
def test_stream_equals_reference_online_objects(tmp_path, case):
    r = pb.Recognizer(online_case_model_dir(tmp_path, case), device=0)
    try:
        r.set_wave_format(case["fmt"])
        if case["penalty"] is not None:
            r.set_penalty(case["penalty"])
        r.stream_open(3)
        a = audio_bytes(case["audio"])[:case["nbytes"]]
        got = stream_one(r, a, case["block"], sid=1)
        assert pb.format_rec(got, r.phonemes) == case["rec"]
        # a different block size, another stream slot: same labels (the result does not depend on how the audio was cut)
        got2 = stream_one(r, a, 2 * (case["block"] // 3) + 2, sid=2)
        assert np.array_equal(got.view(np.uint8), got2.view(np.uint8))
    finally:
        r.close()


def test_bunch_size_that_does_not_divide_the_trap_shift_is_refused(tmp_path):
    cfg = variant_model_dir(tmp_path, "PHN_CZ_SPDAT_LCRC_N1500", {"posteriors/bunch_size": "4"})
    r = pb.Recognizer(cfg, device=0)
    try:
        with pytest.raises(pb.PhnRecError):
            r.stream_open(1)
    finally:
        r.close()


@pytest.mark.parametrize("edits", [{}, {"onlinenorm/estim_interval": "40", "onlinenorm/mean_norm": "true", "onlinenorm/var_norm": "true"}],
                         ids=["no_norm", "live_mean_var"])
def test_many_interleaved_streams_equal_the_restatement(tmp_path, orc, edits):
    """Nine streams, each fed its own utterance(s) in blocks of random size (0, 1 and odd sizes included), a random subset of
    the streams per push; streams end and - without the live normaliser - start a second utterance.  Every utterance's
    labels must equal the oracle's online restatement of the audio the stream actually received (a lin16 block of odd
    length drops its last byte, like ConvertWaveformFormat called per block)."""
    model = "PHN_CZ_SPDAT_LCRC_N1500"
    cfg = variant_model_dir(tmp_path, model, edits) if edits else model_dir(model)
    r = pb.Recognizer(cfg, device=0)
    om = orc.Model(cfg)
    rng = np.random.default_rng(5)
    base = audio_bytes("test.raw")
    n_streams = 9
    lens = [119846, 60000, 9000, 2000 + 2 * 14 * 80, 398, 400 + 160 * 20, 31000, 0, 77777]   # incl. < 15 frames, no frame at all
    utts = {s: [base[2000 * s: 2000 * s + lens[s]]] for s in range(n_streams)}
    if not edits:
        for s in (1, 2, 4):
            utts[s].append(base[30000 + 1000 * s: 30000 + 1000 * s + 25000 + 3000 * s])
    try:
        r.stream_open(n_streams)
        pos = {s: [0, 0] for s in range(n_streams)}            # [utterance index, byte position]
        fed = {s: [bytearray() for _ in utts[s]] for s in range(n_streams)}
        got = {s: [[] for _ in utts[s]] for s in range(n_streams)}
        live = set(range(n_streams))
        while live:
            pick = [s for s in sorted(live) if rng.random() < 0.7] or [min(live)]
            blocks, last = [], []
            for s in pick:
                u, p = pos[s]
                a = utts[s][u]
                k = int(rng.choice([0, 1, 333, 1000, 1600, 2001, 4000, 12000]))
                blk = a[p:p + k]
                pos[s][1] += len(blk)
                fed[s][u] += blk[:len(blk) // 2 * 2]
                blocks.append(blk)
                last.append(pos[s][1] >= len(a))
            labs = r.stream_push(pick, blocks, last)
            for s, l, e in zip(pick, labs, last):
                got[s][pos[s][0]].append(l)
                if e:
                    pos[s] = [pos[s][0] + 1, 0]
                    if pos[s][0] >= len(utts[s]):
                        live.discard(s)
        for s in range(n_streams):
            for u in range(len(utts[s])):
                g = np.concatenate(got[s][u]) if got[s][u] else np.zeros(0, dtype=pb.LABEL_DTYPE)
                want = om.recognize_online(bytes(fed[s][u]), fmt="lin16")
                assert pb.format_rec(g, r.phonemes) == orc.format_rec(want, om.phonemes), (s, u)
    finally:
        r.close()
        om.close()


def test_stream_equals_offline_where_the_reference_paths_coincide():
    """EN system: no sentence normalisation and no live normalisation, so the online and the offline path of the reference
    print the same labels for an utterance of >= 15 frames; here both run on the GPU (exact mode) and must agree bit for bit."""
    r = pb.Recognizer(model_dir("PHN_EN_TIMIT_LCRC_N500"), device=0)
    try:
        a = audio_bytes("test.raw")
        off = r.recognize([a])[0]
        r.stream_open(1)
        on = stream_one(r, a, 4000)
        assert np.array_equal(off.view(np.uint8), on.view(np.uint8))
    finally:
        r.close()


def test_stream_tensor_core_mode_close_to_exact_mode():
    r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
    try:
        a = audio_bytes("test.raw")
        r.stream_open(2)
        ex = stream_one(r, a, 4000, sid=0)
        r.set_mlp_mode(pb.MLP_TC_F16)
        tc = stream_one(r, a, 4000, sid=1)
        seg = lambda l: {(int(x["start"]), int(x["end"]), int(x["phn"])) for x in l}
        assert len(seg(ex) & seg(tc)) >= 0.9 * len(ex)
    finally:
        r.close()


def test_stream_push_argument_errors():
    r = pb.Recognizer(model_dir("PHN_CZ_SPDAT_LCRC_N1500"), device=0)
    try:
        with pytest.raises(pb.PhnRecError):
            r.stream_push([0], [b"\0" * 100])          # no streams opened
        r.stream_open(2)
        with pytest.raises(pb.PhnRecError):
            r.stream_push([2], [b"\0" * 100])
        with pytest.raises(pb.PhnRecError):
            r.stream_push([1, 1], [b"\0" * 100, b"\0" * 100])
        assert [len(x) for x in r.stream_push([0, 1], [b"", b""])] == [0, 0]
    finally:
        r.close()
