"""GPU tests of the TRAPS systems no shipped model uses: posteriors/system = 1BT, 3BT, 1BT_DCT
(phnrec_b200/csrc/k_trap.cu; reference traps.cpp:222-283, 344-361, 405-433).  Synthetic model directories with random nets
(tests/conftest.py:synthetic_trap_model); the expected posteriors and labels are the REFERENCE BINARY's
(tests/golden/ref_trap_systems.npz, written by tests/golden/make_golden.py from oracle/_ref/phnrec_ref) - bit for bit."""
import numpy as np
import pytest

from conftest import GOLDEN, TRAP_CASES, audio_bytes, synthetic_trap_model
import phnrec_b200 as pb

pytestmark = pytest.mark.gpu

Z = np.load(GOLDEN / "ref_trap_systems.npz")


@pytest.mark.parametrize("name,system,seed,opt", TRAP_CASES, ids=[c[0] for c in TRAP_CASES])
def test_trap_system_equals_reference_binary(tmp_path, name, system, seed, opt):
    r = pb.Recognizer(synthetic_trap_model(tmp_path / name, system, seed, **opt), device=0)
    try:
        a = audio_bytes("test.raw")[:int(Z[f"nbytes_{name}"])]
        post = r.posteriors(r.mel([a]))[0]
        want = Z[f"post_{name}"]
        assert post.shape == want.shape
        assert np.array_equal(post.view(np.uint32), want.view(np.uint32))
        assert pb.format_rec(r.recognize([a])[0], r.phonemes) == str(Z[f"rec_{name}"])
        # a ragged batch (utterances shorter than the trajectory, one frame): batch == singletons
        utts = [a[:3000], a, a[:402], a[10000:30000]]
        for u, b in zip(utts, r.recognize(utts)):
            assert np.array_equal(b.view(np.uint8), r.recognize([u])[0].view(np.uint8))
        with pytest.raises(pb.PhnRecError):
            r.set_mlp_mode(pb.MLP_TC_F16)        # the tensor-core mode implements LCRC only
    finally:
        r.close()


@pytest.mark.parametrize("name", ["1bt_dct_len51", "1bt_len21_en", "1bt"])
def test_trap_system_streaming_equals_online_restatement(tmp_path, orc, name):
    """The streaming mode with another trap length (shift 25 / 10 instead of 15): history, right-context wait and the tail's
    warm-up rows follow the trap shift; labels == the oracle's online restatement (pinned to the reference's online objects for
    the LCRC systems, and to the reference binary for these systems' posteriors)."""
    _, system, seed, opt = next(c for c in TRAP_CASES if c[0] == name)
    cfg = synthetic_trap_model(tmp_path / name, system, seed, **opt)
    r = pb.Recognizer(cfg, device=0)
    om = orc.Model(cfg)
    try:
        r.stream_open(2)
        full = audio_bytes("test.raw")
        for n_bytes, block, sid in ((40000, 3200, 0), (2 * (om.vs + 7 * om.step), 700, 1), (9001, 1000, 0)):
            a = full[:n_bytes]
            out, pos = [], 0
            fed = bytearray()
            while True:
                blk = a[pos:pos + block]
                pos += len(blk)
                fed += blk[:len(blk) // 2 * 2]
                last = pos >= len(a)
                out.append(r.stream_push([sid], [blk], [last])[0])
                if last:
                    break
            got = np.concatenate(out)
            want = om.recognize_online(bytes(fed), fmt="lin16")
            assert pb.format_rec(got, r.phonemes) == orc.format_rec(want, om.phonemes), (name, n_bytes)
    finally:
        r.close()
        om.close()
