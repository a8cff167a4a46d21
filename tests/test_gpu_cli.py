"""GPU tests of the process-level drop-in boundary: the `phnrec` executable (phnrec.cpp:26-259 re-implemented in
phnrec_b200/csrc/cli_phnrec.cpp on top of the C ABI) and the fork's `vadalize` personality.  The exact fp32 mode is the
CLI's default, so every output must equal the reference binary's byte for byte (fixtures: tests/golden, generated from
oracle/_ref/phnrec_ref and oracle/_ref/vadalize_ref by tests/golden/make_golden.py)."""
import json
import subprocess

import numpy as np
import pytest

from conftest import AUDIO, GOLDEN, ROOT, model_dir, ref_run

pytestmark = pytest.mark.gpu

BIN = ROOT / "phnrec_b200" / "bin"


def run(tool, *args, env=None):
    import os
    e = dict(os.environ)
    if env:
        e.update(env)
    return subprocess.run([str(BIN / tool), *map(str, args)], capture_output=True, text=True, env=e, timeout=300)


@pytest.mark.parametrize("model,audio", [("PHN_CZ_SPDAT_LCRC_N1500", "test.raw"), ("PHN_EN_TIMIT_LCRC_N500", "test.raw"),
                                         ("PHN_HU_SPDAT_LCRC_N1500", "test.raw"), ("PHN_ES", "es.wav")])
def test_phnrec_rec_file_equals_reference_binary(tmp_path, model, audio):
    out = tmp_path / "o.rec"
    r = run("phnrec", "-c", model_dir(model), "-i", AUDIO / audio, "-o", out)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ""                                   # silent unless -v
    assert out.read_text() == str(ref_run(model, audio)["rec"])


def test_phnrec_saved_stages_and_penalty_sweep(tmp_path):
    """-t par / -t post write the reference's HTK files (bit-identical matrices); -s post -p decodes from them."""
    model, audio = "PHN_CZ_SPDAT_LCRC_N1500", "test.raw"
    ref = ref_run(model, audio)
    from oracle import oracle as orc   # checker: HTK reader
    par, post, rec = tmp_path / "o.par", tmp_path / "o.post", tmp_path / "p.rec"
    assert run("phnrec", "-c", model_dir(model), "-t", "par", "-i", AUDIO / audio, "-o", par).returncode == 0
    assert run("phnrec", "-c", model_dir(model), "-t", "post", "-i", AUDIO / audio, "-o", post).returncode == 0
    assert np.array_equal(orc.read_htk(par).view(np.uint32), np.asarray(ref["mel"]).view(np.uint32))
    got = orc.read_htk(post)
    assert np.array_equal(got[ref["post_rows"]].view(np.uint32), np.asarray(ref["post"]).view(np.uint32))
    r = run("phnrec", "-c", model_dir(model), "-s", "post", "-p", "-1.5", "-i", post, "-o", rec)
    assert r.returncode == 0, r.stderr
    assert rec.read_text() == str(ref["rec_p15"])
    # -s par continues from the saved mel-banks
    rec2 = tmp_path / "q.rec"
    assert run("phnrec", "-c", model_dir(model), "-s", "par", "-i", par, "-o", rec2).returncode == 0
    assert rec2.read_text() == str(ref["rec"])


def test_phnrec_list_to_mlf_matches_shipped_golden(tmp_path):
    """-l list -m mlf: the MLF the reference ships for test/PHN_ES + 8580.wav (test/test): labels and boundaries exact,
    scores to the golden's own build-to-build spread."""
    labels = json.loads((GOLDEN / "ref_labels.json").read_text())["test/test"]
    wav = tmp_path / "8580.wav"
    wav.write_bytes((AUDIO / "8580.wav").read_bytes())
    lst, mlf = tmp_path / "l.scp", tmp_path / "o.mlf"
    lst.write_text(f"{wav}\n")
    r = run("phnrec", "-c", model_dir("PHN_ES"), "-l", lst, "-m", mlf)
    assert r.returncode == 0, r.stderr
    lines = mlf.read_text().splitlines()
    assert lines[0] == "#!MLF!#" and lines[1].startswith('"') and lines[1].endswith('8580.rec"') and lines[-1] == "."
    got = [ln.split() for ln in lines[2:-1]]
    want = [ln.split() for ln in labels["text"].splitlines() if len(ln.split()) == 4]
    assert [g[:3] for g in got] == [w[:3] for w in want]
    assert np.allclose([float(g[3]) for g in got], [float(w[3]) for w in want], rtol=1e-4, atol=1e-4)


def test_phnrec_errors_like_the_reference(tmp_path):
    r = run("phnrec", "-c", tmp_path / "nowhere", "-i", AUDIO / "test.raw", "-o", tmp_path / "o.rec")
    assert r.returncode == 1 and r.stderr.startswith("ERROR: ")
    r = run("phnrec", "-c", model_dir("PHN_CZ_SPDAT_LCRC_N1500"), "-o", tmp_path / "o.rec")
    assert r.returncode == 1 and "input file is not specified" in r.stderr


@pytest.mark.parametrize("key", ["PHN_CZ_SPDAT_LCRC_N1500/test.raw", "PHN_EN_TIMIT_LCRC_N500/test.raw", "PHN_ES/es.wav"])
def test_vadalize_equals_reference_tool(tmp_path, key):
    want = json.loads((GOLDEN / "ref_vad.json").read_text())[key]
    model, audio = key.split("/")
    out = tmp_path / "v.rec"
    r = run("vadalize", "-c", model_dir(model), "-i", AUDIO / audio, "-o", out)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == want


def test_phnrec_tensor_core_mode_switch(tmp_path):
    """PHNREC_MLP=tc (not a reference switch) selects the fast path: same segments on the golden utterance."""
    model, audio = "PHN_CZ_SPDAT_LCRC_N1500", "test.raw"
    out = tmp_path / "o.rec"
    r = run("phnrec", "-c", model_dir(model), "-i", AUDIO / audio, "-o", out, env={"PHNREC_MLP": "tc"})
    assert r.returncode == 0, r.stderr
    got = [ln.split()[:3] for ln in out.read_text().splitlines()]
    want = [ln.split()[:3] for ln in str(ref_run(model, audio)["rec"]).splitlines()]
    assert sum(g == w for g, w in zip(got, want)) >= 0.9 * len(want)


def test_model_directory_with_ascii_weights_only(tmp_path):
    """NeuralNet::Load (nn.cpp:594-621): no .nbin -> the ASCII .weights/.norms are parsed and the .nbin cache is written
    next to them; results are those of the shipped cache (EN is the system that ships the ASCII files)."""
    import shutil
    src = model_dir("PHN_EN_TIMIT_LCRC_N500")
    if not (src / "weights" / "band0.weights").exists():
        pytest.skip("ASCII model files not staged")
    dst = tmp_path / "PHN_EN"
    shutil.copytree(src, dst)
    for f in (dst / "weights").glob("*.nbin"):
        f.unlink()
    out = tmp_path / "o.rec"
    r = run("phnrec", "-c", dst, "-i", AUDIO / "test.raw", "-o", out)
    assert r.returncode == 0, r.stderr
    assert out.read_text() == str(ref_run("PHN_EN_TIMIT_LCRC_N500", "test.raw")["rec"])
    for n in ("band0", "band1", "merger"):
        assert (dst / "weights" / f"{n}.nbin").read_bytes() == (src / "weights" / f"{n}.nbin").read_bytes()
