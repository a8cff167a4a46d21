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


# ------------------------------------------------------------------------------------------------ list mode on 1..N GPUs
REF_BIN = ROOT / "oracle" / "_ref" / "phnrec_ref"


def _ragged_list(tmp_path, n, seed, with_targets):
    """n short synthetic A-law files of ragged length + a list file (the scaling test's input)."""
    import sys
    sys.path.insert(0, str(ROOT))
    from tools.synth_host import synth_audio
    a = synth_audio(16000, n, seed=seed, fmt="alaw", fs=8000)
    d = tmp_path / "wav"
    d.mkdir(exist_ok=True)
    lines = []
    for i in range(n):
        f = d / f"u{i:03d}.raw"
        f.write_bytes(a[i].tobytes()[: 2400 + ((i * 7919) % 13600)])
        lines.append(f"{f} {tmp_path / f'u{i:03d}.lab'}" if with_targets else str(f))
    lst = tmp_path / "list.scp"
    lst.write_text("\n".join(lines) + "\n")
    return lst


def test_list_mode_mlf_equals_reference_and_does_not_depend_on_gpu_count(tmp_path):
    """`phnrec -l list -m out.mlf` over a 64-utterance ragged list (SpeechRec::ProcessFileList, srec.cpp:1246-1291): the MLF
    must be the reference binary's byte for byte (exact mode), whatever the batching (7 files per batch: ten batches, two in
    flight per GPU, handed out dynamically) and whatever the number of GPUs (all visible ones vs one)."""
    if not REF_BIN.exists():
        pytest.skip("reference binary not staged")
    model = model_dir("PHN_CZ_SPDAT_LCRC_N1500")
    lst = _ragged_list(tmp_path, 64, 77, with_targets=False)
    ref_mlf = tmp_path / "ref.mlf"
    r = subprocess.run([str(REF_BIN), "-c", str(model), "-l", str(lst), "-m", str(ref_mlf), "-w", "alaw"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    outs = {}
    for tag, env in [("one", {"PHNREC_DEVICES": "0", "PHNREC_BATCH": "7"}), ("all", {"PHNREC_DEVICES": "all", "PHNREC_BATCH": "7"}),
                     ("big", {"PHNREC_DEVICES": "all"})]:
        mlf = tmp_path / f"{tag}.mlf"
        r = run("phnrec", "-c", model, "-l", lst, "-m", mlf, "-w", "alaw", env=env)
        assert r.returncode == 0, r.stderr
        outs[tag] = mlf.read_bytes()
    assert outs["one"] == ref_mlf.read_bytes()
    assert outs["all"] == outs["one"] and outs["big"] == outs["one"]
    # the fast path: not the reference's bits, but still independent of batching and of the number of GPUs
    tc = {}
    for tag, env in [("one", {"PHNREC_DEVICES": "0", "PHNREC_BATCH": "7"}), ("all", {"PHNREC_DEVICES": "all"})]:
        mlf = tmp_path / f"tc_{tag}.mlf"
        r = run("phnrec", "-c", model, "-l", lst, "-m", mlf, "-w", "alaw", env={**env, "PHNREC_MLP": "tc"})
        assert r.returncode == 0, r.stderr
        tc[tag] = mlf.read_bytes()
    assert tc["one"] == tc["all"]
    a, b = tc["one"].decode().splitlines(), outs["one"].decode().splitlines()
    same = sum(x.split()[:3] == y.split()[:3] for x, y in zip(a, b))
    assert len(a) == len(b) or abs(len(a) - len(b)) < 0.02 * len(b)
    assert same >= 0.9 * len(b) or len(a) != len(b)


def test_list_mode_rec_files_and_first_bad_file_like_the_reference(tmp_path):
    """'source target' list lines write one .rec per file; a file that cannot be read stops the run THERE with the
    reference's message and exit code 1 - everything before it has been written, nothing after it."""
    if not REF_BIN.exists():
        pytest.skip("reference binary not staged")
    model = model_dir("PHN_CZ_SPDAT_LCRC_N1500")
    lst = _ragged_list(tmp_path, 24, 78, with_targets=True)
    lines = lst.read_text().splitlines()
    lines[17] = f"{tmp_path / 'missing.raw'} {tmp_path / 'missing.lab'}"
    lst.write_text("\n".join(lines) + "\n")
    r = run("phnrec", "-c", model, "-l", lst, "-w", "alaw", env={"PHNREC_DEVICES": "all", "PHNREC_BATCH": "5"})
    assert r.returncode == 1 and r.stderr.startswith("ERROR: Can not open waveform file"), r.stderr
    got = {p.name: p.read_text() for p in tmp_path.glob("u*.lab")}
    assert sorted(got) == [f"u{i:03d}.lab" for i in range(17)]
    for p in tmp_path.glob("u*.lab"):
        p.unlink()
    rr = subprocess.run([str(REF_BIN), "-c", str(model), "-l", str(lst), "-w", "alaw"], capture_output=True, text=True, timeout=600)
    assert rr.returncode == 1 and rr.stderr.startswith("ERROR: Can not open waveform file")
    want = {p.name: p.read_text() for p in tmp_path.glob("u*.lab")}
    assert got == want


def _live_expected(rec: str, fmt: str) -> str:
    """live_callback's text (phnrec.cpp:71-110) for the labels of a .rec body (the decoder hands the same start/end/score
    to the callback and to the label file, phndec.cpp:217-231,281-293)."""
    out = []
    for line in rec.splitlines():
        s, e, w, sc = line.split()
        s, e = int(s), int(e)
        out.append({"lab": f"{s} {e} {w} {sc}\n", "str": f" {w}\n", "strlen": f" {w}({(e - s) // 100000 + 1})\n"}[fmt])
    return "".join(out)


@pytest.mark.parametrize("name,fmt", [("cz_2000", "lab"), ("en_odd_len", "strlen"), ("cz_alaw_17_frames", "str"),
                                      ("cz_mean50", "lab"), ("cz_2000", None)])
def test_live_mode_prints_the_reference_online_labels(tmp_path, name, fmt):
    """`phnrec -a`: raw samples on stdin, cut into RunLive's 125 ms blocks (srec.cpp:1450), through the streaming API;
    the labels must be those of the reference's own online objects (ref_online_stream.json) in live_callback's formats,
    and the estimation banner appears exactly when onlinenorm/estim_interval != 0 (phnrec.cpp:289-292)."""
    import os
    from conftest import audio_bytes, online_case_model_dir
    case = next(c for c in json.loads((GOLDEN / "ref_online_stream.json").read_text()) if c["name"] == name)
    cfg = online_case_model_dir(tmp_path, case)
    a = audio_bytes(case["audio"])[:case["nbytes"]]
    args = [str(BIN / "phnrec"), "-c", str(cfg), "-a", "-w", case["fmt"]] + (["-f", fmt] if fmt else [])
    if case["penalty"] is not None:
        args += ["-p", str(case["penalty"])]
    r = subprocess.run(args, input=a, capture_output=True, env=dict(os.environ), timeout=300)
    assert r.returncode == 0, r.stderr.decode()
    banner = "Estimation of normalization parameters, please speak ...\n" if case["edits"].get("onlinenorm/estim_interval") else ""
    assert r.stdout.decode() == banner + _live_expected(case["rec"], fmt or "str")


def test_live_mode_empty_input_and_bad_format(tmp_path):
    import os
    cfg = model_dir("PHN_CZ_SPDAT_LCRC_N1500")
    r = subprocess.run([str(BIN / "phnrec"), "-c", str(cfg), "-a"], input=b"", capture_output=True, env=dict(os.environ), timeout=300)
    assert r.returncode == 0 and r.stdout == b""
    r = run("phnrec", "-c", cfg, "-a", "-f", "words")
    assert r.returncode == 1 and "Invalid output format: words" in r.stderr
