#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/.

Run HERE (the build container), where /root/reference exists and
`make -C oracle ref` has produced oracle/_ref/phnrec_ref (the reference's own
sources compiled in place, canonicalised fexp.h — see oracle/Makefile).

Outputs
  ref_labels.json   the 7 golden label files the reference ships (parsed verbatim:
                    start/end frame, phoneme, printed score) + which model/audio
                    produced each.
  ref_vad.json      .rec output of the reference's `vadalize` tool (oracle/_ref/vadalize_ref) on three model/audio pairs.
  ref_online_norm.npz  Normalization::ProcessFrame (norm.cpp:216-234) run row by row over random matrices by
                    oracle/_ref/online_ref (oracle/ref_online.cpp linked with the reference's own norm.o):
                    inputs, (interval, mean_norm, var_norm) cases and the outputs.
  ref_front_variants.npz  oracle/_ref/phnrec_ref on EDITED copies of shipped model directories: the optional front-end
                    arithmetic no shipped config switches on (source/dc_shift, source/scale, melbanks/z_mean_source,
                    melbanks/preem_coef, framenorm/shift, framenorm/min_floor): `-t par` mel + the .rec text.
  ref_run_*.npz     outputs of oracle/_ref/phnrec_ref on the reference's own test
                    audio: un-normalised log-mel (`-t par`), linear posteriors
                    (`-t post`, float32, full matrix for CZ / EN, every 8th row for
                    the others) and the `.rec` text incl. the %f scores.
  ref_online_stream.json  the ONLINE path: oracle/_ref/online_ref feeds files block by block through the reference's own
                    SpeechRec::ProcessOnline / ProcessTail (srec.cpp:793-927) - several block sizes, penalties and
                    [onlinenorm] settings (on edited copies of model directories); the .rec text of each run.
                    (`python tests/golden/make_golden.py online` regenerates only this one.)
  ref_trap_systems.npz  oracle/_ref/phnrec_ref on synthetic model directories of the other TRAPS systems (1BT, 3BT, 1BT_DCT):
                    posteriors and .rec text (`python tests/golden/make_golden.py trap`).
  ref_plp.npz       PLP coefficients of the reference's PLPCoefs class (`python tests/golden/make_golden.py plp`).
Nothing in tests/ reads /root/reference at run time; only these files.
"""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent

GOLDENS = [  # (golden file, model dir under oracle/_ref/models, audio under oracle/_ref/audio, is_mlf)
    ("test_en.rec", "PHN_EN_TIMIT_LCRC_N500", "test.raw", False),
    ("test.rec.org", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", False),
    ("test_hu.rec", "PHN_HU_SPDAT_LCRC_N1500", "test.raw", False),
    ("test_ru.rec", "PHN_RU_SPDAT_LCRC_N1500", "test.raw", False),
    ("test.rec", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", False),
    ("test/8580.rec", "PHN_ES", "8580.wav", False),
    ("test/test", "PHN_ES", "8580.wav", True),
    ("es.rec", "PHN_ES", "es.wav", False),
]


ONLINE_CASES = [  # (name, model, audio, bytes used, wave format, block bytes, penalty or None, config edits)
    ("en_4000", "PHN_EN_TIMIT_LCRC_N500", "test.raw", 119846, "lin16", 4000, None, {}),
    ("en_1000", "PHN_EN_TIMIT_LCRC_N500", "test.raw", 119846, "lin16", 1000, None, {}),
    ("en_odd_len", "PHN_EN_TIMIT_LCRC_N500", "test.raw", 60001, "lin16", 2500, -3.5, {}),
    ("cz_2000", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", 119846, "lin16", 2000, None, {}),
    ("cz_mean50", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", 119846, "lin16", 2000, None,
     {"onlinenorm/estim_interval": "50", "onlinenorm/mean_norm": "true"}),
    ("cz_meanvar120", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", 90000, "lin16", 3200, -2.0,
     {"onlinenorm/estim_interval": "120", "onlinenorm/mean_norm": "true", "onlinenorm/var_norm": "true"}),
    ("cz_never_estimated", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", 30000, "lin16", 2000, None,
     {"onlinenorm/estim_interval": "500", "onlinenorm/mean_norm": "true"}),
    ("cz_alaw_short", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", 3000, "alaw", 700, None, {}),
    ("cz_alaw_35_frames", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", 200 + 34 * 80 + 13, "alaw", 500, None, {}),
    ("cz_alaw_9_frames", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", 200 + 8 * 80, "alaw", 300, None, {}),
    ("cz_alaw_17_frames", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", 200 + 16 * 80, "alaw", 400, None, {}),
    ("cz_bunch4", "PHN_CZ_SPDAT_LCRC_N1500", "test.raw", 40000, "lin16", 1600, None, {"posteriors/bunch_size": "4"}),
    ("cz_bunch4_14_frames", "PHN_CZ_SPDAT_LCRC_N1500", "8580.wav", 200 + 13 * 80, "alaw", 400, None, {"posteriors/bunch_size": "4"}),
    # synthetic models of the other TRAPS systems (tests/conftest.py): trap shifts 25 and 10 instead of 15
    ("syn_dct51", "synthetic:1bt_dct_len51", "test.raw", 40000, "lin16", 3200, None, {}),
    ("syn_dct51_8_frames", "synthetic:1bt_dct_len51", "test.raw", 2 * (200 + 7 * 80), "lin16", 700, None, {}),
    ("syn_1bt21_en", "synthetic:1bt_len21_en", "test.raw", 50000, "lin16", 2500, None, {}),
    ("syn_1bt21_en_7_frames", "synthetic:1bt_len21_en", "test.raw", 2 * (400 + 6 * 160), "lin16", 900, None, {}),
    ("hu_framenorm", "PHN_HU_SPDAT_LCRC_N1500", "test.raw", 64000, "lin16", 1600, None,
     {"framenorm/shift": "0.75", "onlinenorm/estim_interval": "30", "onlinenorm/mean_norm": "true"}),
]


def online_stream():
    """ref_online_stream.json: the reference's online path, block by block (oracle/_ref/online_ref stream)."""
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import variant_model_dir, synthetic_trap_model, TRAP_CASES  # noqa: E402
    online_ref = orc.REF_BIN.parent / "online_ref"
    out = []
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, model, audio, nbytes, fmt, block, pen, edits in ONLINE_CASES:
            if model.startswith("synthetic:"):
                _, system, seed, opt = next(c for c in TRAP_CASES if c[0] == model.split(":")[1])
                cfg = synthetic_trap_model(td / name, system, seed, **opt)
            else:
                cfg = variant_model_dir(td / name, model, edits) if edits else orc.REF_MODELS / model
            (td / "a.raw").write_bytes((orc.REF_AUDIO / audio).read_bytes()[:nbytes])
            r = subprocess.run([str(online_ref), "stream", str(cfg), str(td / "a.raw"), str(block), fmt, "-" if pen is None else str(pen),
                                str(td / "o.rec")], capture_output=True, text=True)
            assert r.returncode == 0, (name, r.returncode, r.stderr[-300:])
            rec = (td / "o.rec").read_text()
            out.append({"name": name, "model": model, "audio": audio, "nbytes": nbytes, "fmt": fmt, "block": block, "penalty": pen,
                        "edits": edits, "rec": rec})
            print("online", name, len(rec.splitlines()), "labels")
    (OUT / "ref_online_stream.json").write_text(json.dumps(out, indent=1))


PLP_CASES = [  # (name, model whose [melbanks]/[source] settings are used, config edits incl. the [plp] section, bytes of test.raw)
    ("cz_default", "PHN_CZ_SPDAT_LCRC_N1500", {"params/kind": "plp"}, 60000),
    ("cz_order8_c0_nolifter", "PHN_CZ_SPDAT_LCRC_N1500", {"params/kind": "plp", "plp/order": "8", "plp/add_c0": "true", "plp/cep_lifter": "0",
                                                          "plp/cep_scale": "1", "plp/compress_fact": "0.5"}, 40000),
    ("en_preem_zmean", "PHN_EN_TIMIT_LCRC_N500", {"params/kind": "plp", "plp/order": "13", "plp/add_c0": "true", "melbanks/preem_coef": "0.97",
                                                  "melbanks/z_mean_source": "true"}, 100000),
]


def plp():
    """ref_plp.npz: the reference's PLPCoefs class (plp.cpp, compiled out of the PHNREC_ONLY binary) driven by oracle/_ref/online_ref plp
    with the [melbanks] / [plp] settings of edited model directories; rows of coefficients."""
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import variant_model_dir  # noqa: E402
    online_ref = orc.REF_BIN.parent / "online_ref"
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, model, edits, nbytes in PLP_CASES:
            cfg = variant_model_dir(td / name, model, edits)
            m = orc.Model(cfg)   # (only to read the settings the way the reference's Init would)
            cget = lambda sec, var, dflt: next((ln.split("=", 1)[1].strip() for ln in _section((cfg / "config").read_text(), sec) if ln.split("=")[0].strip() == var), dflt)
            (td / "a.raw").write_bytes((orc.REF_AUDIO / "test.raw").read_bytes()[:nbytes])
            args = [str(online_ref), "plp", str(td / "a.raw"), str(td / "o.f32"), str(m.fs), str(m.vs), str(m.step), str(m.nbanks),
                    cget("melbanks", "lower_freq", "0"), cget("melbanks", "higher_freq", "4000"), cget("melbanks", "preem_coef", "0.0"),
                    "1" if cget("melbanks", "z_mean_source", "false") == "true" else "0", cget("plp", "order", "12"),
                    cget("plp", "compress_fact", "0.3333333"), cget("plp", "cep_lifter", "22"), cget("plp", "cep_scale", "10"),
                    "1" if cget("plp", "add_c0", "false") == "true" else "0"]
            subprocess.run(args, check=True, capture_output=True)
            np_ = int(cget("plp", "order", "12")) + (1 if cget("plp", "add_c0", "false") == "true" else 0)
            out[name] = np.fromfile(td / "o.f32", dtype=np.float32).reshape(-1, np_)
            m.close()
            print("plp", name, out[name].shape)
    np.savez_compressed(OUT / "ref_plp.npz", **out)


def _section(text, sec):
    on = False
    for ln in text.splitlines():
        st = ln.strip()
        if st.startswith("["):
            on = st == f"[{sec}]"
        elif on and "=" in st:
            yield st


def trap_systems():
    """ref_trap_systems.npz: oracle/_ref/phnrec_ref on SYNTHETIC model directories (tests/conftest.py: synthetic_trap_model)
    for the TRAPS systems no shipped model uses - 1BT, 3BT, 1BT_DCT (traps.cpp:249-283, 413-433): `-t post` + .rec."""
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import synthetic_trap_model, TRAP_CASES  # noqa: E402
    out = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name, system, seed, opt in TRAP_CASES:
            cfg = synthetic_trap_model(td / name, system, seed, **opt)
            nbytes = 60000 if opt.get("fs", 8000) == 8000 else 100000
            (td / "a.raw").write_bytes((orc.REF_AUDIO / "test.raw").read_bytes()[:nbytes])
            orc.run_ref(["-c", cfg, "-t", "post", "-i", td / "a.raw", "-o", td / "o.post"])
            orc.run_ref(["-c", cfg, "-i", td / "a.raw", "-o", td / "o.rec"])
            out[f"post_{name}"] = orc.read_htk(td / "o.post")
            out[f"rec_{name}"] = np.array((td / "o.rec").read_text())
            out[f"nbytes_{name}"] = np.array(nbytes)
            print("trap", name, out[f"post_{name}"].shape, len(str(out[f"rec_{name}"]).splitlines()), "labels")
    np.savez_compressed(OUT / "ref_trap_systems.npz", **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "online":
        online_stream()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "trap":
        trap_systems()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "plp":
        plp()
        return
    plp()
    trap_systems()
    online_stream()
    labels = {}
    for g, model, audio, mlf in GOLDENS:
        txt = (REF / g).read_text()
        labels[g] = {"model": model, "audio": audio, "mlf": mlf, "text": txt,
                     "labels": [list(x) for x in orc.parse_rec(txt)]}
    (OUT / "ref_labels.json").write_text(json.dumps(labels, indent=1))

    # the fork's `vadalize` tool (vadalize.cpp + phndecalize.cpp, built as oracle/_ref/vadalize_ref): its .rec output
    vad = {}
    with tempfile.TemporaryDirectory() as td:
        for model, audio in (("PHN_CZ_SPDAT_LCRC_N1500", "test.raw"), ("PHN_EN_TIMIT_LCRC_N500", "test.raw"), ("PHN_ES", "es.wav")):
            out = Path(td) / "v.rec"
            subprocess.run([str(orc.REF_BIN.parent / "vadalize_ref"), "-c", str(orc.REF_MODELS / model), "-i", str(orc.REF_AUDIO / audio),
                            "-o", str(out)], check=True, capture_output=True)
            vad[f"{model}/{audio}"] = out.read_text()
    (OUT / "ref_vad.json").write_text(json.dumps(vad, indent=1))

    # the online normaliser (row N2): the reference's own Normalization object, driven by oracle/_ref/online_ref
    online_ref = orc.REF_BIN.parent / "online_ref"
    rng = np.random.default_rng(2)
    norm = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for cols in (15, 23):
            x = (rng.standard_normal((300, cols)) * 3 + 11).astype(np.float32)
            x[7] = 0.0                                        # a frame of digital silence (ln energy 0)
            norm[f"x{cols}"] = x
            x.tofile(td / "in.f32")
            for interval, mean, var in ((100, 1, 0), (100, 1, 1), (100, 0, 0), (1, 1, 0), (300, 1, 1), (301, 1, 1), (0, 1, 1), (5, 1, 1)):
                subprocess.run([str(online_ref), "norm", str(td / "in.f32"), "300", str(cols), str(interval), str(mean), str(var),
                                str(td / "out.f32")], check=True, capture_output=True)
                norm[f"y{cols}_{interval}_{mean}_{var}"] = np.fromfile(td / "out.f32", dtype=np.float32).reshape(300, cols)
    np.savez_compressed(OUT / "ref_online_norm.npz", **norm)
    print("ref_online_norm.npz", len(norm), "arrays")

    # optional front-end arithmetic (srec.cpp:780-788, 1594-1620; melbanks.cpp:111-149) on edited model directories
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import variant_model_dir  # noqa: E402
    variants = [
        ("dc_scale", "PHN_CZ_SPDAT_LCRC_N1500", {"source/dc_shift": "-37.5", "source/scale": "0.5"}),
        ("preem", "PHN_CZ_SPDAT_LCRC_N1500", {"melbanks/preem_coef": "0.97"}),
        ("zmean", "PHN_CZ_SPDAT_LCRC_N1500", {"melbanks/z_mean_source": "true"}),
        ("framenorm", "PHN_CZ_SPDAT_LCRC_N1500", {"framenorm/shift": "1.25", "framenorm/min_floor": "12.5"}),
        ("all_cz", "PHN_CZ_SPDAT_LCRC_N1500", {"source/dc_shift": "11", "source/scale": "1.5", "melbanks/preem_coef": "0.95",
                                               "melbanks/z_mean_source": "true", "framenorm/shift": "-0.5", "framenorm/min_floor": "9"}),
        ("en_zmean_preem", "PHN_EN_TIMIT_LCRC_N500", {"melbanks/z_mean_source": "true", "melbanks/preem_coef": "0.9"}),
    ]
    fv, meta = {}, []
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        nbytes = 48000
        (td / "a.raw").write_bytes((orc.REF_AUDIO / "test.raw").read_bytes()[:nbytes])
        for i, (name, model, edits) in enumerate(variants):
            cfg = variant_model_dir(td / name, model, edits)
            orc.run_ref(["-c", cfg, "-t", "par", "-i", td / "a.raw", "-o", td / "o.par"])
            orc.run_ref(["-c", cfg, "-i", td / "a.raw", "-o", td / "o.rec"])
            fv[f"mel{i}"] = orc.read_htk(td / "o.par")
            fv[f"rec{i}"] = np.array((td / "o.rec").read_text())
            meta.append({"name": name, "model": model, "audio": "test.raw", "nbytes": nbytes, "edits": edits})
            print("variant", name, fv[f"mel{i}"].shape, len(str(fv[f"rec{i}"]).splitlines()), "labels")
    np.savez_compressed(OUT / "ref_front_variants.npz", meta=np.array(json.dumps(meta)), **fv)

    runs = sorted({(m, a) for _, m, a, _ in GOLDENS})
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for model, audio in runs:
            cfg = orc.REF_MODELS / model
            wav = orc.REF_AUDIO / audio
            orc.run_ref(["-c", cfg, "-i", wav, "-o", td / "o.rec"])
            orc.run_ref(["-c", cfg, "-t", "par", "-i", wav, "-o", td / "o.par"])
            orc.run_ref(["-c", cfg, "-t", "post", "-i", wav, "-o", td / "o.post"])
            # decode from the saved posteriors with a non-default penalty (-s post -p)
            orc.run_ref(["-c", cfg, "-s", "post", "-p", "-1.5", "-i", td / "o.post", "-o", td / "p.rec"])
            mel = orc.read_htk(td / "o.par")
            post = orc.read_htk(td / "o.post")
            full = model in ("PHN_CZ_SPDAT_LCRC_N1500", "PHN_EN_TIMIT_LCRC_N500") and audio == "test.raw"
            rows = np.arange(post.shape[0]) if full else np.arange(0, post.shape[0], 8)
            name = f"ref_run_{model}_{audio.replace('.', '_')}.npz"
            np.savez_compressed(OUT / name, mel=mel, post_rows=rows.astype(np.int32), post=post[rows],
                                rec=np.array((td / "o.rec").read_text()),
                                rec_p15=np.array((td / "p.rec").read_text()))
            print(name, mel.shape, post.shape, "full" if full else "subsampled")


if __name__ == "__main__":
    main()
