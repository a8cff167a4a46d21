"""Shared fixtures.  GPU tests are marked `@pytest.mark.gpu`; everything else runs on CPU."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
MODELS = ROOT / "oracle" / "_ref" / "models"
AUDIO = ROOT / "oracle" / "_ref" / "audio"

ALL_MODELS = ["PHN_CZ_SPDAT_LCRC_N1500", "PHN_HU_SPDAT_LCRC_N1500", "PHN_RU_SPDAT_LCRC_N1500",
              "PHN_EN_TIMIT_LCRC_N500", "PHN_ES"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    # checker only: C restatement always; the reference binary + staged model data when
    # /root/reference is present (no-op on the GPU box, which uses the prebuilt oracle/_ref)
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "restatement"], check=True, capture_output=True)
    if Path("/root/reference").is_dir() and not (ROOT / "oracle" / "_ref" / "phnrec_ref").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, capture_output=True)


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as o
    return o


@pytest.fixture(scope="session")
def ref_labels():
    return json.loads((GOLDEN / "ref_labels.json").read_text())


def model_dir(name: str) -> Path:
    p = MODELS / name
    if not (p / "config").exists():
        pytest.skip(f"model data {p} not staged (run `make -C oracle ref` where /root/reference exists)")
    return p


def audio_bytes(name: str) -> bytes:
    p = AUDIO / name
    if not p.exists():
        pytest.skip(f"audio {p} not staged")
    return p.read_bytes()


def ref_run(model: str, audio: str):
    return np.load(GOLDEN / f"ref_run_{model}_{audio.replace('.', '_')}.npz")


@pytest.fixture(scope="session")
def oracle_models(orc):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = orc.Model(model_dir(name))
        return cache[name]
    return get


def patch_config(text: str, edits: dict) -> str:
    """`edits` = {"section/variable": "value"}: replace the variable's line inside its [section], or add it right behind
    the section header (a new section is appended)."""
    lines = text.splitlines()
    for key, val in edits.items():
        sec, var = key.split("/")
        out, in_sec, done, seen_sec = [], False, False, False
        for ln in lines:
            st = ln.strip()
            if st.startswith("["):
                if in_sec and not done:
                    out.append(f"{var}={val}")
                    done = True
                in_sec = st == f"[{sec}]"
                seen_sec = seen_sec or in_sec
            elif in_sec and not done and st.split("=")[0].strip() == var:
                ln = f"{var}={val}"
                done = True
            out.append(ln)
        if not done:
            if not (in_sec and seen_sec):
                out.append(f"[{sec}]")
            out.append(f"{var}={val}")
        lines = out
    return "\n".join(lines) + "\n"


def variant_model_dir(tmp_path, name: str, edits: dict) -> Path:
    """A copy of a staged model directory (weights, windows, dictionary symlinked) whose config carries `edits`."""
    src = model_dir(name)
    dst = Path(tmp_path) / (name + "_variant")
    dst.mkdir(parents=True, exist_ok=True)
    for sub in ("weights", "windows", "dicts", "norms"):
        if (src / sub).exists() and not (dst / sub).exists():
            (dst / sub).symlink_to(src / sub)
    (dst / "config").write_text(patch_config((src / "config").read_text(), edits))
    return dst


def front_end_variants():
    """Fixture tests/golden/ref_front_variants.npz: the reference binary on edited copies of shipped model directories
    (srec.cpp:780-788 dc_shift / scale, melbanks.cpp:111-149 z_mean_source / preem_coef, srec.cpp:1594-1620 framenorm)."""
    z = np.load(GOLDEN / "ref_front_variants.npz")
    meta = json.loads(str(z["meta"]))
    for i, m in enumerate(meta):
        yield m["name"], m["model"], m["audio"], m["nbytes"], m["edits"], z[f"mel{i}"], str(z[f"rec{i}"])


# ---------------------------------------------------------------------------------------------------------------------
# Synthetic model directories for the TRAPS systems no shipped model uses (posteriors/system = 1BT, 3BT, 1BT_DCT; SURVEY
# section 8(f) rank 4): random nets written as the reference's .nbin cache files (nn.cpp:464-531), a config, a phoneme list.
# Deterministic in `seed`, so the generator of the fixtures (tests/golden/make_golden.py) and the tests build the same files.
def write_nbin(path, rng, nin, nhid, nout, w_scale=1.0):
    a4 = lambda n: (n + 3) // 4 * 4
    nin4, nhid4, nout4 = a4(nin), a4(nhid), a4(nout)
    w1 = np.zeros((nhid4, nin4), dtype=np.float32)
    w2 = np.zeros((nout4, nhid4), dtype=np.float32)
    b1, b2 = np.zeros(nhid4, dtype=np.float32), np.zeros(nout4, dtype=np.float32)
    mean, dev = np.zeros(nin4, dtype=np.float32), np.ones(nin4, dtype=np.float32)
    w1[:nhid, :nin] = (rng.standard_normal((nhid, nin)) * w_scale / np.sqrt(nin)).astype(np.float32)
    w2[:nout, :nhid] = (rng.standard_normal((nout, nhid)) * 2.0 / np.sqrt(nhid)).astype(np.float32)
    b1[:nhid] = (rng.standard_normal(nhid) * 0.3).astype(np.float32)
    b2[:nout] = (rng.standard_normal(nout) * 0.3).astype(np.float32)
    mean[:nin] = (rng.standard_normal(nin) * 0.5 + 1.0).astype(np.float32)
    dev[:nin] = (1.0 / (0.5 + rng.random(nin))).astype(np.float32)
    with open(path, "wb") as f:
        np.array([2, nin, nhid, nout], dtype=np.int32).tofile(f)
        for a in (w1, w2, b1, b2, mean, dev):
            a.tofile(f)


def synthetic_trap_model(dst, system: str, seed: int, nb: int = 15, length: int = 31, hamming: bool = False, add_c0: bool = True,
                         band_hid: int = 24, band_out: int = 9, merger_hid: int = 40, n_phn: int = 7, shift: int = 6, fs: int = 8000,
                         sent_mean_norm: bool = True) -> Path:
    """A model directory for posteriors/system `system` with random nets; returns its path."""
    dst = Path(dst)
    (dst / "weights").mkdir(parents=True, exist_ok=True)
    (dst / "dicts").mkdir(exist_ok=True)
    rng = np.random.default_rng(seed)
    nout = 3 * n_phn + 3                       # 3 states per phoneme + the omitted last model, like the shipped systems
    if system == "1BT_DCT":
        merger_in = nb * shift
    else:
        tb = nb - 2 if system == "3BT" else nb
        for i in range(tb):
            write_nbin(dst / "weights" / f"band{i}.nbin", rng, length, band_hid, band_out, w_scale=2.5)
        merger_in = tb * band_out
    write_nbin(dst / "weights" / "merger.nbin", rng, merger_in, merger_hid, nout, w_scale=3.0)
    (dst / "dicts" / "phonemes").write_text("".join(f"p{i}\n" for i in range(n_phn)))
    vs, step = (200, 80) if fs == 8000 else (400, 160)
    b = lambda v: "true" if v else "false"
    (dst / "config").write_text(f"""[source]
format=lin16
sample_freq={fs}

[posteriors]
system={system}
length={length}
add_c0={b(add_c0)}
hamming={b(hamming)}
suffix=lop
bunch_size=5
softening_func=none 0 0 0

[params]
kind=fbanks
suffix=mel

[melbanks]
nbanks={nb}
lower_freq=64
higher_freq={fs // 2}
vector_size={vs}
vector_step={step}
preem_coef=0.0

[decoder]
type=phndec
num_states_per_phn=3
softening_func=log 0 0 0
wpenalty=-1.5
lm_scale=1
time_pruning=40
mode=decode

[offlinenorm]
sent_mean_norm={b(sent_mean_norm)}
sent_var_norm=false

[dirs]
tmp=$C/tmp

[models]
hmm_defs=$T/models
nstates=3
gen_from_phn_list=true

[dicts]
phoneme_list=$C/dicts/phonemes

[networks]
default=$C/net/network
omit_phn=oth

[labels]
suffix=rec
remove_path=true
""")
    return dst


TRAP_CASES = [  # (name, system, seed, options)
    ("1bt", "1BT", 11, {}),
    ("1bt_hamming", "1BT", 12, {"hamming": True}),
    ("1bt_len21_en", "1BT", 13, {"length": 21, "nb": 23, "fs": 16000, "sent_mean_norm": False}),
    ("1bt_dct", "1BT_DCT", 14, {}),
    ("1bt_dct_noc0_hamming", "1BT_DCT", 15, {"add_c0": False, "hamming": True, "shift": 5}),
    ("1bt_dct_len51", "1BT_DCT", 16, {"length": 51, "shift": 12}),
    ("3bt", "3BT", 17, {}),
]


def online_case_model_dir(tmp_path, case) -> Path:
    """The model directory of one tests/golden/ref_online_stream.json run: a shipped model, an edited copy of one, or a synthetic one."""
    if case["model"].startswith("synthetic:"):
        _, system, seed, opt = next(c for c in TRAP_CASES if c[0] == case["model"].split(":")[1])
        return synthetic_trap_model(Path(tmp_path) / case["name"], system, seed, **opt)
    return variant_model_dir(Path(tmp_path) / case["name"], case["model"], case["edits"]) if case["edits"] else model_dir(case["model"])
