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
