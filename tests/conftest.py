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
