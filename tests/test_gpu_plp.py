"""GPU test of params/kind = plp (phnrec_b200/csrc/k_plp.cu; reference plp.cpp:38-165, dspc.cpp:275-335): phn_mel / `-t par`.
Expected values: the reference's own PLPCoefs class (tests/golden/ref_plp.npz, oracle/_ref/online_ref plp).  Every operation is the
reference's single fp32 rounding except powf (device: pow in double, rounded once; glibc's powf is accurate to 0.8 ulp), so the
stated bound is a few last-bit differences: |delta| <= 5e-5 * max(1, |c|) (observed 2.2e-5), and most coefficients bit-identical."""
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, audio_bytes, variant_model_dir
import phnrec_b200 as pb

pytestmark = pytest.mark.gpu

sys.path.insert(0, str(ROOT / "tests" / "golden"))
from make_golden import PLP_CASES  # noqa: E402

Z = np.load(GOLDEN / "ref_plp.npz")
PLP_TOL = 5e-5      # observed: 2.2e-5 (23 banks, pre-emphasis), 0 on most coefficients


@pytest.mark.parametrize("name,model,edits,nbytes", PLP_CASES, ids=[c[0] for c in PLP_CASES])
def test_plp_coefficients_within_stated_bound_of_reference_class(tmp_path, name, model, edits, nbytes):
    r = pb.Recognizer(variant_model_dir(tmp_path / name, model, edits), device=0)
    try:
        a = audio_bytes("test.raw")[:nbytes]
        got = r.mel([a, a[:3000]])
        want = Z[name]
        assert got[0].shape == want.shape and r.n_params == want.shape[1]
        err = np.abs(got[0] - want) / np.maximum(1.0, np.abs(want))
        assert err.max() <= PLP_TOL, err.max()
        assert (got[0].view(np.uint32) == want.view(np.uint32)).mean() >= 0.5      # (most values are the reference's bits)
        assert np.array_equal(got[1], got[0][:got[1].shape[0]])                    # frames do not depend on what follows
        with pytest.raises(pb.PhnRecError):
            r.recognize([a])          # the TRAPS nets take mel-bank energies: plp is for `-t par` only
    finally:
        r.close()


def test_plp_through_the_cli(tmp_path):
    import subprocess
    from oracle import oracle as orc   # checker: HTK reader
    name, model, edits, nbytes = PLP_CASES[0]
    cfg = variant_model_dir(tmp_path / name, model, edits)
    wav, par = tmp_path / "a.raw", tmp_path / "o.par"
    wav.write_bytes(audio_bytes("test.raw")[:nbytes])
    r = subprocess.run([str(ROOT / "phnrec_b200" / "bin" / "phnrec"), "-c", str(cfg), "-t", "par", "-i", str(wav), "-o", str(par)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = orc.read_htk(par)
    assert got.shape == Z[name].shape
    assert (np.abs(got - Z[name]) / np.maximum(1.0, np.abs(Z[name]))).max() <= PLP_TOL
