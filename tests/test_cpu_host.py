"""CPU-side tests: the C-ABI library loads and exports everything include/*.h declares, the host
logic (config dialect, error classes, sharding + ordered gather over gloo) behaves like the reference.
No compute entry point is called here (there is no GPU in the build container)."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, model_dir

import phnrec_b200 as pb
from phnrec_b200 import api, shard


def test_library_exports_every_declared_symbol():
    hdr = (ROOT / "include" / "phnrec_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(phn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    L = api.load_library()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    bound = {n for n, _, _ in api._SYMS}
    assert declared == bound, f"python binding out of sync: {declared ^ bound}"
    assert b"sm_100a" in L.phn_version()


def test_library_carries_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", str(pb.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def _create_err(cfg_dir):
    with pytest.raises(pb.PhnRecError) as e:
        pb.Recognizer(cfg_dir)
    return e.value


def test_missing_config_is_a_file_error(tmp_path):
    err = _create_err(tmp_path / "nope")
    assert err.code == 10 and "Can not open configuration file" in err.message  # srec.cpp:256


def _copy_model(tmp_path, edit):
    src = model_dir("PHN_CZ_SPDAT_LCRC_N1500")
    tmp_path.mkdir(parents=True, exist_ok=True)
    dst = tmp_path / "m"
    subprocess.run(["cp", "-r", str(src), str(dst)], check=True)
    cfg = (dst / "config").read_text()
    (dst / "config").write_text(edit(cfg))
    return dst


def test_unknown_variable_is_rejected_with_line_number(tmp_path):
    d = _copy_model(tmp_path, lambda c: c.replace("nbanks=15", "nbanks=15\nbogus_var=1"))
    err = _create_err(d)
    assert err.code == 11 and "Unknown variable in configuration file" in err.message and "line 20" in err.message


def test_bad_value_and_bad_notation(tmp_path):
    err = _create_err(_copy_model(tmp_path / "a", lambda c: c.replace("nbanks=15", "nbanks=abc")))
    assert err.code == 12
    err = _create_err(_copy_model(tmp_path / "b", lambda c: c.replace("nbanks=15", "nbanks")))
    assert err.code == 13


def test_out_of_scope_configurations_fail_loudly(tmp_path):
    for i, (a, b) in enumerate([("system=LCRC", "system=5BT"), ("type=phndec", "type=stkint"), ("kind=fbanks", "kind=mfcc")]):
        err = _create_err(_copy_model(tmp_path / str(i), lambda c: c.replace(a, b)))
        assert err.code == 30, (a, b, err)


def test_crlf_and_comments_parse(tmp_path):
    d = _copy_model(tmp_path, lambda c: "# leading comment\r\n" + c.replace("\n", "\r\n").replace("nbanks=15", "nbanks=15# trailing"))
    err = _create_err(d) if api.load_library().phn_device_count() == 0 else None
    if err is not None:
        assert err.code == 40  # parsed fine, then: no CUDA device and no CPU path


def test_no_cpu_fallback_without_gpu():
    if api.load_library().phn_device_count() > 0:
        pytest.skip("a GPU is visible")
    err = _create_err(model_dir("PHN_CZ_SPDAT_LCRC_N1500"))
    assert err.code == 40 and "no CPU path" in err.message


def test_missing_weights_is_nn_file_error(tmp_path):
    d = _copy_model(tmp_path, lambda c: c)
    os.remove(d / "weights" / "band1.nbin")
    err = _create_err(d)
    assert err.code == 1 and "band1.nbin" in err.message


def test_product_never_touches_the_oracle():
    for p in list((ROOT / "phnrec_b200").rglob("*.py")) + list((ROOT / "phnrec_b200" / "csrc").glob("*.*")):
        if p.suffix in (".py", ".cu", ".cpp", ".h", ".cuh") or p.name == "Makefile":
            txt = p.read_text()
            assert "phn_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, p


def test_label_text_formats():
    lab = np.array([(3, 0, 69, -71.169685), (7, 69, 75, -10.34726)], dtype=pb.LABEL_DTYPE)
    names = ["a", "b", "c", "spk", "e", "f", "g0", "g"]
    assert pb.format_rec(lab, names) == "000000 6900000 spk -71.169685\n6900000 7500000 g -10.347260\n"
    assert pb.format_mlf_entry("x.rec", lab, names) == '"x.rec"\n0 6900000 spk -71.169685\n6900000 7500000 g -10.347260\n.\n'


def test_htk_roundtrip(tmp_path):
    m = np.random.default_rng(0).standard_normal((7, 15)).astype(np.float32)
    pb.write_htk(tmp_path / "x.mel", m)
    raw = (tmp_path / "x.mel").read_bytes()
    assert raw[:12] == bytes([0, 0, 0, 7, 0, 1, 0x86, 0xA0, 0, 60, 0, 6])
    assert np.array_equal(pb.read_htk(tmp_path / "x.mel"), m)


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(1)
    for W in (1, 2, 4, 8):
        for n in (0, 1, 5, 1000):
            f = rng.integers(1, 3000, size=n)
            b = shard.shard_bounds(f, W)
            assert len(b) == W and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            if n == 1000:
                tot = [int(f[s:e].sum()) for s, e in b]
                assert max(tot) - min(tot) <= 2 * 3000
    assert shard.shard_bounds([998] * 1000, 8) == [(125 * r, 125 * (r + 1)) for r in range(8)]


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from phnrec_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
frames = [int(x) for x in np.random.default_rng(7).integers(1, 2000, size=37)]
bounds = shard.shard_bounds(frames, 2)
b, e = bounds[rank]
local = [("utt%d" % u, frames[u] * 3 + 1) for u in range(b, e)]      # stand-in for per-utterance label arrays
full = shard.gather_in_list_order(local, bounds, rank, 2)
if rank == 0:
    json.dump(full, open({out!r}, "w"))
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_gather_restores_list_order(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    out = tmp_path / "full.json"
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=str(ROOT), port=port, out=str(out)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=180) == 0
    import json
    full = json.loads(out.read_text())
    frames = [int(x) for x in np.random.default_rng(7).integers(1, 2000, size=37)]
    assert full == [["utt%d" % u, frames[u] * 3 + 1] for u in range(37)]


def test_ascii_weights_give_the_shipped_nbin_byte_for_byte(tmp_path):
    """NeuralNet::LoadAscii + SaveBinary (nn.cpp:199-462, 533-592).  The EN system ships both the ASCII
    .weights/.norms and the .nbin cache the reference wrote from them: converting the ASCII pair must reproduce the
    cache exactly (sizes from the bias vectors, rows padded to 4 floats, means 0 / devs 1 in the padding)."""
    from phnrec_b200.api import load_library
    L = load_library()
    d = model_dir("PHN_EN_TIMIT_LCRC_N500")
    if not (d / "weights" / "band0.weights").exists():
        pytest.skip("ASCII model files not staged")
    for n in ("band0", "band1", "merger"):
        out = tmp_path / f"{n}.nbin"
        rc = L.phn_convert_weights(str(d / "weights" / f"{n}.weights").encode(), str(d / "norms" / f"{n}.norms").encode(), str(out).encode())
        assert rc == 0
        assert out.read_bytes() == (d / "weights" / f"{n}.nbin").read_bytes()
    # without a norms file: means 0, inverse devs 1 (NeuralNet::Alloc defaults, nn.cpp:633-682)
    out = tmp_path / "nonorm.nbin"
    assert L.phn_convert_weights(str(d / "weights" / "band0.weights").encode(), None, str(out).encode()) == 0
    raw = np.frombuffer(out.read_bytes(), dtype=np.float32)
    nin4 = 256
    assert (raw[-2 * nin4:-nin4] == 0).all() and (raw[-nin4:] == 1).all()
    # error codes: missing file, malformed file (NN_NOWEIGHTS / NN_BADWEIGHTS, nn.h:35-42)
    assert L.phn_convert_weights(str(tmp_path / "absent.weights").encode(), None, str(out).encode()) != 0
    bad = tmp_path / "bad.weights"
    bad.write_text("weigvec 4\n1\n2\n3\n")
    assert L.phn_convert_weights(str(bad).encode(), None, str(out).encode()) != 0
