"""CPU checks of the drop-in boundary: the library builds, loads, exports every symbol declared in
include/transfer_em_b200.h, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "transfer_em_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tem_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from transfer_em_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(names), set(_lib.EXPORTED_SYMBOLS) ^ set(names)
    assert _lib.load().tem_abi_version() == _lib.ABI_VERSION


def test_config_struct_layout_matches_header():
    from transfer_em_b200 import _lib
    cfg = _lib.TemConfig()
    _lib.load().tem_default_config(ctypes.byref(cfg))
    assert (cfg.abi_version, cfg.is3d, cfg.wf, cfg.dimsize, cfg.dropout, cfg.train) == (1, 1, 8, 74, 1, 1)
    assert abs(cfg.lr - 2e-4) < 1e-9 and abs(cfg.beta1 - 0.5) < 1e-9 and abs(cfg.eps - 1e-7) < 1e-12     # cgan.py:69-73


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from transfer_em_b200 import EM2EM, _lib
    with pytest.raises(_lib.TemError):
        EM2EM(74, "nogpu")


def test_reference_argument_validation_without_gpu():
    from transfer_em_b200 import EM2EM, unet_generator
    with pytest.raises(RuntimeError):
        EM2EM(64, "x")                    # cgan.py:52-53
    with pytest.raises(RuntimeError):
        unet_generator(76)                # generator.py:37-38


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under transfer_em_b200/ may import or execute it."""
    pk = os.path.join(ROOT, "transfer_em_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "tem_oracle" not in src, f


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract keys,
    times the oracle port on the host cores and never touches a GPU."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_voxels_per_s" and d["unit"] == "voxels/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["dimsize"] == 74 and d["config"]["wf"] == 8 and "config 3" in d["config"]["workload"]


def test_tf_checkpoint_converter_packs_the_restore_format():
    """tools/convert_tf_checkpoint.py: its numpy half must write exactly the keys, sizes and variable order
    that EM2EM.make_checkpoint / restore use (flat vectors in layer order, Keras layouts; 129 480 / 181 369 parameters)."""
    import importlib.util
    import numpy as np
    from oracle import tem_oracle as O
    spec = importlib.util.spec_from_file_location("convert_tf_checkpoint", os.path.join(ROOT, "tools", "convert_tf_checkpoint.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    r = np.random.default_rng(0)
    W = {}
    for name in mod.NETS:
        layers = O.generator_layers(8) if name.startswith("generator") else O.discriminator_layers(8, True)
        W[name] = O.init_params(layers, True, r)
    m = {k: [np.full_like(a, 0.5) for a in v] for k, v in W.items()}
    data = mod.pack_checkpoint(W, m, None, step=7, wf=8, is3d=True, dimsize=74)
    assert int(data["step"]) == 7 and int(data["wf"]) == 8 and int(data["is3d"]) == 1 and int(data["dimsize"]) == 74
    for name in mod.NETS:
        n = 129480 if name.startswith("generator") else 181369
        assert data[name].shape == (n,) and data[name].dtype == np.float32
        assert data[name + "_optimizer_m"].shape == (n,) and np.all(data[name + "_optimizer_m"] == 0.5)
        assert data[name + "_optimizer_v"].shape == (n,) and not data[name + "_optimizer_v"].any()
        off = 0
        for a in W[name]:                                  # variable order and layout are preserved
            np.testing.assert_array_equal(data[name][off:off + a.size], np.asarray(a, np.float32).reshape(-1)); off += a.size
    src = open(os.path.join(ROOT, "transfer_em_b200", "cgan.py")).read()
    for key in ("_optimizer_m", "_optimizer_v", '"step"', '"wf"', '"is3d"', '"dimsize"'):
        assert key in src                                  # the reader's keys
