"""The oracle against golden vectors produced by the REFERENCE'S OWN code (tools/make_reference_golden.py runs
/root/reference/transfer_em on top of oracle/tf_shim in the build container and commits tests/golden/ref_*.npz).

This pins the oracle's restatement of everything the reference itself decides -- layer order / widths / shared kernels,
crop / pad / concat arithmetic, the train-step dataflow and loss constants (cgan.py:110-230), which variables every gradient
is taken for, the optimizer wiring, predict_ng_cube's tiling / crop / uint8 conversion (utils.py:68-130), the uint8 <-> float
conventions (datasets.py:157-202), warp_tensor / accuracy (debug.py) -- to that code.  TensorFlow's own kernels stay
unpinned (the shim restates them; see its header).  tests/test_gpu_golden.py holds the CUDA path to the same files."""
import os

import numpy as np
import pytest
import torch

from oracle import tem_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
PROBE_SEED = 977
NETS = ("g", "f", "dx", "dy")


PREFIXES = ["ref", "tf"]      # ref_: reference code on oracle/tf_shim (committed); tf_: reference code on real TensorFlow (when provided)


def gold(name, prefix="ref"):
    p = os.path.join(GOLD, f"{prefix}_{name}")
    if not os.path.exists(p):
        if prefix == "tf":
            pytest.skip("no real-TensorFlow golden (tools/make_reference_golden.py --real-tf needs a TensorFlow install)")
        pytest.fail(f"{p} is missing: run tools/make_reference_golden.py where /root/reference exists")
    return np.load(p)


def params_for(wf, is3d, seed, scale):
    r = np.random.default_rng(seed)
    P = {}
    for k in NETS:
        layers = O.generator_layers(wf) if k in ("g", "f") else O.discriminator_layers(wf, is3d)
        P[k] = [p * scale for p in O.init_params(layers, is3d, r)]
    P["dx"][9] = np.array([0.1], np.float32); P["dy"][9] = np.array([-0.2], np.float32)
    return P


def inputs_for(is3d, B, seed):
    r = np.random.default_rng(seed + 1)
    shape = (B,) + (74,) * (3 if is3d else 2) + (1,)
    rx = np.clip(r.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32)
    ry = np.clip(r.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)
    return rx, ry


def probes(arrs):
    r = np.random.default_rng(PROBE_SEED)
    out = []
    for a in arrs:
        a = np.asarray(a, np.float64).reshape(-1)
        out.append([np.linalg.norm(a), float(a @ r.standard_normal(a.size))])
    return np.array(out)


def on_path(k, is3d, arrs):
    """2-D discriminators: block "1" owns variables in the flat layout but is not part of the reference's Keras graph."""
    return arrs[2:] if (k in ("dx", "dy") and not is3d) else arrs


def check_weights_checksum(z, P):
    cs = np.array([[float(np.sum(np.asarray(w, np.float64))), float(np.sum(np.asarray(w, np.float64) ** 2))] for k in NETS for w in P[k]])
    np.testing.assert_allclose(cs, z["weights_checksum"], rtol=1e-12)      # the recipe regenerates the golden's weights


@pytest.mark.parametrize("prefix", PREFIXES)
def test_structure_matches_reference_builders(prefix):
    z = gold("structure.npz", prefix)
    for is3d, tag in ((True, "3d"), (False, "2d")):
        for k, layers in (("g", O.generator_layers(8)), ("dx", O.discriminator_layers(8, is3d))):
            shapes = []
            for L in on_path(k, is3d, layers):
                shapes.append(L.kernel_shape(is3d))
                if L.bias:
                    shapes.append((L.cout,))
            ref = [tuple(int(v) for v in row if v) for row in z[f"shapes_{k}_{tag}"]]
            assert shapes == ref, (k, tag)
            assert sum(int(np.prod(s)) for s in shapes) == int(z[f"count_{k}_{tag}"])
        assert int(z[f"buffer_{tag}"]) == 17 and int(z[f"outdimsize_{tag}"]) == O.generator_out_dim(74) == 40
        nd = 3 if is3d else 2
        assert tuple(z[f"gen_out_shape_{tag}"]) == (1,) + (40,) * nd + (1,)
        P = params_for(8, is3d, 3, 1.0)
        with torch.no_grad():
            lg = O.discriminator_forward([torch.tensor(p) for p in P["dx"]], torch.zeros((1,) + (40,) * nd + (1,)), 8, is3d)
        assert tuple(lg.shape) == tuple(z[f"disc_out_shape_{tag}"])
    assert list(z["raises_runtime_error_70_76_132"]) == [1, 1, 1]
    with pytest.raises(RuntimeError):
        O.OracleEM2EM(70)                                   # cgan.py:52-53
    from transfer_em_b200 import EM2EM, unet_generator     # the product's argument checks run before any device work
    for d in (70, 76, 132):
        with pytest.raises(RuntimeError):
            EM2EM(d, "golden_err")
    for d in (76, 132):
        with pytest.raises(RuntimeError):
            unet_generator(d)                               # generator.py:37-38


@pytest.mark.parametrize("prefix", PREFIXES)
def test_uint8_conventions_match_reference_functions(prefix):
    z = gold("conversions.npz", prefix)
    u = np.arange(256, dtype=np.uint8)
    for i in range(4):
        got = O.standardize_population(O.scale_tensor(u), tuple(z[f"ms_{i}"]))[:, 0]
        assert np.array_equal(got, z[f"std_{i}"]), i                      # bit-exact
    assert np.array_equal(O.to_uint8_reference(z["y_sweep"], tuple(z["ms_y"])), z["y_u8"])
    v = (O.unstandardize_population(z["y_sweep"], tuple(z["ms_y"])) + np.float32(1)) * np.float32(127.5)
    inr = (v > -1) & (v < 256)                                            # the truncating cast is only defined in range
    assert np.array_equal(v[inr].astype(np.uint8), z["y_trunc"][inr])
    r = np.random.default_rng(5)
    tensors = [r.standard_normal((9, 11, 13, 1)).astype(np.float32) * (1 + 0.1 * i) + 0.01 * i for i in range(5)]
    np.testing.assert_allclose(np.array(O.get_meanstd(tensors)), z["meanstd"], rtol=2e-6)
    for nd in (3, 2):
        w = O.warp_tensor(z[f"warp_in_{nd}"], z[f"warp_uniform_{nd}"])
        np.testing.assert_allclose(w, z[f"warp_out_{nd}"], rtol=0, atol=2e-6)
        assert np.unique(z[f"warp_out_{nd}"], return_counts=True)[1].max() > 30          # the golden does contain holes (one repeated value)
    rmse = np.sqrt(np.mean((z["warp_in_3"].astype(np.float64) - z["warp_out_3"]) ** 2))
    np.testing.assert_allclose(rmse, float(z["accuracy"]), rtol=1e-6)


@pytest.mark.parametrize("prefix", PREFIXES)
def test_predict_ng_cube_indexing_bit_exact_vs_reference(prefix):
    z = gold("predict_ng_cube.npz", prefix)
    vol = np.random.default_rng(int(z["vol_seed"])).integers(0, 256, (110, 110, 110), dtype=np.uint8)
    start, size = tuple(int(v) for v in z["start"]), tuple(int(v) for v in z["size"])
    ms_x, ms_y = tuple(z["ms_x"]), tuple(z["ms_y"])

    def exact(x):
        return np.float32(0.5) * np.asarray(x, np.float32)[:, 17:-17, 17:-17, 17:-17, :] + np.float32(0.1)
    inb, out = O.predict_ng_cube_oracle(vol, start, size, exact, ms_x, ms_y, fetch_input=True)
    assert np.array_equal(out, z["exact_out"]) and np.array_equal(inb, z["exact_in"])
    # the same request with a real generator: the oracle's forward differs from the shim's only by fp32 summation order
    P = params_for(8, True, 33, 5.0)
    P["g"][11] = (P["g"][11] * np.float32(z["g11_scale"])).astype(np.float32)
    G = [torch.tensor(p) for p in P["g"]]

    def cpu_predict(t):
        with torch.no_grad():
            return O.generator_forward(G, torch.tensor(t), 8, True).numpy()
    out2 = O.predict_ng_cube_oracle(vol, start, size, cpu_predict, ms_x, ms_y)
    diff = np.abs(out2.astype(int) - z["gen_out"].astype(int)); diff = np.minimum(diff, 256 - diff)
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


@pytest.mark.parametrize("prefix", PREFIXES)
@pytest.mark.parametrize("name", ["2d", "3d", "3d_dropout"])
def test_train_step_matches_reference_train_step(name, prefix):
    if prefix == "tf" and name == "3d_dropout":
        pytest.skip("mask injection needs the shim's Dropout layer")
    z = gold(f"train_{name}.npz", prefix)
    is3d, B, seed, scale, wf = bool(z["is3d"]), int(z["B"]), int(z["seed"]), float(z["scale"]), int(z["wf"])
    P = params_for(wf, is3d, seed, scale)
    check_weights_checksum(z, P)
    rx, ry = inputs_for(is3d, B, seed)
    masks = None
    if "dropout_keys" in z.files:
        keys = [int(k) for k in z["dropout_keys"]]
        d = O.generator_dims(74)
        names = ('g_realx', 'f_fakey', 'f_realy', 'g_fakex', 'f_realx', 'g_realy')
        masks = {nm: {'g6': O.dropout_keep_mask(keys[2 * p], (B, d['g6'], d['g6'], d['g6'], 16)),
                      'g9': O.dropout_keep_mask(keys[2 * p + 1], (B, d['g9'], d['g9'], d['g9'], 8))} for p, nm in enumerate(names)}
    orc = O.OracleEM2EM(74, is3d=is3d, wf=wf)
    orc.P = {k: [p.copy() for p in v] for k, v in P.items()}
    orc.M = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    orc.V = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    ref = O.train_step_grads(P, rx, ry, wf, is3d, masks=masks, dtype=torch.float32)
    np.testing.assert_allclose(np.array(ref.losses), z["losses"][0], rtol=2e-5, atol=1e-7)
    for k in NETS:
        g = on_path(k, is3d, ref.grads[k])
        np.testing.assert_allclose(probes(g), z[f"grad_probe_{k}"], rtol=2e-3, atol=1e-9)
        for i, gi in enumerate(g):
            if f"grad_{k}_{i}" in z.files:
                full = z[f"grad_{k}_{i}"]
            else:
                full = z[f"grad16_{k}_{i}"].astype(np.float64) * float(z[f"gradmax_{k}_{i}"])
            tol = 1e-4 if f"grad_{k}_{i}" in z.files else 1e-3
            assert np.linalg.norm(gi - full) <= tol * max(np.linalg.norm(full), 1e-30), (k, i)
    # the optimizer wiring: Adam deltas and the losses of the following steps
    for step in range(z["losses"].shape[0]):
        before = {k: [p.copy() for p in v] for k, v in orc.P.items()}
        l = orc.train_step(rx, ry, masks=masks)
        np.testing.assert_allclose(np.array(l), z["losses"][step], rtol=5e-4, atol=1e-6)
        for k in NETS:
            delta = on_path(k, is3d, [a - b for a, b in zip(orc.P[k], before[k])])
            np.testing.assert_allclose(probes(delta)[:, 0], z[f"delta_{k}_step{step + 1}"][:, 0], rtol=5e-3, atol=1e-9)
    y = orc.predict(rx)
    np.testing.assert_allclose(y[:1, ::3, ::3], z["predict_after"], rtol=0, atol=2e-3 * float(np.abs(z["predict_after"]).max()))
