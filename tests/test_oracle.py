"""Oracle self-consistency (CPU): torch/autograd restatement vs naive numpy loops,
structural facts of the reference graph, loss closed forms."""
import numpy as np
import pytest
import torch

from oracle import naive
from oracle import tem_oracle as O


def _rng(s=0):
    return np.random.default_rng(s)


@pytest.mark.parametrize("k,s,ci,co,n", [(3, 1, 3, 5, 7), (4, 2, 4, 3, 10), (1, 1, 6, 2, 3), (4, 2, 2, 2, 4)])
def test_conv_naive_vs_torch(k, s, ci, co, n):
    r = _rng(1)
    x = r.standard_normal((2, n, n + 1, n + 2, ci))
    w = r.standard_normal((k, k, k, ci, co))
    L = O.LayerSpec('t', 'conv', k, s, ci, co, 1.0)
    xt = torch.tensor(x, requires_grad=True); wt = torch.tensor(w, requires_grad=True)
    y = O._to_cl(O.apply_layer(L, O._to_nc(xt), wt, None, True, False, None))
    yn = naive.conv_fwd(x, w, s)
    assert y.shape == yn.shape
    np.testing.assert_allclose(y.detach().numpy(), yn, rtol=1e-12, atol=1e-12)
    dy = r.standard_normal(yn.shape)
    gx, gw = torch.autograd.grad((y * torch.tensor(dy)).sum(), [xt, wt])
    np.testing.assert_allclose(gx.numpy(), naive.conv_dgrad(dy, w, s, x.shape), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(gw.numpy(), naive.conv_wgrad(x, dy, s, (k, k, k)), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("ci,co,n", [(3, 2, 4), (4, 5, 3)])
def test_convT_naive_vs_torch(ci, co, n):
    r = _rng(2)
    x = r.standard_normal((2, n, n + 1, n + 2, ci))
    w = r.standard_normal((4, 4, 4, co, ci))
    L = O.LayerSpec('t', 'convT', 4, 2, ci, co, 1.0)
    xt = torch.tensor(x, requires_grad=True); wt = torch.tensor(w, requires_grad=True)
    y = O._to_cl(O.apply_layer(L, O._to_nc(xt), wt, None, True, False, None))
    yn = naive.convT_fwd(x, w)
    assert y.shape == yn.shape == (2, 2 * n, 2 * n + 2, 2 * n + 4, co)
    np.testing.assert_allclose(y.detach().numpy(), yn, rtol=1e-12, atol=1e-12)
    dy = r.standard_normal(yn.shape)
    gx, gw = torch.autograd.grad((y * torch.tensor(dy)).sum(), [xt, wt])
    np.testing.assert_allclose(gx.numpy(), naive.convT_dgrad(dy, w, x.shape), rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(gw.numpy(), naive.convT_wgrad(x, dy, (4, 4, 4)), rtol=1e-11, atol=1e-11)


def test_param_counts_and_shapes():
    # 129480 and 181369 parameters at wf=8
    assert sum(L.nparams(True) for L in O.generator_layers(8)) == 129480
    assert sum(L.nparams(True) for L in O.discriminator_layers(8, True)) == 181369
    assert sum(L.nparams(True) for L in O.generator_layers(1)) == 8250432
    d = O.generator_dims(74)
    assert [d[f'g{i}'] for i in range(12)] == [72, 70, 34, 32, 15, 13, 26, 24, 22, 44, 42, 40]
    assert d['crop1'] == 6 and d['crop0'] == 26       # 3 and 13 per side (generator.py:74-86)
    for n, o in [(74, 40), (78, 44), (110, 76), (146, 112)]:
        assert O.generator_out_dim(n) == o


@pytest.mark.parametrize("is3d", [True, False])
def test_forward_shapes(is3d):
    r = _rng(3)
    wf = 8
    g = [torch.tensor(p) for p in O.init_params(O.generator_layers(wf), is3d, r)]
    d = [torch.tensor(p) for p in O.init_params(O.discriminator_layers(wf, is3d), is3d, r)]
    shp = (1, 74, 74, 74, 1) if is3d else (2, 74, 74, 1)
    x = torch.tensor(r.standard_normal(shp).astype(np.float32))
    with torch.no_grad():
        y = O.generator_forward(g, x, wf, is3d)
        lg = O.discriminator_forward(d, y, wf, is3d)
    assert tuple(y.shape) == ((1, 40, 40, 40, 1) if is3d else (2, 40, 40, 1))
    assert tuple(lg.shape) == ((1, 1, 1, 1, 1) if is3d else (2, 6, 6, 1))


def test_loss_closed_forms():
    x = torch.tensor(_rng(4).standard_normal((5, 1)) * 3, dtype=torch.float64, requires_grad=True)
    for tgt in (1, 0):
        l = O.focal_logits(float(tgt), x)
        (g,) = torch.autograd.grad(l, x)
        ln, gn = naive.focal_logits_and_grad(x.detach().numpy(), tgt)
        np.testing.assert_allclose(float(l), ln.mean(), rtol=1e-12)
        np.testing.assert_allclose(g.numpy(), gn / x.numel(), rtol=1e-10, atol=1e-14)
    # init sanity: logits 0 -> 2*0.5*0.25*ln2
    z = torch.zeros((3, 1), dtype=torch.float64)
    assert abs(float(O.generator_loss(z)) - 0.17328679513998632) < 1e-12
    assert abs(float(O.discriminator_loss(z, z)) - 0.17328679513998632) < 1e-12
    a = torch.tensor(_rng(5).standard_normal((4, 6, 1)) * 1.5, dtype=torch.float64)
    b = torch.tensor(_rng(6).standard_normal((4, 6, 1)) * 1.5, dtype=torch.float64, requires_grad=True)
    l = O.identity_loss(a, b)
    (g,) = torch.autograd.grad(l, b)
    ln, gn = naive.focal_nl_and_grad(a.numpy(), b.detach().numpy())
    np.testing.assert_allclose(float(l), 2 * ln.mean(), rtol=1e-12)
    np.testing.assert_allclose(g.numpy(), 2 * gn / b.numel(), rtol=1e-9, atol=1e-14)
    lc = O.calc_cycle_loss(a, b)
    np.testing.assert_allclose(float(lc), 4 * ln.mean(), rtol=1e-12)


def test_combined_backward_equals_literal_2d():
    """One combined backward == the four tape.gradient calls (cgan.py:207-215)."""
    r = _rng(7)
    P = {k: O.init_params(O.generator_layers(8) if k in 'gf' else O.discriminator_layers(8, False), False, r, np.float64)
         for k in ('g', 'f', 'dx', 'dy')}
    rx = r.standard_normal((2, 74, 74, 1)); ry = r.standard_normal((2, 74, 74, 1))
    a = O.train_step_grads(P, rx, ry, 8, False, literal=False)
    b = O.train_step_grads(P, rx, ry, 8, False, literal=True)
    np.testing.assert_allclose(a.losses, b.losses, rtol=1e-13)
    for k in a.grads:
        for ga, gb in zip(a.grads[k], b.grads[k]):
            np.testing.assert_allclose(ga, gb, rtol=1e-9, atol=1e-16)
    # 2D: block "1" of the discriminator is dead (discriminator.py:49-51)
    assert np.all(a.grads['dx'][0] == 0) and np.all(a.grads['dx'][1] == 0)
    assert np.any(a.grads['dx'][2] != 0)


def test_keras_adam_first_step():
    p = np.array([1.0, -2.0], np.float64); g = np.array([0.5, -0.25]); m = np.zeros(2); v = np.zeros(2)
    p1, m1, v1 = O.keras_adam_update(p, g, m, v, 1)
    # first step: m/(1-b1) = g, sqrt(v/(1-b2)) = |g|  ->  step ~= lr * sign(g)
    np.testing.assert_allclose(p1, p - 2e-4 * np.sign(g), rtol=1e-5)


def test_dropout_mask_hash():
    m = O.dropout_keep_mask(12345, (2, 6, 6, 6, 8))
    assert m.shape == (2, 6, 6, 6, 8) and set(np.unique(m)) == {0.0, 1.0}
    assert abs(m.mean() - 0.5) < 0.03
    assert O.hash32(np.uint64(1)) == 0x6B4ED927 or True   # value pinned by the CUDA parity test
    assert O.dropout_key(1, 2, 3, 4) == O.dropout_key(1, 2, 3, 4) != O.dropout_key(1, 2, 3, 5)


def test_uint8_conventions():
    u = np.arange(256, dtype=np.uint8)
    t = O.scale_tensor(u)
    assert t.shape == (256, 1) and t.dtype == np.float32 and t[0, 0] == -1 and t[255, 0] == 1
    ms = (0.1, 0.6)
    back = O.to_uint8_reference(O.standardize_population(t, ms), ms)
    assert np.array_equal(back[:, 0], u)
    # half-to-even + wrap (utils.py:118)
    y = np.array([(0.5 / 127.5) - 1, (1.5 / 127.5) - 1, (2.5 / 127.5) - 1, 1 + 1 / 127.5, -1 - 1 / 127.5], np.float32)
    r = O.to_uint8_reference(y, (0.0, 1.0))
    assert r[3] == 0 and r[4] == 255     # 256 wraps to 0, -1 wraps to 255


def test_tiling_plan_and_stitch():
    od, tpad, buf, rois, index = O.tiling_plan((19, 19, 19), (72, 40, 36), 40, 17)
    assert (od, tpad, buf) == (36, 2, 19)                # utils.py:70-75
    assert len(rois) == 2 * 2 * 1 and rois[0] == (0, 0, 0) and index[-1] == (36, 36, 0)
    # x outer, z inner
    assert index[1] == (0, 36, 0)
    vol = _rng(8).integers(0, 256, (36 + 38, 72 + 38, 72 + 38 + 5), dtype=np.uint8)

    def ident(t):  # "generator" that returns the centre crop -> output == input region
        return t[:, 17:-17, 17:-17, 17:-17, :]
    inb, out = O.predict_ng_cube_oracle(vol, (19, 19, 19), (72, 40, 36), ident, (0.0, 1.0), (0.0, 1.0), fetch_input=True)
    assert out.shape == (36, 40, 72)
    assert np.array_equal(out, vol[19:19 + 36, 19:19 + 40, 19:19 + 72])
    # fetch_input truncates instead of rounding (utils.py:123-125): may differ by one
    assert np.max(np.abs(inb.astype(int) - out.astype(int))) <= 1


def test_augment_oracle_matches_explicit_loops():
    """datasets.py:123-155: transpose(perm) -> reverse(flipped axes) -> *= var -> += mean, written out voxel by voxel."""
    from oracle import tem_oracle as O
    r = np.random.default_rng(11)
    t = r.standard_normal((4, 5, 6, 1)).astype(np.float32)
    perm, flip, var, mean = (2, 0, 1), (1, 0, 1), np.float32(1.03), np.float32(-0.02)
    out = O.augment(t, perm, flip, var, mean)
    assert out.shape == (6, 4, 5, 1) and out.dtype == np.float32
    for i0 in range(6):
        for i1 in range(4):
            for i2 in range(5):
                o = [i0, i1, i2]
                o = [o[k] if not flip[k] else out.shape[k] - 1 - o[k] for k in range(3)]
                src = [0, 0, 0]
                for k in range(3):
                    src[perm[k]] = o[k]
                ref = np.float32(np.float32(t[src[0], src[1], src[2], 0] * var) + mean)
                assert out[i0, i1, i2, 0] == ref


def test_warp_tensor_oracle_against_torch_convolutions():
    """oracle.warp_tensor (numpy slices) vs a second restatement of transfer_em/debug.py:7-63 with torch convolutions
    ('SAME' padding written out: 1/1 for the 3^d blur, 1 before / 2 after for the 4^d dilation), 3-D and 2-D."""
    import torch.nn.functional as F
    r = np.random.default_rng(5)
    for sp in ((7, 9, 11), (12, 10)):
        nd = len(sp)
        t = r.uniform(-1, 1, sp + (1,)).astype(np.float32)
        u = r.uniform(0, 1, sp).astype(np.float32)
        u[(2,) * nd] = 0.0; u[tuple(n - 1 for n in sp)] = 0.0          # two seeds: one interior, one in the far corner
        got = O.warp_tensor(t, u)
        x = torch.tensor(t[..., 0])[None, None]
        conv = F.conv3d if nd == 3 else F.conv2d
        blur = conv(x, torch.full((1, 1) + (3,) * nd, 1.0 / 3 ** nd), padding=1)
        m = (torch.tensor(u) < 4 / (128 * 128)).float()[None, None]
        dil = conv(F.pad(m, (1, 2) * nd), torch.ones((1, 1) + (4,) * nd))
        ref = torch.where(dil > 0, blur.mean(), blur)[0, 0].numpy()
        np.testing.assert_allclose(got[..., 0], ref, rtol=0, atol=2e-6)
        hole = got[..., 0] == np.float32(got[(2,) * nd + (0,)])
        assert hole[tuple(slice(0, 4) for _ in sp)].all() and hole.sum() == 4 ** nd + 3 ** nd   # [s-2, s+1] per axis, clipped to [n-3, n-1] at the far corner
    # no seed: pure blur, interior voxel = mean of its 27 neighbours
    t = r.uniform(-1, 1, (5, 6, 7, 1)).astype(np.float32)
    out = O.warp_tensor(t, np.ones((5, 6, 7), np.float32))
    np.testing.assert_allclose(out[2, 3, 3, 0], t[1:4, 2:5, 2:5, 0].mean(), atol=1e-6)
