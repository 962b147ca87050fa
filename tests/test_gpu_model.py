"""GPU parity of the full path through the reference-shaped Python API: generator / discriminator
forward, the CycleGAN train step (losses, gradients, Adam), dropout mask injection, tiled inference."""
import os

import numpy as np
import pytest
import torch

from oracle import tem_oracle as O
from transfer_em_b200 import EM2EM, Engine, predict_ng_cube, unet_generator, discriminator
from transfer_em_b200._lib import NET_G, NET_F, NET_DX, NET_DY
from tests.gpu_helpers import rel_l2

pytestmark = pytest.mark.gpu
NETS = {'g': NET_G, 'f': NET_F, 'dx': NET_DX, 'dy': NET_DY}
# north-star tolerance: relative L2 <= 1e-2 for bf16 against the fp32 reference.  It is applied (a) per convolution on
# identical operands (tests/test_gpu_ops.py, measured <= 4e-3) and (b) end to end against the oracle evaluated at the
# values the kernels stored (bf16 activations + bf16 weight shadows).  Against the *plain* fp32 oracle a whole
# generator is 12 chained bf16-stored layers with bf16 weights: each adds ~2.5e-3 and the chain lands at ~1e-2, so the
# end-to-end fp32 comparisons use TOL_E2E = 2e-2 (one generator deep) and 3e-2 (two generators deep / logit losses).
TOL = 1e-2
TOL_E2E = 2e-2


def _params(wf, is3d, seed, scale=1.0):
    r = np.random.default_rng(seed)
    P = {}
    for k in NETS:
        layers = O.generator_layers(wf) if k in ('g', 'f') else O.discriminator_layers(wf, is3d)
        P[k] = [p * scale for p in O.init_params(layers, is3d, r)]
    return P


def _load(engine, P):
    for k, net in NETS.items():
        engine.set_weights(net, P[k])


def _tt(P, k):
    return [torch.tensor(p, dtype=torch.float32) for p in P[k]]


@pytest.mark.parametrize("is3d,wf,n,B", [(True, 8, 74, 2), (True, 8, 78, 1), (False, 8, 74, 3), (True, 4, 74, 1), (True, 16, 74, 1),
                                          (True, 2, 74, 1),    # wf=2: wide tcgen05 kernel incl. its two-source (crop-and-concat) passes
                                          (True, 8, 110, 1), (True, 1, 110, 1)])   # BASELINE config 4 geometry: n = 110, wf = 1
def test_generator_forward(is3d, wf, n, B):
    # weights 5x the init scale so that activations are O(0.1..1) and the comparison is meaningful
    P = _params(wf, is3d, 11, scale=5.0)
    eng = Engine(dimsize=max(n, 74), is3d=is3d, wf=wf, max_batch=B, train=False)
    _load(eng, P)
    r = np.random.default_rng(12)
    shape = (B,) + (n,) * (3 if is3d else 2) + (1,)
    x = r.standard_normal(shape).astype(np.float32)
    y = eng.gen_forward(NET_G, x)
    acts = {}
    with torch.no_grad():
        ref = O.generator_forward(_tt(P, 'g'), torch.tensor(x), wf, is3d, acts=acts).numpy()
        refq = O.generator_forward(_tt(P, 'g'), torch.tensor(x), wf, is3d, quant=O.bf16_round, qweights=True).numpy()
    assert y.shape == ref.shape == (B,) + (n - 34,) * (3 if is3d else 2) + (1,)
    for li in range(11):
        a = eng.last_activation(NET_G, li).reshape(acts[f'g{li}'].shape)
        assert rel_l2(a, acts[f'g{li}'].numpy()) < TOL_E2E, f"layer g{li}"
    assert rel_l2(y, ref) < TOL_E2E
    assert rel_l2(y, refq) < 8e-3          # vs an oracle that rounds activations to bf16 like the kernels


def test_generator_uint8_input_matches_float_path():
    P = _params(8, True, 13, 5.0)
    eng = Engine(dimsize=74, max_batch=1, train=False)
    _load(eng, P)
    u = np.random.default_rng(14).integers(0, 256, (1, 74, 74, 74, 1), dtype=np.uint8)
    ms = (0.02, 0.55)
    y8 = eng.gen_forward(NET_G, u, meanstd=ms)
    yf = eng.gen_forward(NET_G, O.standardize_population(O.scale_tensor(u[..., 0]), ms))
    assert np.array_equal(y8, yf)            # fused standardise is bit-identical to the float path


@pytest.mark.parametrize("is3d,B", [(True, 2), (False, 2)])
def test_discriminator_forward(is3d, B):
    P = _params(8, is3d, 15, 5.0)
    P['dx'][9] = np.array([0.25], np.float32)     # non-zero bias
    eng = Engine(dimsize=74, is3d=is3d, max_batch=B, train=False)
    _load(eng, P)
    x = np.random.default_rng(16).standard_normal((B,) + (40,) * (3 if is3d else 2) + (1,)).astype(np.float32)
    lg = eng.disc_forward(NET_DX, x)
    with torch.no_grad():
        ref = O.discriminator_forward(_tt(P, 'dx'), torch.tensor(x), 8, is3d).numpy()
    assert lg.shape == ref.shape == ((B, 1, 1, 1, 1) if is3d else (B, 6, 6, 1))
    assert rel_l2(lg, ref) < TOL_E2E


def test_api_errors_match_reference():
    with pytest.raises(RuntimeError):
        EM2EM(70, "t")                 # cgan.py:52-53
    with pytest.raises(RuntimeError):
        unet_generator(76)             # generator.py:37-38
    m, od = unet_generator(74)
    assert od == 40 and m.count_params() == 129480
    d = discriminator()
    assert d.count_params() == 181369
    shapes = [s for _, _, s in m.variable_info()]
    assert shapes[0] == (3, 3, 3, 1, 8) and shapes[6] == (4, 4, 4, 16, 32) and shapes[11] == (3, 3, 3, 16, 1)


def _train_case(is3d, B, dropout, seed, loss_mode='focal', scale=2.0, wf=8):
    # weights at 2x the init scale: activations O(0.1-1) but D logits not saturated (at 4x the focal-loss derivative
    # (1-p)^2 amplifies a 1e-3 logit difference into a 10 % change of the whole adversarial gradient, for any implementation)
    P = _params(wf, is3d, seed, scale)
    P['dx'][9] = np.array([0.1], np.float32); P['dy'][9] = np.array([-0.2], np.float32)
    model = EM2EM(74, "parity", is3d=is3d, wf=wf, max_batch=B, dropout=dropout, loss_mode=loss_mode,
                  checkpoint_dir="/tmp/tem_parity_ckpt_none")
    _load(model.engine, P)
    r = np.random.default_rng(seed + 1)
    shape = (B,) + (74,) * (3 if is3d else 2) + (1,)
    # inputs bounded by 0.9: the reference's identity / cycle loss is focal CE of t = 1 - |a-b|/2, whose derivative
    # carries 1/(t + 1e-7): an element with |a-b| -> 2 has a gradient ~1e6 x the typical one and a jump at the clip,
    # so a handful of such elements would dominate (and randomise) any gradient comparison.  _check_step asserts the
    # regime (max |a-b| < 1.9); real training data does visit the singular region.
    rx = np.clip(r.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32)
    ry = np.clip(r.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)
    return model, P, rx, ry


def _cos(a, b):
    a = np.asarray(a, np.float64).reshape(-1); b = np.asarray(b, np.float64).reshape(-1)
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def _disc_tail_signs(engine, rx, ry, ov, P, wf, is3d, layers=(4, 5, 6, 7)):
    """The stored activations of the discriminators' small tail layers (d4..d7: 6^3 .. 1 voxels in 3-D) for the four D passes
    of the step, re-run through the deterministic forward API.  A 3-D discriminator ends in ONE voxel per sample: its 32 + 32
    tail activations gate the whole backward pass through LeakyReLU' (slope 0.09 / 0.3 vs 1), so a single pre-activation that
    lies within bf16 noise of zero (|v| < ~5e-3 of the typical magnitude: measured with tools/diag_wf4_dy.py, one flip among
    the 32 values of d6 moved every upstream gradient of that network by 20 %) decides a fifth of the gradient.  "Evaluated at
    the stored values" therefore includes the stored SIGNS of those layers (apply_layer(sign_ref=...)); the flips found
    are returned and printed so that the claim is checkable: every flipped element must be a near-zero pre-activation."""
    b = (rx.shape[1] - ov['fake_y'].shape[1]) // 2
    crop = (slice(None),) + (slice(b, -b),) * (rx.ndim - 2) + (slice(None),)
    inputs = {'dx_real': (NET_DX, 'dx', rx[crop]), 'dy_real': (NET_DY, 'dy', ry[crop]),
              'dx_fake': (NET_DX, 'dx', ov['fake_x']), 'dy_fake': (NET_DY, 'dy', ov['fake_y'])}
    signs, flips = {}, {}
    for name, (net, key, x) in inputs.items():
        x = np.ascontiguousarray(x, np.float32)
        engine.disc_forward(net, x)
        acts = {}
        with torch.no_grad():
            O.discriminator_forward(_tt(P, key), torch.tensor(x), wf, is3d, quant=O.bf16_round, qweights=True, acts=acts)
        signs[name] = {}
        for li in layers:
            ref = acts[f'd{li}'].numpy()
            got = engine.last_activation(net, li).reshape(ref.shape)
            signs[name][f'd{li}'] = got
            bad = (got > 0) != (ref > 0)
            if bad.any():
                flips[(name, li)] = int(bad.sum())
                assert np.abs(ref[bad]).max() < 2e-2 * np.abs(ref).mean(), (name, li)     # only near-zero values may flip
    if flips:
        print("LeakyReLU' sign flips in the discriminator tails (near-zero pre-activations):", flips)
    return signs, flips


def _check_step(model, P, rx, ry, is3d, masks=None, loss_mode='focal', loss_rtol=3 * TOL, wf=8, grad_tol=TOL, var_tol=8 * TOL,
                repeat=1, keys=None):
    """Forward outputs and losses are compared with the plain fp32 oracle (north-star tolerance 1e-2; the loss
    vector gets 2e-2 because the adversarial terms hang off a single logit per sample that sits behind
    21 bf16-stored layers -- against the oracle at stored values they agree to 2e-3).

    Weight gradients are compared at that tolerance with the oracle evaluated at the values the kernels
    actually stored: activations rounded to bf16 (quant=bf16_round) and the two fakes that feed the second
    generator passes taken from the GPU (override_fakes, straight-through).  Reason: LeakyReLU' is
    discontinuous at 0, so a 1e-3 relative difference in a layer input flips the sign of ~0.1 % of its
    activations and moves the L2 norm of any gradient by a few percent -- for every 16-bit implementation,
    however exact its kernels (measured: 3-9e-2 against the fp32 oracle, <= 6e-3 at identical stored values).
    Against the fp32 oracle the gradient's direction and norm are bounded instead."""
    for _ in range(repeat):        # the first step of a handle runs on one stream; from the second on, the four-stream overlap path
        if keys is not None:
            model.engine.set_dropout_keys(keys)
        losses = model.engine.train_grads(rx, ry)
    ref = O.train_step_grads(P, rx, ry, wf, is3d, masks=masks, dtype=torch.float32, loss_mode=loss_mode, keep_outputs=True)
    b = (rx.shape[1] - ref.outputs["same_x"].shape[1]) // 2
    crop = (slice(None),) + (slice(b, -b),) * (rx.ndim - 2) + (slice(None),)
    assert max(np.abs(rx[crop] - ref.outputs["same_x"]).max(), np.abs(ry[crop] - ref.outputs["same_y"]).max()) < 1.9
    for name in ("fake_y", "fake_x", "same_x", "same_y"):
        assert rel_l2(model.engine.train_output(name), ref.outputs[name]) < TOL_E2E, name
    for name in ("cycled_x", "cycled_y"):      # two generators deep (24 bf16-stored layers)
        assert rel_l2(model.engine.train_output(name), ref.outputs[name]) < 3 * TOL, name
    np.testing.assert_allclose(np.array(losses), np.array(ref.losses), rtol=loss_rtol, atol=1e-4)
    ov = {'fake_y': model.engine.train_output('fake_y'), 'fake_x': model.engine.train_output('fake_x')}
    signs, tail_flips = _disc_tail_signs(model.engine, rx, ry, ov, P, wf, is3d)
    refq = O.train_step_grads(P, rx, ry, wf, is3d, masks=masks, dtype=torch.float32, loss_mode=loss_mode,
                              quant=O.bf16_round, qweights=True, override_fakes=ov, keep_outputs=True, disc_signs=signs)
    for name in ("fake_y", "cycled_x", "fake_x", "cycled_y", "same_x", "same_y"):
        assert rel_l2(model.engine.train_output(name), refq.outputs[name]) < 8e-3, name
    np.testing.assert_allclose(np.array(losses), np.array(refq.losses), rtol=1e-2, atol=1e-5)
    worst = {}
    for k, net in NETS.items():
        got = model.engine.get_weights(net, which=1)
        flat_g = np.concatenate([g.reshape(-1) for g in got])
        flat_q = np.concatenate([g.reshape(-1) for g in refq.grads[k]])
        flat_r = np.concatenate([g.reshape(-1) for g in ref.grads[k]])
        worst[k] = rel_l2(flat_g, flat_q)
        assert worst[k] < grad_tol, f"gradient of net {k}: rel-L2 {worst[k]} vs oracle at stored values"
        assert _cos(flat_g, flat_r) > 0.995 and abs(np.linalg.norm(flat_g) / np.linalg.norm(flat_r) - 1) < 2e-2, k
        for (vname, _, _), a, b in zip(model.engine.variables(net), got, refq.grads[k]):
            if np.linalg.norm(b) > 0:
                # single small variables carry the residual sign-flip noise (stored activations still differ from the
                # oracle's by ~1e-3 through bf16 rounding-boundary cascades): 8e-2 each, 1e-2 for the whole network
                assert rel_l2(a, b) < var_tol, f"{k}/{vname}: {rel_l2(a, b)}"
            else:
                assert np.all(a == 0)
    return losses, ref, worst


@pytest.mark.parametrize("is3d,B", [(True, 1), (False, 2)])
def test_train_step_gradients_no_dropout(is3d, B):
    model, P, rx, ry = _train_case(is3d, B, False, 21)
    _check_step(model, P, rx, ry, is3d)


def test_train_step_gradients_batch2_stream_overlap():
    """The benchmarked configuration shape: 3-D, batch > 1, and the SECOND step of a handle, i.e. with the independent passes
    overlapped on the four internal streams (tem_runtime.cu train_fwd_bwd); same oracle, same tolerances."""
    model, P, rx, ry = _train_case(True, 2, False, 31)
    _check_step(model, P, rx, ry, True, repeat=2)


def test_train_step_gradients_batch2_overlap_with_dropout():
    model, P, rx, ry = _train_case(True, 2, True, 33)
    keys = [0x2003 + 104729 * i for i in range(12)]
    d = O.generator_dims(74)
    names = ('g_realx', 'f_fakey', 'f_realy', 'g_fakex', 'f_realx', 'g_realy')
    masks = {nm: {'g6': O.dropout_keep_mask(keys[2 * p], (2, d['g6'], d['g6'], d['g6'], 16)),
                  'g9': O.dropout_keep_mask(keys[2 * p + 1], (2, d['g9'], d['g9'], d['g9'], 8))} for p, nm in enumerate(names)}
    _check_step(model, P, rx, ry, True, masks=masks, repeat=2, keys=keys)


@pytest.mark.slow
def test_train_step_gradients_config4_width():
    """wf = 1 (64 / 128 / 256 channels: the BASELINE config 4 model) at n = 74, batch 1: the whole train step on the wide
    tcgen05 kernels against the oracle (about 10 TFLOP of CPU work for the two oracle evaluations)."""
    model, P, rx, ry = _train_case(True, 1, False, 35, scale=1.0, wf=1)
    _, _, worst = _check_step(model, P, rx, ry, True, wf=1)
    print("wf=1 gradient rel-L2 per network:", worst)


def test_train_step_gradients_wide_model():
    """wf = 4 (16 / 32 / 64 channels; d4 64 -> 64, g6 64 -> 32): the train step on the wide-layer kernels -- conv_upw_tc /
    conv_downw_tc (streamed weights, swizzled tiles), conv3_tcw, wgrad_tcw in both forms -- against the same oracle."""
    model, P, rx, ry = _train_case(True, 1, False, 27, scale=1.5, wf=4)
    # Round 1 measured d_y at 1.4e-2 here (0.20 on each of d0..d6) and allowed it.  tools/diag_wf4_dy.py isolated it (round 2,
    # gpurun_out/diag_wf4_dy.txt): in D_y(real_y) exactly ONE of the 32 activations of the one-voxel layer d6 has the other
    # sign on the GPU (reference value 2.4e-5 against a typical 4e-3, i.e. inside the bf16 noise of the 20 stored layers in
    # front of it); LeakyReLU' of d6 is 0.09 vs 1, so that single gate scales everything upstream.  No kernel is involved:
    # with the stored signs of the tail handed to the oracle (_disc_tail_signs) the standard tolerances hold.
    _, _, worst = _check_step(model, P, rx, ry, True, wf=4)
    print("wide-model gradient rel-L2 per network:", worst)


def test_train_step_gradients_with_injected_dropout_masks():
    model, P, rx, ry = _train_case(True, 1, True, 23)
    keys = [0x1001 + 7919 * i for i in range(12)]
    model.engine.set_dropout_keys(keys)
    d = O.generator_dims(74)
    names = ('g_realx', 'f_fakey', 'f_realy', 'g_fakex', 'f_realx', 'g_realy')     # pass order of the C ABI
    masks = {}
    for p, nm in enumerate(names):
        masks[nm] = {'g6': O.dropout_keep_mask(keys[2 * p], (1, d['g6'], d['g6'], d['g6'], 16)),
                     'g9': O.dropout_keep_mask(keys[2 * p + 1], (1, d['g9'], d['g9'], d['g9'], 8))}
    _check_step(model, P, rx, ry, True, masks=masks)


def test_train_step_lsgan_l1_mode():
    model, P, rx, ry = _train_case(False, 2, False, 25, loss_mode='lsgan_l1')
    # (x-1)^2 on O(1) logits that sit behind 21 bf16-stored layers: twice the one-pass band for the loss values
    _check_step(model, P, rx, ry, False, loss_mode='lsgan_l1')


def test_adam_update_integrated():
    """tem_apply_adam on the step's own gradients == Keras Adam (cgan.py:69-73,218-228), two consecutive steps."""
    model, P, rx, ry = _train_case(False, 2, False, 27)
    eng = model.engine
    for t in (1, 2):
        p0 = {k: eng.get_vector(n, 0) for k, n in NETS.items()}
        m0 = {k: eng.get_vector(n, 2) for k, n in NETS.items()}
        v0 = {k: eng.get_vector(n, 3) for k, n in NETS.items()}
        eng.train_grads(rx, ry)
        g = {k: eng.get_vector(n, 1) for k, n in NETS.items()}
        eng.apply_adam(1.0)
        assert eng.step == t
        for k, n in NETS.items():
            pe, me, ve = O.keras_adam_update(p0[k].astype(np.float64), g[k].astype(np.float64), m0[k].astype(np.float64), v0[k].astype(np.float64), t)
            np.testing.assert_allclose(eng.get_vector(n, 0), pe, rtol=2e-6, atol=2e-9)
            np.testing.assert_allclose(eng.get_vector(n, 2), me, rtol=1e-5, atol=1e-12)
            np.testing.assert_allclose(eng.get_vector(n, 3), ve, rtol=5e-5, atol=1e-16)   # fp32 (1 - 0.999f)


def test_loss_curve_200_steps_2d():
    """north-star: loss curves over 200 steps against the fp32 reference.

    The CycleGAN dynamics are chaotic: an fp32 oracle whose initial weights are perturbed by 1e-3 (relative)
    leaves the band of the unperturbed oracle after ~50 steps and deviates by O(100 %) afterwards (measured).
    So the band can only be held while the trajectories have not decorrelated: the first 30 steps must stay
    within the 1e-2 tolerance; afterwards the GPU run must stay as close to the reference as the reference
    stays to its own perturbed copy (time-averaged, factor 3), and must train (finite, cycle loss decreasing)."""
    scale = 1.0           # the reference's own initialisation N(0, 0.02)
    wf, is3d, B = 8, False, 2
    P = _params(wf, is3d, 41, scale)
    model = EM2EM(74, "curve", is3d=is3d, wf=wf, max_batch=B, dropout=False, checkpoint_dir="/tmp/tem_parity_ckpt_none")
    _load(model.engine, P)

    def mk(perturb):
        o = O.OracleEM2EM(74, is3d=False, wf=8)
        rr = np.random.default_rng(7)
        o.P = {k: [(p * (1 + perturb * rr.standard_normal(p.shape))).astype(np.float32) for p in v] for k, v in P.items()}
        o.M = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
        o.V = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
        return o
    o_ref, o_prt = mk(0.0), mk(1e-3)
    r = np.random.default_rng(42)
    shape = (B, 74, 74, 1)
    data = [(r.standard_normal(shape).astype(np.float32), (r.standard_normal(shape) * 0.8).astype(np.float32)) for _ in range(8)]
    G, R, Pt = [], [], []
    for step in range(200):
        bx, by = data[step % 8]
        G.append(model.train_step(bx, by)); R.append(o_ref.train_step(bx, by)); Pt.append(o_prt.train_step(bx, by))
    G, R, Pt = (np.array(a, np.float64) for a in (G, R, Pt))
    assert np.isfinite(G).all()
    rel = np.abs(G - R) / np.maximum(np.abs(R), 1e-3)
    print("loss-curve: max rel deviation over the first 30 steps", rel[:30].max(), "| mean |dev| all steps gpu",
          rel.mean(), "perturbed oracle", (np.abs(Pt - R) / np.maximum(np.abs(R), 1e-3)).mean())
    assert rel[:30].max() < TOL
    relp = np.abs(Pt - R) / np.maximum(np.abs(R), 1e-3)
    assert relp.max() > TOL                   # the reference itself cannot hold the band against a 1e-3 perturbation
    for col in (0, 1, 6):                     # long-run health: late-window mean losses comparable to the reference
        ratio = G[-50:, col].mean() / R[-50:, col].mean()
        assert 0.5 < ratio < 2.0, (col, ratio)
    assert G[-8:, 6].mean() < G[:8, 6].mean()            # the cycle loss went down


def test_loss_curve_3d_first_steps():
    model, P, rx, ry = _train_case(True, 1, False, 45)
    orc = O.OracleEM2EM(74, is3d=True, wf=8)
    orc.P = {k: [p.copy() for p in v] for k, v in P.items()}
    orc.M = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    orc.V = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    r = np.random.default_rng(46)
    for step in range(4):
        bx = r.standard_normal(rx.shape).astype(np.float32); by = r.standard_normal(rx.shape).astype(np.float32)
        np.testing.assert_allclose(np.array(model.train_step(bx, by)), np.array(orc.train_step(bx, by)), rtol=3 * TOL, atol=1e-4)


def test_loss_curve_3d_30_steps():
    """north-star band on the 3-D model: 30 consecutive train steps (Adam applied) against the fp32 oracle, every one of the
    7 reported losses within 1e-2 (the chaotic divergence of the 2-D test sets in after ~50 steps)."""
    P = _params(8, True, 47, 1.0)        # the reference's own initialisation N(0, 0.02)
    model = EM2EM(74, "curve3d", is3d=True, wf=8, max_batch=1, dropout=False, checkpoint_dir="/tmp/tem_parity_ckpt_none")
    _load(model.engine, P)
    orc = O.OracleEM2EM(74, is3d=True, wf=8)
    orc.P = {k: [p.copy() for p in v] for k, v in P.items()}
    orc.M = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    orc.V = {k: [np.zeros_like(a) for a in v] for k, v in P.items()}
    r = np.random.default_rng(48)
    shape = (1, 74, 74, 74, 1)
    data = [(np.clip(r.standard_normal(shape) * 0.4, -0.9, 0.9).astype(np.float32),
             np.clip(r.standard_normal(shape) * 0.35 + 0.1, -0.9, 0.9).astype(np.float32)) for _ in range(4)]
    G, R = [], []
    for step in range(30):
        bx, by = data[step % 4]
        G.append(model.train_step(bx, by)); R.append(orc.train_step(bx, by))
    G, R = np.array(G, np.float64), np.array(R, np.float64)
    rel = np.abs(G - R) / np.maximum(np.abs(R), 1e-3)
    print("3-D loss curve: max rel deviation over 30 steps", rel.max(), "per loss", rel.max(0))
    assert np.isfinite(G).all() and rel.max() < TOL


def test_save_model_roundtrip_and_saved_model_predict(tmp_path):
    """save_model -> meta.json (the reference's four keys, utils.py:148-167) -> predict_cube_from_saved_model (utils.py:12-38)
    gives the same stitched volume as predict_ng_cube on the live model."""
    import json
    from transfer_em_b200 import save_model, predict_cube_from_saved_model
    ck = str(tmp_path / "ck")
    model = EM2EM(74, "exp", max_batch=2, checkpoint_dir=ck)
    P = _params(8, True, 51, 5.0)
    _load(model.engine, P)
    path = model.make_checkpoint(1)
    ms_x, ms_y = (0.01, 0.57), (0.03, 0.4)
    out_dir = save_model(str(tmp_path / "saved"), path, ms_x, ms_y, size=74)
    meta = json.load(open(out_dir + "/meta.json"))
    assert set(meta) == {"buffer", "outdimsize", "meanstd_x", "meanstd_y"}
    assert meta["buffer"] == 17 and meta["outdimsize"] == 40 and np.allclose(meta["meanstd_x"], ms_x) and np.allclose(meta["meanstd_y"], ms_y)
    vol = np.random.default_rng(52).integers(0, 256, (74, 74, 110), dtype=np.uint8)
    a = predict_cube_from_saved_model(vol, (19, 19, 19), (72, 36, 36), None, out_dir)
    b = predict_ng_cube(vol, (19, 19, 19), (72, 36, 36), model, ms_x, ms_y)
    assert a.shape == (36, 36, 72) and np.array_equal(a, b)
    ia, oa = predict_cube_from_saved_model(vol, (19, 19, 19), (72, 36, 36), None, out_dir, fetch_input=True)
    assert np.array_equal(oa, a) and ia.shape == a.shape
    with pytest.raises(TypeError):
        predict_ng_cube(vol.astype(np.int16), (19, 19, 19), (72, 36, 36), model, ms_x, ms_y)     # byte kernels never reinterpret


def test_train_loop_two_epochs(tmp_path, capsys):
    """EM2EM.train (cgan.py:242-287): epoch loop, mean 7-loss line, checkpoint every check_freq epochs, sample RMSE."""
    model = EM2EM(74, "loop", is3d=False, max_batch=2, dropout=False, checkpoint_dir=str(tmp_path / "ck"))
    r = np.random.default_rng(53)
    xs = [r.standard_normal((2, 74, 74, 1)).astype(np.float32) * 0.4 for _ in range(3)]
    ys = [r.standard_normal((2, 74, 74, 1)).astype(np.float32) * 0.4 for _ in range(3)]
    model.train(xs, ys, epochs=2, debug=True, sample=xs[0], sample_gt=ys[0], check_freq=2)
    out = capsys.readouterr().out
    assert out.count("loss [g_gen_total, f_gen_total, disc_y, disc_x, g_gen_only, f_gen_only, cycle]") == 2
    assert "Saving checkpoint for epoch 2" in out and "Saving checkpoint for epoch 1 " not in out
    assert "Accuracy on sample:" in out and out.count("Time taken for epoch") == 2
    assert model.engine.step == 6 and model.latest_checkpoint().endswith("ckpt-1.npz")


def test_block_models_are_callable():
    """downsample / upsample (models/utils.py:41-137) return callable models sharing the reference's structure."""
    from oracle import naive
    from transfer_em_b200.models import downsample, upsample
    from tests.gpu_helpers import bf16r
    r = np.random.default_rng(55)
    down, skip = downsample("1", 8, 16, True, seed=3)
    assert down.trainable_variables[0] is skip.trainable_variables[0] and down.count_params() == 27 * 8 * 16 + 64 * 16 * 16
    for v in down._vars:
        v.assign(v.value * 5)
    x = bf16r(r.standard_normal((1, 14, 14, 14, 8)))
    w0, w1 = (bf16r(w) for w in down.get_weights())
    s_ref = bf16r(naive.lrelu(naive.conv_fwd(x, w0, 1), 0.3))
    d_ref = naive.lrelu(naive.conv_fwd(s_ref, w1, 2), 0.3)
    assert rel_l2(skip(x.astype(np.float32)), s_ref) < TOL and rel_l2(down(x.astype(np.float32)), d_ref) < TOL
    up = upsample("2", 16, 8, True, seed=4)
    for v in up._vars:
        v.assign(v.value * 5)
    x = bf16r(r.standard_normal((1, 7, 7, 7, 16)))
    w0, w1 = (bf16r(w) for w in up.get_weights())
    u_ref = naive.lrelu(naive.convT_fwd(bf16r(naive.lrelu(naive.conv_fwd(x, w0, 1), 0.3)), w1), 0.3)
    y = up(x.astype(np.float32))
    assert y.shape == (1, 10, 10, 10, 8) and rel_l2(y, u_ref) < TOL          # inference: no dropout
    yt = up(x.astype(np.float32), training=True, dropout_key=77)
    keep = O.dropout_keep_mask(77, y.shape)
    assert rel_l2(yt, naive.lrelu(2.0 * keep * naive.convT_fwd(bf16r(naive.lrelu(naive.conv_fwd(x, w0, 1), 0.3)), w1), 0.3)) < TOL
    d2, s2 = downsample("2d", 1, 8, False, seed=5)
    assert d2(r.standard_normal((2, 20, 20, 1)).astype(np.float32)).shape == (2, 8, 8, 8)


def test_uint8_train_inputs():
    model, P, rx, ry = _train_case(True, 1, False, 29)
    r = np.random.default_rng(30)
    ux = r.integers(0, 256, (1, 74, 74, 74, 1), dtype=np.uint8); uy = r.integers(0, 256, (1, 74, 74, 74, 1), dtype=np.uint8)
    model.meanstd_x, model.meanstd_y = (0.0, 0.58), (0.05, 0.6)
    l8 = model.engine.train_grads(ux, uy, model.meanstd_x, model.meanstd_y)
    g8 = model.engine.get_vector(NET_G, 1)
    fx = O.standardize_population(O.scale_tensor(ux[..., 0]), model.meanstd_x); fy = O.standardize_population(O.scale_tensor(uy[..., 0]), model.meanstd_y)
    lf = model.engine.train_grads(fx, fy)
    gf = model.engine.get_vector(NET_G, 1)
    np.testing.assert_allclose(np.array(l8), np.array(lf), rtol=1e-5)
    assert rel_l2(g8, gf) < 1e-4        # identical math; only atomic summation order differs


def test_checkpoint_roundtrip(tmp_path):
    model = EM2EM(74, "ck", is3d=False, max_batch=1, checkpoint_dir=str(tmp_path / "ck"))
    x = np.random.default_rng(31).standard_normal((1, 74, 74, 1)).astype(np.float32)
    model.train_step(x, x[::-1].copy())
    path = model.make_checkpoint(1)
    m2 = EM2EM(74, "ck", is3d=False, max_batch=1, checkpoint_dir=str(tmp_path / "ck"))     # auto-restore latest
    assert m2.engine.step == 1
    for net in NETS.values():
        for which in (0, 2, 3):
            assert np.array_equal(model.engine.get_vector(net, which), m2.engine.get_vector(net, which))
    assert path.endswith("ckpt-1.npz")


def test_predict_ng_cube_tiling_bit_exact_and_parity():
    """predict_ng_cube (utils.py:41-130): tile indexing / crop / uint8 conversion bit-exact; values vs oracle."""
    P = _params(8, True, 33, 5.0)
    with torch.no_grad():      # rescale the last layer so that the standardised output is O(1) like a trained model
        probe = O.generator_forward(_tt(P, 'g'), torch.randn(1, 74, 74, 74, 1, generator=torch.Generator().manual_seed(1)), 8, True)
    P['g'][11] = (P['g'][11] * (0.8 / float(probe.std()))).astype(np.float32)
    model = EM2EM(74, "tile", max_batch=5, train=False, checkpoint_dir="/tmp/tem_parity_ckpt_none")
    _load(model.engine, P)
    r = np.random.default_rng(34)
    vol = r.integers(0, 256, (36 + 38 + 3, 50 + 38, 72 + 38), dtype=np.uint8)
    start, size = (19, 19, 19), (72, 50, 36 + 3)          # ragged in y and z -> 2 x 2 x 2 tiles
    ms_x, ms_y = (0.0, 0.5774), (0.03, 0.4)
    inb, out = predict_ng_cube(vol, start, size, model, ms_x, ms_y, fetch_input=True)
    assert out.shape == (39, 50, 72) and out.dtype == np.uint8
    # (1) index/crop/uint8 math: rebuild the stitched volume from the GPU's own per-tile generator outputs
    def gpu_predict(t):
        return model.engine.gen_forward(NET_G, t)
    in_ref, out_ref_gpu = O.predict_ng_cube_oracle(vol, start, size, gpu_predict, ms_x, ms_y, fetch_input=True)
    assert np.array_equal(out, out_ref_gpu)
    assert np.array_equal(inb, in_ref)
    # (2) values vs the fp32 oracle generator: +-1 grey level from bf16 on a few voxels
    G = _tt(P, 'g')
    def cpu_predict(t):
        with torch.no_grad():
            return O.generator_forward(G, torch.tensor(t), 8, True).numpy()
    out_ref = O.predict_ng_cube_oracle(vol, start, size, cpu_predict, ms_x, ms_y)
    diff = np.abs(out.astype(int) - out_ref.astype(int)); diff = np.minimum(diff, 256 - diff)
    assert diff.max() <= 2 and (diff > 0).mean() < 0.35
    # (3) z-slab sharding: two "ranks" fill disjoint slabs whose union is the full result
    a = predict_ng_cube(vol, start, size, model, ms_x, ms_y, rank=0, world=2)
    b = predict_ng_cube(vol, start, size, model, ms_x, ms_y, rank=1, world=2)
    assert np.array_equal(np.maximum(a, b), out) and np.all((a == 0) | (b == 0))


def test_concat_layer_merged_launches_match_split_launches(tmp_path):
    """The one-launch data / weight gradients of the concat layers g10, g7 (conv_tc3.cu two-destination epilogue, wgrad_tc.cu
    two-source tiles), the input-slice-major stride-2 kernel and the tcgen05 single-channel kernels against the launches
    they replaced (debug knobs, read once per process): same MMAs, so gradients agree to fp32 summation order."""
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(__file__), "cat_worker.py")

    def run(tag, **env):
        out = str(tmp_path / f"{tag}.npz")
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, worker, out], check=True, env=e, timeout=600)
        return np.load(out)
    ref = run("merged")
    for tag, env in (("split", {"TEM_NO_DGRAD_CAT": "1", "TEM_NO_WGRAD_CAT": "1"}), ("down_v1", {"TEM_CONV_DOWN_V1": "1"})):
        got = run(tag, **env)
        np.testing.assert_allclose(got["losses"], ref["losses"], rtol=1e-6, err_msg=tag)
        for k in ("g", "f", "dx", "dy"):
            assert rel_l2(got[k], ref[k]) < 1e-5, (tag, k, rel_l2(got[k], ref[k]))
    # the CUDA-core first-layer kernels are different arithmetic (fp32 input x bf16 weights against bf16 hi + lo on the tensor
    # cores): a few bf16 roundings of a0 differ and, through the one-voxel discriminator tail, can flip a
    # LeakyReLU' gate that carries ~20 % of a gradient: losses are compared tightly, gradients by direction
    got = run("c1_cuda_cores", TEM_NO_CONV_C1TC="1", TEM_NO_WGRAD_C1TC="1")
    np.testing.assert_allclose(got["losses"], ref["losses"], rtol=2e-3)
    for k in ("g", "f", "dx", "dy"):
        assert _cos(got[k], ref[k]) > 0.95, (k, _cos(got[k], ref[k]))
