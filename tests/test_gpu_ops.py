"""GPU parity, per op, through the C ABI: direct conv fwd / dgrad / wgrad vs the naive fp64 oracle on
identical bf16-rounded operands; loss, Adam and uint8 kernels vs the oracle formulas."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import naive
from oracle import tem_oracle as O
from transfer_em_b200 import _lib
from tests.gpu_helpers import (bf16r, conv_dgrad, conv_forward, conv_wgrad, make_desc, ptr, rel_l2, stream)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# tolerance of one bf16 output rounding (2^-9 relative) + accumulation-order noise
BF16_RTOL, BF16_ATOL = 2.0 ** -8, 2e-3


def _cuda(a, dt):
    return torch.tensor(np.asarray(a), dtype=torch.float32).to(dt).to(DEV).contiguous()


CASES = [
    # k, s, cin, cout, transposed, dims(z,y,x), is3d
    (3, 1, 8, 8, False, (9, 10, 11), True),
    (3, 1, 16, 16, False, (7, 8, 13), True),
    (3, 1, 32, 32, False, (6, 7, 9), True),
    (3, 1, 8, 16, False, (8, 8, 8), True),
    (3, 1, 16, 1, False, (8, 9, 10), True),      # g11: Cout = 1
    (4, 2, 8, 8, False, (10, 12, 14), True),     # strided downsample
    (4, 2, 32, 32, False, (6, 6, 8), True),
    (4, 2, 32, 16, True, (5, 6, 7), True),       # convT k4 s2 SAME
    (4, 2, 16, 8, True, (6, 5, 4), True),
    (1, 1, 32, 32, False, (1, 1, 1), True),      # d7
    (3, 1, 4, 2, False, (6, 6, 7), True),        # wf=32 widths (scalar-channel path)
    (3, 1, 8, 8, False, (1, 12, 13), False),     # 2-D
    (4, 2, 8, 8, False, (1, 12, 14), False),
    (4, 2, 16, 8, True, (1, 7, 6), False),
]


@pytest.mark.parametrize("k,s,cin,cout,tr,dims,is3d", CASES)
def test_conv_fwd_dgrad_wgrad(k, s, cin, cout, tr, dims, is3d):
    r = np.random.default_rng(hash((k, s, cin, cout, tr, dims)) & 0xFFFF)
    B = 2
    kk = (k, k, k) if is3d else (1, k, k)
    x = bf16r(r.standard_normal((B,) + dims + (cin,)))
    wshape = kk + ((cout, cin) if tr else (cin, cout))
    w = bf16r(r.standard_normal(wshape) * 0.2)
    slope = 0.3
    out_dt = torch.float32 if cout == 1 else torch.bfloat16
    d = make_desc(B, dims, cin, cout, k, s, tr, slope, 0, torch.bfloat16, out_dt, is3d=is3d)
    xg, wg = _cuda(x, torch.bfloat16), _cuda(w, torch.float32)
    # ---- forward
    y = conv_forward(xg, wg, d).float().cpu().numpy()
    pre = naive.convT_fwd(x, w) if tr else naive.conv_fwd(x, w, s)
    ref = naive.lrelu(pre, slope)
    assert y.shape == ref.shape
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    assert rel_l2(y, ref) < 4e-3
    # ---- data gradient, fused with LeakyReLU' of the producer's stored activation
    dy = bf16r(r.standard_normal(ref.shape))
    act = bf16r(r.standard_normal(x.shape))
    dyg, actg = _cuda(dy, torch.bfloat16), _cuda(act, torch.bfloat16)
    dx = conv_dgrad(dyg, wg, d, actg, 0.3).float().cpu().numpy()
    dref = (naive.convT_dgrad(dy, w, x.shape) if tr else naive.conv_dgrad(dy, w, s, x.shape)) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx, dref, rtol=BF16_RTOL, atol=BF16_ATOL * 4)
    assert rel_l2(dx, dref) < 4e-3
    # ---- weight gradient (fp32 accumulate, atomics)
    dw = conv_wgrad(xg, dyg, d, wshape).cpu().numpy()
    wref = naive.convT_wgrad(x, dy, kk) if tr else naive.conv_wgrad(x, dy, s, kk)
    assert rel_l2(dw, wref) < 1e-5
    np.testing.assert_allclose(dw, wref, rtol=1e-4, atol=1e-3)


def test_first_layer_uint8_fused_standardize():
    """k1: uint8 input standardised on load (datasets.py:157-163,193-202) -> conv3 -> LReLU."""
    r = np.random.default_rng(3)
    u = r.integers(0, 256, (2, 9, 10, 11, 1), dtype=np.uint8)
    ms = (0.07, 0.61)
    w = bf16r(r.standard_normal((3, 3, 3, 1, 8)) * 0.3)
    d = make_desc(2, (9, 10, 11), 1, 8, 3, 1, False, 0.3, 0, torch.uint8, torch.bfloat16, ms)
    y = conv_forward(torch.tensor(u).to(DEV), _cuda(w, torch.float32), d).float().cpu().numpy()
    xs = O.standardize_population(O.scale_tensor(u[..., 0]), ms).astype(np.float64)
    ref = naive.lrelu(naive.conv_fwd(xs, w, 1), 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    # wgrad w.r.t. the same uint8 input
    dy = bf16r(r.standard_normal(ref.shape))
    dw = conv_wgrad(torch.tensor(u).to(DEV), _cuda(dy, torch.bfloat16), d, (3, 3, 3, 1, 8)).cpu().numpy()
    assert rel_l2(dw, naive.conv_wgrad(xs, dy, 1, (3, 3, 3))) < 1e-5


def test_dropout_mask_and_forward():
    key = 0xC0FFEE11
    n = 2 * 6 * 8 * 10 * 8
    m = torch.empty(n, dtype=torch.float32, device=DEV)
    _lib.check(_lib.load().tem_dropout_mask(key, ptr(m), n, stream()))
    torch.cuda.synchronize()
    ref = O.dropout_keep_mask(key, (n,))
    assert np.array_equal(m.cpu().numpy(), ref)          # bit-exact hash parity with the oracle
    r = np.random.default_rng(5)
    x = bf16r(r.standard_normal((2, 3, 4, 5, 16)))
    w = bf16r(r.standard_normal((4, 4, 4, 8, 16)) * 0.2)
    d = make_desc(2, (3, 4, 5), 16, 8, 4, 2, True, 0.3, key)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w, torch.float32), d).float().cpu().numpy()
    pre = naive.convT_fwd(x, w)
    mask = O.dropout_keep_mask(key, pre.shape)
    ref = naive.lrelu(pre * mask * 2.0, 0.3)              # Dropout(0.5) then LeakyReLU (utils.py:134-135)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_bias_last_layer():
    r = np.random.default_rng(6)
    x = bf16r(r.standard_normal((3, 1, 1, 1, 32)))
    w = bf16r(r.standard_normal((1, 1, 1, 32, 1)))
    b = np.array([0.37], np.float32)
    d = make_desc(3, (1, 1, 1), 32, 1, 1, 1, False, 1.0, 0, torch.bfloat16, torch.float32)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w, torch.float32), d, bias=torch.tensor(b).to(DEV)).cpu().numpy()
    np.testing.assert_allclose(y, naive.conv_fwd(x, w, 1) + 0.37, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("target", [1.0, 0.0])
def test_focal_logits_kernel(target):
    r = np.random.default_rng(7)
    x = (r.standard_normal(37) * 3).astype(np.float32)
    xg = torch.tensor(x).to(DEV)
    loss = torch.zeros(1, device=DEV); grad = torch.empty(37, device=DEV)
    _lib.check(_lib.load().tem_focal_logits(ptr(xg), 37, target, 2.0, 2.0, ptr(loss), ptr(grad), stream()))
    torch.cuda.synchronize()
    l, g = naive.focal_logits_and_grad(x, int(target))
    np.testing.assert_allclose(loss.item(), 2 * l.mean(), rtol=2e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), 2 * g / 37, rtol=2e-4, atol=1e-7)
    xt = torch.tensor(x[:, None], dtype=torch.float64)
    assert abs(loss.item() - 2 * float(O.focal_logits(target, xt))) < 1e-5


def test_focal_probs_kernel():
    r = np.random.default_rng(8)
    a = (r.standard_normal(4099) * 1.2).astype(np.float32)
    b = (r.standard_normal(4099) * 1.5).astype(np.float32)
    a[:5] = b[:5]                       # exact ties: zero gradient
    ag, bg = torch.tensor(a).to(DEV), torch.tensor(b).to(DEV)
    loss = torch.zeros(1, device=DEV); grad = torch.empty(4099, device=DEV)
    _lib.check(_lib.load().tem_focal_probs(ptr(ag), ptr(bg), 4099, 2.0, 4.0, ptr(loss), ptr(grad), stream()))
    torch.cuda.synchronize()
    l, g = naive.focal_nl_and_grad(a, b)
    np.testing.assert_allclose(loss.item(), 4 * l.mean(), rtol=5e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), 4 * g / 4099, rtol=5e-4, atol=1e-8)
    ref = float(O.calc_cycle_loss(torch.tensor(a, dtype=torch.float64)[:, None], torch.tensor(b, dtype=torch.float64)[:, None]))
    assert abs(loss.item() - ref) / ref < 5e-5


def test_keras_adam_kernel():
    r = np.random.default_rng(9)
    n = 10007
    p = r.standard_normal(n).astype(np.float32); g = (r.standard_normal(n) * 1e-3).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    pg, gg, mg, vg = (torch.tensor(t).to(DEV) for t in (p, g, m, v))
    for t in (1, 2, 3):
        _lib.check(_lib.load().tem_adam(ptr(pg), ptr(gg), ptr(mg), ptr(vg), n, t, 2e-4, 0.5, 0.999, 1e-7, 1.0, stream()))
        p, m, v = O.keras_adam_update(p.astype(np.float64), g.astype(np.float64), m.astype(np.float64), v.astype(np.float64), t)
    torch.cuda.synchronize()
    np.testing.assert_allclose(pg.cpu().numpy(), p, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(mg.cpu().numpy(), m, rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(vg.cpu().numpy(), v, rtol=1e-5, atol=1e-12)


def test_uint8_conventions_bit_exact():
    """scale/standardise and (y*std+mean+1)*127.5 -> rint -> wrap must be bit-exact (utils.py:109,118)."""
    from transfer_em_b200 import datasets as D
    r = np.random.default_rng(10)
    u = r.integers(0, 256, (3, 17, 19), dtype=np.uint8)
    for ms in [(0.0, 1.0), (0.0, 0.5774), (-0.113, 0.731)]:
        t = D.scale_and_standardize(u, ms)
        ref = O.standardize_population(O.scale_tensor(u), ms)
        assert t.dtype == np.float32 and np.array_equal(t, ref)
        y = (r.standard_normal(200003) * 1.3).astype(np.float32)
        y[:9] = np.array([(0.5 / 127.5 - 1 - ms[0]) / ms[1], (1.5 / 127.5 - 1 - ms[0]) / ms[1], 5.0, -5.0, 0.0, 1e-3, -1e-3, 2.0, -2.0], np.float32)
        q = D.unstandardize_to_uint8(y, ms)
        assert q.dtype == np.uint8 and np.array_equal(q, O.to_uint8_reference(y, ms))
        back = D.unstandardize_to_uint8(ref, ms)
        assert np.array_equal(back[..., 0], u)           # round trip uint8 -> float -> uint8 is the identity


TC_CASES = [
    # cin, cout, dims(z,y,x): tcgen05 implicit-GEMM path (3x3x3, stride 1)
    (8, 8, (9, 20, 19)),        # Cin == 8: tap-pair k-steps, N padded 8 -> 16
    (16, 16, (7, 18, 13)),
    (32, 32, (8, 21, 12)),      # NPAD = 32
    (8, 16, (12, 9, 10)),
    (32, 16, (5, 35, 27)),      # several y / x tiles with ragged edges
    (16, 32, (23, 10, 10)),     # z chunks
    # wide layers (conv_tcw.cu): 32-channel output groups, Cin swept in 32-channel passes, several z chunks
    (64, 64, (15, 19, 11)),
    (128, 32, (6, 9, 10)),
    (64, 96, (5, 18, 9)),
    (96, 40, (14, 8, 8)),
]


@pytest.mark.parametrize("cin,cout,dims", TC_CASES)
def test_tcgen05_conv_fwd_dgrad(cin, cout, dims):
    """tcgen05/TMA implicit GEMM vs the naive fp64 oracle on identical bf16 operands, and vs the direct kernel."""
    r = np.random.default_rng(cin * 1000 + cout)
    B = 2
    x = bf16r(r.standard_normal((B,) + dims + (cin,)))
    w = bf16r(r.standard_normal((3, 3, 3, cin, cout)) * 0.2)
    xg, wg = _cuda(x, torch.bfloat16), _cuda(w, torch.float32)
    d_tc = make_desc(B, dims, cin, cout, 3, 1, False, 0.3, 0, tc=1)
    d_dir = make_desc(B, dims, cin, cout, 3, 1, False, 0.3, 0, tc=0)
    y = conv_forward(xg, wg, d_tc).float().cpu().numpy()
    ref = naive.lrelu(naive.conv_fwd(x, w, 1), 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    y_dir = conv_forward(xg, wg, d_dir).float().cpu().numpy()
    assert rel_l2(y, y_dir) < 2e-3
    dy = bf16r(r.standard_normal(ref.shape))
    act = bf16r(r.standard_normal(x.shape))
    dyg, actg = _cuda(dy, torch.bfloat16), _cuda(act, torch.bfloat16)
    dx = conv_dgrad(dyg, wg, d_tc, actg, 0.3).float().cpu().numpy()
    dref = naive.conv_dgrad(dy, w, 1, x.shape) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx, dref, rtol=BF16_RTOL, atol=BF16_ATOL * 4)
    assert rel_l2(dx, dref) < 4e-3


def test_tcgen05_dropout_epilogue():
    key = 0xABCDEF01
    r = np.random.default_rng(77)
    x = bf16r(r.standard_normal((1, 6, 18, 9, 16)))
    w = bf16r(r.standard_normal((3, 3, 3, 16, 8)) * 0.2)
    d = make_desc(1, (6, 18, 9), 16, 8, 3, 1, False, 0.3, key, tc=1)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w, torch.float32), d).float().cpu().numpy()
    pre = naive.conv_fwd(x, w, 1)
    ref = naive.lrelu(pre * O.dropout_keep_mask(key, pre.shape) * 2.0, 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)


WG_CASES = [
    # k, s, cin, cout, transposed, dims
    (3, 1, 8, 8, False, (9, 10, 21)),
    (3, 1, 16, 16, False, (7, 8, 13)),
    (3, 1, 32, 16, False, (6, 7, 19)),
    (3, 1, 16, 32, False, (5, 9, 35)),
    (4, 2, 8, 8, False, (10, 12, 38)),
    (4, 2, 16, 16, False, (8, 10, 14)),
    (4, 2, 32, 32, False, (6, 6, 8)),
    (4, 2, 32, 16, True, (5, 6, 7)),
    (4, 2, 16, 8, True, (6, 5, 20)),
    (1, 1, 32, 32, False, (1, 1, 1)),
    # tcgen05 weight gradient (wgrad_tc.cu): every (planes, rows) plan, ragged rows / runs, z chunks
    (3, 1, 8, 8, False, (12, 23, 37)),
    (3, 1, 8, 16, False, (9, 12, 20)),
    (3, 1, 8, 32, False, (6, 9, 19)),
    (3, 1, 16, 8, False, (7, 11, 18)),
    (3, 1, 16, 32, False, (10, 10, 10)),
    (3, 1, 32, 8, False, (6, 8, 35)),
    (3, 1, 32, 32, False, (5, 7, 21)),
    (3, 1, 16, 16, False, (26, 9, 50)),
    # tcgen05 stride-2 weight gradient (wgrad_tc_s2.cu): conv and transposed conv, every (planes, rows) plan
    (4, 2, 8, 8, False, (22, 30, 38)),
    (4, 2, 8, 16, False, (12, 20, 70)),
    (4, 2, 16, 32, False, (10, 12, 14)),
    (4, 2, 32, 8, False, (8, 10, 36)),
    (4, 2, 8, 8, True, (9, 11, 19)),
    (4, 2, 32, 32, True, (4, 5, 6)),
    (4, 2, 8, 32, True, (7, 6, 17)),
    # wide tcgen05 weight gradient (wgrad_tcw.cu): M = 64 / 128 / two 128-channel blocks, ragged rows / column blocks / z chunks
    (3, 1, 64, 32, False, (7, 9, 37)),
    (3, 1, 128, 64, False, (6, 11, 20)),
    (3, 1, 256, 32, False, (5, 7, 10)),
    (3, 1, 64, 96, False, (13, 6, 35)),
    (4, 2, 64, 32, False, (10, 14, 38)),      # wide stride-2 (parity-class form of wgrad_tcw)
    (4, 2, 128, 64, False, (8, 10, 12)),
    (4, 2, 64, 64, True, (5, 7, 9)),
    (4, 2, 32, 128, True, (6, 5, 18)),
    # N = 64 / 128 output channels per CTA: (dz, dy) split in the stride-1 form, two 64-channel swizzled chunks per g tile
    (3, 1, 128, 128, False, (6, 9, 37)),
    (3, 1, 64, 256, False, (5, 7, 20)),
    (3, 1, 256, 64, False, (7, 6, 11)),
    (4, 2, 128, 128, False, (8, 10, 38)),
    (4, 2, 256, 128, True, (4, 5, 9)),
]


@pytest.mark.parametrize("k,s,cin,cout,tr,dims", WG_CASES)
def test_tensor_core_wgrad(k, s, cin, cout, tr, dims):
    """mma.sync weight gradient vs the naive fp64 oracle and vs the direct kernel."""
    r = np.random.default_rng(k * 100 + cin + cout)
    B = 2
    x = bf16r(r.standard_normal((B,) + dims + (cin,)))
    wshape = (k, k, k) + ((cout, cin) if tr else (cin, cout))
    d_tc = make_desc(B, dims, cin, cout, k, s, tr, 1.0, 0, tc=1)
    d_dir = make_desc(B, dims, cin, cout, k, s, tr, 1.0, 0, tc=0)
    w0 = np.zeros(wshape)
    yshape = (naive.convT_fwd(x, w0) if tr else naive.conv_fwd(x, w0, s)).shape
    dy = bf16r(r.standard_normal(yshape))
    xg, dyg = _cuda(x, torch.bfloat16), _cuda(dy, torch.bfloat16)
    dw = conv_wgrad(xg, dyg, d_tc, wshape).cpu().numpy()
    wref = naive.convT_wgrad(x, dy, (k, k, k)) if tr else naive.conv_wgrad(x, dy, s, (k, k, k))
    assert rel_l2(dw, wref) < 1e-5
    np.testing.assert_allclose(dw, wref, rtol=1e-4, atol=2e-3)
    assert rel_l2(conv_wgrad(xg, dyg, d_dir, wshape).cpu().numpy(), dw) < 1e-5


def test_single_channel_wgrads_fast_path():
    """g0-type (Cin = 1, uint8 / fp32 input) and g11-type (Cout = 1, fp32 dy) weight gradients, dedicated kernels."""
    r = np.random.default_rng(123)
    B, dims = 2, (9, 11, 37)
    u = r.integers(0, 256, (B,) + dims + (1,), dtype=np.uint8)
    ms = (0.05, 0.6)
    xs = O.standardize_population(O.scale_tensor(u[..., 0]), ms).astype(np.float64)
    dy = bf16r(r.standard_normal((B, 7, 9, 35, 8)))
    for tc in (1, 0):
        d = make_desc(B, dims, 1, 8, 3, 1, False, 1.0, 0, torch.uint8, torch.bfloat16, ms, tc=tc)
        dw = conv_wgrad(torch.tensor(u).to(DEV), _cuda(dy, torch.bfloat16), d, (3, 3, 3, 1, 8)).cpu().numpy()
        assert rel_l2(dw, naive.conv_wgrad(xs, dy, 1, (3, 3, 3))) < 1e-5
    xf = r.standard_normal((B,) + dims + (1,)).astype(np.float32)
    d = make_desc(B, dims, 1, 8, 3, 1, False, 1.0, 0, torch.float32, torch.bfloat16, tc=1)
    dw = conv_wgrad(torch.tensor(xf).to(DEV), _cuda(dy, torch.bfloat16), d, (3, 3, 3, 1, 8)).cpu().numpy()
    assert rel_l2(dw, naive.conv_wgrad(xf.astype(np.float64), dy, 1, (3, 3, 3))) < 1e-5
    # Cout = 1
    x = bf16r(r.standard_normal((B,) + dims + (16,)))
    dy1 = r.standard_normal((B, 7, 9, 35, 1)).astype(np.float32)
    d = make_desc(B, dims, 16, 1, 3, 1, False, 1.0, 0, torch.bfloat16, torch.float32, tc=1)
    dw = conv_wgrad(_cuda(x, torch.bfloat16), torch.tensor(dy1).to(DEV), d, (3, 3, 3, 16, 1)).cpu().numpy()
    assert rel_l2(dw, naive.conv_wgrad(x, dy1.astype(np.float64), 1, (3, 3, 3))) < 1e-5


CM_CASES = [
    # k, s, cin, cout, transposed, dims
    (4, 2, 8, 8, False, (10, 12, 38)),      # g2
    (4, 2, 16, 16, False, (8, 10, 14)),     # g4
    (4, 2, 32, 32, False, (6, 6, 8)),       # d4
    (4, 2, 32, 32, False, (4, 4, 4)),       # d6
    (4, 2, 32, 16, True, (5, 6, 7)),        # g6
    (4, 2, 16, 8, True, (6, 5, 20)),        # g9
    (1, 1, 32, 32, False, (1, 1, 1)),       # d7
    (3, 1, 16, 24, False, (5, 6, 17)),      # generic shape (NB = 3)
    # tcgen05 stride-2 kernels (conv_tc_s2.cu): several (x,y) tiles with ragged edges, z chunks, odd extents
    (4, 2, 8, 8, False, (22, 40, 38)),
    (4, 2, 8, 16, False, (9, 36, 21)),
    (4, 2, 16, 32, False, (14, 9, 40)),
    (4, 2, 16, 8, True, (11, 19, 10)),
    (4, 2, 32, 16, True, (4, 18, 9)),
    (4, 2, 8, 8, True, (13, 5, 23)),
    (4, 2, 16, 16, False, (32, 32, 32)),
]


@pytest.mark.parametrize("k,s,cin,cout,tr,dims", CM_CASES)
def test_mma_conv_fwd_dgrad(k, s, cin, cout, tr, dims):
    """mma.sync conv kernel (stride-2 / transposed / 1x1 layers) vs the naive fp64 oracle, forward + data gradient."""
    r = np.random.default_rng(k * 1000 + cin * 10 + cout + int(tr))
    B = 2
    x = bf16r(r.standard_normal((B,) + dims + (cin,)))
    wshape = (k, k, k) + ((cout, cin) if tr else (cin, cout))
    w = bf16r(r.standard_normal(wshape) * 0.2)
    d = make_desc(B, dims, cin, cout, k, s, tr, 0.3, 0, tc=1)
    xg, wg = _cuda(x, torch.bfloat16), _cuda(w, torch.float32)
    y = conv_forward(xg, wg, d).float().cpu().numpy()
    ref = naive.lrelu(naive.convT_fwd(x, w) if tr else naive.conv_fwd(x, w, s), 0.3)
    assert y.shape == ref.shape
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    dy = bf16r(r.standard_normal(ref.shape))
    act = bf16r(r.standard_normal(x.shape))
    dx = conv_dgrad(_cuda(dy, torch.bfloat16), wg, d, _cuda(act, torch.bfloat16), 0.3).float().cpu().numpy()
    dref = (naive.convT_dgrad(dy, w, x.shape) if tr else naive.conv_dgrad(dy, w, s, x.shape)) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx, dref, rtol=BF16_RTOL, atol=BF16_ATOL * 4)
    assert rel_l2(dx, dref) < 4e-3


WIDE_S2_CASES = [
    # k, s, cin, cout, transposed, dims: wide 4x4x4 stride-2 layers of the wf <= 4 models (streamed-weight tcgen05 kernels)
    (4, 2, 64, 64, False, (10, 20, 22)),     # g2 / d1 at wf = 1: fwd = wide DOWN (NP 64, 32-slot weight ring), dgrad = wide UP
    (4, 2, 128, 128, False, (8, 6, 38)),     # g4 at wf = 1: eight 16-channel chunks, two 64-channel chunks, ragged x tiles
    (4, 2, 128, 64, True, (5, 9, 18)),       # g9 at wf = 1: fwd = wide UP (Cin 128, two cout groups), dgrad = wide DOWN (NP 128)
    (4, 2, 64, 32, True, (4, 18, 9)),        # g6 at wf = 4: fwd = wide UP (one cout group), dgrad = wide DOWN (Cin 32)
    (4, 2, 64, 256, False, (20, 6, 8)),      # NP = 256: two output slices per TMEM strip, several z chunks
    (4, 2, 64, 40, False, (6, 8, 8)),        # Cout not a multiple of the column groups
    (4, 2, 32, 32, False, (20, 36, 38)),     # g2 at wf = 2: too many weights for the resident kernel, half of the NP = 64 columns used
]


@pytest.mark.parametrize("k,s,cin,cout,tr,dims", WIDE_S2_CASES)
def test_wide_stride2_tcgen05_fwd_dgrad(k, s, cin, cout, tr, dims):
    """conv_upw_tc_kernel / conv_downw_tc_kernel (models/utils.py:80,129-130 at 64-256 channels) vs the naive fp64 oracle."""
    lib = _lib.load()
    r = np.random.default_rng(k * 1000 + cin * 10 + cout + int(tr))
    B = 2
    x = bf16r(r.standard_normal((B,) + dims + (cin,)))
    wshape = (k, k, k) + ((cout, cin) if tr else (cin, cout))
    w = bf16r(r.standard_normal(wshape) * 0.05)
    d = make_desc(B, dims, cin, cout, k, s, tr, 0.3, 0, tc=1)
    xg, wg = _cuda(x, torch.bfloat16), _cuda(w, torch.float32)
    y = conv_forward(xg, wg, d).float().cpu().numpy()
    assert lib.tem_last_kernel().decode() == ("conv_upw_tc_kernel" if tr else "conv_downw_tc_kernel")
    ref = naive.lrelu(naive.convT_fwd(x, w) if tr else naive.conv_fwd(x, w, s), 0.3)
    assert y.shape == ref.shape
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    dy = bf16r(r.standard_normal(ref.shape))
    act = bf16r(r.standard_normal(x.shape))
    dx = conv_dgrad(_cuda(dy, torch.bfloat16), wg, d, _cuda(act, torch.bfloat16), 0.3).float().cpu().numpy()
    if tr or cout % 64 == 0:      # the data gradient reads `cout` channels: wide UP needs 64-channel chunks (else resident kernel)
        assert lib.tem_last_kernel().decode() == ("conv_downw_tc_kernel" if tr else "conv_upw_tc_kernel")
    dref = (naive.convT_dgrad(dy, w, x.shape) if tr else naive.conv_dgrad(dy, w, s, x.shape)) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx, dref, rtol=BF16_RTOL, atol=BF16_ATOL * 4)
    assert rel_l2(dx, dref) < 4e-3


def test_tcgen05_stride2_dropout_epilogue():
    """Conv3DTranspose forward on the tcgen05 UP kernel with the fused Dropout(0.5) mask (models/utils.py:129-135)."""
    key = 0x1234ABCD
    r = np.random.default_rng(91)
    x = bf16r(r.standard_normal((2, 5, 18, 9, 16)))
    w = bf16r(r.standard_normal((4, 4, 4, 8, 16)) * 0.2)
    d = make_desc(2, (5, 18, 9), 16, 8, 4, 2, True, 0.3, key, tc=1)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w, torch.float32), d).float().cpu().numpy()
    pre = naive.convT_fwd(x, w)
    ref = naive.lrelu(pre * O.dropout_keep_mask(key, pre.shape) * 2.0, 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_wide_stride2_dropout_epilogue():
    """Conv3DTranspose forward at 64 -> 32 channels (g6 at wf = 4) on conv_upw_tc_kernel with the fused Dropout(0.5) mask."""
    key = 0x2345BCDE
    r = np.random.default_rng(92)
    x = bf16r(r.standard_normal((2, 5, 18, 9, 64)))
    w = bf16r(r.standard_normal((4, 4, 4, 32, 64)) * 0.05)
    d = make_desc(2, (5, 18, 9), 64, 32, 4, 2, True, 0.3, key, tc=1)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w, torch.float32), d).float().cpu().numpy()
    assert _lib.load().tem_last_kernel().decode() == "conv_upw_tc_kernel"
    pre = naive.convT_fwd(x, w)
    ref = naive.lrelu(pre * O.dropout_keep_mask(key, pre.shape) * 2.0, 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)


def test_single_channel_conv_fast_paths():
    """1 -> C forward (uint8 / fp32), its flipped form (dgrad of a Cout=1 layer) and C -> 1 forward."""
    r = np.random.default_rng(321)
    B, dims = 2, (9, 13, 37)
    u = r.integers(0, 256, (B,) + dims + (1,), dtype=np.uint8)
    ms = (0.05, 0.6)
    xs = O.standardize_population(O.scale_tensor(u[..., 0]), ms).astype(np.float64)
    w = bf16r(r.standard_normal((3, 3, 3, 1, 8)) * 0.3)
    for tc in (1, 0):
        d = make_desc(B, dims, 1, 8, 3, 1, False, 0.3, 0, torch.uint8, torch.bfloat16, ms, tc=tc)
        y = conv_forward(torch.tensor(u).to(DEV), _cuda(w, torch.float32), d).float().cpu().numpy()
        # tensor-core path (conv_c1tc.cu): the standardised input enters as a bf16 hi + lo pair, i.e. with ~16 mantissa bits
        np.testing.assert_allclose(y, naive.lrelu(naive.conv_fwd(xs, w, 1), 0.3), rtol=BF16_RTOL, atol=BF16_ATOL)
        assert _lib.load().tem_last_kernel().decode() == ("conv_c1tc_kernel" if tc else "conv_direct_kernel")
    # C -> 1 forward (fp32 out) and its data gradient (1 -> C, flipped, pad 2, LeakyReLU' of the stored activation)
    x = bf16r(r.standard_normal((B,) + dims + (16,)))
    w1 = bf16r(r.standard_normal((3, 3, 3, 16, 1)) * 0.2)
    d = make_desc(B, dims, 16, 1, 3, 1, False, 1.0, 0, torch.bfloat16, torch.float32, tc=1)
    y = conv_forward(_cuda(x, torch.bfloat16), _cuda(w1, torch.float32), d).cpu().numpy()
    ref = naive.conv_fwd(x, w1, 1)
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-4)
    dy = r.standard_normal(ref.shape).astype(np.float32)
    act = bf16r(r.standard_normal(x.shape))
    dx = conv_dgrad(torch.tensor(dy).to(DEV), _cuda(w1, torch.float32), d, _cuda(act, torch.bfloat16), 0.3).float().cpu().numpy()
    assert _lib.load().tem_last_kernel().decode() == "conv_c1tc_kernel"       # dy enters as a bf16 hi + lo pair
    dref = naive.conv_dgrad(dy.astype(np.float64), w1, 1, x.shape) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx, dref, rtol=BF16_RTOL, atol=BF16_ATOL * 2)


def test_single_channel_conv_wide_models():
    """The single-channel ends of the wf <= 4 models (g0 / d0: 1 -> 64, g11: 128 -> 1) stay on the dedicated kernels:
    the channel dimension is sliced into 16-channel (1 -> C) / 32-channel (C -> 1) launches."""
    lib = _lib.load()
    r = np.random.default_rng(654)
    B, dims = 2, (7, 11, 35)
    u = r.integers(0, 256, (B,) + dims + (1,), dtype=np.uint8)
    ms = (0.05, 0.6)
    xs = O.standardize_population(O.scale_tensor(u[..., 0]), ms).astype(np.float64)
    w = bf16r(r.standard_normal((3, 3, 3, 1, 64)) * 0.3)
    d = make_desc(B, dims, 1, 64, 3, 1, False, 0.3, 0, torch.uint8, torch.bfloat16, ms, tc=1)
    y = conv_forward(torch.tensor(u).to(DEV), _cuda(w, torch.float32), d).float().cpu().numpy()
    assert lib.tem_last_kernel().decode() == "conv_c1_kernel"
    ref = naive.lrelu(naive.conv_fwd(xs, w, 1), 0.3)
    np.testing.assert_allclose(y, ref, rtol=BF16_RTOL, atol=BF16_ATOL)
    # data gradient of the 1 -> 64 layer into its fp32 single-channel input (64 -> 1, flipped taps, zero padding 2)
    dy = bf16r(r.standard_normal(ref.shape))
    df = make_desc(B, dims, 1, 64, 3, 1, False, 0.3, 0, torch.float32, torch.bfloat16, tc=1)
    dx = conv_dgrad(_cuda(dy, torch.bfloat16), _cuda(w, torch.float32), df, None, 1.0, torch.float32).cpu().numpy()
    assert lib.tem_last_kernel().decode() == "conv_c1_kernel"
    dref = naive.conv_dgrad(dy, w, 1, xs.shape)
    np.testing.assert_allclose(dx, dref, rtol=1e-3, atol=1e-3)
    # 128 -> 1 forward (fp32, linear) and its data gradient (1 -> 128 with LeakyReLU' of the stored activation)
    x = bf16r(r.standard_normal((B,) + dims + (128,)))
    w1 = bf16r(r.standard_normal((3, 3, 3, 128, 1)) * 0.1)
    d1 = make_desc(B, dims, 128, 1, 3, 1, False, 1.0, 0, torch.bfloat16, torch.float32, tc=1)
    y1 = conv_forward(_cuda(x, torch.bfloat16), _cuda(w1, torch.float32), d1).cpu().numpy()
    assert lib.tem_last_kernel().decode() == "conv_c1_kernel"
    ref1 = naive.conv_fwd(x, w1, 1)
    np.testing.assert_allclose(y1, ref1, rtol=1e-4, atol=2e-4)
    dy1 = r.standard_normal(ref1.shape).astype(np.float32)
    act = bf16r(r.standard_normal(x.shape))
    dx1 = conv_dgrad(torch.tensor(dy1).to(DEV), _cuda(w1, torch.float32), d1, _cuda(act, torch.bfloat16), 0.3).float().cpu().numpy()
    assert lib.tem_last_kernel().decode() == "conv_c1_kernel"
    dref1 = naive.conv_dgrad(dy1.astype(np.float64), w1, 1, x.shape) * naive.lrelu_grad_from_output(act, 0.3)
    np.testing.assert_allclose(dx1, dref1, rtol=BF16_RTOL, atol=BF16_ATOL * 2)


def test_device_augment_bit_exact():
    """datasets.py:123-155 on device: axis shuffle + flips are index work (bit-exact), the intensity jitter is two fp32
    roundings in the reference's order; the uint8 entry fuses scale_tensor + standardize_population in front."""
    from transfer_em_b200 import datasets as D
    r = np.random.default_rng(42)
    B, n = 5, 12
    ms = (0.04, 0.57)
    u = r.integers(0, 256, (B, n, n, n), dtype=np.uint8)
    choices = D.draw_augmentation(B, 3, r)
    perm, flip, var, mean = choices
    assert sorted(perm[0].tolist()) == [0, 1, 2] and var.min() >= 1.0 and var.max() <= 1.05 and abs(mean).max() <= 0.05
    out = D.augment(u, meanstd=ms, choices=choices)
    assert out.shape == (B, n, n, n, 1) and out.dtype == np.float32
    for b in range(B):
        xs = O.standardize_population(O.scale_tensor(u[b]), ms)          # [n,n,n,1] float32
        ref = O.augment(xs, perm[b], flip[b], var[b], mean[b])
        assert np.array_equal(out[b], ref), b
    # float32 entry, non-cubic patch with one shared permutation, and the 2-D form
    x = r.standard_normal((3, 4, 6, 5, 1)).astype(np.float32)
    pm = np.tile(np.array([[1, 2, 0]], np.int32), (3, 1)); fl = np.array([[0, 1, 1], [1, 0, 0], [1, 1, 1]], np.int32)
    va = np.array([1.0, 1.02, 1.05], np.float32); me = np.array([0.05, -0.05, 0.0], np.float32)
    out = D.augment(x, choices=(pm, fl, va, me))
    for b in range(3):
        assert np.array_equal(out[b], O.augment(x[b], pm[b], fl[b], va[b], me[b]))
    x2 = r.standard_normal((2, 7, 9)).astype(np.float32)
    pm2 = np.array([[0, 2, 1], [0, 2, 1]], np.int32); fl2 = np.array([[0, 1, 0], [0, 0, 1]], np.int32)
    out2 = D.augment(x2, choices=(pm2, fl2, va[:2], me[:2]))
    for b in range(2):
        assert np.array_equal(out2[b], O.augment(x2[b][..., None], pm2[b, 1:] - 1, fl2[b, 1:], va[b], me[b]))


def test_device_get_meanstd():
    """datasets.py:173-190: mean of per-tensor means, sqrt of the mean of per-tensor population variances."""
    from transfer_em_b200 import datasets as D
    r = np.random.default_rng(43)
    ts = [(r.standard_normal((20, 31, 17, 1)) * (0.5 + i) + 0.1 * i).astype(np.float32) for i in range(4)]
    m, s = D.get_meanstd(ts)
    mr, sr = O.get_meanstd(ts)
    m64 = np.mean([t.astype(np.float64).mean() for t in ts]); s64 = np.sqrt(np.mean([t.astype(np.float64).var() for t in ts]))
    assert abs(m - m64) < 2e-6 * max(1.0, abs(m64)) and abs(s - s64) < 2e-6 * s64
    assert abs(m - mr) < 1e-5 and abs(s - sr) < 1e-5 * sr


def test_chunk_volume_matches_reference_blocks(tmp_path):
    """model_cloudrun/transferem.py:171-184: 64^3 blocks of the zyx result (clipped at the edge), C-order bytes, z outer /
    y / x inner, file names with the un-clipped +64 and the request offset.  Byte work: bit-exact."""
    import gzip
    from transfer_em_b200.utils import chunk_volume, write_ng_chunks
    r = np.random.default_rng(5)
    vol = r.integers(0, 256, (70, 130, 65), dtype=np.uint8)
    blocks = chunk_volume(vol, 64)
    ref = []
    for z in range(0, 70, 64):
        for y in range(0, 130, 64):
            for x in range(0, 65, 64):
                ref.append(((x, y, z), vol[z:z + 64, y:y + 64, x:x + 64].tobytes()))
    assert [b[0] for b in blocks] == [b[0] for b in ref]
    assert all(a[1] == b[1] for a, b in zip(blocks, ref))
    names = write_ng_chunks(vol, str(tmp_path / "ng"), offset_xyz=(128, 0, 64))
    assert names[0] == "128-192_0-64_64-128" and names[-1] == "192-256_128-192_128-192" and len(names) == 2 * 3 * 2
    assert gzip.decompress(open(tmp_path / "ng" / names[1], "rb").read()) == vol[0:64, 0:64, 64:65].tobytes()


def test_device_warp_tensor():
    """debug.warp_tensor (tem_warp_tensor: box blur + dilated holes at the blurred mean, transfer_em/debug.py:7-63) vs the oracle,
    3-D and 2-D, with the seeds of the holes given explicitly; accuracy() is the RMSE of debug.py:65-71."""
    from transfer_em_b200 import debug as D
    r = np.random.default_rng(77)
    for sp in ((9, 13, 37), (40, 33)):
        nd = len(sp)
        t = r.uniform(-1, 1, sp + (1,)).astype(np.float32)
        u = r.uniform(0.01, 1, sp).astype(np.float32)
        for pos in ((0,) * nd, tuple(n // 2 for n in sp), tuple(n - 1 for n in sp)):
            u[pos] = 0.0
        got = D.warp_tensor(torch.tensor(t).to(DEV), uniform=u).cpu().numpy()
        ref = O.warp_tensor(t, u)
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)
        assert (got == got[(0,) * nd + (0,)]).sum() >= 2 ** nd + 4 ** nd          # the holes are there
    # default draw: numpy in, numpy out, hole rate 4 / 128^2
    big = r.uniform(-1, 1, (64, 64, 64, 1)).astype(np.float32)
    w = D.warp_tensor(big, rng=np.random.default_rng(3))
    u = np.random.default_rng(3).uniform(0.0, 1.0, 64 ** 3).astype(np.float32)
    np.testing.assert_allclose(w, O.warp_tensor(big, u), rtol=0, atol=2e-6)
    assert abs(D.accuracy(big, w) - float(np.sqrt(np.mean((big - w) ** 2)))) < 1e-6
