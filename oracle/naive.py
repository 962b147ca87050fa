"""Independent naive-loop restatement of the convolution algebra (TEST INFRASTRUCTURE ONLY).

Second, deliberately different implementation of the per-layer formulas of
SURVEY.md Appendix A (derived from transfer_em/models/utils.py:73,80,129-130):
explicit tap loops in numpy fp64, channels-last, with hand-written data- and
weight-gradient formulas (no autograd).  tests/test_oracle.py checks that this
and oracle/tem_oracle.py (torch ops + autograd) agree to fp64 round-off, which is
the only pin available for the oracle (the reference has no tests and TensorFlow
is not installable here: parity unpinned).

All functions are 3-D; a 2-D case is a depth-1 volume with a (1,k,k) kernel.
x: [B,Z,Y,X,Ci]; conv kernel w: [kz,ky,kx,Ci,Co]; convT kernel w: [kz,ky,kx,Co,Ci].
"""
from __future__ import annotations

import itertools
import numpy as np


def conv_fwd(x, w, s):
    """VALID cross-correlation: y[o,co] = sum_{k,ci} x[s*o+k,ci] w[k,ci,co]."""
    B, Z, Y, X, Ci = x.shape
    kz, ky, kx, _, Co = w.shape
    oz, oy, ox = (Z - kz) // s + 1, (Y - ky) // s + 1, (X - kx) // s + 1
    y = np.zeros((B, oz, oy, ox, Co), np.float64)
    for dz, dy, dx in itertools.product(range(kz), range(ky), range(kx)):
        xs = x[:, dz:dz + s * (oz - 1) + 1:s, dy:dy + s * (oy - 1) + 1:s, dx:dx + s * (ox - 1) + 1:s, :]
        y += xs @ w[dz, dy, dx]
    return y


def conv_dgrad(dy, w, s, in_shape):
    """dx[s*o+k,ci] += dy[o,co] w[k,ci,co] (scatter form)."""
    B, Z, Y, X, Ci = in_shape
    kz, ky, kx, _, Co = w.shape
    _, oz, oy, ox, _ = dy.shape
    dx_ = np.zeros(in_shape, np.float64)
    for dz, dyy, dxx in itertools.product(range(kz), range(ky), range(kx)):
        dx_[:, dz:dz + s * (oz - 1) + 1:s, dyy:dyy + s * (oy - 1) + 1:s, dxx:dxx + s * (ox - 1) + 1:s, :] += dy @ w[dz, dyy, dxx].T
    return dx_


def conv_wgrad(x, dy, s, kshape):
    """dw[k,ci,co] = sum_{b,o} x[s*o+k,ci] dy[o,co]."""
    kz, ky, kx = kshape
    _, oz, oy, ox, Co = dy.shape
    Ci = x.shape[-1]
    dw = np.zeros((kz, ky, kx, Ci, Co), np.float64)
    d2 = dy.reshape(-1, Co)
    for dz, dyy, dxx in itertools.product(range(kz), range(ky), range(kx)):
        xs = x[:, dz:dz + s * (oz - 1) + 1:s, dyy:dyy + s * (oy - 1) + 1:s, dxx:dxx + s * (ox - 1) + 1:s, :]
        dw[dz, dyy, dxx] = xs.reshape(-1, Ci).T @ d2
    return dw


def _axis_pairs(n_in, k, n_out):
    """Per-axis (i, j) index arrays of the SAME stride-2 transposed conv: j = 2i + k - 1."""
    i = np.arange(n_in)
    j = 2 * i + k - 1
    ok = (j >= 0) & (j < n_out)
    return i[ok], j[ok]


def convT_fwd(x, w):
    """Keras Conv3DTranspose(k=4, s=2, 'same'): out = 2n, y[2i+k-1,co] += x[i,ci] w[k,co,ci].
    Depth-1 inputs with kz == 1 are treated as 2-D (no z upsampling)."""
    B, Z, Y, X, Ci = x.shape
    kz, ky, kx, Co, _ = w.shape
    oz = Z if kz == 1 else 2 * Z
    y = np.zeros((B, oz, 2 * Y, 2 * X, Co), np.float64)
    for dz, dy, dx in itertools.product(range(kz), range(ky), range(kx)):
        if kz == 1:
            iz = jz = np.arange(Z)
        else:
            iz, jz = _axis_pairs(Z, dz, oz)
        iy, jy = _axis_pairs(Y, dy, 2 * Y)
        ix, jx = _axis_pairs(X, dx, 2 * X)
        contrib = x[:, iz][:, :, iy][:, :, :, ix] @ w[dz, dy, dx].T      # [.., Co]
        y[np.ix_(np.arange(B), jz, jy, jx)] += contrib
    return y


def convT_dgrad(dy, w, in_shape):
    """dx[i,ci] = sum_{k,co} dy[2i+k-1,co] w[k,co,ci] (out-of-range dy = 0)."""
    B, Z, Y, X, Ci = in_shape
    kz, ky, kx, Co, _ = w.shape
    oz = dy.shape[1]
    dx_ = np.zeros(in_shape, np.float64)
    for dz, dyy, dxx in itertools.product(range(kz), range(ky), range(kx)):
        if kz == 1:
            iz = jz = np.arange(Z)
        else:
            iz, jz = _axis_pairs(Z, dz, oz)
        iy, jy = _axis_pairs(Y, dyy, 2 * Y)
        ix, jx = _axis_pairs(X, dxx, 2 * X)
        g = dy[:, jz][:, :, jy][:, :, :, jx] @ w[dz, dyy, dxx]           # [.., Ci]
        dx_[np.ix_(np.arange(B), iz, iy, ix)] += g
    return dx_


def convT_wgrad(x, dy, kshape):
    """dw[k,co,ci] = sum_{b,i} x[i,ci] dy[2i+k-1,co]."""
    kz, ky, kx = kshape
    B, Z, Y, X, Ci = x.shape
    Co = dy.shape[-1]
    oz = dy.shape[1]
    dw = np.zeros((kz, ky, kx, Co, Ci), np.float64)
    for dz, dyy, dxx in itertools.product(range(kz), range(ky), range(kx)):
        if kz == 1:
            iz = jz = np.arange(Z)
        else:
            iz, jz = _axis_pairs(Z, dz, oz)
        iy, jy = _axis_pairs(Y, dyy, 2 * Y)
        ix, jx = _axis_pairs(X, dxx, 2 * X)
        xs = x[:, iz][:, :, iy][:, :, :, ix].reshape(-1, Ci)
        ds = dy[:, jz][:, :, jy][:, :, :, jx].reshape(-1, Co)
        dw[dz, dyy, dxx] = ds.T @ xs
    return dw


def lrelu(x, slope):
    return np.where(x > 0, x, x * slope)


def lrelu_grad_from_output(y, slope):
    """dL/dx = dL/dy * (y > 0 ? 1 : slope): sign of the stored output (slope > 0 keeps sign)."""
    return np.where(y > 0, 1.0, slope)


def focal_logits_and_grad(x, target, gamma=2.0, alpha=0.5):
    """Closed forms of SURVEY Appendix A for gamma=2; general gamma by formula.
    Returns (per-element loss, d loss / d logit) WITHOUT the mean."""
    x = np.asarray(x, np.float64)
    p = 1.0 / (1.0 + np.exp(-x))
    if target == 1:
        pt, a_t = p, alpha
        logpt = -np.logaddexp(0.0, -x)
    else:
        pt, a_t = 1 - p, 1 - alpha
        logpt = -np.logaddexp(0.0, x)
    l = a_t * (1 - pt) ** gamma * (-logpt)
    # d/dx: dpt/dx = +-p(1-p)
    dpt = p * (1 - p) * (1.0 if target == 1 else -1.0)
    dl_dpt = a_t * (gamma * (1 - pt) ** (gamma - 1) * logpt - (1 - pt) ** gamma / pt)
    return l, dl_dpt * dpt


def focal_nl_and_grad(a, b, gamma=2.0, alpha=0.5, eps=1e-7):
    """Identity/cycle element loss l(t), t = 1-|a-b|/2 vs target 1, and dl/db (no mean)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = a - b
    u = np.abs(d) / 2
    t = 1 - u
    tc = np.clip(t, eps, 1 - eps)
    c = -np.log(tc + eps)
    l = alpha * u ** gamma * c
    inside = (t > eps) & (t < 1 - eps)
    dl_dt = alpha * (-gamma * u ** (gamma - 1) * c + u ** gamma * np.where(inside, -1.0 / (tc + eps), 0.0))
    # dt/db = +sign(d)/2
    return l, dl_dt * np.sign(d) / 2
