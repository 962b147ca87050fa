"""CPU oracle for the transfer_em hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (torch-CPU ops, fp32 or fp64) of the reference
graph.  It is imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
(``transfer_em_b200``) never imports it.

PARITY UNPINNED: the arithmetic of the reference lives in TensorFlow 2.x,
Keras and tensorflow_addons, none of which is installed (or installable) in the
build image, and the reference ships no tests / golden vectors.  The upstream
semantics encoded here (marked [upstream]) are restated from the published
behaviour of those libraries; ``tools/export_tf_golden.py`` lets a third party
with TF close the loop.  What *is* pinned: an independent naive numpy-loop
restatement (``oracle/naive.py``) agrees with this one to fp64 round-off, and
the structural facts of the reference graph (shape table, parameter counts
129480 / 181369 at wf=8, initial loss 0.17329) are asserted in the tests.

Reference files restated (paths relative to /root/reference):
  transfer_em/models/utils.py:41-137      downsample / upsample blocks
  transfer_em/models/generator.py:22-117  unet_generator
  transfer_em/models/discriminator.py:14-105  discriminator
  transfer_em/cgan.py:40-81,110-142,144-230,289-293  EM2EM losses / train_step
  transfer_em/utils.py:41-130             predict_ng_cube tiling and uint8 math
  transfer_em/datasets/datasets.py:157-171,193-202  scale / (un)standardize

Layouts: activations are channels-last numpy/torch at the API ([B,Z,Y,X,C] or
[B,Y,X,C]); kernels are in Keras layout ([k..,Cin,Cout]; transposed conv
[k..,Cout,Cin] [upstream]).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import itertools
import numpy as np
import torch
import torch.nn.functional as F

LRELU_ALPHA = 0.3        # tf.keras.layers.LeakyReLU() default [upstream]
KERAS_EPS = 1e-7         # tf.keras.backend.epsilon() [upstream]
INIT_STD = 0.02          # random_normal_initializer(0., 0.02): models/utils.py:58


# --------------------------------------------------------------------------
# layer tables (generator.py:54-110, discriminator.py:39-99, utils.py:73-135)
# --------------------------------------------------------------------------
@dataclass
class LayerSpec:
    name: str
    kind: str            # 'conv' | 'convT'
    k: int
    stride: int
    cin: int
    cout: int
    slope: float         # activation slope for x<0 (1.0 = linear)
    dropout: bool = False
    bias: bool = False

    def kernel_shape(self, is3d: bool) -> Tuple[int, ...]:
        ks = (self.k,) * (3 if is3d else 2)
        if self.kind == 'conv':
            return ks + (self.cin, self.cout)
        return ks + (self.cout, self.cin)     # Keras Conv*DTranspose layout

    def nparams(self, is3d: bool) -> int:
        n = int(np.prod(self.kernel_shape(is3d)))
        return n + (self.cout if self.bias else 0)


def generator_layers(wf: int = 8) -> List[LayerSpec]:
    """generator.py:54-110 with utils.py blocks expanded (g0..g11)."""
    c1, c2, c4 = 64 // wf, 128 // wf, 256 // wf
    a = LRELU_ALPHA
    return [
        LayerSpec('g0', 'conv', 3, 1, 1, c1, a),            # generator.py:54-57
        LayerSpec('g1', 'conv', 3, 1, c1, c1, a),           # down1.conv3 (skip0) utils.py:73-77
        LayerSpec('g2', 'conv', 4, 2, c1, c1, a),           # down1.conv4 s2  utils.py:80-83
        LayerSpec('g3', 'conv', 3, 1, c1, c2, a),           # down2.conv3 (skip1)
        LayerSpec('g4', 'conv', 4, 2, c2, c2, a),           # down2.conv4 s2
        LayerSpec('g5', 'conv', 3, 1, c2, 2 * c2, a),       # up2.conv3 utils.py:122-126
        LayerSpec('g6', 'convT', 4, 2, 2 * c2, c2, a, dropout=True),  # utils.py:129-135
        LayerSpec('g7', 'conv', 3, 1, 2 * c2, c4, a),       # generator.py:96-99 (input = cat)
        LayerSpec('g8', 'conv', 3, 1, c4, 2 * c1, a),       # up1.conv3
        LayerSpec('g9', 'convT', 4, 2, 2 * c1, c1, a, dropout=True),
        LayerSpec('g10', 'conv', 3, 1, 2 * c1, c2, a),      # generator.py:108-109 (input = cat)
        LayerSpec('g11', 'conv', 3, 1, c2, 1, 1.0),         # generator.py:110 linear, no bias
    ]


def discriminator_layers(wf: int = 8, is3d: bool = True) -> List[LayerSpec]:
    """discriminator.py:39-99 (d0..d8).

    The reference hard-codes 16 (HACK conv, :45-46) and dims=32 (:60); both
    equal 128//wf and 256//wf at wf=8, the only width at which the reference
    type-checks.  They are generalised to those expressions here (identical at
    wf=8).  The literal 32 out-filters of block "3" (:72) is kept.
    In 2D the HACK conv is applied to the raw input (:49-51) so block "1" is
    dead; that quirk is reproduced.
    """
    c1, c2, c4 = 64 // wf, 128 // wf, 256 // wf
    a = LRELU_ALPHA
    L = []
    if is3d:
        L += [LayerSpec('d0', 'conv', 3, 1, 1, c1, a),
              LayerSpec('d1', 'conv', 4, 2, c1, c1, a),
              LayerSpec('d2', 'conv', 3, 1, c1, c2, a)]     # HACK conv
    else:
        # block "1" still owns variables (d0,d1) but is not on the path
        L += [LayerSpec('d0', 'conv', 3, 1, 1, c1, a),
              LayerSpec('d1', 'conv', 4, 2, c1, c1, a),
              LayerSpec('d2', 'conv', 3, 1, 1, c2, a)]      # applied to raw input
    L += [LayerSpec('d3', 'conv', 3, 1, c2, c4, a),
          LayerSpec('d4', 'conv', 4, 2, c4, c4, a),
          LayerSpec('d5', 'conv', 3, 1, c4, 32, a),
          LayerSpec('d6', 'conv', 4, 2, 32, 32, a * a),     # LReLU twice: discriminator.py:73-74
          LayerSpec('d7', 'conv', 1, 1, 32, c4, a),
          LayerSpec('d8', 'conv', 1, 1, c4, 1, 1.0, bias=True)]
    return L


def generator_out_dim(n: int) -> int:
    """generator.py:46-112 curr_dim bookkeeping: 74 -> 40; n -> n-34 for n = 2 mod 4."""
    return generator_dims(n)['out']


def generator_dims(n: int) -> Dict[str, int]:
    d = {}
    d['g0'] = n - 2
    d['g1'] = d['g0'] - 2            # skip0
    d['g2'] = (d['g1'] - 4) // 2 + 1
    d['g3'] = d['g2'] - 2            # skip1
    d['g4'] = (d['g3'] - 4) // 2 + 1
    d['g5'] = d['g4'] - 2
    d['g6'] = d['g5'] * 2
    d['crop1'] = d['g3'] - d['g6']   # total crop (both sides)
    d['g7'] = d['g6'] - 2
    d['g8'] = d['g7'] - 2
    d['g9'] = d['g8'] * 2
    d['crop0'] = d['g1'] - d['g9']
    d['g10'] = d['g9'] - 2
    d['g11'] = d['g10'] - 2
    d['out'] = d['g11']
    return d


def init_params(layers: Sequence[LayerSpec], is3d: bool, rng: np.random.Generator,
                dtype=np.float32) -> List[np.ndarray]:
    """N(0, 0.02) kernels, zero bias (utils.py:58, discriminator.py:29,97-99)."""
    out = []
    for L in layers:
        out.append((rng.standard_normal(L.kernel_shape(is3d)) * INIT_STD).astype(dtype))
        if L.bias:
            out.append(np.zeros((L.cout,), dtype))
    return out


def flatten_params(params: Sequence[np.ndarray]) -> np.ndarray:
    return np.concatenate([np.asarray(p).reshape(-1) for p in params])


def unflatten_params(flat: np.ndarray, layers: Sequence[LayerSpec], is3d: bool) -> List[np.ndarray]:
    out, o = [], 0
    for L in layers:
        shp = L.kernel_shape(is3d)
        n = int(np.prod(shp))
        out.append(np.asarray(flat[o:o + n]).reshape(shp)); o += n
        if L.bias:
            out.append(np.asarray(flat[o:o + L.cout])); o += L.cout
    assert o == len(flat)
    return out


# --------------------------------------------------------------------------
# dropout mask: counter hash shared bit-for-bit with the CUDA kernels
# (the reference uses tf.random, whose streams cannot be matched)
# --------------------------------------------------------------------------
def hash32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & 0xFFFFFFFF
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & 0xFFFFFFFF
    x ^= x >> np.uint64(16)
    return x


def dropout_keep_mask(key: int, shape: Sequence[int]) -> np.ndarray:
    """keep[idx] = top bit of hash32(idx ^ key); idx = linear NDHWC index."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    h = hash32(idx ^ np.uint64(key & 0xFFFFFFFF))
    return ((h >> np.uint64(31)) & np.uint64(1)).astype(np.float32).reshape(shape)


def dropout_key(seed: int, step: int, pass_id: int, layer_id: int) -> int:
    """Host-side key derivation (mirrored in csrc/tem_runtime.cu)."""
    k = (seed * 0x9E3779B97F4A7C15 + step * 0xD1B54A32D192ED03
         + pass_id * 0x94D049BB133111EB + layer_id * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    k ^= k >> 31
    return int(hash32(np.uint64((k ^ (k >> 32)) & 0xFFFFFFFF)))


# --------------------------------------------------------------------------
# torch helpers (channels-last API <-> NC... internal)
# --------------------------------------------------------------------------
def _to_nc(x: torch.Tensor) -> torch.Tensor:
    nd = x.dim()
    return x.permute(0, nd - 1, *range(1, nd - 1)).contiguous()


def _to_cl(x: torch.Tensor) -> torch.Tensor:
    nd = x.dim()
    return x.permute(0, *range(2, nd), 1).contiguous()


def _w_conv(w: torch.Tensor) -> torch.Tensor:
    # Keras [k.., Cin, Cout] -> torch [Cout, Cin, k..]
    nd = w.dim()
    return w.permute(nd - 1, nd - 2, *range(0, nd - 2)).contiguous()


def _w_convT(w: torch.Tensor) -> torch.Tensor:
    # Keras [k.., Cout, Cin] -> torch conv_transpose weight [Cin, Cout, k..]
    nd = w.dim()
    return w.permute(nd - 1, nd - 2, *range(0, nd - 2)).contiguous()


def lrelu(x: torch.Tensor, slope: float) -> torch.Tensor:
    if slope == 1.0:
        return x
    return torch.where(x > 0, x, x * slope)


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bf16 and back, differentiable as identity."""
    y = x.detach().to(torch.float32).to(torch.bfloat16).to(x.dtype)
    return x + (y - x.detach())


Quant = Optional[Callable[[torch.Tensor], torch.Tensor]]


def apply_layer(L: LayerSpec, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor],
                is3d: bool, training: bool, mask: Optional[torch.Tensor],
                sign_ref: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One layer on NC... tensors.  mask is the dropout keep-mask (NC... layout).
    sign_ref (NC... layout, optional): take the LeakyReLU branch of every element from the sign of this externally stored
    activation instead of from the freshly computed pre-activation (test hook: "evaluate at the stored values" extended to
    the activation signs; it only matters for pre-activations within rounding noise of zero)."""
    if L.kind == 'conv':
        f = F.conv3d if is3d else F.conv2d
        y = f(x, _w_conv(w), bias=b, stride=L.stride)             # VALID
    else:
        # Keras Conv*DTranspose(k=4, s=2, 'same'): out = 2n, y[2i+k-1] += x[i] w[k] [upstream]
        f = F.conv_transpose3d if is3d else F.conv_transpose2d
        y = f(x, _w_convT(w), stride=L.stride, padding=1)
    if L.dropout and training:
        if mask is not None:
            y = y * mask * 2.0                                     # Dropout(0.5): keep, scale 1/(1-p)
    if sign_ref is not None and L.slope != 1.0:
        return y * torch.where(sign_ref > 0, torch.ones_like(y), torch.full_like(y, L.slope))
    return lrelu(y, L.slope)


def center_crop(x: torch.Tensor, total: int, is3d: bool) -> torch.Tensor:
    """generator.py:74-83: crop1 = total//2 low, crop1 (+1 if odd) high.  NC... layout."""
    lo = total // 2
    hi = total - lo
    if total == 0:
        return x
    if is3d:
        return x[:, :, lo:x.shape[2] - hi, lo:x.shape[3] - hi, lo:x.shape[4] - hi]
    return x[:, :, lo:x.shape[2] - hi, lo:x.shape[3] - hi]


def generator_forward(params: Sequence[torch.Tensor], x_cl: torch.Tensor, wf: int = 8,
                      is3d: bool = True, training: bool = False,
                      masks: Optional[Dict[str, torch.Tensor]] = None,
                      quant: Quant = None, qweights: bool = False,
                      acts: Optional[Dict[str, torch.Tensor]] = None,
                      graph_acts: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """generator.py:22-117.  x_cl: [B,(Z,)Y,X,1].  masks: {'g6','g9'} keep-masks in
    channels-last layout.  quant: applied to every stored activation except the
    final linear output (mimics bf16 storage); qweights: bf16-round the kernels.
    acts (optional dict) receives every layer output in channels-last layout."""
    layers = generator_layers(wf)
    q = quant if quant is not None else (lambda t: t)
    W = [bf16_round(p) if qweights else p for p in params]
    n = x_cl.shape[1]
    dims = generator_dims(n)
    x = _to_nc(x_cl)

    def run(i, t):
        L = layers[i]
        m = None
        if L.dropout and training and masks is not None and L.name in masks:
            m = _to_nc(masks[L.name])
        y = apply_layer(L, t, W[i], None, is3d, training, m)
        if i != len(layers) - 1:
            y = q(y)
        if acts is not None:
            acts[L.name] = _to_cl(y)
        if graph_acts is not None:
            if y.requires_grad:
                y.retain_grad()
            graph_acts[L.name] = y          # NC.. tensor inside the autograd graph
        return y

    a0 = run(0, x)
    a1 = run(1, a0)      # skip0
    a2 = run(2, a1)
    a3 = run(3, a2)      # skip1
    a4 = run(4, a3)
    a5 = run(5, a4)
    a6 = run(6, a5)
    cat1 = torch.cat([a6, center_crop(a3, dims['crop1'], is3d)], dim=1)   # generator.py:84
    a7 = run(7, cat1)
    a8 = run(8, a7)
    a9 = run(9, a8)
    cat0 = torch.cat([a9, center_crop(a1, dims['crop0'], is3d)], dim=1)
    a10 = run(10, cat0)
    a11 = run(11, a10)
    return _to_cl(a11)


def discriminator_forward(params: Sequence[torch.Tensor], x_cl: torch.Tensor, wf: int = 8,
                          is3d: bool = True, quant: Quant = None, qweights: bool = False,
                          acts: Optional[Dict[str, torch.Tensor]] = None,
                          sign_acts: Optional[Dict[str, np.ndarray]] = None) -> torch.Tensor:
    """discriminator.py:14-105 (disc_prior=None).  params order: d0..d8 kernels, d8 bias last.
    sign_acts: {layer name: stored channels-last activation} -> apply_layer(sign_ref=...) for those layers."""
    layers = discriminator_layers(wf, is3d)
    q = quant if quant is not None else (lambda t: t)
    W = [bf16_round(p) if qweights else p for p in params[:9]]
    bias = params[9]
    x = _to_nc(x_cl)
    start = 0 if is3d else 2          # 2D: HACK conv on the raw input, block "1" dead
    t = x
    for i in range(start, 9):
        L = layers[i]
        sr = None
        if sign_acts is not None and L.name in sign_acts:
            sr = _to_nc(torch.as_tensor(np.asarray(sign_acts[L.name]), dtype=t.dtype))
        t = apply_layer(L, t, W[i], bias if L.bias else None, is3d, True, None, sign_ref=sr)
        if i != 8:
            t = q(t)
        if acts is not None:
            acts[L.name] = _to_cl(t)
    return _to_cl(t)


# --------------------------------------------------------------------------
# losses (cgan.py:78-81,110-142 via tfa.losses.SigmoidFocalCrossEntropy [upstream])
# --------------------------------------------------------------------------
def focal_logits(target: float, logits: torch.Tensor, gamma: float = 2.0, alpha: float = 0.5) -> torch.Tensor:
    """tfa sigmoid_focal_crossentropy(from_logits=True), mean reduction.
    ce = max(x,0) - x z + log1p(exp(-|x|)); p_t = z p + (1-z)(1-p);
    loss = mean(alpha_t (1-p_t)^gamma ce)   (last axis has size 1)."""
    x = logits
    z = target
    ce = torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x)))
    p = torch.sigmoid(x)
    p_t = z * p + (1 - z) * (1 - p)
    alpha_t = z * alpha + (1 - z) * (1 - alpha)
    return (alpha_t * torch.pow(1.0 - p_t, gamma) * ce).mean()


def focal_probs_target1(t: torch.Tensor, gamma: float = 2.0, alpha: float = 0.5) -> torch.Tensor:
    """tfa sigmoid_focal_crossentropy(from_logits=False) against ones, mean reduction.
    Keras binary_crossentropy clips to [eps, 1-eps] then -log(p + eps) [upstream]."""
    eps = KERAS_EPS
    ce = -torch.log(torch.clamp(t, eps, 1.0 - eps) + eps)
    return (alpha * torch.pow(1.0 - t, gamma) * ce).mean()


def generator_loss(disc_generated, gamma=2.0):      # cgan.py:119-120
    return focal_logits(1.0, disc_generated, gamma) * 2


def discriminator_loss(real, generated, gamma=2.0):  # cgan.py:110-117
    real_loss = focal_logits(1.0, real, gamma) * 2
    gen_loss = focal_logits(0.0, generated, gamma) * 2
    return (real_loss + gen_loss) * 0.5


def identity_loss(real_image, same_image, gamma=2.0):  # cgan.py:122-131
    LAMBDA = 2
    tconf = 1 - torch.abs(real_image - same_image) / 2
    return LAMBDA * 0.5 * (focal_probs_target1(tconf, gamma) * 2)


def calc_cycle_loss(real_image, cycled_image, gamma=2.0):  # cgan.py:133-142
    LAMBDA = 2
    tconf = 1 - torch.abs(real_image - cycled_image) / 2
    return LAMBDA * (focal_probs_target1(tconf, gamma) * 2)


# north-star "lsgan_l1" mode (dead-code docstrings cgan.py:124-126,135-137 + LSGAN)
def lsgan_generator_loss(d_fake):
    return ((d_fake - 1.0) ** 2).mean()


def lsgan_discriminator_loss(d_real, d_fake):
    return 0.5 * (((d_real - 1.0) ** 2).mean() + (d_fake ** 2).mean())


def l1_identity_loss(real, same):
    return 1 * 0.5 * torch.abs(real - same).mean()


def l1_cycle_loss(real, cycled):
    return 1 * torch.abs(real - cycled).mean()


# --------------------------------------------------------------------------
# train step (cgan.py:144-230)
# --------------------------------------------------------------------------
def crop_cl(x: torch.Tensor, c: int) -> torch.Tensor:
    if c == 0:
        return x
    if x.dim() == 5:
        return x[:, c:-c, c:-c, c:-c, :]
    return x[:, c:-c, c:-c, :]


def pad_cl(x: torch.Tensor, p: int) -> torch.Tensor:
    if x.dim() == 5:
        return F.pad(x, (0, 0, p, p, p, p, p, p))
    return F.pad(x, (0, 0, p, p, p, p))


@dataclass
class StepResult:
    losses: List[float]                  # order of cgan.py:230
    grads: Dict[str, List[np.ndarray]]   # 'g','f','dx','dy'
    outputs: Dict[str, np.ndarray] = field(default_factory=dict)


def train_step_grads(P: Dict[str, List[np.ndarray]], real_x: np.ndarray, real_y: np.ndarray,
                     wf: int = 8, is3d: bool = True, gamma: float = 2.0,
                     masks: Optional[Dict[str, Dict[str, np.ndarray]]] = None,
                     loss_mode: str = 'focal', dtype=torch.float64, literal: bool = False,
                     quant: Quant = None, qweights: bool = False,
                     keep_outputs: bool = False,
                     override_fakes: Optional[Dict[str, np.ndarray]] = None,
                     disc_signs: Optional[Dict[str, Dict[str, np.ndarray]]] = None) -> StepResult:
    """Forward + gradients of cgan.py:148-215.  P = {'g','f','dx','dy'} parameter lists.

    masks: {pass_name: {'g6': keep, 'g9': keep}} for pass_name in
    ('g_realx','f_fakey','f_realy','g_fakex','f_realx','g_realy'); missing => no dropout.
    literal=True evaluates the four tape.gradient calls separately (cgan.py:207-215);
    literal=False does the single combined backward used by the CUDA path.
    """
    T = {k: [torch.tensor(np.asarray(a), dtype=dtype, requires_grad=True) for a in v] for k, v in P.items()}
    rx = torch.tensor(real_x, dtype=dtype)
    ry = torch.tensor(real_y, dtype=dtype)
    n = rx.shape[1]
    out = generator_out_dim(n)
    buf = (n - out) // 2                                 # cgan.py:65
    masks = masks or {}

    def mk(name):
        m = masks.get(name)
        if m is None:
            return None
        return {k: torch.tensor(v, dtype=dtype) for k, v in m.items()}

    def G(x, name):
        m = mk(name)
        return generator_forward(T['g'], x, wf, is3d, training=m is not None, masks=m, quant=quant, qweights=qweights)

    def Fn(x, name):
        m = mk(name)
        return generator_forward(T['f'], x, wf, is3d, training=m is not None, masks=m, quant=quant, qweights=qweights)

    ds = disc_signs or {}     # {'dx_real' | 'dy_real' | 'dx_fake' | 'dy_fake': {layer: stored activation}} (see apply_layer)

    def Dx(x, which):
        return discriminator_forward(T['dx'], x, wf, is3d, quant=quant, qweights=qweights, sign_acts=ds.get('dx_' + which))

    def Dy(x, which):
        return discriminator_forward(T['dy'], x, wf, is3d, quant=quant, qweights=qweights, sign_acts=ds.get('dy_' + which))

    def subst(name, t):
        # test hook: evaluate the downstream graph at externally supplied values of a fake (straight-through
        # for the gradient) so that both implementations see bit-identical second-pass inputs
        if override_fakes and name in override_fakes:
            return t + (torch.tensor(override_fakes[name], dtype=dtype) - t).detach()
        return t

    fake_y = subst('fake_y', G(rx, 'g_realx'))           # cgan.py:152
    cycled_x = Fn(pad_cl(fake_y, buf), 'f_fakey')        # :161-162
    cycled_x_c = crop_cl(cycled_x, buf)                  # :163
    rx_c2 = crop_cl(rx, 2 * buf)                         # :165
    fake_x = subst('fake_x', Fn(ry, 'f_realy'))          # :167
    cycled_y = G(pad_cl(fake_x, buf), 'g_fakex')         # :170-171
    cycled_y_c = crop_cl(cycled_y, buf)
    ry_c2 = crop_cl(ry, 2 * buf)
    same_x = Fn(rx, 'f_realx')                           # :177
    rx_c = crop_cl(rx, buf)
    same_y = G(ry, 'g_realy')                            # :181
    ry_c = crop_cl(ry, buf)
    d_real_x = Dx(rx_c, 'real'); d_real_y = Dy(ry_c, 'real')             # :185-186
    d_fake_x = Dx(fake_x, 'fake'); d_fake_y = Dy(fake_y, 'fake')         # :188-189

    if loss_mode == 'focal':
        gen_g = generator_loss(d_fake_y, gamma); gen_f = generator_loss(d_fake_x, gamma)
        cyc = calc_cycle_loss(rx_c2, cycled_x_c, gamma) + calc_cycle_loss(ry_c2, cycled_y_c, gamma)
        id_g = identity_loss(ry_c, same_y, gamma); id_f = identity_loss(rx_c, same_x, gamma)
        disc_x = discriminator_loss(d_real_x, d_fake_x, gamma)
        disc_y = discriminator_loss(d_real_y, d_fake_y, gamma)
    else:
        gen_g = lsgan_generator_loss(d_fake_y); gen_f = lsgan_generator_loss(d_fake_x)
        cyc = l1_cycle_loss(rx_c2, cycled_x_c) + l1_cycle_loss(ry_c2, cycled_y_c)
        id_g = l1_identity_loss(ry_c, same_y); id_f = l1_identity_loss(rx_c, same_x)
        disc_x = lsgan_discriminator_loss(d_real_x, d_fake_x)
        disc_y = lsgan_discriminator_loss(d_real_y, d_fake_y)
    total_g = gen_g + cyc + id_g                         # :199
    total_f = gen_f + cyc + id_f                         # :200

    if literal:
        gg = torch.autograd.grad(total_g, T['g'], retain_graph=True)
        gf = torch.autograd.grad(total_f, T['f'], retain_graph=True)
    else:
        comb = gen_g + gen_f + cyc + id_g + id_f
        both = torch.autograd.grad(comb, T['g'] + T['f'], retain_graph=True)
        gg, gf = both[:len(T['g'])], both[len(T['g']):]
    gdx = torch.autograd.grad(disc_x, T['dx'], retain_graph=True, allow_unused=True)
    gdy = torch.autograd.grad(disc_y, T['dy'], allow_unused=True)

    def npl(gs, ps):
        return [(g if g is not None else torch.zeros_like(p)).detach().numpy() for g, p in zip(gs, ps)]

    res = StepResult(
        losses=[float(v) for v in (total_g, total_f, disc_y, disc_x, gen_g, gen_f, cyc)],   # cgan.py:230
        grads={'g': npl(gg, T['g']), 'f': npl(gf, T['f']), 'dx': npl(gdx, T['dx']), 'dy': npl(gdy, T['dy'])})
    if keep_outputs:
        res.outputs = {k: v.detach().numpy() for k, v in dict(
            fake_y=fake_y, cycled_x=cycled_x, fake_x=fake_x, cycled_y=cycled_y, same_x=same_x, same_y=same_y,
            d_real_x=d_real_x, d_real_y=d_real_y, d_fake_x=d_fake_x, d_fake_y=d_fake_y).items()}
    return res


# --------------------------------------------------------------------------
# Keras Adam (cgan.py:69-73, 218-228) [upstream defaults beta2=.999, eps=1e-7]
# --------------------------------------------------------------------------
def keras_adam_update(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, t: int,
                      lr=2e-4, b1=0.5, b2=0.999, eps=1e-7):
    """One step; t is the 1-based step count after increment.  Arithmetic in p.dtype."""
    dt = p.dtype.type
    m = dt(b1) * m + dt(1 - b1) * g
    v = dt(b2) * v + dt(1 - b2) * g * g
    lr_t = dt(lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    p = p - lr_t * m / (np.sqrt(v) + dt(eps))
    return p, m, v


class OracleEM2EM:
    """Minimal CPU EM2EM (cgan.py:32-293) used for loss-curve parity and the CPU baseline."""

    def __init__(self, dimsize=74, is3d=True, wf=8, focal_gamma=2.0, seed=0, dtype=np.float32,
                 loss_mode='focal', quant: Quant = None):
        if dimsize < 74:
            raise RuntimeError("minimum dimension allowed is 74")      # cgan.py:52-53
        rng = np.random.default_rng(seed)
        self.is3d, self.wf, self.gamma, self.dtype, self.loss_mode = is3d, wf, focal_gamma, dtype, loss_mode
        self.quant = quant
        self.layers = {'g': generator_layers(wf), 'f': generator_layers(wf),
                       'dx': discriminator_layers(wf, is3d), 'dy': discriminator_layers(wf, is3d)}
        self.P = {k: init_params(self.layers[k], is3d, rng, dtype) for k in ('g', 'f', 'dx', 'dy')}
        self.M = {k: [np.zeros_like(a) for a in v] for k, v in self.P.items()}
        self.V = {k: [np.zeros_like(a) for a in v] for k, v in self.P.items()}
        self.t = 0
        self.outdimsize = generator_out_dim(dimsize)
        self.buffer = (dimsize - self.outdimsize) // 2

    def train_step(self, real_x, real_y, masks=None):
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        r = train_step_grads(self.P, real_x, real_y, self.wf, self.is3d, self.gamma, masks,
                             self.loss_mode, dtype=tdt, quant=self.quant)
        self.t += 1
        for k in self.P:
            for i in range(len(self.P[k])):
                g = r.grads[k][i].astype(self.dtype)
                self.P[k][i], self.M[k][i], self.V[k][i] = keras_adam_update(
                    self.P[k][i], g, self.M[k][i], self.V[k][i], self.t)
        return r.losses

    def predict(self, data):                                      # cgan.py:289-293
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        with torch.no_grad():
            y = generator_forward([torch.tensor(p, dtype=tdt) for p in self.P['g']],
                                  torch.tensor(np.asarray(data), dtype=tdt), self.wf, self.is3d, training=False)
        return y.numpy()


# --------------------------------------------------------------------------
# uint8 conventions + tiling (datasets.py:157-171,193-202; utils.py:41-130)
# --------------------------------------------------------------------------
def scale_tensor(u8: np.ndarray) -> np.ndarray:
    """datasets.py:193-202: float32(u8)/127.5 - 1, add channel."""
    t = u8.astype(np.float32)
    t = (t / np.float32(127.5)) - np.float32(1)
    return t[..., None]


def standardize_population(t: np.ndarray, meanstd) -> np.ndarray:   # datasets.py:157-163
    mean, std = np.float32(meanstd[0]), np.float32(meanstd[1])
    return (t - mean) / std


def unstandardize_population(t: np.ndarray, meanstd) -> np.ndarray:  # datasets.py:165-171
    mean, std = np.float32(meanstd[0]), np.float32(meanstd[1])
    return t * std + mean


def to_uint8_reference(y_std: np.ndarray, meanstd_y) -> np.ndarray:
    """utils.py:109,118: (y*std+mean+1)*127.5 -> np.around (half-even) -> astype(uint8).
    astype(uint8) of out-of-range floats is implementation-defined in numpy; the
    reference behaviour on x86 wraps modulo 256 via an int conversion - restated
    explicitly as int64 -> & 0xFF so it is deterministic."""
    v = (unstandardize_population(y_std.astype(np.float32), meanstd_y) + np.float32(1)) * np.float32(127.5)
    r = np.around(v)
    return (r.astype(np.int64) & 0xFF).astype(np.uint8)


def tiling_plan(start, size, outdimsize: int, buffer: int):
    """utils.py:68-84: returns (outdimsize', tpad, buffer', rois, index) with xyz tuples."""
    tpad = 0
    if (outdimsize // 6) != 0:                      # literal (always true) utils.py:71
        diff = outdimsize % 6
        outdimsize -= diff
        tpad = diff // 2
        buffer += tpad
    rois, index = [], []
    for xi in range(start[0], start[0] + size[0], outdimsize):
        for yi in range(start[1], start[1] + size[1], outdimsize):
            for zi in range(start[2], start[2] + size[2], outdimsize):
                rois.append((xi - buffer, yi - buffer, zi - buffer))
                index.append((xi - start[0], yi - start[1], zi - start[2]))
    return outdimsize, tpad, buffer, rois, index


def predict_ng_cube_oracle(volume_zyx: np.ndarray, start, size, predict_fn, meanstd_x, meanstd_y,
                           outdimsize: int = 40, buffer: int = 17, fetch_input: bool = False):
    """utils.py:41-130 with `location` replaced by an in-memory uint8[z,y,x] volume.

    predict_fn maps a standardised fp32 [1,n,n,n,1] tile to [1,n-34,...,1].
    Tiles that reach outside the volume read zeros (tensorstore would fail; the
    synthetic configs keep every tile in bounds)."""
    od, tpad, buf, rois, index = tiling_plan(start, size, outdimsize, buffer)
    tsz = od + 2 * buf
    z, y, x = size[2], size[1], size[0]
    if size[0] % od: x += od - size[0] % od
    if size[1] % od: y += od - size[1] % od
    if size[2] % od: z += od - size[2] % od
    out_buffer = np.zeros((z, y, x), np.uint8)
    in_buffer = np.zeros((z, y, x), np.uint8) if fetch_input else None
    VZ, VY, VX = volume_zyx.shape
    for roi, idx in zip(rois, index):
        x0, y0, z0 = roi
        tile = np.zeros((tsz, tsz, tsz), np.uint8)
        zs, ze = max(z0, 0), min(z0 + tsz, VZ)
        ys, ye = max(y0, 0), min(y0 + tsz, VY)
        xs, xe = max(x0, 0), min(x0 + tsz, VX)
        if zs < ze and ys < ye and xs < xe:
            tile[zs - z0:ze - z0, ys - y0:ye - y0, xs - x0:xe - x0] = volume_zyx[zs:ze, ys:ye, xs:xe]
        data_x = standardize_population(scale_tensor(tile), meanstd_x)[None]      # [1,n,n,n,1]
        data_y = predict_fn(data_x)
        u8 = to_uint8_reference(data_y, meanstd_y)
        if tpad > 0:
            u8 = u8[:, tpad:-tpad, tpad:-tpad, tpad:-tpad, :]
        out_buffer[idx[2]:idx[2] + od, idx[1]:idx[1] + od, idx[0]:idx[0] + od] = u8[0, :, :, :, 0]
        if fetch_input:
            dx = (unstandardize_population(data_x, meanstd_x) + np.float32(1)) * np.float32(127.5)
            b = dx[0, buf:od + buf, buf:od + buf, buf:od + buf, 0]
            in_buffer[idx[2]:idx[2] + od, idx[1]:idx[1] + od, idx[0]:idx[0] + od] = b.astype(np.uint8)  # truncates utils.py:123-125
    if fetch_input:
        return in_buffer[0:size[2], 0:size[1], 0:size[0]], out_buffer[0:size[2], 0:size[1], 0:size[0]]
    return out_buffer[0:size[2], 0:size[1], 0:size[0]]


def augment(t: np.ndarray, perm, flip, var_adj, mean_adj) -> np.ndarray:
    """datasets.py:123-155 with the random choices made explicit: t [*spatial, 1] float32 -> transpose(perm + [channel]) ->
    reverse the flipped axes -> *= var_adj -> += mean_adj (two float32 roundings)."""
    nd = t.ndim - 1
    out = np.transpose(t, tuple(perm) + (nd,))
    for d in range(nd):
        if flip[d]:
            out = np.flip(out, axis=d)
    out = out.astype(np.float32) * np.float32(var_adj)
    out = out + np.float32(mean_adj)
    return np.ascontiguousarray(out.astype(np.float32))


def warp_tensor(t: np.ndarray, uniform: np.ndarray, rate: float = 4 / (128 * 128)) -> np.ndarray:
    """transfer_em/debug.py:7-63 with the random draw made explicit.  t: float32 [Z,Y,X,1] or [Y,X,1] -> conv 'SAME' (zero
    padding) with ones(3^d)/3^d -> mask = uniform < rate -> conv 'SAME' with ones(4^d) (TF pads 1 before, 2 after: window
    offsets -1..+2) -> where(mask > 0, mean(blurred), blurred).  fp32 throughout; TF's summation order inside the convolution
    is not specified, so parity with this restatement is to ~1e-6, not bit-exact (parity unpinned, see the module header)."""
    nd = t.ndim - 1
    x = np.asarray(t, np.float32)[..., 0]
    sp = x.shape
    w = np.float32(1.0) / np.float32(3 ** nd)
    xp = np.pad(x, 1)
    blur = np.zeros(sp, np.float32)
    for off in itertools.product(range(3), repeat=nd):
        sl = tuple(slice(o, o + n) for o, n in zip(off, sp))
        blur = (blur + xp[sl] * w).astype(np.float32)
    seeds = (np.asarray(uniform, np.float32).reshape(sp) < np.float32(rate))
    sp_ = np.pad(seeds, [(1, 2)] * nd)
    hole = np.zeros(sp, bool)
    for off in itertools.product(range(4), repeat=nd):
        sl = tuple(slice(o, o + n) for o, n in zip(off, sp))
        hole |= sp_[sl]
    mean = np.float32(blur.mean(dtype=np.float64))
    return np.where(hole, mean, blur).astype(np.float32)[..., None]


def get_meanstd(tensors: Sequence[np.ndarray]):
    """datasets.py:173-190: mean of per-tensor means, sqrt of mean of per-tensor variances."""
    mean = np.float32(0); var = np.float32(0)
    for t in tensors:
        mean += np.float32(np.mean(t, dtype=np.float32)); var += np.float32(np.var(t, dtype=np.float32))
    mean /= len(tensors); var /= len(tensors)
    return float(mean), float(np.sqrt(var))
