"""Model objects and block helpers mirroring transfer_em/models/utils.py.

The reference returns Keras models; here a model is a view (engine, net id) onto a tem_handle whose
arithmetic runs in libtem_b200.  ``downsample`` / ``upsample`` (utils.py:41,89) return callable
block models with their own N(0, 0.02) kernels, executed layer by layer through the per-op entry
point ``tem_conv_forward``; inside a generator / discriminator the same layers run as part of the
handle's fused pass.
"""
from dataclasses import dataclass

import numpy as np
import torch

from .._lib import NET_G, NET_F, NET_DX, NET_DY  # noqa: F401


@dataclass
class BlockLayer:
    kind: str          # 'conv' | 'convT'
    kernel: int
    stride: int
    padding: str       # 'valid' | 'same'
    filters: int
    activation: str    # 'leaky_relu(0.3)'
    dropout: float = 0.0


class _Var:
    """One kernel variable (Keras layout), shared between the two models `downsample` returns (utils.py:85: both Keras
    models are built over the same first Conv layer)."""

    def __init__(self, shape, rng):
        self.value = (rng.standard_normal(shape) * 0.02).astype(np.float32)      # random_normal_initializer(0., 0.02)
        self._dev = None

    def device(self, dev):
        if self._dev is None or self._dev.device != dev:
            self._dev = torch.from_numpy(self.value).to(dev)
        return self._dev

    def assign(self, w):
        w = np.asarray(w, np.float32)
        if w.shape != self.value.shape:
            raise ValueError(f"expected shape {self.value.shape}, got {w.shape}")
        self.value[...] = w
        self._dev = None


class Block:
    """Callable block model: the stand-in for the Keras models `downsample` / `upsample` return (utils.py:85,137).
    `block(x, training=False)` runs its convolutions through `tem_conv_forward` (one launch per layer, LeakyReLU and the
    inverted dropout fused into the epilogue); x is [B, (n,) n, n, infilters] float32 / bfloat16, numpy or torch."""
    _ctr = 0

    def __init__(self, name, infilters, layers, is3d, variables):
        self.name, self.infilters, self.layers, self.is3d, self._vars = name, infilters, list(layers), is3d, list(variables)

    # Keras-like surface
    @property
    def trainable_variables(self):
        return [v.value for v in self._vars]

    def get_weights(self):
        return [v.value.copy() for v in self._vars]

    def set_weights(self, weights):
        if len(weights) != len(self._vars):
            raise ValueError(f"expected {len(self._vars)} arrays")
        for v, w in zip(self._vars, weights):
            v.assign(w)

    def count_params(self):
        return int(sum(v.value.size for v in self._vars))

    def output_dim(self, n):
        for L in self.layers:
            n = n * L.stride if L.kind == 'convT' else (n - L.kernel) // L.stride + 1
        return n

    def __call__(self, x, training=False, dropout_key=None):
        import ctypes as C
        from .. import _lib
        from ..engine import _stream
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TemError("no CUDA device: transfer_em_b200 has no CPU fallback")
        was_np = isinstance(x, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(x)) if was_np else x
        dev = t.device if t.device.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        nd = 3 if self.is3d else 2
        if t.dim() != nd + 2 or t.shape[-1] != self.infilters:
            raise ValueError(f"expected [B,{'n,' * nd}{self.infilters}], got {tuple(t.shape)}")
        cur = t.to(dev).to(torch.float32 if self.infilters == 1 else torch.bfloat16).contiguous()
        B = int(cur.shape[0])
        dims = [1] * (3 - nd) + [int(v) for v in cur.shape[1:-1]]
        cin = self.infilters
        for li, (L, v) in enumerate(zip(self.layers, self._vars)):
            d = _lib.TemConvDesc()
            d.B = B
            for i in range(3):
                d.in_dims[i] = dims[i]
                d.k[i] = L.kernel if (self.is3d or i > 0) else 1
                d.stride[i] = L.stride if (self.is3d or i > 0) else 1
            d.cin, d.cout, d.transposed, d.slope = cin, L.filters, int(L.kind == 'convT'), 0.3
            d.dropout_key = 0
            if training and L.dropout > 0:
                if dropout_key is None:
                    Block._ctr += 1
                    dropout_key = (0x9E3779B1 * Block._ctr) & 0xFFFFFFFF or 1
                d.dropout_key = int(dropout_key)
            d.in_dtype = _lib.TEM_F32 if cur.dtype == torch.float32 else _lib.TEM_BF16
            d.out_dtype = _lib.TEM_BF16
            d.meanstd[0], d.meanstd[1] = 0.0, 1.0
            d.use_tensor_cores = 1
            od = (C.c_int32 * 3)()
            _lib.check(lib.tem_conv_forward(C.byref(d), None, None, None, None, od, _stream()))
            out = torch.empty((B, od[0], od[1], od[2], L.filters), dtype=torch.bfloat16, device=dev)
            _lib.check(lib.tem_conv_forward(C.byref(d), C.c_void_p(cur.data_ptr()), C.c_void_p(v.device(dev).data_ptr()), None,
                                            C.c_void_p(out.data_ptr()), od, _stream()))
            cur, dims, cin = out, [od[0], od[1], od[2]], L.filters
        res = cur.to(torch.float32)
        if not self.is3d:
            res = res.reshape((B,) + tuple(dims[1:]) + (cin,))
        return res.cpu().numpy() if was_np else res

    predict = __call__


def _kshape(kind, k, cin, cout, is3d):
    nd = 3 if is3d else 2
    return (k,) * nd + ((cin, cout) if kind == 'conv' else (cout, cin))      # Keras: convT kernels are [k.., Cout, Cin]


def downsample(id, infilters, outfilters, is3d, filter_size=4, norm_type='instancenorm', apply_norm=True, *, seed=None):
    """transfer_em/models/utils.py:41-85: conv3 VALID + LReLU (skip output) -> conv(filter_size) stride 2 VALID + LReLU.
    Returns (down_model, skip_model) like the reference: two callables sharing the first convolution's kernel.
    norm_type / apply_norm are accepted and ignored, as in the reference (every normalisation call is commented out there)."""
    rng = np.random.default_rng(seed)
    l0 = BlockLayer('conv', 3, 1, 'valid', outfilters, 'leaky_relu(0.3)')
    l1 = BlockLayer('conv', filter_size, 2, 'valid', outfilters, 'leaky_relu(0.3)')
    v0 = _Var(_kshape('conv', 3, infilters, outfilters, is3d), rng)
    v1 = _Var(_kshape('conv', filter_size, outfilters, outfilters, is3d), rng)
    skip = Block(f"Downsample_{id}_skip", infilters, [l0], is3d, [v0])
    down = Block(f"Downsample_{id}", infilters, [l0, l1], is3d, [v0, v1])
    return down, skip


def upsample(id, infilters, outfilters, is3d, filter_size=4, norm_type='instancenorm', apply_dropout=True, *, seed=None):
    """transfer_em/models/utils.py:89-137: conv3 VALID (2*outfilters) + LReLU -> convT(filter_size) stride 2 SAME
    -> Dropout(0.5) -> LReLU.  Returns a callable model; dropout is live only with training=True."""
    if filter_size != 4:
        raise NotImplementedError("the transposed convolution kernels take filter_size=4 (the only size the reference uses)")
    rng = np.random.default_rng(seed)
    l0 = BlockLayer('conv', 3, 1, 'valid', outfilters * 2, 'leaky_relu(0.3)')
    l1 = BlockLayer('convT', filter_size, 2, 'same', outfilters, 'leaky_relu(0.3)', 0.5 if apply_dropout else 0.0)
    v0 = _Var(_kshape('conv', 3, infilters, outfilters * 2, is3d), rng)
    v1 = _Var(_kshape('convT', filter_size, outfilters * 2, outfilters, is3d), rng)
    return Block(f"Upsample_{id}", infilters, [l0, l1], is3d, [v0, v1])


class NetModel:
    """Callable stand-in for the Keras model of one network of an Engine."""

    def __init__(self, engine, net, kind):
        self.engine, self.net, self.kind = engine, net, kind

    @property
    def trainable_variables(self):
        return self.engine.get_weights(self.net)

    def get_weights(self):
        return self.engine.get_weights(self.net)

    def set_weights(self, weights):
        self.engine.set_weights(self.net, weights)

    def count_params(self):
        return self.engine.param_count(self.net)

    def variable_info(self):
        return self.engine.variables(self.net)

    def __call__(self, x, training=False, meanstd=None, dropout_key=None):
        if self.kind == 'generator':
            key = 0
            if training and self.engine_dropout:
                key = dropout_key if dropout_key is not None else self._next_key()
            return self.engine.gen_forward(self.net, x, meanstd=meanstd, dropout_key=key)
        return self.engine.disc_forward(self.net, x)

    predict = __call__
    engine_dropout = True
    _ctr = 0

    def _next_key(self):
        NetModel._ctr += 1
        return (0x9E3779B1 * NetModel._ctr) & 0xFFFFFFFF or 1
