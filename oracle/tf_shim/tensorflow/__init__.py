"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the `tensorflow` API surface that the reference's hot path touches,
executed with torch-CPU ops, so that the reference's OWN Python (transfer_em/cgan.py, models/*.py, utils.py,
datasets/datasets.py, debug.py under /root/reference) can be imported and run in a container that has no TensorFlow.

What this pins and what it does not.  Running the reference's code through this shim pins everything the reference itself
decides: layer order and channel widths, which layers share kernels, crop / pad / concat arithmetic, the train-step dataflow
(cgan.py:144-230), loss composition and constants (cgan.py:110-142), which variables each gradient is taken for, optimizer
wiring, tile origins / crops / uint8 conversion of predict_ng_cube (utils.py:68-130).  It does NOT pin TensorFlow's own
kernels: the op semantics below are restated from the published TF 2.x / Keras / tensorflow_addons behaviour
(README.md:31 names tensorflow 2.2; tfa is un-pinned in the reference, 0.10.x pairs with TF 2.2):
  * Conv2D/3D 'valid' = cross-correlation, kernel [k..,Cin,Cout], out = (n-k)//s + 1; use_bias default True, zeros
  * Conv2D/3DTranspose 'same' = gradient of a SAME strided conv: out = n*s, y[s*i + k - pad_before] += x[i] w[k],
    pad_before = (k - s)//2, kernel [k..,Cout,Cin]
  * LeakyReLU() alpha 0.3; Dropout(rate): training only, keep mask scaled by 1/(1-rate)
  * tfa SigmoidFocalCrossEntropy: alpha_t (1-p_t)^gamma * K.binary_crossentropy, summed over the last axis, then
    reduction AUTO = mean over the remaining elements; K.binary_crossentropy(from_logits=False) clips to
    [eps, 1-eps] and takes log(p + eps), eps = 1e-7
  * Keras Adam: lr_t = lr sqrt(1-b2^t)/(1-b1^t); var -= lr_t m / (sqrt(v) + eps), eps = 1e-7
  * tf.nn.convNd 'SAME' stride 1: pad_total = k-1, pad_before = pad_total//2
Only tools/make_reference_golden.py (run in the build container, writes tests/golden/) and tests may import this package;
the product never does."""
import math
import types

import numpy as np
import torch

float32 = torch.float32
uint8 = torch.uint8
newaxis = None
DROPOUT = {"enabled": True, "keys": None, "mask_fn": None, "calls": 0}    # injected by the golden generator


class TensorShape(tuple):
    @property
    def rank(self):
        return len(self)


class Tensor(torch.Tensor):
    """torch tensor with the two attributes of tf.Tensor the reference reads (.shape.rank, .numpy())."""

    @property
    def shape(self):
        return TensorShape(super().shape)

    def numpy(self):
        return self.detach().as_subclass(torch.Tensor).numpy()


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.as_tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.as_subclass(Tensor)


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def function(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


def ones_like(x): return torch.ones_like(_t(x)).as_subclass(Tensor)
def zeros_like(x): return torch.zeros_like(_t(x)).as_subclass(Tensor)
def ones(shape, dtype=float32): return torch.ones(tuple(shape), dtype=dtype).as_subclass(Tensor)
def abs(x): return torch.abs(_t(x))   # noqa: A001
def cast(x, dtype): return _t(x).to(dtype)
def expand_dims(x, axis): return _t(x).unsqueeze(axis)
def squeeze(x, axis): return _t(x).squeeze(axis[0] if isinstance(axis, (list, tuple)) else axis)
def reshape(x, shape): return _t(x).reshape(tuple(shape))
def where(c, a, b): return torch.where(c, a, b).as_subclass(Tensor)
def constant(v, shape=None, dtype=None): return _t(np.full(shape, v) if shape is not None else v, dtype)


def _reduce_mean(x): return _t(x).mean()
def _reduce_variance(x): return _t(x).var(unbiased=False)
math_ns = types.SimpleNamespace(reduce_mean=_reduce_mean, reduce_variance=_reduce_variance, sqrt=lambda x: torch.sqrt(_t(x)),
                                rsqrt=lambda x: torch.rsqrt(_t(x)), less=lambda a, b: a < b)
globals()["math"] = math_ns      # tf.math (the stdlib module stays reachable as _pymath)
import math as _pymath           # noqa: E402


class _Random:
    def __init__(self):
        self.gen = torch.Generator().manual_seed(0)

    def set_seed(self, s):
        self.gen.manual_seed(int(s))

    def uniform(self, shape, lo=0.0, hi=1.0):
        return (torch.rand(tuple(shape), generator=self.gen) * (hi - lo) + lo).as_subclass(Tensor)


random = _Random()


def random_normal_initializer(mean=0.0, stddev=0.05):
    def init(shape):
        return torch.randn(tuple(shape), generator=random.gen) * stddev + mean
    return init


def _same_pad(k):
    tot = k - 1
    return tot // 2, tot - tot // 2


def _nn_conv(x, filters, strides, padding, nd):
    x = _t(x); w = _t(filters)
    assert padding == "SAME" and all(s == 1 for s in strides)
    xc = x.movedim(-1, 1)
    ks = w.shape[:nd]
    pads = []
    for k in reversed(ks):
        b, a = _same_pad(k); pads += [b, a]
    xc = torch.nn.functional.pad(xc, pads)
    wt = w.permute(nd + 1, nd, *range(nd))
    y = (torch.nn.functional.conv3d if nd == 3 else torch.nn.functional.conv2d)(xc, wt)
    return y.movedim(1, -1).as_subclass(Tensor)


nn = types.SimpleNamespace(conv3d=lambda x, f, s, p: _nn_conv(x, f, s, p, 3), conv2d=lambda x, f, s, p: _nn_conv(x, f, s, p, 2),
                           moments=lambda x, axes, keepdims=False: (x.mean(axes, keepdim=keepdims), x.var(axes, unbiased=False, keepdim=keepdims)))
data = types.SimpleNamespace(experimental=types.SimpleNamespace(AUTOTUNE=-1))
config = types.SimpleNamespace(experimental_run_functions_eagerly=lambda flag: None, run_functions_eagerly=lambda flag: None)


# ---------------------------------------------------------------------------------------------------------------------
# autodiff
# ---------------------------------------------------------------------------------------------------------------------
class GradientTape:
    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def gradient(self, target, sources):
        gs = torch.autograd.grad(target, list(sources), retain_graph=True, allow_unused=True)
        return [None if g is None else g.detach() for g in gs]


# ---------------------------------------------------------------------------------------------------------------------
# keras: functional graph of symbolic nodes, layers, Model, Adam
# ---------------------------------------------------------------------------------------------------------------------
class _Sym:
    """Symbolic tensor of the functional API: (producer, inputs) + static channel count."""

    def __init__(self, shape, producer=None, inputs=()):
        self.shape = TensorShape(shape); self.producer = producer; self.inputs = tuple(inputs)


def _eval(sym, feed, training, memo):
    if id(sym) in memo:
        return memo[id(sym)]
    if id(sym) in feed:
        v = feed[id(sym)]
    else:
        args = [_eval(s, feed, training, memo) for s in sym.inputs]
        v = sym.producer._run(args if sym.producer._multi else args[0], training)
    memo[id(sym)] = v
    return v


class Layer:
    _multi = False

    def __init__(self, name=None, **kw):
        self.name = name; self._weights = []; self.built = False

    def add_weight(self, name, shape, initializer, trainable=True):
        init = initializer if callable(initializer) else (lambda s: torch.zeros(tuple(s)))
        w = init(shape).detach().to(torch.float32).requires_grad_(trainable)
        w._tf_name = name
        self._weights.append(w)
        return w

    def build(self, input_shape):
        pass

    def out_shape(self, shape):
        return shape

    def _layers(self):
        return [self]

    def _run(self, x, training):
        return self.call(x)

    def __call__(self, x, training=False):
        shp = [s.shape for s in x] if self._multi else x.shape
        if not self.built:
            self.build(TensorShape(shp) if not self._multi else shp); self.built = True
        if isinstance(x, _Sym) or (self._multi and isinstance(x[0], _Sym)):
            return _Sym(self.out_shape(shp), self, x if self._multi else (x,))
        xx = [_t(v) for v in x] if self._multi else _t(x)
        return self._run(xx, training).as_subclass(Tensor)


class _Conv(Layer):
    nd = 3; transposed = False

    def __init__(self, filters, kernel_size, strides=1, padding="valid", kernel_initializer=None, use_bias=True, name=None):
        super().__init__(name)
        self.filters, self.k, self.s, self.padding, self.init, self.use_bias = filters, kernel_size, strides, padding.lower(), kernel_initializer, use_bias

    def build(self, shape):
        cin = shape[-1]
        ks = (self.k,) * self.nd
        kshape = ks + ((cin, self.filters) if not self.transposed else (self.filters, cin))
        self.kernel = self.add_weight("kernel", kshape, self.init or random_normal_initializer(0.0, 0.05))
        self.bias = self.add_weight("bias", (self.filters,), "zeros") if self.use_bias else None

    def out_shape(self, shape):
        return tuple(shape[:-1]) + (self.filters,)

    def call(self, x):
        nd = self.nd
        xc = x.movedim(-1, 1)
        f = torch.nn.functional
        if not self.transposed:
            assert self.padding == "valid"
            w = self.kernel.permute(nd + 1, nd, *range(nd))                      # [Cout,Cin,k..]
            y = (f.conv3d if nd == 3 else f.conv2d)(xc, w, self.bias, stride=self.s)
        else:
            assert self.padding == "same"
            w = self.kernel.permute(nd + 1, nd, *range(nd))                      # [Cin,Cout,k..]
            y = (f.conv_transpose3d if nd == 3 else f.conv_transpose2d)(xc, w, self.bias, stride=self.s)   # full: (n-1)s + k
            before = (self.k - self.s) // 2
            sl = (slice(None), slice(None)) + tuple(slice(before, before + n * self.s) for n in xc.shape[2:])
            y = y[sl]
        return y.movedim(1, -1)


class Conv3D(_Conv): nd = 3
class Conv2D(_Conv): nd = 2
class Conv3DTranspose(_Conv): nd = 3; transposed = True
class Conv2DTranspose(_Conv): nd = 2; transposed = True


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kw):
        super().__init__(**kw); self.alpha = alpha

    def call(self, x):
        return torch.nn.functional.leaky_relu(x, self.alpha)


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super().__init__(**kw); self.rate = rate

    def _run(self, x, training):
        if not training or not DROPOUT["enabled"]:
            return x
        i = DROPOUT["calls"]; DROPOUT["calls"] += 1
        if DROPOUT["keys"] is not None:          # injected counter-hash masks (same keys go to the CUDA path)
            keep = torch.as_tensor(DROPOUT["mask_fn"](DROPOUT["keys"][i], tuple(x.shape)), dtype=x.dtype)
        else:
            keep = (torch.rand(x.shape, generator=random.gen) >= self.rate).to(x.dtype)
        return x * keep / (1.0 - self.rate)


class _Crop(Layer):
    nd = 3

    def __init__(self, cropping, **kw):
        super().__init__(**kw)
        self.c = [(cropping, cropping)] * self.nd if isinstance(cropping, int) else [(c, c) if isinstance(c, int) else tuple(c) for c in cropping]

    def call(self, x):
        sl = (slice(None),) + tuple(slice(a, x.shape[1 + i] - b) for i, (a, b) in enumerate(self.c)) + (slice(None),)
        return x[sl]


class Cropping3D(_Crop): nd = 3
class Cropping2D(_Crop): nd = 2


class _Pad(Layer):
    nd = 3

    def __init__(self, padding, **kw):
        super().__init__(**kw); self.p = padding

    def call(self, x):
        return torch.nn.functional.pad(x, [0, 0] + [self.p, self.p] * self.nd)


class ZeroPadding3D(_Pad): nd = 3
class ZeroPadding2D(_Pad): nd = 2


class Concatenate(Layer):
    _multi = True

    def out_shape(self, shapes):
        return tuple(shapes[0][:-1]) + (sum(s[-1] for s in shapes),)

    def call(self, xs):
        return torch.cat(list(xs), dim=-1)


class BatchNormalization(Layer):
    def __init__(self, *a, **kw):
        raise NotImplementedError("BatchNormalization is never applied on the reference's hot path")


def Input(shape, name=None):
    return _Sym((None,) + tuple(shape))


class Model(Layer):
    def __init__(self, inputs, outputs, name=None):
        super().__init__(name); self.inputs_sym, self.outputs_sym = inputs, outputs; self.built = True; self.trainable = True

    def out_shape(self, shape):
        return self.outputs_sym.shape

    def _layers(self):
        seen, order = set(), []

        def walk(s):
            if id(s) in seen or s is self.inputs_sym:
                return
            seen.add(id(s))
            for i in s.inputs:
                walk(i)
            if s.producer is not None:
                for l in s.producer._layers():
                    if all(l is not o for o in order):
                        order.append(l)
        walk(self.outputs_sym)
        return order

    @property
    def trainable_variables(self):
        return [w for l in self._layers() for w in l._weights if w.requires_grad]

    def get_weights(self):
        return [w.detach().numpy().copy() for l in self._layers() for w in l._weights]

    def set_weights(self, ws):
        cur = [w for l in self._layers() for w in l._weights]
        assert len(cur) == len(ws)
        with torch.no_grad():
            for c, w in zip(cur, ws):
                c.copy_(torch.as_tensor(np.asarray(w, np.float32)).reshape(c.shape))

    def _run(self, x, training):
        return _eval(self.outputs_sym, {id(self.inputs_sym): x}, training, {})

    def predict(self, x):
        with torch.no_grad():
            return self(x, training=False).numpy()


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps, self.t, self.state = learning_rate, beta_1, beta_2, epsilon, 0, {}

    def apply_gradients(self, grads_and_vars):
        self.t += 1
        lr_t = self.lr * _pymath.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        with torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                m, s = self.state.setdefault(id(v), (torch.zeros_like(v), torch.zeros_like(v)))
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                s.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                v.sub_(lr_t * m / (torch.sqrt(s) + self.eps))


class _RMSE:
    def __init__(self): self.se, self.n = 0.0, 0
    def update_state(self, a, b):
        d = (_t(a).double() - _t(b).double()); self.se += float((d * d).sum()); self.n += d.numel()
    def result(self): return _t(np.float32(_pymath.sqrt(self.se / max(self.n, 1))))


class _Checkpoint:
    def __init__(self, **objs): self.objs = objs
    def restore(self, path): raise NotImplementedError("TF checkpoint bundles are out of scope")


class _CheckpointManager:
    def __init__(self, ckpt, directory, max_to_keep=None): self.latest_checkpoint = None; self.n = 0; self.directory = directory
    def save(self): self.n += 1; return f"{self.directory}/ckpt-{self.n}"


train = types.SimpleNamespace(Checkpoint=_Checkpoint, CheckpointManager=_CheckpointManager)
keras = types.SimpleNamespace(
    layers=types.SimpleNamespace(Layer=Layer, Input=Input, Conv3D=Conv3D, Conv2D=Conv2D, Conv3DTranspose=Conv3DTranspose,
                                 Conv2DTranspose=Conv2DTranspose, LeakyReLU=LeakyReLU, Dropout=Dropout, Cropping3D=Cropping3D,
                                 Cropping2D=Cropping2D, ZeroPadding3D=ZeroPadding3D, ZeroPadding2D=ZeroPadding2D,
                                 Concatenate=Concatenate, BatchNormalization=BatchNormalization),
    Model=Model,
    optimizers=types.SimpleNamespace(Adam=Adam),
    losses=types.SimpleNamespace(Reduction=types.SimpleNamespace(AUTO="auto", SUM_OVER_BATCH_SIZE="sum_over_batch_size")),
    metrics=types.SimpleNamespace(RootMeanSquaredError=_RMSE),
    models=types.SimpleNamespace(load_model=None),
    utils=types.SimpleNamespace(plot_model=None),
)
