"""Thin object wrapper over a tem_handle: owns the handle, marshals torch / numpy buffers."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import TEM_F32, TEM_U8, NET_G, NET_F, NET_DX, NET_DY, check

PASS_NAMES = ("fake_y", "cycled_x", "fake_x", "cycled_y", "same_x", "same_y")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _as_device(x, device, dtypes=(torch.float32, torch.uint8)):
    """numpy / torch (cpu or cuda) -> contiguous cuda tensor.  Returns (tensor, was_numpy)."""
    was_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x)) if was_np else x
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(x))
        was_np = True
    if t.dtype not in dtypes:
        if torch.float32 not in dtypes:      # byte kernels read raw uint8: never reinterpret another dtype
            raise TypeError(f"expected a {' / '.join(str(d) for d in dtypes)} array, got {t.dtype}")
        t = t.to(torch.float32)
    if t.device.type != "cuda":
        t = t.to(device, non_blocking=True)
    return t.contiguous(), was_np


class Engine:
    """One tem_handle = two generators + two discriminators + optimizer state + workspace."""

    def __init__(self, dimsize=74, is3d=True, wf=8, max_batch=1, train=True, device=None, seed=0,
                 dropout=True, loss_mode="focal", focal_gamma=2.0, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-7,
                 use_tensor_cores=True):
        lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.TemError("no CUDA device: transfer_em_b200 has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.set_device(self.device)
        torch.cuda.init()
        cfg = _lib.TemConfig()
        lib.tem_default_config(C.byref(cfg))
        cfg.device = self.device.index
        cfg.is3d = 1 if is3d else 0
        cfg.wf = wf
        cfg.dimsize = dimsize
        cfg.max_batch = max_batch
        cfg.loss_mode = _lib.LOSS_FOCAL if loss_mode == "focal" else _lib.LOSS_LSGAN_L1
        cfg.dropout = 1 if dropout else 0
        cfg.focal_gamma = focal_gamma
        cfg.lr, cfg.beta1, cfg.beta2, cfg.eps = lr, beta1, beta2, eps
        cfg.seed = seed
        cfg.train = 1 if train else 0
        cfg.use_tensor_cores = 1 if use_tensor_cores else 0
        h = C.c_void_p()
        status = lib.tem_create(C.byref(cfg), C.byref(h))
        if status != 0:
            msg = lib.tem_last_error().decode()
            if status == -1:
                raise RuntimeError(msg)          # reference raises RuntimeError for bad dims (cgan.py:53, generator.py:38)
            raise _lib.TemError(msg)
        self._h = h
        self._lib = lib
        self.is3d, self.wf, self.dimsize, self.max_batch, self.train_enabled = is3d, wf, dimsize, max_batch, train
        od, buf = C.c_int32(), C.c_int32()
        check(lib.tem_out_dim(h, C.byref(od), C.byref(buf)))
        self.outdimsize, self.buffer = od.value, buf.value
        self._losses = torch.zeros(8, dtype=torch.float32).pin_memory()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.tem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------
    def param_count(self, net):
        return int(self._lib.tem_param_count(self._h, net))

    def variables(self, net):
        """[(name, offset, shape)] in layer order, Keras layouts."""
        out = []
        for v in range(self._lib.tem_num_variables(self._h, net)):
            name = C.create_string_buffer(32)
            off, nd = C.c_int64(), C.c_int32()
            shape = (C.c_int64 * 6)()
            check(self._lib.tem_variable_info(self._h, net, v, name, C.byref(off), C.byref(nd), shape))
            out.append((name.value.decode(), off.value, tuple(shape[i] for i in range(nd.value))))
        return out

    def get_vector(self, net, which=0):
        out = np.empty(self.param_count(net), np.float32)
        check(self._lib.tem_get_vector(self._h, net, which, out.ctypes.data_as(C.c_void_p), _stream()))
        return out

    def set_vector(self, net, vec, which=0):
        vec = np.ascontiguousarray(vec, np.float32).reshape(-1)
        if vec.size != self.param_count(net):
            raise ValueError(f"expected {self.param_count(net)} values, got {vec.size}")
        check(self._lib.tem_set_vector(self._h, net, which, vec.ctypes.data_as(C.c_void_p), _stream()))

    def get_weights(self, net, which=0):
        flat = self.get_vector(net, which)
        return [flat[o:o + int(np.prod(s))].reshape(s).copy() for _, o, s in self.variables(net)]

    def set_weights(self, net, weights, which=0):
        self.set_vector(net, np.concatenate([np.asarray(w, np.float32).reshape(-1) for w in weights]), which)

    @property
    def step(self):
        s = C.c_int64()
        check(self._lib.tem_get_step(self._h, C.byref(s)))
        return s.value

    @step.setter
    def step(self, v):
        check(self._lib.tem_set_step(self._h, int(v)))

    # ---- forward ----------------------------------------------------------------------------
    def gen_forward(self, net, x, meanstd=None, dropout_key=0):
        t, was_np = _as_device(x, self.device)
        nd = 3 if self.is3d else 2
        if t.dim() != nd + 2 or t.shape[-1] != 1:
            raise ValueError(f"expected [B,{'n,' * nd}1], got {tuple(t.shape)}")
        B, n = t.shape[0], t.shape[1]
        dt = TEM_U8 if t.dtype == torch.uint8 else TEM_F32
        outs = []
        for b0 in range(0, B, self.max_batch):
            tb = t[b0:b0 + self.max_batch]
            ob = torch.empty((tb.shape[0],) + (n - 34,) * nd + (1,), dtype=torch.float32, device=self.device)
            check(self._lib.tem_gen_forward(self._h, net, C.c_void_p(tb.data_ptr()), dt, _lib.fptr2(meanstd),
                                            tb.shape[0], n, dropout_key, C.c_void_p(ob.data_ptr()), _stream()))
            outs.append(ob)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return out.cpu().numpy() if was_np else out

    def disc_forward(self, net, x):
        t, was_np = _as_device(x, self.device, (torch.float32,))
        nd = 3 if self.is3d else 2
        B, m = t.shape[0], t.shape[1]
        l = C.c_int32()
        check(self._lib.tem_disc_out_dim(self._h, m, C.byref(l)))
        outs = []
        for b0 in range(0, B, self.max_batch):
            tb = t[b0:b0 + self.max_batch]
            ob = torch.empty((tb.shape[0],) + (l.value,) * nd + (1,), dtype=torch.float32, device=self.device)
            check(self._lib.tem_disc_forward(self._h, net, C.c_void_p(tb.data_ptr()), tb.shape[0], m,
                                             C.c_void_p(ob.data_ptr()), _stream()))
            outs.append(ob)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return out.cpu().numpy() if was_np else out

    def last_activation(self, net, layer):
        cnt = C.c_int64()
        check(self._lib.tem_last_activation(self._h, net, layer, None, C.byref(cnt), _stream()))
        out = torch.empty(cnt.value, dtype=torch.float32, device=self.device)
        check(self._lib.tem_last_activation(self._h, net, layer, C.c_void_p(out.data_ptr()), C.byref(cnt), _stream()))
        return out.cpu().numpy()

    # ---- training ---------------------------------------------------------------------------
    def _train_call(self, fn, real_x, real_y, meanstd_x, meanstd_y):
        tx, _ = _as_device(real_x, self.device)
        ty, _ = _as_device(real_y, self.device)
        if tx.shape != ty.shape or tx.dtype != ty.dtype:
            raise ValueError("real_x and real_y must have the same shape and dtype")
        dt = TEM_U8 if tx.dtype == torch.uint8 else TEM_F32
        check(fn(self._h, C.c_void_p(tx.data_ptr()), C.c_void_p(ty.data_ptr()), dt, _lib.fptr2(meanstd_x),
                 _lib.fptr2(meanstd_y), tx.shape[0], C.c_void_p(self._losses.data_ptr()), _stream()))
        self._keep = (tx, ty)       # keep inputs alive until the stream has consumed them
        return self._losses

    def train_step_async(self, real_x, real_y, meanstd_x=None, meanstd_y=None):
        """Enqueue one train step; returns the pinned 7-float loss buffer (valid after a stream sync)."""
        return self._train_call(self._lib.tem_train_step, real_x, real_y, meanstd_x, meanstd_y)

    def train_step(self, real_x, real_y, meanstd_x=None, meanstd_y=None):
        l = self.train_step_async(real_x, real_y, meanstd_x, meanstd_y)
        torch.cuda.current_stream().synchronize()
        return tuple(np.float32(v) for v in l[:7].tolist())

    def train_grads(self, real_x, real_y, meanstd_x=None, meanstd_y=None):
        l = self._train_call(self._lib.tem_train_grads, real_x, real_y, meanstd_x, meanstd_y)
        torch.cuda.current_stream().synchronize()
        return tuple(np.float32(v) for v in l[:7].tolist())

    def apply_adam(self, grad_scale=1.0):
        check(self._lib.tem_apply_adam(self._h, grad_scale, _stream()))

    def train_output(self, name):
        p = PASS_NAMES.index(name)
        cnt = C.c_int64()
        check(self._lib.tem_train_output(self._h, p, None, C.byref(cnt), _stream()))
        out = torch.empty(cnt.value, dtype=torch.float32, device=self.device)
        check(self._lib.tem_train_output(self._h, p, C.c_void_p(out.data_ptr()), C.byref(cnt), _stream()))
        nd = 3 if self.is3d else 2
        return out.cpu().numpy().reshape((-1,) + (self.outdimsize,) * nd + (1,))

    def debug_backward_scratch(self, is_gen, layer):
        cnt = C.c_int64()
        check(self._lib.tem_debug_backward_scratch(self._h, int(is_gen), layer, None, C.byref(cnt), _stream()))
        out = torch.empty(cnt.value, dtype=torch.float32, device=self.device)
        check(self._lib.tem_debug_backward_scratch(self._h, int(is_gen), layer, C.c_void_p(out.data_ptr()), C.byref(cnt), _stream()))
        return out.cpu().numpy()

    def set_dropout_keys(self, keys):
        arr = (C.c_uint32 * 12)(*[int(k) for k in keys])
        check(self._lib.tem_set_dropout_keys(self._h, arr))

    def get_dropout_keys(self):
        arr = (C.c_uint32 * 12)()
        check(self._lib.tem_get_dropout_keys(self._h, arr))
        return list(arr)

    # ---- tiled inference --------------------------------------------------------------------
    def predict_volume(self, vol, start, size, meanstd_x, meanstd_y, net=NET_G, outdimsize=None, buffer=None,
                       fetch_input=False, tile_z_range=None, out=None):
        """vol: uint8 [z,y,x] (numpy or torch).  Returns uint8 [size_z,size_y,size_x] cuda tensor(s)."""
        tv, _ = _as_device(vol, self.device, (torch.uint8,))
        if tv.dim() != 3:
            raise ValueError("volume must be uint8 [z,y,x]")
        sz = (int(size[2]), int(size[1]), int(size[0]))
        if out is None:
            out = torch.zeros(sz, dtype=torch.uint8, device=self.device)
        inb = torch.zeros(sz, dtype=torch.uint8, device=self.device) if fetch_input else None
        vd = (C.c_int64 * 3)(*tv.shape)
        st = (C.c_int64 * 3)(*[int(s) for s in start])
        si = (C.c_int64 * 3)(*[int(s) for s in size])
        zb, ze = (-1, -1) if tile_z_range is None else tile_z_range
        check(self._lib.tem_predict_volume(self._h, net, C.c_void_p(tv.data_ptr()), vd, st, si, _lib.fptr2(meanstd_x),
                                           _lib.fptr2(meanstd_y), -1 if outdimsize is None else outdimsize,
                                           -1 if buffer is None else buffer, zb, ze, C.c_void_p(out.data_ptr()),
                                           C.c_void_p(inb.data_ptr()) if inb is not None else None, _stream()))
        return (inb, out) if fetch_input else out

    # ---- measurement -----------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.tem_launch_count())

    def profile(self, on):
        check(self._lib.tem_profile_enable(self._h, 1 if on else 0))

    def profile_report(self):
        """{tag: dict(count, ms, bytes, flops, kernel)} - per-launch algorithmic bytes / flops, total ms, kernel chosen."""
        buf = C.create_string_buffer(1 << 16)
        check(self._lib.tem_profile_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            t, n, ms, by, fl, kn = line.split()
            out[t] = dict(count=int(n), ms=float(ms), bytes=float(by), flops=float(fl), kernel=kn)
        return out

    # ---- data parallel ----------------------------------------------------------------------
    def init_comm(self, rank, world, broadcast_id):
        """broadcast_id(bytes_or_None) -> bytes: ships rank 0's 128-byte NCCL id to every rank."""
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            check(self._lib.tem_comm_unique_id(ident))
        raw = broadcast_id(bytes(ident) if rank == 0 else None)
        ident = (C.c_uint8 * 128)(*raw)
        check(self._lib.tem_comm_init(self._h, ident, rank, world))
        check(self._lib.tem_comm_sync_params(self._h, _stream()))
        torch.cuda.current_stream().synchronize()
