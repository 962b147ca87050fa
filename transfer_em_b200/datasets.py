"""uint8 <-> float conventions of transfer_em/datasets/datasets.py (157-171, 193-202), on device."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _as_device, _stream


def scale_and_standardize(u8, meanstd, device=None):
    """scale_tensor + standardize_population: (float32(u8)/127.5 - 1 - mean)/std, channel added."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    t, was_np = _as_device(u8, dev, (torch.uint8,))
    out = torch.empty(tuple(t.shape) + (1,), dtype=torch.float32, device=t.device)
    _lib.check(lib.tem_standardize_u8(C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), t.numel(), _lib.fptr2(meanstd), _stream()))
    return out.cpu().numpy() if was_np else out


def scale_tensor(u8, device=None):
    return scale_and_standardize(u8, (0.0, 1.0), device)


def unstandardize_to_uint8(y, meanstd, device=None):
    """(y*std + mean + 1)*127.5 -> round-half-even -> uint8 with wrap (transfer_em/utils.py:109,118)."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    t, was_np = _as_device(y, dev, (torch.float32,))
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    _lib.check(lib.tem_unstandardize_to_u8(C.c_void_p(t.data_ptr()), C.c_void_p(out.data_ptr()), t.numel(), _lib.fptr2(meanstd), _stream()))
    return out.cpu().numpy() if was_np else out


def get_meanstd(tensors, device=None):
    """datasets.py:173-190 on device: mean of the per-tensor means, sqrt of the mean of the per-tensor (population)
    variances; every tensor is reduced by one kernel (fp64 accumulation), the two running sums stay fp32 as in the reference."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    scratch = torch.zeros(4, dtype=torch.float64, device=dev)
    out = torch.empty(2, dtype=torch.float32, device=dev)
    mean = np.float32(0); var = np.float32(0); count = 0
    for t in tensors:
        tt, _ = _as_device(t, dev, (torch.float32,))
        _lib.check(lib.tem_mean_var(C.c_void_p(tt.data_ptr()), tt.numel(), C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), _stream()))
        m, v = out.tolist()
        mean += np.float32(m); var += np.float32(v); count += 1
    return float(mean / np.float32(count)), float(np.sqrt(var / np.float32(count)))


def draw_augmentation(batch, ndims, rng):
    """The random choices of datasets.py:123-155 for `batch` samples: a shuffled axis order, a coin per axis, the intensity
    shift U(-0.05, 0.05) and the variance scale U(1, 1.05).  Returned as int32 [B,3] / [B,3] and float32 [B] / [B]
    (2-D data keeps axis 0 fixed)."""
    perm = np.zeros((batch, 3), np.int32); flip = np.zeros((batch, 3), np.int32)
    off = 3 - ndims
    for b in range(batch):
        perm[b] = np.arange(3)
        perm[b, off:] = off + rng.permutation(ndims)
        flip[b, off:] = rng.uniform(0.0, 1.0, ndims) < 0.5
    mean_adj = rng.uniform(-0.05, 0.05, batch).astype(np.float32)
    var_adj = rng.uniform(1.0, 1.05, batch).astype(np.float32)
    return perm, flip, var_adj, mean_adj


def augment(batch, rng=None, meanstd=None, choices=None, device=None):
    """datasets.py:123-155 for a whole batch on device.  batch: uint8 [B,(z,)y,x] raw patches (scale + standardise with
    `meanstd` fused in front, as the reference pipeline orders them) or float32 [B,(z,)y,x(,1)] standardised tensors.
    Returns float32 [B,(z,)y,x,1].  `choices` = (perm, flip, var_adj, mean_adj) overrides the random draw."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    t, was_np = _as_device(batch, dev, (torch.uint8, torch.float32))
    if t.dtype == torch.float32 and t.shape[-1] == 1 and t.dim() in (4, 5):
        t = t[..., 0]
    t = t.contiguous()
    B = t.shape[0]; sp = tuple(t.shape[1:]); nd = len(sp)
    if nd not in (2, 3):
        raise ValueError("augment expects [B,y,x] or [B,z,y,x] patches")
    in_dims = (1,) * (3 - nd) + sp
    if choices is None:
        choices = draw_augmentation(B, nd, rng if rng is not None else np.random.default_rng())
    perm, flip, var_adj, mean_adj = [np.ascontiguousarray(c) for c in choices]
    out_dims = [tuple(in_dims[perm[b, k]] for k in range(3)) for b in range(B)]
    if any(d != out_dims[0] for d in out_dims):
        raise ValueError("samples of one batch must keep one output shape (use cubic patches or one permutation)")
    od = (C.c_int32 * 3)(*out_dims[0])
    dperm = torch.from_numpy(perm.astype(np.int32)).to(dev); dflip = torch.from_numpy(flip.astype(np.int32)).to(dev)
    dvar = torch.from_numpy(var_adj.astype(np.float32)).to(dev); dmean = torch.from_numpy(mean_adj.astype(np.float32)).to(dev)
    out = torch.empty((B,) + out_dims[0][3 - nd:] + (1,), dtype=torch.float32, device=dev)
    if t.dtype == torch.uint8 and meanstd is None:
        meanstd = (0.0, 1.0)
    _lib.check(lib.tem_augment(C.c_void_p(t.data_ptr()), _lib.TEM_U8 if t.dtype == torch.uint8 else _lib.TEM_F32,
                               _lib.fptr2(meanstd) if meanstd is not None else None, C.c_void_p(out.data_ptr()), B, od,
                               C.c_void_p(dperm.data_ptr()), C.c_void_p(dflip.data_ptr()), C.c_void_p(dvar.data_ptr()),
                               C.c_void_p(dmean.data_ptr()), _stream()))
    return out.cpu().numpy() if was_np else out
