"""Helpers for the GPU parity tests: call the per-op C-ABI entry points on torch buffers."""
import ctypes as C

import numpy as np
import torch

from transfer_em_b200 import _lib

DT = {torch.uint8: _lib.TEM_U8, torch.bfloat16: _lib.TEM_BF16, torch.float32: _lib.TEM_F32}


def bf16r(a):
    """round a float array to bf16 precision (returned as float64 numpy)."""
    return torch.tensor(np.asarray(a, np.float32)).to(torch.bfloat16).to(torch.float64).numpy()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def make_desc(B, in_dims, cin, cout, k, s, transposed=False, slope=1.0, key=0, in_dtype=torch.bfloat16,
              out_dtype=torch.bfloat16, meanstd=(0.0, 1.0), is3d=True, tc=0):
    d = _lib.TemConvDesc()
    d.B = B
    for i in range(3):
        d.in_dims[i] = in_dims[i]
        d.k[i] = k if (is3d or i > 0) else 1
        d.stride[i] = s if (is3d or i > 0) else 1
    d.cin, d.cout, d.transposed, d.slope, d.dropout_key = cin, cout, int(transposed), slope, key
    d.in_dtype, d.out_dtype = DT[in_dtype], DT[out_dtype]
    d.meanstd[0], d.meanstd[1] = meanstd
    d.use_tensor_cores = tc
    return d


def conv_forward(x, w, desc, bias=None):
    lib = _lib.load()
    od = (C.c_int32 * 3)()
    _lib.check(lib.tem_conv_forward(C.byref(desc), None, None, None, None, od, stream()))
    out_dt = torch.bfloat16 if desc.out_dtype == _lib.TEM_BF16 else torch.float32
    out = torch.empty((desc.B, od[0], od[1], od[2], desc.cout), dtype=out_dt, device=x.device)
    wb = w
    if bias is not None:       # bias must live in the same allocation as w (net-relative offset)
        wb = torch.cat([w.reshape(-1), bias.reshape(-1)])
        w = wb[:w.numel()]
        bias = wb[w.numel():]
    _lib.check(lib.tem_conv_forward(C.byref(desc), ptr(x), ptr(w), ptr(bias), ptr(out), od, stream()))
    torch.cuda.synchronize()
    return out


def conv_dgrad(dy, w, desc, x_act=None, x_slope=1.0, dx_dtype=torch.bfloat16):
    lib = _lib.load()
    dx = torch.empty((desc.B, desc.in_dims[0], desc.in_dims[1], desc.in_dims[2], desc.cin), dtype=dx_dtype, device=dy.device)
    _lib.check(lib.tem_conv_dgrad(C.byref(desc), ptr(dy), DT[dy.dtype], ptr(w), ptr(x_act), x_slope, ptr(dx), DT[dx_dtype], stream()))
    torch.cuda.synchronize()
    return dx


def conv_wgrad(x, dy, desc, wshape):
    lib = _lib.load()
    dw = torch.zeros(wshape, dtype=torch.float32, device=x.device)
    _lib.check(lib.tem_conv_wgrad(C.byref(desc), ptr(x), ptr(dy), DT[dy.dtype], ptr(dw), stream()))
    torch.cuda.synchronize()
    return dw


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
