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


def get_meanstd(tensors):
    """datasets.py:173-190 (host-side; not on the hot path)."""
    mean = np.float32(0); var = np.float32(0)
    for t in tensors:
        t = np.asarray(t, np.float32)
        mean += t.mean(dtype=np.float32); var += t.var(dtype=np.float32)
    return float(mean / len(tensors)), float(np.sqrt(var / len(tensors)))
