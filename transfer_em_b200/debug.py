"""transfer_em/debug.py on device: the artificial source domain of the self-comparison tests (warp_tensor, :7-63) and the
RMSE metric (accuracy, :65-71).  Same names and argument meaning as the reference; the random draw of the hole seeds, which the
reference takes from an unseeded tf.random.uniform, comes from `rng` (numpy Generator) or an explicit `uniform` array."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _as_device, _stream

HOLE_RATE = 4 / (128 * 128)          # debug.py:20: fraction of voxels that seed a hole


def warp_tensor(tensor, rng=None, uniform=None, device=None):
    """tensor: float32 [Z,Y,X,1] (3-D) or [Y,X,1] (2-D), already scaled to [-1,1] (datasets.scale_tensor).
    Returns the warped tensor of the same shape: 3^d box blur ('SAME'), then 4^d-dilated holes set to the mean of the
    blurred tensor (debug.py:22-60)."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    t, was_np = _as_device(tensor, dev, (torch.float32,))
    if t.dim() not in (3, 4) or t.shape[-1] != 1:
        raise RuntimeError("warp_tensor expects [Z,Y,X,1] or [Y,X,1]")
    nd = t.dim() - 1
    spatial = tuple(int(v) for v in t.shape[:-1])
    dims = (C.c_int32 * 3)(*((1,) * (3 - nd) + spatial))
    if uniform is None:
        rng = np.random.default_rng() if rng is None else rng
        uniform = rng.uniform(0.0, 1.0, int(np.prod(spatial))).astype(np.float32)
    u, _ = _as_device(np.asarray(uniform, np.float32).reshape(spatial) if not isinstance(uniform, torch.Tensor) else uniform.reshape(spatial),
                      dev, (torch.float32,))
    t = t.contiguous(); u = u.contiguous()
    out = torch.empty_like(t)
    scratch = torch.zeros(1, dtype=torch.float64, device=t.device)
    _lib.check(lib.tem_warp_tensor(C.c_void_p(t.data_ptr()), C.c_void_p(u.data_ptr()), C.c_void_p(out.data_ptr()), dims, nd,
                                   C.c_float(HOLE_RATE), C.c_void_p(scratch.data_ptr()), _stream()))
    return out.cpu().numpy() if was_np else out


def accuracy(unwarped_orig_tensor, predicted_tensor):
    """Root-mean-squared error between the two tensors (tf.keras.metrics.RootMeanSquaredError, debug.py:65-71)."""
    a = unwarped_orig_tensor.detach().cpu().numpy() if isinstance(unwarped_orig_tensor, torch.Tensor) else np.asarray(unwarped_orig_tensor)
    b = predicted_tensor.detach().cpu().numpy() if isinstance(predicted_tensor, torch.Tensor) else np.asarray(predicted_tensor)
    return float(np.sqrt(np.mean((a.astype(np.float32) - b.astype(np.float32)) ** 2, dtype=np.float64)))
