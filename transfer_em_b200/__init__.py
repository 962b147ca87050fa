"""transfer_em_b200: B200-native (sm_100a) implementation of the transfer_em hot path.

Python keeps the reference's call signatures (transfer_em/cgan.py, transfer_em/models/*.py,
transfer_em/utils.py); the arithmetic runs in hand-written CUDA behind the C ABI of
include/transfer_em_b200.h (libtem_b200.so).  PyTorch is used only for device memory, streams and
torch.distributed rendezvous.  No TensorFlow, no Triton, no CPU fallback.
"""
from . import _lib                                  # noqa: F401
from .engine import Engine                          # noqa: F401
from .cgan import EM2EM                             # noqa: F401
from .models.generator import unet_generator, VALID_DIMS, VALID_OUT       # noqa: F401
from .models.discriminator import discriminator     # noqa: F401
from .utils import predict_ng_cube, predict_cube_from_saved_model, save_model   # noqa: F401

__all__ = ["EM2EM", "Engine", "unet_generator", "discriminator", "predict_ng_cube",
           "predict_cube_from_saved_model", "save_model"]
