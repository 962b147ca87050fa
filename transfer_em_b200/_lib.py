"""ctypes binding of libtem_b200.so (the C ABI declared in include/transfer_em_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, every
entry point raises.  The library is built in-tree by ``python -m transfer_em_b200.build``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtem_b200.so")
if os.environ.get("TEM_ABLATION_LIB") == "1":      # profiling tools only: the -DTEM_ABLATION build (python -m transfer_em_b200.build --ablation)
    LIB_PATH = os.path.join(_HERE, "libtem_b200_abl.so")

TEM_U8, TEM_BF16, TEM_F32 = 0, 1, 2
NET_G, NET_F, NET_DX, NET_DY = 0, 1, 2, 3
LOSS_FOCAL, LOSS_LSGAN_L1 = 0, 1
ABI_VERSION = 1


class TemError(RuntimeError):
    pass


class TemConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("is3d", C.c_int32), ("wf", C.c_int32),
                ("dimsize", C.c_int32), ("max_batch", C.c_int32), ("loss_mode", C.c_int32), ("dropout", C.c_int32),
                ("focal_gamma", C.c_float), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("seed", C.c_uint64), ("train", C.c_int32), ("use_tensor_cores", C.c_int32)]


class TemConvDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("in_dims", C.c_int32 * 3), ("cin", C.c_int32), ("cout", C.c_int32),
                ("k", C.c_int32 * 3), ("stride", C.c_int32 * 3), ("transposed", C.c_int32), ("slope", C.c_float),
                ("dropout_key", C.c_uint32), ("in_dtype", C.c_int32), ("out_dtype", C.c_int32),
                ("meanstd", C.c_float * 2), ("use_tensor_cores", C.c_int32)]


_P = C.c_void_p
_SIGS = {
    "tem_last_error": (C.c_char_p, []),
    "tem_abi_version": (C.c_int, []),
    "tem_default_config": (None, [C.POINTER(TemConfig)]),
    "tem_create": (C.c_int, [C.POINTER(TemConfig), C.POINTER(_P)]),
    "tem_destroy": (C.c_int, [_P]),
    "tem_out_dim": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "tem_param_count": (C.c_int64, [_P, C.c_int]),
    "tem_num_variables": (C.c_int, [_P, C.c_int]),
    "tem_variable_info": (C.c_int, [_P, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "tem_get_vector": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "tem_set_vector": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "tem_get_step": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "tem_set_step": (C.c_int, [_P, C.c_int64]),
    "tem_gen_forward": (C.c_int, [_P, C.c_int, _P, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_uint32, _P, _P]),
    "tem_disc_forward": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "tem_disc_out_dim": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32)]),
    "tem_last_activation": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int64), _P]),
    "tem_debug_graph_replay": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "tem_train_step": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, _P, _P]),
    "tem_train_grads": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, _P, _P]),
    "tem_apply_adam": (C.c_int, [_P, C.c_float, _P]),
    "tem_train_output": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_int64), _P]),
    "tem_debug_backward_scratch": (C.c_int, [_P, C.c_int, C.c_int, _P, C.POINTER(C.c_int64), _P]),
    "tem_set_dropout_keys": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "tem_get_dropout_keys": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "tem_predict_volume": (C.c_int, [_P, C.c_int, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "tem_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "tem_comm_init": (C.c_int, [_P, C.POINTER(C.c_uint8), C.c_int, C.c_int]),
    "tem_comm_world": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tem_comm_sync_params": (C.c_int, [_P, _P]),
    "tem_launch_count": (C.c_uint64, []),
    "tem_last_kernel": (C.c_char_p, []),
    "tem_profile_enable": (C.c_int, [_P, C.c_int]),
    "tem_profile_report": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "tem_standardize_u8": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_float), _P]),
    "tem_unstandardize_to_u8": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_float), _P]),
    "tem_chunk_volume": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32, _P, _P]),
    "tem_augment": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P, C.c_int32, C.POINTER(C.c_int32), _P, _P, _P, _P, _P]),
    "tem_mean_var": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "tem_warp_tensor": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_float, _P, _P]),
    "tem_conv_forward": (C.c_int, [C.POINTER(TemConvDesc), _P, _P, _P, _P, C.POINTER(C.c_int32), _P]),
    "tem_conv_dgrad": (C.c_int, [C.POINTER(TemConvDesc), _P, C.c_int, _P, _P, C.c_float, _P, C.c_int, _P]),
    "tem_conv_wgrad": (C.c_int, [C.POINTER(TemConvDesc), _P, _P, C.c_int, _P, _P]),
    "tem_focal_logits": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, C.c_float, _P, _P, _P]),
    "tem_focal_probs": (C.c_int, [_P, _P, C.c_int64, C.c_float, C.c_float, _P, _P, _P]),
    "tem_adam": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P]),
    "tem_dropout_mask": (C.c_int, [C.c_uint32, _P, C.c_int64, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def load():
    """Load the shared library (once).  Raises TemError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TemError(f"{LIB_PATH} not found: build it with `python -m transfer_em_b200.build` "
                       "(transfer_em_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.tem_abi_version() != ABI_VERSION:
        raise TemError("libtem_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().tem_last_error()
        raise TemError(f"libtem_b200 error {status}: {msg.decode() if msg else '?'}")


def fptr2(ms):
    if ms is None:
        return None
    return (C.c_float * 2)(float(ms[0]), float(ms[1]))
