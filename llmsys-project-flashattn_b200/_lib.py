"""ctypes loader for the three C-ABI libraries (include/flashattn_b200.h).

The reference loads ``minitorch/cuda_kernels/*.so`` relative to the cwd at import time
(minitorch/cuda_kernel_ops.py:26-29); here the paths are resolved next to this file and
a missing library is a loud ImportError -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_bool, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
KERNEL_DIR = os.environ.get("FLASHATTN_B200_KERNEL_DIR", os.path.join(_HERE, "minitorch", "cuda_kernels"))

FA_OK, FA_ERR_INVALID, FA_ERR_UNSUPPORTED, FA_ERR_CUDA = 0, 1, 2, 3
FA_DTYPE_F32, FA_DTYPE_BF16 = 0, 1
FA_MODE_FP32, FA_MODE_BF16 = 0, 1


class FlashAttnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[flashattn_b200 status {code}] {msg}")
        self.code = code


class fa_attn_desc(ctypes.Structure):
    _fields_ = [
        ("B", c_int), ("H", c_int), ("N", c_int), ("d", c_int),
        ("dtype", c_int), ("causal", c_int),
        ("stride_b", c_longlong), ("stride_h", c_longlong), ("stride_n", c_longlong),
        ("kv_len", c_void_p), ("key_mask", c_void_p),
    ]


class fa_decode_desc(ctypes.Structure):
    _fields_ = [
        ("B", c_int), ("H", c_int), ("d", c_int), ("L", c_int), ("L_cap", c_int), ("dtype", c_int),
        ("q_stride_b", c_longlong), ("q_stride_h", c_longlong),
        ("o_stride_b", c_longlong), ("o_stride_h", c_longlong),
        ("cache_stride_b", c_longlong), ("cache_stride_h", c_longlong), ("cache_stride_n", c_longlong),
        ("kv_len", c_void_p),
    ]


_f32 = np.ctypeslib.ndpointer(dtype=np.float32, ndim=1, flags="C_CONTIGUOUS")

# symbol -> (restype, argtypes); shared by all four libraries
_COMMON = {
    "fa_last_status": (c_int, []),
    "fa_last_error": (c_char_p, []),
    "fa_device_count": (c_int, []),
    "fa_set_device": (c_int, [c_int]),
    "fa_malloc": (c_void_p, [c_size_t]),
    "fa_malloc_host": (c_void_p, [c_size_t]),
    "fa_malloc_async": (c_void_p, [c_size_t, c_void_p]),
    "fa_free_async": (c_int, [c_void_p, c_void_p]),
    "fa_memset_async": (c_int, [c_void_p, c_int, c_size_t, c_void_p]),
    "fa_free": (c_int, [c_void_p]),
    "fa_free_host": (c_int, [c_void_p]),
    "fa_memset": (c_int, [c_void_p, c_int, c_size_t]),
    "fa_h2d": (c_int, [c_void_p, c_void_p, c_size_t]),
    "fa_d2h": (c_int, [c_void_p, c_void_p, c_size_t]),
    "fa_sync": (c_int, []),
    "fa_event_create": (c_void_p, []),
    "fa_event_record": (c_int, [c_void_p, c_void_p]),
    "fa_event_elapsed_ms": (c_float, [c_void_p, c_void_p]),
    "fa_event_destroy": (c_int, [c_void_p]),
    "fa_flush_l2": (c_int, []),
    "fa_launch_count": (ctypes.c_ulonglong, []),
}
_HOST4 = [_f32] * 6 + [c_int] * 4
_HOST4B = [_f32] * 10 + [c_int] * 4
SYMBOLS = {
    "flashattention_kernel": {
        **_COMMON,
        "fa_set_mode": (None, [c_int]),
        "fa_get_mode": (c_int, []),
        "fa_set_deterministic": (None, [c_int]),
        "fa_set_legacy_chunk_bytes": (None, [c_size_t]),
        "fa_set_keep_forward_mb": (None, [c_longlong]),
        "fa_forward_cache_stats": (None, [POINTER(ctypes.c_ulonglong), POINTER(ctypes.c_ulonglong)]),
        "fa_staging_fallbacks": (ctypes.c_ulonglong, []),
        "fa_set_transfer_policy": (None, [ctypes.c_double, c_longlong]),
        "fa_plan_transfer_preview": (c_int, [c_int, POINTER(c_int), c_int, POINTER(c_int), ctypes.c_double, POINTER(c_int),
                                             POINTER(c_int)]),
        "fa_wire_bytes": (None, [POINTER(ctypes.c_ulonglong), POINTER(ctypes.c_ulonglong)]),
        "fa_release_staging": (c_int, []),
        "launch_flashattention_forward": (None, _HOST4),
        "launch_flashattention_forward_causal": (None, _HOST4),
        "launch_flashattention_backward": (None, _HOST4B),
        "launch_flashattention_backward_causal": (None, _HOST4B),
        "launch_flashattention_forward_masked": (None, [_f32] * 6 + [c_void_p, c_int] + [c_int] * 4),
        "launch_flashattention_backward_masked": (None, [_f32] * 10 + [c_void_p, c_int] + [c_int] * 4),
        "fa_flash_fwd_dev": (c_int, [POINTER(fa_attn_desc)] + [c_void_p] * 6 + [c_void_p]),
        "fa_flash_bwd_dev": (c_int, [POINTER(fa_attn_desc)] + [c_void_p] * 10 + [c_void_p]),
        "fa_flash_decode_dev": (c_int, [POINTER(fa_decode_desc)] + [c_void_p] * 5 + [c_void_p]),
        "fa_cast_f32_to_bf16_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
        "fa_cast_bf16_to_f32_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
        "fa_attn_flops": (c_double, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_int]),
    },
    "softmax_kernel": {
        **_COMMON,
        "launch_attn_softmax": (None, [_f32, c_void_p, c_int, c_int, c_int, c_int, c_bool, c_void_p]),
        "launch_attn_softmax_bw": (None, [_f32, _f32, c_int, c_int, c_void_p]),
        "fa_attn_softmax_dev": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
        "fa_attn_softmax_bw_dev": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    },
    "layernorm_kernel": {
        **_COMMON,
        "launch_layernorm": (None, [_f32] * 6 + [c_int, c_int, c_void_p]),
        "launch_layernorm_bw": (None, [_f32] * 9 + [c_int, c_int, c_void_p, c_void_p]),
        "fa_layernorm_dev": (c_int, [c_void_p] * 6 + [c_longlong, c_int, c_void_p]),
        "fa_layernorm_bw_dev": (c_int, [c_void_p] * 9 + [c_longlong, c_int, c_void_p]),
    },
}

_i32 = np.ctypeslib.ndpointer(dtype=np.int32, ndim=1, flags="C_CONTIGUOUS")
SYMBOLS["combine"] = {
    **_COMMON,
    "tensorMap": (None, [_f32, _i32, _i32, c_int, _f32, _i32, _i32, c_int, c_int, c_int]),
    "tensorZip": (None, [_f32, _i32, _i32, c_int, c_int, _f32, _i32, _i32, c_int, c_int, _f32, _i32, _i32, c_int,
                         c_int, c_int]),
    "tensorReduce": (None, [_f32, _i32, _i32, c_int, _f32, _i32, _i32, c_int, c_double, c_int, c_int]),
    "MatrixMultiply": (None, [_f32, _i32, _i32, _f32, _i32, _i32, _f32, _i32, _i32, c_int, c_int, c_int]),
    "fa_gemm_bf16_dev": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong,
                                 c_int, c_int, c_int, c_void_p]),
    "fa_qkv_proj_bf16_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fa_map_dev": (c_int, [c_void_p, _i32, _i32, c_int, c_void_p, _i32, _i32, c_int, c_int, c_void_p]),
    "fa_zip_dev": (c_int, [c_void_p, _i32, _i32, c_int, c_void_p, _i32, _i32, c_int, c_void_p, _i32, _i32, c_int, c_int,
                           c_void_p]),
    "fa_reduce_dev": (c_int, [c_void_p, _i32, _i32, c_void_p, _i32, _i32, c_int, c_int, c_double, c_int, c_void_p]),
    "fa_matmul_dev": (c_int, [c_void_p, _i32, _i32, c_void_p, _i32, _i32, c_void_p, _i32, _i32, c_void_p]),
    "launch_embedding_fw": (None, [_f32, _f32, _f32, c_longlong, c_int, c_int]),
    "launch_embedding_bw": (None, [_f32, _f32, _f32, c_longlong, c_int, c_int]),
    "launch_softmax_xent_fw": (None, [_f32, _f32, _f32, _f32, c_longlong, c_int]),
    "launch_softmax_xent_bw": (None, [_f32, _f32, _f32, _f32, _f32, c_longlong, c_int]),
    "fa_embedding_fw_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "fa_embedding_bw_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "fa_softmax_xent_fw_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
    "fa_softmax_xent_bw_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]),
}

_libs = {}


def load(name: str) -> ctypes.CDLL:
    """Load (once) ``<KERNEL_DIR>/<name>.so`` and bind every symbol the header declares."""
    lib = _libs.get(name)
    if lib is not None:
        return lib
    path = os.path.join(KERNEL_DIR, name + ".so")
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build the CUDA libraries with "
            f"`bash {os.path.join(_HERE, 'compile_cuda.sh')}` (or __graft_entry__.build()). "
            "There is no CPU / PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(path)
    for sym, (restype, argtypes) in SYMBOLS[name].items():
        fn = getattr(lib, sym)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _libs[name] = lib
    return lib


def check(lib: ctypes.CDLL, rc=None) -> None:
    """Raise if the last call into `lib` failed (legacy symbols return void)."""
    code = lib.fa_last_status() if rc is None else rc
    if code != FA_OK:
        msg = lib.fa_last_error()
        raise FlashAttnError(code, msg.decode() if msg else "")


def as_f32_ptr(arr):
    """Optional host fp32 array -> c_void_p (None stays NULL)."""
    if arr is None:
        return None
    assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(c_void_p)
