"""Device-resident buffers and the device-pointer entry points (``fa_*_dev``).

This is the zero-copy side of the C ABI: tensors stay in HBM between calls, which is what
the throughput numbers in bench.py are measured on and what batch x head sharding across
GPUs uses.  Only ctypes + numpy; no torch.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib

_ITEMSIZE = {"f32": 4, "bf16": 2, "i32": 4}


def _fa():
    return _lib.load("flashattention_kernel")


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round-to-nearest-even (host-side packing helper)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16).reshape(np.shape(x))


def from_bf16_bits(u: np.ndarray) -> np.ndarray:
    return (np.asarray(u, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


class DeviceArray:
    """A typed allocation in device memory (freed on garbage collection)."""

    def __init__(self, shape: Sequence[int], dtype: str):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * _ITEMSIZE[dtype]
        lib = _fa()
        self.ptr = lib.fa_malloc(self.nbytes)
        if not self.ptr:
            _lib.check(lib)
            raise MemoryError(f"fa_malloc({self.nbytes}) failed")

    @classmethod
    def from_numpy(cls, arr: np.ndarray, dtype: Optional[str] = None) -> "DeviceArray":
        if dtype is None:
            dtype = "i32" if arr.dtype.kind == "i" else "f32"
        if dtype == "bf16":
            host = to_bf16_bits(arr)
        elif dtype == "i32":
            host = np.ascontiguousarray(arr, dtype=np.int32)
        else:
            host = np.ascontiguousarray(arr, dtype=np.float32)
        out = cls(arr.shape, dtype)
        lib = _fa()
        _lib.check(lib, lib.fa_h2d(out.ptr, host.ctypes.data_as(ctypes.c_void_p), out.nbytes))
        return out

    def to_numpy(self) -> np.ndarray:
        lib = _fa()
        host = np.empty(self.size, dtype={"f32": np.float32, "bf16": np.uint16, "i32": np.int32}[self.dtype])
        _lib.check(lib, lib.fa_d2h(host.ctypes.data_as(ctypes.c_void_p), self.ptr, self.nbytes))
        if self.dtype == "bf16":
            host = from_bf16_bits(host)
        return host.reshape(self.shape)

    def fill_bytes(self, byte: int = 0) -> None:
        lib = _fa()
        _lib.check(lib, lib.fa_memset(self.ptr, byte, self.nbytes))

    def free(self) -> None:
        if getattr(self, "ptr", None):
            _fa().fa_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _desc(B, H, N, d, dtype, causal, strides, kv_len, key_mask):
    a = _lib.fa_attn_desc()
    a.B, a.H, a.N, a.d = B, H, N, d
    a.dtype = _lib.FA_DTYPE_BF16 if dtype == "bf16" else _lib.FA_DTYPE_F32
    a.causal = int(bool(causal))
    a.stride_b, a.stride_h, a.stride_n = strides if strides else (0, 0, 0)
    a.kv_len = kv_len.ptr if kv_len is not None else None
    a.key_mask = key_mask.ptr if key_mask is not None else None
    return a


def flash_fwd(Q: DeviceArray, K: DeviceArray, V: DeviceArray, causal=False, kv_len: Optional[DeviceArray] = None,
              key_mask: Optional[DeviceArray] = None, shape=None, strides=None, out=None, stream=None):
    """O, m, l = flash attention forward on device tensors.  `shape` = (B,H,N,d) defaults to Q.shape;
    `strides` = (stride_b, stride_h, stride_n) in elements for non-contiguous layouts."""
    B, H, N, d = shape if shape else Q.shape
    lib = _fa()
    if out is None:
        out = (DeviceArray(Q.shape, Q.dtype), DeviceArray((B, H, N), "f32"), DeviceArray((B, H, N), "f32"))
    O, m, l = out
    a = _desc(B, H, N, d, Q.dtype, causal, strides, kv_len, key_mask)
    _lib.check(lib, lib.fa_flash_fwd_dev(ctypes.byref(a), Q.ptr, K.ptr, V.ptr, O.ptr, m.ptr, l.ptr, stream))
    return O, m, l


def flash_bwd(Q, K, V, O, dO, m, l, causal=False, kv_len=None, key_mask=None, shape=None, strides=None, out=None,
              stream=None):
    B, H, N, d = shape if shape else Q.shape
    lib = _fa()
    if out is None:
        out = tuple(DeviceArray(Q.shape, Q.dtype) for _ in range(3))
    dQ, dK, dV = out
    a = _desc(B, H, N, d, Q.dtype, causal, strides, kv_len, key_mask)
    _lib.check(lib, lib.fa_flash_bwd_dev(ctypes.byref(a), Q.ptr, K.ptr, V.ptr, O.ptr, dO.ptr, m.ptr, l.ptr, dQ.ptr,
                                         dK.ptr, dV.ptr, stream))
    return dQ, dK, dV


def attn_flops(B, H, N, d, causal=False, kv_len: Optional[np.ndarray] = None, backward=False) -> float:
    lib = _fa()
    kp = None
    if kv_len is not None:
        kv = np.ascontiguousarray(kv_len, dtype=np.int32)
        kp = kv.ctypes.data_as(ctypes.c_void_p)
    return float(lib.fa_attn_flops(B, H, N, d, int(bool(causal)), kp, int(bool(backward))))


def sync() -> None:
    lib = _fa()
    _lib.check(lib, lib.fa_sync())


class Timer:
    """CUDA-event timer on the stream the kernels are launched on (the default stream)."""

    def __init__(self):
        lib = _fa()
        self.a, self.b = lib.fa_event_create(), lib.fa_event_create()

    def start(self, stream=None):
        _fa().fa_event_record(self.a, stream)

    def stop(self, stream=None) -> float:
        lib = _fa()
        lib.fa_event_record(self.b, stream)
        ms = lib.fa_event_elapsed_ms(self.a, self.b)
        _lib.check(lib)
        return float(ms)
