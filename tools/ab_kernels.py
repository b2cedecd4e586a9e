#!/usr/bin/env python
"""Interleaved A/B timing of alternative builds of flashattention_kernel.so IN ONE PROCESS (same inputs, same clocks
and temperature history): usage  ab_kernels.py [fwd|bwd] dir1 dir2 ...   -> per build: median / min ms over rounds."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import _lib  # noqa: E402

what, dirs = sys.argv[1], sys.argv[2:]
libs = []
for d in dirs:
    lib = ctypes.CDLL(os.path.join(ROOT, d, "flashattention_kernel.so"))
    for sym, (restype, argtypes) in _lib.SYMBOLS["flashattention_kernel"].items():
        if hasattr(lib, sym):
            getattr(lib, sym).restype, getattr(lib, sym).argtypes = restype, argtypes
    libs.append(lib)
L0 = libs[0]
d = int(os.environ.get("AB_D", "128"))
B, H, N = int(os.environ.get("AB_B", "8")), int(os.environ.get("AB_H", "32")), int(os.environ.get("AB_N", "4096"))
CAUSAL, PAD = int(os.environ.get("AB_CAUSAL", "0")), int(os.environ.get("AB_PAD", "1"))
n = B * H * N * d
rng = np.random.default_rng(0)
kv = rng.integers(N // 2, N + 1, B).astype(np.int32)


def dev_bf16():
    p = L0.fa_malloc(n * 2)
    per = H * N * d
    for b in range(B):
        u = fb.device.to_bf16_bits(rng.standard_normal(per, dtype=np.float32))
        L0.fa_h2d(ctypes.c_void_p(p + b * per * 2), u.ctypes.data_as(ctypes.c_void_p), per * 2)
    return p


Q, K, V, dO = (dev_bf16() for _ in range(4))
O, dQ, dK, dV = (L0.fa_malloc(n * 2) for _ in range(4))
m, l = L0.fa_malloc(B * H * N * 4), L0.fa_malloc(B * H * N * 4)
dkv = L0.fa_malloc(B * 4)
L0.fa_h2d(dkv, kv.ctypes.data_as(ctypes.c_void_p), B * 4)
a = _lib.fa_attn_desc()
a.B, a.H, a.N, a.d, a.dtype, a.causal, a.kv_len = B, H, N, d, _lib.FA_DTYPE_BF16, CAUSAL, (dkv if PAD else None)


def run(lib):
    if what == "fwd":
        rc = lib.fa_flash_fwd_dev(ctypes.byref(a), Q, K, V, O, m, l, None)
    else:
        rc = lib.fa_flash_bwd_dev(ctypes.byref(a), Q, K, V, O, dO, m, l, dQ, dK, dV, None)
    assert rc == 0, lib.fa_last_error()


L0.fa_flash_fwd_dev(ctypes.byref(a), Q, K, V, O, m, l, None)
for lib in libs:
    for _ in range(5):
        run(lib)
L0.fa_sync()
ROUNDS, ITERS = int(os.environ.get("AB_ROUNDS", "12")), int(os.environ.get("AB_ITERS", "20"))
flops = _lib.load("flashattention_kernel").fa_attn_flops(B, H, N, d, CAUSAL, kv.ctypes.data_as(ctypes.c_void_p) if PAD else None,
                                                         0 if what == "fwd" else 1)
times = [[] for _ in libs]
for r in range(ROUNDS):
    order = list(range(len(libs)))
    if r % 2:
        order.reverse()
    for i in order:
        lib = libs[i]
        e0, e1 = lib.fa_event_create(), lib.fa_event_create()
        lib.fa_event_record(e0, None)
        for _ in range(ITERS):
            run(lib)
        lib.fa_event_record(e1, None)
        times[i].append(lib.fa_event_elapsed_ms(e0, e1) / ITERS)
for dname, t in zip(dirs, times):
    print(f"{dname:24s} {what} B{B} H{H} N{N} d{d} causal{CAUSAL} pad{PAD}: median {np.median(t):.4f} ms  min {min(t):.4f}  "
          f"max {max(t):.4f}  -> {flops / np.median(t) / 1e9:.0f} TFLOP/s")
