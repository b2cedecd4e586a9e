#!/usr/bin/env python
"""Bring-up tool: per-KV-tile timeline (clock64) of one forward CTA from the FA_TRACE build."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["FLASHATTN_B200_KERNEL_DIR"] = os.path.join(ROOT, "build", "trace")
sys.path.insert(0, ROOT)
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402

lib = fb._lib.load("flashattention_kernel")
B, H, N, d = 2, 148, 4096, 128
rng = np.random.default_rng(0)
x = rng.standard_normal((1, 1, N, d)).astype(np.float32)
def mk():
    a = dev.DeviceArray((B, H, N, d), "bf16")
    bits = dev.to_bf16_bits(np.broadcast_to(x, (1, H, N, d)).copy())
    for b in range(B):
        lib.fa_h2d(ctypes.c_void_p(a.ptr + b * H * N * d * 2), bits.ctypes.data_as(ctypes.c_void_p), H * N * d * 2)
    return a
Q, K, V = mk(), mk(), mk()
tbuf = lib.fa_malloc(48 * 32 * 8)
lib.fa_memset(tbuf, 0, 48 * 32 * 8)
lib.fa_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.fa_debug_set_trace(tbuf)
for _ in range(2):
    O, m, l = dev.flash_fwd(Q, K, V)
dev.sync()
host = np.zeros(48 * 32, dtype=np.int64)
lib.fa_d2h(host.ctypes.data_as(ctypes.c_void_p), tbuf, host.nbytes)
T = host.reshape(48, 32)
t0 = T[4, 0]
mma = ["p0>", "PV0 iss", "k>", "QK0 iss", "p1>", "PV1 iss", "k>", "QK1 iss"]
print("MMA issuer (per KV tile j): wait-satisfied / issue-done stamps, clk relative to j=4")
print(" j  " + " ".join(f"{n:>8s}" for n in mma))
for j in range(4, 14):
    print(f"{j:2d}  " + " ".join(f"{T[j, k] - t0:8d}" for k in range(8)))
print("softmax groups (lane 0 of warp 0 of each): s_full seen, max published, barrier passed, exp start, P_lo arrived, P_hi arrived")
for wg, name in enumerate(["g0 h0", "g1 h0", "g0 h1", "g1 h1"]):
    print(name)
    for j in range(4, 10):
        cols = [8 + 4 * wg, 9 + 4 * wg, 10 + 4 * wg, 28 + wg, 24 + wg, 11 + 4 * wg]
        print(f"  {j:2d}  " + " ".join(f"{T[j, k] - t0:8d}" for k in cols))
print("period per KV tile:", (T[12, 0] - T[4, 0]) / 8)
