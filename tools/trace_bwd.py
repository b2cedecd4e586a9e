#!/usr/bin/env python
"""Bring-up tool: run the FA_TRACE build of the backward on a cfg4-like problem and print the per-iteration
timeline (clock64 deltas) of one CTA: MMA-issuer events and compute-group events."""
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
B, H, N, d = 2, 148, 4096, 128   # 2*148*32 CTAs: every SM busy for several waves
rng = np.random.default_rng(0)
x = rng.standard_normal((1, 1, N, d)).astype(np.float32)
def mk():
    a = dev.DeviceArray((B, H, N, d), "bf16")
    bits = dev.to_bf16_bits(np.broadcast_to(x, (1, H, N, d)).copy())
    for b in range(B):
        lib.fa_h2d(ctypes.c_void_p(a.ptr + b * H * N * d * 2), bits.ctypes.data_as(ctypes.c_void_p), H * N * d * 2)
    return a
Q, K, V, dO = mk(), mk(), mk(), mk()
O, m, l = dev.flash_fwd(Q, K, V)
trace = dev.DeviceArray((48 * 16,), "f32")  # reused as raw bytes below
tbuf = lib.fa_malloc(48 * 32 * 8)
lib.fa_memset(tbuf, 0, 48 * 32 * 8)
lib.fa_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.fa_debug_set_trace(tbuf)
for _ in range(2):
    g = dev.flash_bwd(Q, K, V, O, dO, m, l)
dev.sync()
host = np.zeros(48 * 32, dtype=np.int64)
lib.fa_d2h(host.ctypes.data_as(ctypes.c_void_p), tbuf, host.nbytes)
T = host.reshape(48, 32)
t0 = T[2, 0]
names = ["p_full>", "dV iss", "q_full>", "S iss", "ds_full>", "dQdK iss", "dq_free>", "dP iss",
         "s_full>", "S ld", "P done", "dp_full>", "dS done", "dq_full>", "dq ld", "stg done"]
print("iteration rows; times relative to iteration 2's p_full (clk). '>' = wait satisfied")
print("it  " + " ".join(f"{n:>9s}" for n in names))
for it in range(2, 14):
    print(f"{it:2d}  " + " ".join(f"{(T[it, k] - t0) if T[it, k] else -1:9d}" for k in range(16)))
per = (T[12, 0] - T[4, 0]) / 8
print("avg period per iteration:", per, "clk")
print("per-warp (w0..w3): [P done, dS done, s_full seen, dp_full seen] relative to the same origin")
for it in range(6, 12):
    print(f"{it:2d}  " + " | ".join(" ".join(f"{T[it, 16 + 4 * w + k] - t0:7d}" for k in (2, 0, 3, 1)) for w in range(4)))
print("columns per warp: s_full-seen, P-done, dp_full-seen, dS-done")
