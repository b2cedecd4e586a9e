#!/usr/bin/env python
"""One tcgen05 GEMM shape, a few launches (for ncu): python tools/gemm_one.py M N K [a_mn b_mn]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashattn_b200 import _lib  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
a_mn, b_mn = (int(x) for x in sys.argv[4:6]) if len(sys.argv) > 5 else (0, 1)
lib = _lib.load("combine")
fl = _lib.load("flashattention_kernel")
a, b = dev.DeviceArray((M * K,), "bf16"), dev.DeviceArray((K * N,), "bf16")
a.fill_bytes(0x3c), b.fill_bytes(0x3c)
out = dev.DeviceArray((M, N), "f32")
t = dev.Timer()
for i in range(6):
    fl.fa_flush_l2()
    t.start()
    _lib.check(lib, lib.fa_gemm_bf16_dev(out.ptr, 0, N, a.ptr, a_mn, M if a_mn else K, b.ptr, b_mn, N if b_mn else K, M, N, K, None))
    ms = t.stop()
    print(f"launch {i}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
