#!/usr/bin/env python
"""Throughput of the tcgen05 bf16 GEMM (csrc/gemm_sm100.cuh) at Linear-layer shapes, CUDA events, L2 flushed when the
operands are small; prints one JSON line per shape with the fraction of the measured cuBLAS bf16 peak."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flashattn_b200 import _lib  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402
from tools.bench_extra import peaks, time_call  # noqa: E402

lib = _lib.load("combine")
fl = _lib.load("flashattention_kernel")
P = peaks()
for (M, N, K, a_mn, b_mn, what) in [(8192, 8192, 8192, 0, 1, "square"), (131072, 2048, 2048, 0, 1, "Linear fwd, reference grid max"),
                                    (131072, 2048, 2048, 0, 0, "Linear dx = dy.W^T"), (2048, 2048, 131072, 1, 1, "Linear dW = x^T.dy"),
                                    (4992, 256, 256, 0, 1, "cfg2 projection"), (4992, 10000, 256, 0, 1, "cfg2 lm_head"),
                                    (32768, 12288, 4096, 0, 1, "fused QKV, E=4096")]:
    a = dev.DeviceArray((M * K,), "bf16")
    b = dev.DeviceArray((K * N,), "bf16")
    a.fill_bytes(0x3c)
    b.fill_bytes(0x3c)
    out = dev.DeviceArray((M, N), "f32")
    lda = M if a_mn else K
    ldb = N if b_mn else K
    fn = lambda: _lib.check(lib, lib.fa_gemm_bf16_dev(out.ptr, 0, N, a.ptr, a_mn, lda, b.ptr, b_mn, ldb, M, N, K, None))  # noqa: E731
    avg, best = time_call(fn, fl, reps=5, flush=(M * K + K * N) * 2 < (256 << 20))
    tf = 2.0 * M * N * K / (avg * 1e-3) / 1e12
    print(json.dumps(dict(shape=[M, N, K], a_mn=a_mn, b_mn=b_mn, what=what, ms=avg, tflops=tf,
                          frac_of_measured_burst=tf / P["tf_burst"])), flush=True)
