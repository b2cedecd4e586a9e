#!/usr/bin/env python
"""SM clock and board power while ONE kernel family runs back to back for a few seconds (cfg4 shapes): where does the
power cap put the forward, the backward and the step?  One JSON line per arm."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402
from bench import ClockSampler  # noqa: E402

lib = fb._lib.load("flashattention_kernel")
B, H, N, d = 8, 32, 4096, 128
rng = np.random.default_rng(0)


def mk():
    a = dev.DeviceArray((B, H, N, d), "bf16")
    per = H * N * d
    for b in range(B):
        u = dev.to_bf16_bits(rng.standard_normal(per, dtype=np.float32))
        lib.fa_h2d(ctypes.c_void_p(a.ptr + b * per * 2), u.ctypes.data_as(ctypes.c_void_p), per * 2)
    return a


Q, K, V, dO = mk(), mk(), mk(), mk()
O, m, l = dev.flash_fwd(Q, K, V)
grads = dev.flash_bwd(Q, K, V, O, dO, m, l)
dev.sync()
ff = dev.attn_flops(B, H, N, d, False, None, False)
fbw = dev.attn_flops(B, H, N, d, False, None, True)
arms = {"fwd": (lambda: dev.flash_fwd(Q, K, V, out=(O, m, l)), ff),
        "bwd": (lambda: dev.flash_bwd(Q, K, V, O, dO, m, l, out=grads), fbw),
        "fwd+bwd": (lambda: (dev.flash_fwd(Q, K, V, out=(O, m, l)), dev.flash_bwd(Q, K, V, O, dO, m, l, out=grads)), ff + fbw)}
for name, (fn, flops) in arms.items():
    s = ClockSampler(0)
    s.start()
    time.sleep(1.2)
    mark = s.mark()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 4.0:
        for _ in range(20):
            fn()
        dev.sync()
        n += 20
    dt = time.perf_counter() - t0
    c = s.stop(mark)
    print(json.dumps({"arm": name, "tflops_sustained_4s": flops * n / dt / 1e12, "ms_per_call": dt / n * 1e3, **c}), flush=True)
    time.sleep(2.0)
