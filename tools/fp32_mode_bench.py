#!/usr/bin/env python
"""Throughput of the fp32-accurate mode (FA_MODE_FP32, CUDA-core kernels in flash_fp32.cuh) through the device API:
fwd and bwd at a few shapes, CUDA-event timed.  One JSON line per shape."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402

lib = fb._lib.load("flashattention_kernel")
rng = np.random.default_rng(0)
for (B, H, N, d, causal) in [(2, 4, 64, 32, False), (128, 8, 39, 32, True), (8, 16, 1024, 64, True), (8, 16, 4096, 128, False),
                             (2, 2, 4096, 1024, True)]:
    Q, K, V, dO = (dev.DeviceArray.from_numpy(rng.standard_normal((B, H, N, d), dtype=np.float32), "f32") for _ in range(4))
    O, m, l = dev.flash_fwd(Q, K, V, causal=causal)
    g = dev.flash_bwd(Q, K, V, O, dO, m, l, causal=causal)
    dev.sync()
    reps = 5 if N >= 4096 else 20
    ev = [lib.fa_event_create() for _ in range(3)]
    lib.fa_event_record(ev[0], None)
    for _ in range(reps):
        dev.flash_fwd(Q, K, V, causal=causal, out=(O, m, l))
    lib.fa_event_record(ev[1], None)
    for _ in range(reps):
        dev.flash_bwd(Q, K, V, O, dO, m, l, causal=causal, out=g)
    lib.fa_event_record(ev[2], None)
    f_ms = lib.fa_event_elapsed_ms(ev[0], ev[1]) / reps
    b_ms = lib.fa_event_elapsed_ms(ev[1], ev[2]) / reps
    ff, fb_ = dev.attn_flops(B, H, N, d, causal, None, False), dev.attn_flops(B, H, N, d, causal, None, True)
    print(json.dumps({"mode": "fp32", "B": B, "H": H, "N": N, "d": d, "causal": causal, "fwd_ms": f_ms, "bwd_ms": b_ms,
                      "fwd_tflops": ff / f_ms / 1e9, "bwd_tflops": fb_ / b_ms / 1e9}), flush=True)
