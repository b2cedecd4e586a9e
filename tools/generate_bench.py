#!/usr/bin/env python
"""Decode regime: greedy generation as the reference does it (full-prefix recompute per token,
batch 1) with BASELINE config #2's model, host storage vs device-resident storage.  One JSON line per arm."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402

out = {}
for residency in ("device", "host"):
    backend = fb.TensorBackend(fb.DeviceKernelOps) if residency == "device" else fb.default_backend()
    for branch in ("flash", "composed"):
        np.random.seed(5)
        model = fb.DecoderLM(n_vocab=10000, n_embd=256, n_head=8, n_positions=40, p_dropout=0.0, ln_eps=1e-5, bias=True,
                             backend=backend, use_flash_attention=branch == "flash",
                             use_fused_embedding=residency == "device")
        prompt = list(np.random.default_rng(1).integers(0, 10000, 10))
        fb.generate(model, prompt, model_max_length=12)            # warm-up
        t0 = time.perf_counter()
        ids = fb.generate(model, prompt, model_max_length=40)
        dt = time.perf_counter() - t0
        out[(residency, branch)] = ids
        print(json.dumps({"workload": "greedy generate, DecoderLM cfg2, batch 1, prefix 10 -> 41 tokens (full-prefix recompute)",
                          "storage": residency, "attention": branch, "new_tokens": len(ids) - len(prompt),
                          "seconds": dt, "tokens_per_s": (len(ids) - len(prompt)) / dt}), flush=True)
# cache-aware generation: prefill once, then one position per token through the split-KV decode kernel
backend = fb.TensorBackend(fb.DeviceKernelOps)
np.random.seed(5)
model = fb.DecoderLM(n_vocab=10000, n_embd=256, n_head=8, n_positions=40, p_dropout=0.0, ln_eps=1e-5, bias=True,
                     backend=backend, use_flash_attention=True, use_fused_embedding=True)
prompt = list(np.random.default_rng(1).integers(0, 10000, 10))
fb.generate_cached(model, prompt, model_max_length=12)
t0 = time.perf_counter()
ids = fb.generate_cached(model, prompt, model_max_length=40)
dt = time.perf_counter() - t0
out[("device", "cached")] = ids
print(json.dumps({"workload": "greedy generate_cached, DecoderLM cfg2, batch 1, prefix 10 -> 41 tokens (KV cache + split-KV decode)",
                  "storage": "device", "attention": "flash + decode kernel", "new_tokens": len(ids) - len(prompt),
                  "seconds": dt, "tokens_per_s": (len(ids) - len(prompt)) / dt}), flush=True)
assert out[("device", "cached")] == out[("device", "flash")], "cached generation differs from the full-prefix loop"

# the decode-attention kernel alone at long-cache shapes: bytes of K and V cache read / time vs the measured HBM peak
import ctypes  # noqa: E402
from flashattn_b200 import _lib, device as dev  # noqa: E402
from tools.bench_extra import peaks, time_call  # noqa: E402
P = peaks()
fl = _lib.load("flashattention_kernel")
for (B, H, L, d, dt_) in [(1, 32, 8192, 128, "bf16"), (8, 32, 8192, 128, "bf16"), (64, 32, 8192, 128, "bf16"),
                          (64, 32, 8192, 128, "f32"), (128, 8, 40, 32, "f32")]:
    q = dev.DeviceArray((B, H, d), dt_)
    kc, vc = dev.DeviceArray((B, H, L, d), dt_), dev.DeviceArray((B, H, L, d), dt_)
    for t in (q, kc, vc):
        t.fill_bytes(0x3c)
    o = dev.DeviceArray((B, H, d), dt_)
    a = _lib.fa_decode_desc()
    a.B, a.H, a.d, a.L, a.L_cap = B, H, d, L, L
    a.dtype = _lib.FA_DTYPE_BF16 if dt_ == "bf16" else _lib.FA_DTYPE_F32
    fn = lambda: _lib.check(fl, fl.fa_flash_decode_dev(ctypes.byref(a), q.ptr, kc.ptr, vc.ptr, o.ptr, None, None))  # noqa: E731
    avg, best = time_call(fn, fl, reps=10, flush=kc.nbytes < (128 << 20))
    gbs = 2.0 * kc.nbytes / (avg * 1e-3) / 1e9
    print(json.dumps({"op": "flash_decode", "shape": [B, H, L, d], "dtype": dt_, "ms": avg, "gbs": gbs,
                      "frac_of_hbm_peak": gbs / P["hbm"]}), flush=True)
assert out[("device", "flash")] == out[("host", "flash")] == out[("device", "composed")] == out[("host", "composed")], \
    "generated ids differ between arms"
print(json.dumps({"same_tokens_in_all_arms": True}))
