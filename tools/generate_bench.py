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
assert out[("device", "flash")] == out[("host", "flash")] == out[("device", "composed")] == out[("host", "composed")], \
    "generated ids differ between arms"
print(json.dumps({"same_tokens_in_all_arms": True}))
