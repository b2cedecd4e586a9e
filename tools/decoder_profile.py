#!/usr/bin/env python
"""cProfile of one DecoderLM cfg2 step per attention branch (where does the host time go?)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from tests.test_host_modules import decoder_loss  # noqa: E402

backend = fb.TensorBackend(fb.DeviceKernelOps) if os.environ.get("RESIDENCY") == "device" else fb.default_backend()
n_vocab, n_embd, n_head, n_pos, B = 10000, 256, 8, 40, 128
rng = np.random.default_rng(11111)
ids = rng.integers(0, n_vocab, (B, n_pos))
w = np.zeros((B, n_pos), np.float32)
w[:, n_pos // 2:] = 1.0
z = dict(input_ids=ids[:, :-1], labels=ids[:, 1:], label_token_weights=w[:, 1:])
order = sys.argv[1:] or ["composed", "flash", "composed", "flash"]
for branch in order:
    np.random.seed(5)
    model = fb.DecoderLM(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_positions=n_pos, p_dropout=0.0, ln_eps=1e-5,
                         bias=True, backend=backend, use_flash_attention=branch == "flash")

    def step():
        _, total = decoder_loss(model, z, backend=backend)
        total.backward()

    step()
    pr = cProfile.Profile()
    pr.enable()
    step()
    pr.disable()
    print("=====", branch)
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
