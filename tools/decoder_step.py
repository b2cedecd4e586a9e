#!/usr/bin/env python
"""BASELINE config #2: one training step (forward + loss + backward) of the run_machine_translation decoder
(DecoderLM n_vocab 10000, n_embd 256, 8 heads, seq 40 -> 39 after the label shift, batch 128, causal) with the
attention core on each of the module's three branches.  Everything else goes through combine.so's host-pointer
plumbing (one H2D/D2H round trip per op, as in the reference), which is what dominates the step -- see
SURVEY.md 8(f)-1.  Prints one JSON line per branch."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from tests.test_host_modules import decoder_loss  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    residency = sys.argv[2] if len(sys.argv) > 2 else "host"      # "host": reference-style host storage; "device": HBM
    gemm = sys.argv[3] if len(sys.argv) > 3 else "fp32"           # "bf16": Linear fwd / bwd on the tcgen05 GEMM (device only)
    flash = sys.argv[4] if len(sys.argv) > 4 else "fp32"          # "bf16": tcgen05 attention kernels (head dim 32 -> D = 64 tiles)
    if residency == "device":
        fb.DeviceKernelOps.set_gemm_mode(gemm)
        fb.DeviceKernelOps.set_flash_mode(flash)
    backend = fb.TensorBackend(fb.DeviceKernelOps) if residency == "device" else fb.default_backend()
    n_vocab, n_embd, n_head, n_pos = 10000, 256, 8, 40
    rng = np.random.default_rng(11111)
    ids = rng.integers(0, n_vocab, (B, n_pos))
    w = np.zeros((B, n_pos), np.float32)
    w[:, n_pos // 2:] = 1.0
    z = dict(input_ids=ids[:, :-1], labels=ids[:, 1:], label_token_weights=w[:, 1:])
    models = {}
    branches = ("flash", "fused", "composed") + (("flash+lookup",) if residency == "device" else ())
    for branch in branches:
        np.random.seed(5)
        models[branch] = fb.DecoderLM(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_positions=n_pos, p_dropout=0.0,
                                      ln_eps=1e-5, bias=True, backend=backend, use_flash_attention=branch.startswith("flash"),
                                      use_fused_kernel=branch == "fused", use_fused_embedding=branch == "flash+lookup")
    times = {b: [] for b in models}
    loss = {}
    for rnd in range(4):                      # round 0 = warm-up; branches interleaved so none owns the cold start
        for branch, model in models.items():
            t0 = time.perf_counter()
            _, total = decoder_loss(model, z, backend=backend, fused_loss=branch == "flash+lookup")
            total.backward()
            if rnd:
                times[branch].append(time.perf_counter() - t0)
            loss[branch] = float(total.to_numpy().reshape(-1)[0])
    for branch in models:
        print(json.dumps({"workload": f"DecoderLM cfg2 step (fwd + loss + bwd), batch {B}, seq 39, fp32",
                          "storage": residency, "gemm": gemm if residency == "device" else "fp32",
                          "flash_mode": flash if residency == "device" else "fp32",
                          "attention": branch, "step_s_best": min(times[branch]), "step_s_all": times[branch],
                          "loss": loss[branch]}), flush=True)
    fb.CudaKernelOps.set_flash_mode("fp32")
    fb.DeviceKernelOps.set_gemm_mode("fp32")
    fb.DeviceKernelOps.set_flash_mode("fp32")


if __name__ == "__main__":
    main()
