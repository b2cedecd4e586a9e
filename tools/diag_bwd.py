#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 backward (bring-up tool, not part of the product)."""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402
from oracle import attention_ref as R  # noqa: E402


def run(B, H, N, d, causal, kv=None, seed=0):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = (R.round_bf16(rng.standard_normal((B, H, N, d)).astype(np.float32)) for _ in range(4))
    kv_len = np.asarray(kv, dtype=np.int32) if kv is not None else None
    ge = R.attention_bwd(Q, K, V, dO, causal=causal, kv_len=kv_len)
    dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(x, "bf16") for x in (Q, K, V, dO))
    dkv = dev.DeviceArray.from_numpy(kv_len) if kv_len is not None else None
    O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal, kv_len=dkv)
    g = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal, kv_len=dkv)
    dev.sync()
    tag = f"B{B} H{H} N{N} d{d} causal={int(causal)} kv={kv}"
    out = []
    for name, got, want in zip(("dQ", "dK", "dV"), g, ge):
        gg = got.to_numpy()
        err = np.abs(gg - want)
        out.append(f"{name}: max={np.nanmax(err):.3e} (|ref|max {np.abs(want).max():.2f}) nan={int(np.isnan(gg).sum())}")
        if not (np.nanmax(err) < 5e-2) or np.isnan(gg).any():
            e = np.nan_to_num(err[0, 0], nan=9.0)
            rb = [(r, float(e[r:r + 32].max())) for r in range(0, N, 32)]
            cb = [(c, float(e[:, c:c + 16].max())) for c in range(0, d, 16)]
            print(f"   {name} row-block max err:", " ".join(f"{r}:{v:.2g}" for r, v in rb[:20]))
            print(f"   {name} col-block max err:", " ".join(f"{c}:{v:.2g}" for c, v in cb))
            print(f"   {name} got[0,0,0,:6] ", np.round(gg[0, 0, 0, :6], 4), " want ", np.round(want[0, 0, 0, :6], 4))
            print(f"   {name} got[0,0,N-1,:6] ", np.round(gg[0, 0, N - 1, :6], 4), " want ", np.round(want[0, 0, N - 1, :6], 4))
    print(tag, " | ".join(out), flush=True)


if __name__ == "__main__":
    cases = [
        (1, 1, 128, 128, False, None), (1, 1, 128, 64, False, None), (1, 2, 256, 128, False, None),
        (1, 2, 256, 128, True, None), (2, 2, 512, 128, False, None), (2, 2, 512, 64, True, None),
        (1, 2, 200, 128, False, None), (1, 2, 1000, 64, True, None), (2, 2, 512, 128, False, [300, 512]),
        (2, 2, 384, 64, True, [129, 384]), (1, 4, 2048, 128, True, None),
    ]
    for c in cases:
        try:
            run(*c)
        except Exception:
            traceback.print_exc()
            print("ABORT: CUDA context is likely dead", flush=True)
            sys.exit(2)
