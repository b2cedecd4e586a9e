#!/usr/bin/env python
"""GPU diagnostic for the tcgen05 forward: runs a handful of shapes through the device API and
prints where (which row block / column chunk) the result departs from the oracle.  Used while
bringing the kernel up; not part of the product or the test-suite."""
import ctypes
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402
from flashattn_b200 import device as dev  # noqa: E402
from oracle import attention_ref as R  # noqa: E402


def run(B, H, N, d, causal, pmode, kv=None, mask=False, seed=0):
    lib = fb._lib.load("flashattention_kernel")
    rng = np.random.default_rng(seed)
    Q, K, V = (R.round_bf16(rng.standard_normal((B, H, N, d)).astype(np.float32)) for _ in range(3))
    kv_len = None
    km = None
    if kv is not None:
        kv_len = np.asarray(kv, dtype=np.int32)
    if mask:
        valid = rng.integers(1, N + 1, B)
        km = np.where(np.arange(N)[None, :] < valid[:, None], 0.0, -1e8).astype(np.float32)
    Oe, me, le = R.attention_fwd(Q, K, V, causal=causal, kv_len=kv_len, key_mask=km)
    dQ, dK, dV = (dev.DeviceArray.from_numpy(x, "bf16") for x in (Q, K, V))
    dkv = dev.DeviceArray.from_numpy(kv_len) if kv_len is not None else None
    dkm = dev.DeviceArray.from_numpy(km) if km is not None else None
    O, m, l = dev.flash_fwd(dQ, dK, dV, causal=causal, kv_len=dkv, key_mask=dkm)
    dev.sync()
    Og, mg, lg = O.to_numpy(), m.to_numpy(), l.to_numpy()
    err = np.abs(Og - Oe)
    lse_err = np.abs((mg + np.log(lg)) - (me + np.log(le)))
    tag = f"B{B} H{H} N{N} d{d} causal={int(causal)} pmode={pmode} kv={kv} mask={int(mask)}"
    print(f"{tag}: max|dO|={err.max():.3e} mean={err.mean():.3e} max|dLSE|={np.nanmax(lse_err):.3e} "
          f"nan={int(np.isnan(Og).sum())}", flush=True)
    if err.max() > 2e-2 or np.isnan(Og).any():
        e = np.nan_to_num(err[0, 0], nan=9.0)
        rb = [(r, float(e[r:r + 32].max())) for r in range(0, N, 32)]
        cb = [(c, float(e[:, c:c + 16].max())) for c in range(0, d, 16)]
        print("   row-block max err:", " ".join(f"{r}:{v:.2g}" for r, v in rb[:24]))
        print("   col-block max err:", " ".join(f"{c}:{v:.2g}" for c, v in cb))
        print("   got[0,0,0,:8]  ", np.round(Og[0, 0, 0, :8], 4))
        print("   want[0,0,0,:8] ", np.round(Oe[0, 0, 0, :8], 4))
        print("   got[0,0,1,:8]  ", np.round(Og[0, 0, 1, :8], 4))
        print("   want[0,0,1,:8] ", np.round(Oe[0, 0, 1, :8], 4))
        print("   m got/want", mg[0, 0, :4], me[0, 0, :4], " l got/want", lg[0, 0, :4], le[0, 0, :4])
    return float(err.max())


if __name__ == "__main__":
    pmodes = [0]
    cases = [
        (1, 1, 128, 128, False, None, False),
        (1, 1, 128, 64, False, None, False),
        (1, 2, 256, 128, False, None, False),
        (2, 2, 512, 128, False, None, False),
        (2, 2, 512, 128, True, None, False),
        (1, 2, 200, 128, False, None, False),
        (1, 2, 1000, 64, True, None, False),
        (2, 2, 512, 128, False, [300, 512], False),
        (2, 2, 384, 64, True, [129, 384], False),
        (2, 1, 384, 128, False, None, True),
        (1, 4, 2048, 128, True, None, False),
    ]
    for pm in pmodes:
        for (B, H, N, d, c, kv, mk) in cases:
            try:
                run(B, H, N, d, c, pm, kv, mk)
            except Exception:
                traceback.print_exc()
                print("ABORT: CUDA context is likely dead", flush=True)
                sys.exit(2)
