"""Diagnose test_bf16_lazy_rescale_divergent_rows: where does dQ/dK/dV leave its bound?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb
from flashattn_b200 import device as dev
from oracle import attention_ref as R

for causal in (False, True):
    for d in (128, 64):
        B, H, N = 1, 2, 640
        rng = np.random.default_rng(77)
        Q, K, V, dO = (rng.standard_normal((B, H, N, d)).astype(np.float32) for _ in range(4))
        Q[..., 0] = 0.0
        K[..., 0] = 0.0
        for rows, key in (((0, 32), 300), ((200, 211), 420), ((500, 541), 600), ((5, 9), 639)):
            Q[:, :, rows[0]:rows[1], 0] = 4.0
            K[:, :, key, 0] = 40.0 if key != 639 else 80.0
        Q, K, V, dO = (R.round_bf16(x) for x in (Q, K, V, dO))
        dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(x, "bf16") for x in (Q, K, V, dO))
        O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal)
        gq, gk, gv = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal)
        ge = R.attention_bwd(Q, K, V, dO, causal=causal)
        # same backward but fed the ORACLE's O (rounded to bf16) and statistics: separates fwd rounding from bwd
        Oe, me, le = R.attention_fwd(Q, K, V, causal=causal)
        O2 = dev.DeviceArray.from_numpy(R.round_bf16(Oe.astype(np.float32)), "bf16")
        m2 = dev.DeviceArray.from_numpy(me.astype(np.float32)); l2 = dev.DeviceArray.from_numpy(le.astype(np.float32))
        hq, hk, hv = dev.flash_bwd(dq, dk, dv, O2, ddo, m2, l2, causal=causal)
        for name, got, got2, want in zip(("dQ", "dK", "dV"), (gq, gk, gv), (hq, hk, hv), ge):
            g = got.to_numpy().astype(np.float64); g2 = got2.to_numpy().astype(np.float64)
            err = np.abs(g - want)
            idx = np.unravel_index(np.argmax(err), err.shape)
            print(f"causal={causal} d={d} {name}: max err {err.max():.4f} at {idx} got {g[idx]:.4f} want {want[idx]:.4f} "
                  f"(with oracle O/m/l: got {g2[idx]:.4f}, max err {np.abs(g2 - want).max():.4f})")
            if name == "dQ":
                rows = np.argsort(-err.max(-1)[0, idx[1]])[:8]
                print("   worst rows:", [(int(r), round(float(err[0, idx[1], r].max()), 3), int(err[0, idx[1], r].argmax())) for r in rows])
