#!/usr/bin/env python
"""The split-KV decode-attention kernel alone at long-cache shapes: bytes of K and V cache read / time against the
measured HBM peak.  One JSON line per shape.  MINITORCH_FA_DECODE_SPLITS forces the split
count (see fa_flash_decode_dev)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flashattn_b200 import _lib, device as dev  # noqa: E402
from tools.bench_extra import peaks, time_call  # noqa: E402

P = peaks()
fl = _lib.load("flashattention_kernel")
tag = {k: os.environ[k] for k in ("MINITORCH_FA_DECODE_SPLITS",) if k in os.environ}
for (B, H, L, d, dt_) in [(1, 32, 8192, 128, "bf16"), (8, 32, 8192, 128, "bf16"), (64, 32, 8192, 128, "bf16"),
                          (64, 32, 8192, 64, "bf16"), (1, 32, 8192, 128, "f32"), (64, 32, 8192, 128, "f32"),
                          (128, 8, 40, 32, "f32")]:
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
                      "frac_of_hbm_peak": gbs / P["hbm"], **tag}), flush=True)
    for t in (q, kc, vc, o):
        t.free()
