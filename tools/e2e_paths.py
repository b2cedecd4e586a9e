#!/usr/bin/env python
"""End-to-end (host fp32 buffers in and out) time of one config-#4 step through the two public entry points
(bench.run_e2e), then the forward and the backward call of the raw C ABI timed separately on page-locked buffers, and
the raw link rates (one 512 MiB page-locked tensor each way) they are to be compared with.
MINITORCH_FA_HYBRID / _HYBRID_COST / _COPY_THREADS select the transfer policy (legacy_pipeline.cuh).
One JSON line per measurement."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import flashattn_b200 as fb  # noqa: E402

lib = fb._lib.load("flashattention_kernel")
B, H, N, d = 8, 32, 4096, 128
kv_len = np.random.default_rng(2).integers(N // 2, N + 1, B).astype(np.int32)
tag = {k: os.environ[k] for k in ("MINITORCH_FA_HYBRID", "MINITORCH_FA_HYBRID_COST", "MINITORCH_FA_COPY_THREADS")
       if k in os.environ}
what = sys.argv[1:] or ["ops", "pinned", "calls", "link"]
for path in [w for w in what if w in ("ops", "pinned")]:
    r = bench.run_e2e(fb, lib, B, H, N, d, False, kv_len, 4, None, path)
    print(json.dumps({"path": path, "ms_per_step": round(r["ms_per_step"], 2), "tflops": round(r["value"], 1),
                      "h2d_MB": r["h2d_bytes_per_step"] >> 20, "d2h_MB": r["d2h_bytes_per_step"] >> 20, **tag}), flush=True)

n, r_ = B * H * N * d, B * H * N
ptrs = []


def pinned(count):
    p = lib.fa_malloc_host(count * 4)
    assert p
    ptrs.append(p)
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(count,))


def wire():
    a, b = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    lib.fa_wire_bytes(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


if "calls" in what:
    A = {k: pinned(n) for k in ("Q", "K", "V", "O", "dO", "dQ", "dK", "dV")}
    S = {k: pinned(r_) for k in ("l", "m")}
    base = np.random.default_rng(5).standard_normal(H * N * d, dtype=np.float32)
    for k in ("Q", "K", "V", "dO"):
        A[k].reshape(B, -1)[:] = base[None, :]
    mask = np.where(np.arange(N)[None, :] < kv_len[:, None], 0.0, -1e8).astype(np.float32)
    mptr = mask.ctypes.data_as(ctypes.c_void_p)
    lib.fa_set_mode(fb._lib.FA_MODE_BF16)
    tf, tb, wf, wb = [], [], None, None
    for it in range(5):
        w0 = wire()
        t0 = time.perf_counter()
        lib.launch_flashattention_forward_masked(A["Q"], A["K"], A["V"], A["O"], S["l"], S["m"], mptr, 0, B, H, N, d)
        fb._lib.check(lib)
        t1 = time.perf_counter()
        w1 = wire()
        lib.launch_flashattention_backward_masked(A["Q"], A["K"], A["V"], A["O"], A["dQ"], A["dK"], A["dV"], A["dO"],
                                                  S["l"], S["m"], mptr, 0, B, H, N, d)
        fb._lib.check(lib)
        t2 = time.perf_counter()
        w2 = wire()
        if it:
            tf.append(t1 - t0), tb.append(t2 - t1)
            wf, wb = [(b - a) >> 20 for a, b in zip(w0, w1)], [(b - a) >> 20 for a, b in zip(w1, w2)]
    lib.fa_set_mode(fb._lib.FA_MODE_FP32)
    lib.fa_release_staging()
    print(json.dumps({"calls": "pinned", "fwd_ms": round(1e3 * float(np.mean(tf)), 2), "bwd_ms": round(1e3 * float(np.mean(tb)), 2),
                      "fwd_h2d_d2h_MB": wf, "bwd_h2d_d2h_MB": wb, **tag}), flush=True)

if "link" in what:
    from flashattn_b200 import device as dev
    h = pinned(n)
    h[:] = 1.0
    dbuf = dev.DeviceArray((n,), "f32")
    res = {}
    for name, fn in (("h2d", lambda: lib.fa_h2d(ctypes.c_void_p(dbuf.ptr), h.ctypes.data_as(ctypes.c_void_p), n * 4)),
                     ("d2h", lambda: lib.fa_d2h(h.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(dbuf.ptr), n * 4))):
        fn()
        lib.fa_sync()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        lib.fa_sync()
        res[name + "_GBps"] = round(3 * n * 4 / (time.perf_counter() - t0) / 1e9, 1)
    # the host threads' conversion rates (one thread; the staging pool runs up to 16 of them)
    src = np.ones(1 << 26, dtype=np.float32)
    dst = np.empty(1 << 26, dtype=np.uint16)
    lib.fa_host_narrow_f32_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    lib.fa_host_widen_bf16_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    for name, fn in (("narrow_1thread", lambda: lib.fa_host_narrow_f32_bf16(dst.ctypes.data, src.ctypes.data, src.size)),
                     ("widen_1thread", lambda: lib.fa_host_widen_bf16_f32(src.ctypes.data, dst.ctypes.data, src.size))):
        fn()
        t0 = time.perf_counter()
        fn()
        res[name + "_GBps_fp32_side"] = round(src.nbytes / (time.perf_counter() - t0) / 1e9, 1)
    print(json.dumps({"link": res, "cores": os.cpu_count()}), flush=True)
for p in ptrs:
    lib.fa_free_host(p)
