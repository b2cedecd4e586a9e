#!/usr/bin/env python
"""Aggregate fp32 -> bf16 narrowing / bf16 -> fp32 widening rate of the staging threads' conversion routine
(csrc/host_convert.cpp) against the number of threads, on THIS host: is the legacy ABI's staged route bound by the
cores or by the host's memory system?  (ctypes releases the GIL, so Python threads run the routine concurrently.)
One JSON line; rates in GB/s counted on the fp32 side."""
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb  # noqa: E402

lib = fb._lib.load("flashattention_kernel")
lib.fa_host_narrow_f32_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
lib.fa_host_widen_bf16_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
PER = 1 << 24                      # 64 MiB of fp32 per thread and pass
NMAX = min(16, os.cpu_count() or 1)
src = np.ones(PER * NMAX, dtype=np.float32)
dst = np.zeros(PER * NMAX, dtype=np.uint16)
out = {"cores": os.cpu_count(), "narrow_GBps": {}, "widen_GBps": {}}
for name, key in (("narrow", "narrow_GBps"), ("widen", "widen_GBps")):
    for nt in [n for n in (1, 2, 4, 8, 16) if n <= NMAX]:
        def work(i):
            s, d = src.ctypes.data + 4 * PER * i, dst.ctypes.data + 2 * PER * i
            for _ in range(3):
                if name == "narrow":
                    lib.fa_host_narrow_f32_bf16(d, s, PER)
                else:
                    lib.fa_host_widen_bf16_f32(s, d, PER)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(nt)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        out[key][str(nt)] = round(3 * nt * PER * 4 / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(out))
