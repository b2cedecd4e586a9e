#!/usr/bin/env python
"""Small driver for profiling the companion kernels under ncu: one launch of each at a large shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flashattn_b200 as fb
from flashattn_b200 import device as dev
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extra import dev_rand
sm = fb._lib.load("softmax_kernel"); ln = fb._lib.load("layernorm_kernel")
B, H, F, T = 8, 16, 2048, 2048
x = dev_rand((B, H, F, T), 3); y = dev_rand((B, H, F, T), 4, 0.0, 1.0)
mask = dev.DeviceArray.from_numpy(np.zeros((B, T), np.float32))
for _ in range(2):
    sm.fa_attn_softmax_dev(x.ptr, mask.ptr, B, H, F, T, 0, None)
    sm.fa_attn_softmax_bw_dev(x.ptr, y.ptr, B * H * F, T, None)
rows, h = 32768, 4096
a = dev_rand((rows, h), 6); dy = dev_rand((rows, h), 7); yb = dev.DeviceArray((rows, h), "f32"); dx = dev.DeviceArray((rows, h), "f32")
g = dev_rand((h,), 8); b = dev_rand((h,), 9)
var, mean = dev.DeviceArray((rows,), "f32"), dev.DeviceArray((rows,), "f32")
dg, db = dev.DeviceArray((h,), "f32"), dev.DeviceArray((h,), "f32")
for _ in range(2):
    ln.fa_layernorm_dev(yb.ptr, var.ptr, mean.ptr, a.ptr, g.ptr, b.ptr, rows, h, None)
    ln.fa_layernorm_bw_dev(dg.ptr, db.ptr, dx.ptr, dy.ptr, a.ptr, g.ptr, b.ptr, var.ptr, mean.ptr, rows, h, None)
dev.sync()
print("ok")
