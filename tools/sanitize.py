#!/usr/bin/env python
"""Small-shape pass over every kernel family for compute-sanitizer (memcheck): fp32 + bf16 attention
fwd/bwd (causal, ragged N, kv_len), softmax and layernorm fw/bw."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import flashattn_b200 as fb
from flashattn_b200 import device as dev
ops, T = fb.CudaKernelOps, fb.tensor_from_numpy
rng = np.random.default_rng(0)
for dtype, shapes in (("f32", [(1, 2, 39, 32, True), (1, 1, 130, 64, False)]),
                      ("bf16", [(1, 2, 200, 128, True), (1, 1, 300, 64, False), (2, 1, 256, 128, False)])):
    for (B, H, N, d, causal) in shapes:
        Q, K, V, dO = (dev.DeviceArray.from_numpy(rng.standard_normal((B, H, N, d)).astype(np.float32), dtype) for _ in range(4))
        kv = dev.DeviceArray.from_numpy(np.array([N - 7] * B, dtype=np.int32))
        O, m, l = dev.flash_fwd(Q, K, V, causal=causal, kv_len=kv)
        g = dev.flash_bwd(Q, K, V, O, dO, m, l, causal=causal, kv_len=kv)
        dev.sync()
        print("attention", dtype, (B, H, N, d, causal), "ok", float(np.abs(g[0].to_numpy()).max()))
x = rng.uniform(-1, 1, (2, 2, 9, 100)).astype(np.float32)
y = ops.attn_softmax_fw(T(x), T(np.zeros((2, 100), np.float32)))
ops.attn_softmax_bw(T(x), y)
for rows, h in ((33, 64), (5, 2048)):
    a = rng.uniform(-1, 1, (rows, h)).astype(np.float32)
    gm, bt = np.ones(h, np.float32), np.zeros(h, np.float32)
    ln, var, mean = ops.layernorm_fw(T(a), T(gm), T(bt))
    ops.layernorm_bw(T(a), T(a), T(gm), T(bt), var, mean)
print("companions ok")
# legacy host-pointer pipeline (several chunks) and the combine / lookup / loss kernels on device storage
lib = fb._lib.load("flashattention_kernel")
lib.fa_set_legacy_chunk_bytes(2 * 100 * 32 * 4)
for mode in ("fp32", "bf16"):
    ops.set_flash_mode(mode)
    d = 32 if mode == "fp32" else 64
    Q, K, V, dO = (T(rng.standard_normal((2, 5, 100, d)).astype(np.float32)) for _ in range(4))
    O, m, l = ops.flash_attention_causal_fw(Q, K, V)
    ops.flash_attention_causal_bw(Q, K, V, O, dO, m, l)
ops.set_flash_mode("fp32")
lib.fa_set_legacy_chunk_bytes(0)
print("legacy pipeline ok")
DEV = fb.TensorBackend(fb.DeviceKernelOps)
D = lambda a, g=False: fb.tensor_from_numpy(np.asarray(a, np.float32), backend=DEV, requires_grad=g)
a, b = D(rng.standard_normal((3, 4, 5)), True), D(rng.standard_normal((1, 5)))
((a * b).permute(2, 0, 1).contiguous().sum(1) + 1.0).sum().backward()
w, ids = D(rng.standard_normal((11, 6)), True), D(rng.integers(0, 11, (2, 7)))
emb = fb.EmbeddingLookup.apply(ids, w)
logits = D(rng.standard_normal((14, 9)), True)
loss = fb.softmax_loss(logits, D(rng.integers(0, 9, (14,))), fused=True)
(loss.sum() + (emb * emb).sum()).backward()
layer = fb.MultiHeadAttention(64, 2, causal=True, p_dropout=0.0, bias=True, backend=DEV, use_flash_attention=True)
x = D(rng.standard_normal((2, 33, 64)), True)
layer(x).sum().backward()
print("device-resident ops ok", float(np.abs(x.grad.to_numpy()).max()))
# tcgen05 GEMM (coalesced and ragged epilogues, both output dtypes) and split-KV decode attention
import ctypes
cl = fb._lib.load("combine")
for (M, N, K, obf, pad) in ((200, 304, 136, 0, 0), (130, 296, 72, 0, 1), (256, 512, 64, 1, 0), (100, 88, 40, 1, 3)):
    a_ = dev.DeviceArray.from_numpy(rng.standard_normal((M, K)).astype(np.float32), "bf16")
    b_ = dev.DeviceArray.from_numpy(rng.standard_normal((K, N)).astype(np.float32), "bf16")
    o_ = dev.DeviceArray((M, N + pad), "bf16" if obf else "f32")     # pad != 0: row stride not a multiple of 4 -> ragged stores
    fb._lib.check(cl, cl.fa_gemm_bf16_dev(o_.ptr, obf, N + pad, a_.ptr, 0, K, b_.ptr, 1, N, M, N, K, None))
    dev.sync()
    print("gemm", (M, N, K, obf, pad), "ok", float(np.abs(o_.to_numpy()[:, :N]).max()))
for dt, (B, H, L, d) in (("f32", (2, 3, 700, 64)), ("bf16", (1, 2, 3000, 128)), ("bf16", (3, 1, 50, 40))):
    q = dev.DeviceArray.from_numpy(rng.standard_normal((B, H, d)).astype(np.float32), dt)
    kc = dev.DeviceArray.from_numpy(rng.standard_normal((B, H, L, d)).astype(np.float32), dt)
    vc = dev.DeviceArray.from_numpy(rng.standard_normal((B, H, L, d)).astype(np.float32), dt)
    o = dev.DeviceArray((B, H, d), dt)
    dd = fb._lib.fa_decode_desc()
    dd.B, dd.H, dd.d, dd.L, dd.L_cap = B, H, d, L - 3, L
    dd.dtype = fb._lib.FA_DTYPE_BF16 if dt == "bf16" else fb._lib.FA_DTYPE_F32
    fb._lib.check(lib, lib.fa_flash_decode_dev(ctypes.byref(dd), q.ptr, kc.ptr, vc.ptr, o.ptr, None, None))
    dev.sync()
    print("decode", dt, (B, H, L, d), "ok", float(np.abs(o.to_numpy()).max()))
