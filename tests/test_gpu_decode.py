"""Decode shapes: the split-KV decode-attention kernel against the fp64 composed formula, and the
cache-aware generate() against the reference's full-prefix loop (project/run_machine_translation.py:299-325, mirrored
by generate()): same tokens, same logits.

Tolerances: fp32 caches 1e-5 max-abs (the fp32 bar of north_star), bf16 caches 2e-2."""
import ctypes

import numpy as np
import pytest

import flashattn_b200 as fb
from flashattn_b200 import _lib
from flashattn_b200 import device as dev
from oracle import attention_ref as R

pytestmark = pytest.mark.gpu


def _oracle(q, K, V, lens):
    B, H, d = q.shape
    out = np.zeros((B, H, d))
    lse = np.full((B, H), -np.inf)
    for b in range(B):
        n = int(lens[b])
        if n == 0:
            continue
        s = np.einsum("hd,hnd->hn", q[b].astype(np.float64), K[b, :, :n].astype(np.float64)) / np.sqrt(d)
        mx = s.max(axis=1, keepdims=True)
        p = np.exp(s - mx)
        lse[b] = (mx + np.log(p.sum(axis=1, keepdims=True)))[:, 0]
        out[b] = np.einsum("hn,hnd->hd", p / p.sum(axis=1, keepdims=True), V[b, :, :n].astype(np.float64))
    return out, lse


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("B,H,d,cap,lens", [
    (1, 8, 32, 41, [41]),              # config #2's model at its last position
    (3, 2, 32, 40, [1, 5, 40]),        # ragged kv_len, one key only
    (2, 4, 64, 700, [700, 333]),
    (1, 2, 128, 9000, [8192]),         # batch 1, long cache: split-KV + combine
    (2, 3, 96, 130, [130, 0]),         # head_dim that is not a power of two; an empty cache row
    (1, 1, 256, 300, [300]),           # two vectors per lane in fp32
    (4, 16, 8, 64, [64, 3, 17, 50]),
])
def test_decode_kernel_vs_oracle(dtype, B, H, d, cap, lens):
    lib = _lib.load("flashattention_kernel")
    rng = np.random.default_rng(B * 1000 + d)
    rnd = (lambda s: R.round_bf16(rng.standard_normal(s).astype(np.float32))) if dtype == "bf16" else \
        (lambda s: rng.standard_normal(s).astype(np.float32))
    q, K, V = rnd((B, H, d)), rnd((B, H, cap, d)), rnd((B, H, cap, d))
    dq, dK, dV = (dev.DeviceArray.from_numpy(x, dtype) for x in (q, K, V))
    out = dev.DeviceArray((B, H, d), dtype)
    lse = dev.DeviceArray((B, H), "f32")
    kv = dev.DeviceArray.from_numpy(np.asarray(lens, dtype=np.int32))
    a = _lib.fa_decode_desc()
    a.B, a.H, a.d, a.L, a.L_cap = B, H, d, max(lens), cap
    a.dtype = _lib.FA_DTYPE_BF16 if dtype == "bf16" else _lib.FA_DTYPE_F32
    a.kv_len = kv.ptr
    _lib.check(lib, lib.fa_flash_decode_dev(ctypes.byref(a), dq.ptr, dK.ptr, dV.ptr, out.ptr, lse.ptr, None))
    want, lse_want = _oracle(q, K, V, lens)
    tol = 2e-2 if dtype == "bf16" else 1e-5
    assert np.abs(out.to_numpy() - want).max() < tol
    got_lse = lse.to_numpy()
    fin = np.isfinite(lse_want)
    assert np.abs(got_lse[fin] - lse_want[fin]).max() < 1e-4
    assert np.all(np.isneginf(got_lse[~fin]))
    if len(set(lens)) == 1:      # common length passed as a scalar instead of kv_len[]
        a.kv_len = None
        a.L = lens[0]
        out2 = dev.DeviceArray((B, H, d), dtype)
        _lib.check(lib, lib.fa_flash_decode_dev(ctypes.byref(a), dq.ptr, dK.ptr, dV.ptr, out2.ptr, None, None))
        if cap == lens[0]:       # same split of the cache in both calls: bit-identical
            np.testing.assert_array_equal(out2.to_numpy(), out.to_numpy())
        else:                    # the split follows the capacity when only kv_len[] is known: same result up to rounding
            assert np.abs(out2.to_numpy() - out.to_numpy()).max() < tol / 4


def test_decode_rejects_unsupported_head_dim():
    lib = _lib.load("flashattention_kernel")
    a = _lib.fa_decode_desc()
    a.B, a.H, a.d, a.L, a.L_cap, a.dtype = 1, 1, 6, 4, 4, _lib.FA_DTYPE_F32
    x = dev.DeviceArray((64,), "f32")
    assert lib.fa_flash_decode_dev(ctypes.byref(a), x.ptr, x.ptr, x.ptr, x.ptr, None, None) == _lib.FA_ERR_UNSUPPORTED


@pytest.mark.timeout(300)
@pytest.mark.parametrize("flash", [True, False])
def test_generate_cached_matches_full_prefix_generate(flash):
    """Same greedy tokens as the reference-style loop, and the cached logits of every step equal the last-position
    logits of a full forward over the prefix (fp32, 2e-4: different summation orders over <= 25 positions)."""
    backend = fb.TensorBackend(fb.DeviceKernelOps)
    np.random.seed(7)
    model = fb.DecoderLM(n_vocab=97, n_embd=64, n_head=2, n_positions=24, p_dropout=0.0, backend=backend,
                         use_flash_attention=flash)
    model.eval()
    prompt = [5, 17, 3, 88, 41]
    full = fb.generate(model, prompt, 24)
    cached = fb.generate_cached(model, prompt, 24)
    assert cached == full and len(full) == 25
    # step-by-step logits
    attn = model.t_layer_1.attention
    caches = [fb.DeviceKernelOps.kv_cache_new(1, attn.n_head, 32, attn.attn_hidden_dim) for _ in range(4)]
    ids = full[:12]
    lg = fb.decode_step(model, np.asarray(ids[:7], dtype=np.float32).reshape(1, 7), caches).to_numpy()
    ref = model(fb.tensor_from_numpy(np.asarray(ids[:7], dtype=np.float32).reshape(1, 7), backend=backend)).to_numpy()
    assert np.abs(lg - ref).max() < 2e-4
    for t in range(7, 12):
        lg = fb.decode_step(model, np.asarray([[ids[t]]], dtype=np.float32), caches).to_numpy()
        ref = model(fb.tensor_from_numpy(np.asarray(ids[:t + 1], dtype=np.float32).reshape(1, t + 1),
                                         backend=backend)).to_numpy()
        assert np.abs(lg[0, 0] - ref[0, t]).max() < 2e-4, t
    assert caches[0].len == 12
