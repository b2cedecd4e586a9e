"""The legacy host-pointer entry points (reference ABI, src/flashattention_kernel.cu:259,352,694,761) stream
the (batch, head) units through an H2D / kernels / D2H pipeline in chunks.  Chunking must not change a
single bit of the results (units are independent problems; the kernels are deterministic per unit in fp32
mode, and in bf16 mode for O, dK, dV), and padding masks / kv_len must follow the chunk's batch index."""
import numpy as np
import pytest

import flashattn_b200 as fb
from oracle import attention_ref as R
from tests.gpu_util import maxabs

pytestmark = pytest.mark.gpu
ops = fb.CudaKernelOps
T = fb.tensor_from_numpy


def _run(Q, K, V, dO, causal, key_mask):
    fw = ops.flash_attention_causal_fw if causal else ops.flash_attention_fw
    bw = ops.flash_attention_causal_bw if causal else ops.flash_attention_bw
    km = T(key_mask) if key_mask is not None else None
    q, k, v = T(Q), T(K), T(V)
    O, m, l = fw(q, k, v, key_mask=km)
    dQ, dK, dV = bw(q, k, v, O, T(dO), m, l, key_mask=km)
    return [t.to_numpy().copy() for t in (O, m, l, dQ, dK, dV)]


@pytest.fixture
def lib():
    lib = fb._lib.load("flashattention_kernel")
    yield lib
    lib.fa_set_legacy_chunk_bytes(0)
    ops.set_flash_mode("fp32")


@pytest.mark.parametrize("mode,d", [("fp32", 32), ("bf16", 128), ("bf16", 64)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("masked", ["none", "padding", "generic"])
def test_chunked_equals_single_chunk_and_oracle(lib, mode, d, causal, masked):
    B, H, N = 3, 5, 200
    rng = np.random.default_rng(31)
    Q, K, V, dO = (R.round_bf16(rng.standard_normal((B, H, N, d)).astype(np.float32)) for _ in range(4))
    key_mask = None
    if masked == "padding":      # recognised as kv_len[] by the library (0 ... 0, -1e8 ...)
        kv = np.array([200, 77, 130])
        key_mask = np.where(np.arange(N)[None, :] < kv[:, None], 0.0, -1e8).astype(np.float32)
    elif masked == "generic":    # arbitrary additive mask, different per batch
        key_mask = np.where(rng.random((B, N)) < 0.3, -1e8, 0.0).astype(np.float32)
        key_mask[:, 0] = 0.0
    ops.set_flash_mode(mode)
    lib.fa_set_legacy_chunk_bytes(1 << 40)           # the whole problem in one chunk
    whole = _run(Q, K, V, dO, causal, key_mask)
    lib.fa_set_legacy_chunk_bytes(2 * H * N * d * 4)  # two whole batches per chunk (2 + 1): masks index by batch
    pairs = _run(Q, K, V, dO, causal, key_mask)
    lib.fa_set_legacy_chunk_bytes(2 * N * d * 4)     # two heads per chunk -> 3 chunks per batch, ragged last
    parts = _run(Q, K, V, dO, causal, key_mask)
    lib.fa_set_legacy_chunk_bytes(1)                 # one head per chunk
    single = _run(Q, K, V, dO, causal, key_mask)
    for name, a, b in zip(("O", "m", "l", "dQ", "dK", "dV"), whole, pairs):
        if mode == "bf16" and name == "dQ":
            assert maxabs(a, b) <= 2.0 ** -7 * max(1.0, float(np.abs(a).max())), name
        else:
            np.testing.assert_array_equal(a, b, err_msg=name)
    names = ("O", "m", "l", "dQ", "dK", "dV")
    for name, a, b, c in zip(names, whole, parts, single):
        if mode == "bf16" and name == "dQ":          # dQ is an fp32 add-reduction across CTAs: rounding-level only
            assert maxabs(a, b) <= 2.0 ** -7 * max(1.0, float(np.abs(a).max())), name
            assert maxabs(a, c) <= 2.0 ** -7 * max(1.0, float(np.abs(a).max())), name
        else:
            np.testing.assert_array_equal(a, b, err_msg=name)
            np.testing.assert_array_equal(a, c, err_msg=name)
    tol = 1e-5 if mode == "fp32" else 2e-2
    Oe, _, _ = R.attention_fwd(Q, K, V, causal=causal, key_mask=key_mask)
    ge = R.attention_bwd(Q, K, V, dO, causal=causal, key_mask=key_mask)
    assert maxabs(parts[0], Oe) < tol
    for got, want in zip(parts[3:], ge):
        assert maxabs(got, want) < tol * max(1.0, float(np.abs(want).max()))


@pytest.mark.timeout(120)
@pytest.mark.parametrize("B,H", [(2, 300), (300, 2), (17, 40)])
def test_many_chunks_cap(lib, B, H):
    """More (batch, head) units than pipeline slots (256): heads / batches are regrouped, nothing is dropped."""
    N, d = 16, 8
    rng = np.random.default_rng(5)
    Q, K, V, dO = (rng.standard_normal((B, H, N, d)).astype(np.float32) for _ in range(4))
    ops.set_flash_mode("fp32")
    lib.fa_set_legacy_chunk_bytes(1)
    got = _run(Q, K, V, dO, True, None)
    Oe, _, _ = R.attention_fwd(Q, K, V, causal=True)
    ge = R.attention_bwd(Q, K, V, dO, causal=True)
    assert maxabs(got[0], Oe) < 1e-5
    for g, w in zip(got[3:], ge):
        assert maxabs(g, w) < 1e-5 * max(1.0, float(np.abs(w).max()))


def test_pinned_and_pageable_caller_buffers_give_identical_results(lib):
    """Pageable (numpy) buffers are staged through the library's pinned ring by host threads; page-locked buffers go
    straight to the copy engines.  Same kernels, same bits -- forward and backward, several chunks."""
    import ctypes
    B, H, N, d = 2, 6, 384, 64
    n, r = B * H * N * d, B * H * N
    rng = np.random.default_rng(9)
    data = {k: rng.standard_normal(n).astype(np.float32) for k in ("Q", "K", "V", "dO")}
    ops.set_flash_mode("bf16")
    lib.fa_set_legacy_chunk_bytes(2 * N * d * 4)          # two heads per chunk -> 6 chunks, ring slots get reused

    def run(alloc):
        a = {k: alloc(n) for k in ("Q", "K", "V", "dO", "O", "dQ", "dK", "dV")}
        s = {k: alloc(r) for k in ("l", "m")}
        for k, v in data.items():
            a[k][1][:] = v
        lib.launch_flashattention_forward_causal(a["Q"][1], a["K"][1], a["V"][1], a["O"][1], s["l"][1], s["m"][1], B, H, N, d)
        fb._lib.check(lib)
        lib.launch_flashattention_backward_causal(a["Q"][1], a["K"][1], a["V"][1], a["O"][1], a["dQ"][1], a["dK"][1],
                                                  a["dV"][1], a["dO"][1], s["l"][1], s["m"][1], B, H, N, d)
        fb._lib.check(lib)
        out = {k: np.array(a[k][1]) for k in ("O", "dK", "dV", "dQ")}
        out.update({k: np.array(s[k][1]) for k in ("l", "m")})
        for p, _ in list(a.values()) + list(s.values()):
            if p:
                lib.fa_free_host(p)
        return out

    def pinned(count):
        p = lib.fa_malloc_host(count * 4)
        assert p
        return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(count,))

    def wire():
        a, b = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        lib.fa_wire_bytes(ctypes.byref(a), ctypes.byref(b))
        return a.value + b.value

    pageable = run(lambda count: (None, np.zeros(count, dtype=np.float32)))
    w0 = wire()
    locked = run(pinned)
    w1 = wire()
    # page-locked tensors SPLIT between the two routes chunk by chunk (the policy that large tensors get by default,
    # forced onto these small ones): the conversions on the host and on the device round alike, so the bits are the same
    try:
        mixed = {}
        for cost in (0.3, 0.8, 2.0):
            lib.fa_set_transfer_policy(cost, 0)
            w2 = wire()
            mixed[cost] = run(pinned)
            assert wire() - w2 < w1 - w0, "no tensor took the bf16 route"
    finally:
        lib.fa_set_transfer_policy(0.0, -1)
    for other in [locked] + list(mixed.values()):
        for k in ("O", "l", "m", "dK", "dV"):
            np.testing.assert_array_equal(pageable[k], other[k], err_msg=k)
        assert maxabs(pageable["dQ"], other["dQ"]) <= 2.0 ** -7 * max(1.0, float(np.abs(locked["dQ"]).max()))
    Q, K, V, dO = (R.round_bf16(data[k]).reshape(B, H, N, d) for k in ("Q", "K", "V", "dO"))
    Oe, _, _ = R.attention_fwd(Q, K, V, causal=True)
    assert maxabs(locked["O"].reshape(B, H, N, d), Oe) < 2e-2
