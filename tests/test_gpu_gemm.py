"""bf16 tensor-core GEMM (csrc/gemm_sm100.cuh) against an fp64 product of the same
bf16-rounded operands: every operand-order combination (the forward and the two backward GEMMs of a Linear layer),
ragged M / N / K (TMA zero-fill + predicated stores), fp32 and bf16 outputs, the fused Q/K/V projection, and the
Linear / MultiHeadAttention modules in bf16 GEMM mode against the fp32 CUDA-core path.

Tolerance: bf16 products are exact in fp32 and the accumulation is fp32, so the error of an element is bounded by
K * 2^-24 * sum|a||b| -- asserted as 4e-6 * (|A| @ |B|) (+ one bf16 ulp when the output is bf16)."""
import ctypes

import numpy as np
import pytest

import flashattn_b200 as fb
from flashattn_b200 import _lib
from flashattn_b200 import device as dev
from oracle import attention_ref as R

pytestmark = pytest.mark.gpu


def _gemm(A, B, a_mn, b_mn, out_bf16=False):
    """A (M,K), B (K,N) logical; stored transposed when *_mn asks for the other memory order."""
    lib = _lib.load("combine")
    M, K = A.shape
    N = B.shape[1]
    a_store = np.ascontiguousarray(A.T) if a_mn else np.ascontiguousarray(A)       # [K][M] or [M][K]
    b_store = np.ascontiguousarray(B) if b_mn else np.ascontiguousarray(B.T)       # [K][N] or [N][K]
    da, db = dev.DeviceArray.from_numpy(a_store, "bf16"), dev.DeviceArray.from_numpy(b_store, "bf16")
    out = dev.DeviceArray((M, N), "bf16" if out_bf16 else "f32")
    rc = lib.fa_gemm_bf16_dev(out.ptr, int(out_bf16), N, da.ptr, int(a_mn), a_store.shape[1], db.ptr, int(b_mn),
                              b_store.shape[1], M, N, K, None)
    _lib.check(lib, rc)
    return out.to_numpy()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 512), (304, 200, 136), (5000, 256, 256), (64, 10000, 256),
                                   (1000, 72, 8)])
def test_gemm_vs_fp64(a_mn, b_mn, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A = R.round_bf16(rng.standard_normal((M, K)).astype(np.float32))
    B = R.round_bf16(rng.standard_normal((K, N)).astype(np.float32))
    got = _gemm(A, B, a_mn, b_mn)
    want = A.astype(np.float64) @ B.astype(np.float64)
    bound = 4e-6 * (np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)) + 1e-6
    assert np.all(np.abs(got - want) <= bound), float(np.abs(got - want).max())


def test_gemm_bf16_output_and_rejects_misaligned():
    rng = np.random.default_rng(9)
    A = R.round_bf16(rng.standard_normal((200, 96)).astype(np.float32))
    B = R.round_bf16(rng.standard_normal((96, 136)).astype(np.float32))
    got = _gemm(A, B, 0, 1, out_bf16=True)
    want = A.astype(np.float64) @ B.astype(np.float64)
    assert np.all(np.abs(got - want) <= 2.0 ** -8 * np.abs(want) + 1e-3)
    lib = _lib.load("combine")
    x = dev.DeviceArray((64 * 64,), "bf16")
    o = dev.DeviceArray((64, 64), "f32")
    assert lib.fa_gemm_bf16_dev(o.ptr, 0, 64, x.ptr, 0, 60, x.ptr, 1, 64, 64, 64, 60, None) == _lib.FA_ERR_UNSUPPORTED


@pytest.mark.parametrize("M,E", [(5000, 256), (333, 64), (4096, 1024)])
def test_fused_qkv_projection(M, E):
    lib = _lib.load("combine")
    rng = np.random.default_rng(E)
    X = R.round_bf16(rng.standard_normal((M, E)).astype(np.float32))
    W = R.round_bf16((rng.standard_normal((E, 3 * E)) / np.sqrt(E)).astype(np.float32))
    dx, dw = dev.DeviceArray.from_numpy(X, "bf16"), dev.DeviceArray.from_numpy(W, "bf16")
    outs = [dev.DeviceArray((M, E), "f32") for _ in range(3)]
    _lib.check(lib, lib.fa_qkv_proj_bf16_dev(outs[0].ptr, outs[1].ptr, outs[2].ptr, 0, dx.ptr, dw.ptr, M, E, None))
    want = X.astype(np.float64) @ W.astype(np.float64)
    bound = 4e-6 * (np.abs(X).astype(np.float64) @ np.abs(W).astype(np.float64)) + 1e-6
    for j in range(3):
        assert np.all(np.abs(outs[j].to_numpy() - want[:, j * E:(j + 1) * E]) <= bound[:, j * E:(j + 1) * E]), j


@pytest.mark.timeout(300)
def test_mha_module_in_bf16_gemm_mode_matches_fp32_path():
    """MultiHeadAttention fwd + bwd with the projections on the tensor-core GEMMs (fused Q/K/V forward, transposed-view
    backward GEMMs) against the same module on the fp32 CUDA-core matmul: bf16 operand rounding only (2e-2)."""
    ops = fb.DeviceKernelOps
    backend = fb.TensorBackend(ops)
    np.random.seed(3)
    B, N, E, nh = 4, 128, 256, 4
    layer = fb.MultiHeadAttention(E, nh, causal=True, p_dropout=0.0, bias=False, backend=backend, use_flash_attention=True)
    X = np.random.rand(B, N, E).astype(np.float32)
    res = {}
    try:
        for mode in ("fp32", "bf16"):
            ops.set_gemm_mode(mode)
            for p in layer.parameters():
                p.value.grad = None
            x = fb.tensor_from_numpy(X, backend=backend, requires_grad=True)
            y = layer(x)
            y.sum().backward()
            res[mode] = [y.to_numpy(), x.grad.to_numpy()] + [p.value.grad.to_numpy() for p in layer.parameters()]
    finally:
        ops.set_gemm_mode("fp32")
    for a, b in zip(res["fp32"], res["bf16"]):
        scale = max(1.0, float(np.abs(a).max()))
        assert np.abs(a - b).max() <= 2e-2 * scale, (a.shape, float(np.abs(a - b).max()), scale)
