"""combine.so: map / zip / reduce / batched matmul through the reference's
host-pointer C ABI (src/combine.cu:315,385,443,523) and the CudaKernelOps.map/zip/reduce/matrix_multiply
surface (minitorch/cuda_kernel_ops.py:58-437), against (a) golden vectors produced by the reference's own
FastOps implementation of the same four operations and (b) the numpy oracle at larger / odd shapes."""
import numpy as np
import pytest

import flashattn_b200 as fb
from oracle import combine_ref as C
from tests.gpu_util import golden, have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]

ops, T = fb.CudaKernelOps, fb.tensor_from_numpy


def close(a, b, rtol=3e-6, atol=2e-6):
    """fp32 elementwise tolerance: exp / tanh / pow use CUDA libm (<= 2 ulp)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.all(np.abs(a - b) <= atol + rtol * np.abs(b)), float(np.max(np.abs(a - b)))


@pytest.fixture(scope="module")
def g():
    return np.load(golden("combine_ops.npz")[0])


def test_map_golden(g):
    for name in C.UNARY:
        src = {"log": "pos", "inv": "nz"}.get(name, "a")
        close(ops.map(name)(T(g[src])).to_numpy(), g[f"map_{name}"])
    # strided (permuted) input view, contiguous output
    close(ops.map("neg")(T(g["a"]).permute(2, 0, 1)).to_numpy(), g["map_neg_perm"])


def test_zip_golden(g):
    for name in C.BINARY:
        a = {"log_back": "pos", "pow": "pos", "inv_back": "nz"}.get(name, "a")
        close(ops.zip(name)(T(g[a]), T(g["b"])).to_numpy(), g[f"zip_{name}"])
        if f"zipb_{name}" in g.files:    # right-aligned broadcast of a (1,5) operand against (3,4,5)
            close(ops.zip(name)(T(g[a]), T(g["brow"])).to_numpy(), g[f"zipb_{name}"])
    close(ops.zip("add")(T(g["a"]).permute(1, 0, 2), T(g["b"]).permute(1, 0, 2)).to_numpy(), g["zip_add_perm"])


def test_reduce_golden(g):
    for dim in (0, 1, 2):
        close(ops.reduce("add", 0.0)(T(g["a"]), dim).to_numpy(), g[f"red_add_{dim}"], atol=5e-6)
        close(ops.reduce("mul", 1.0)(T(g["a"]), dim).to_numpy(), g[f"red_mul_{dim}"], rtol=1e-5)
        close(ops.reduce("max", -1e9)(T(g["a"]), dim).to_numpy(), g[f"red_max_{dim}"])
    close(ops.reduce("add", 0.0)(T(g["big"]), 1).to_numpy(), g["red_add_big"], atol=1e-4)


def test_matmul_golden(g):
    A, Bm, W, Kt = (T(g[k]) for k in ("mm_A", "mm_B", "mm_W", "mm_Kt"))
    close(ops.matrix_multiply(A, Bm).to_numpy(), g["mm_batched"], atol=2e-5, rtol=1e-5)
    close(ops.matrix_multiply(A, W).to_numpy(), g["mm_bcast"], atol=2e-5, rtol=1e-5)
    close(ops.matrix_multiply(A, Kt.permute(0, 2, 1)).to_numpy(), g["mm_transposed"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("shape", [(1,), (7,), (1024,), (3, 1, 5), (2, 3, 4, 5, 6), (257, 129)])
def test_map_zip_shapes_vs_oracle(shape):
    rng = np.random.default_rng(7)
    a = rng.uniform(-2, 2, shape).astype(np.float32)
    b = rng.uniform(-2, 2, shape).astype(np.float32)
    for name in ("neg", "sigmoid", "relu", "exp", "tanh"):
        close(ops.map(name)(T(a)).to_numpy(), C.tensor_map(C.FN_IDS[name], a))
    for name in ("add", "mul", "lt", "max", "relu_back"):
        close(ops.zip(name)(T(a), T(b)).to_numpy(), C.tensor_zip(C.FN_IDS[name], a, b))


def test_zip_broadcast_both_sides():
    rng = np.random.default_rng(8)
    a = rng.standard_normal((4, 1, 6)).astype(np.float32)
    b = rng.standard_normal((5, 1)).astype(np.float32)
    close(ops.zip("mul")(T(a), T(b)).to_numpy(), C.tensor_zip(2, a, b))
    s = rng.standard_normal((1,)).astype(np.float32)            # scalar-like operand
    close(ops.zip("add")(T(a), T(s)).to_numpy(), C.tensor_zip(1, a, s))


def test_map_into_broadcast_out():
    a = np.arange(5, dtype=np.float32).reshape(1, 5)
    out = T(np.zeros((3, 5), np.float32))
    ops.map("id")(T(a), out)
    close(out.to_numpy(), np.broadcast_to(a, (3, 5)))


@pytest.mark.parametrize("shape,dim", [((64, 1000), 1), ((1000, 64), 0), ((8, 16, 33, 65), 3), ((8, 16, 33, 65), 1)])
def test_reduce_vs_oracle(shape, dim):
    rng = np.random.default_rng(9)
    a = rng.uniform(-1, 1, shape).astype(np.float32)
    close(ops.reduce("add", 0.0)(T(a), dim).to_numpy(), C.tensor_reduce(1, a, dim, 0.0), atol=2e-4)
    close(ops.reduce("max", -1e9)(T(a), dim).to_numpy(), C.tensor_reduce(16, a, dim, -1e9))
    # reduce_value is honoured (the reference's C side read the double as a float and lost it)
    close(ops.reduce("add", 2.5)(T(a), dim).to_numpy(), C.tensor_reduce(1, a, dim, 2.5), atol=2e-4)


@pytest.mark.parametrize("B,M,K,N", [(1, 1, 1, 1), (3, 64, 64, 64), (2, 130, 257, 67), (5120, 39, 32, 39)])
def test_matmul_vs_oracle(B, M, K, N):
    rng = np.random.default_rng(10)
    a = rng.standard_normal((B, M, K)).astype(np.float32)
    b = rng.standard_normal((B, K, N)).astype(np.float32)
    close(ops.matrix_multiply(T(a), T(b)).to_numpy(), C.matrix_multiply(a, b), atol=1e-4, rtol=1e-5)


def test_matmul_4d_and_2d_like_mha():
    """The shapes MultiHeadAttention drives through matrix_multiply (modules_transfomer.py:87-107,181-192)."""
    rng = np.random.default_rng(11)
    q = rng.standard_normal((2, 4, 64, 32)).astype(np.float32)
    k = rng.standard_normal((2, 4, 64, 32)).astype(np.float32)
    s = ops.matrix_multiply(T(q), T(k).permute(0, 1, 3, 2))
    assert s.shape == (2, 4, 64, 64)
    close(s.to_numpy(), C.matrix_multiply(q, np.swapaxes(k, -1, -2)), atol=1e-4)
    x = rng.standard_normal((128, 256)).astype(np.float32)
    w = rng.standard_normal((256, 96)).astype(np.float32)
    y = ops.matrix_multiply(T(x), T(w))
    assert y.shape == (128, 96)
    close(y.to_numpy(), C.matrix_multiply(x, w), atol=2e-4)


def test_backend_binds_reference_attribute_names():
    be = fb.default_backend()
    for name in ("neg_map", "sigmoid_map", "relu_map", "log_map", "exp_map", "id_map", "id_cmap", "inv_map",
                 "tanh_map", "add_zip", "mul_zip", "lt_zip", "eq_zip", "is_close_zip", "relu_back_zip",
                 "log_back_zip", "inv_back_zip", "pow_scalar_zip", "add_reduce", "mul_reduce", "matrix_multiply"):
        assert callable(getattr(be, name)), name
    a = np.array([[1.0, -2.0], [3.0, 4.0]], np.float32)
    close(be.add_reduce(T(a), 1).to_numpy(), a.sum(1, keepdims=True))


def test_errors_are_reported_not_fatal():
    with pytest.raises(KeyError):
        ops.map("sqrt")
    with pytest.raises(fb.FlashAttnError):
        ops.map("neg")(T(np.zeros((1,) * 9, np.float32)))     # > 8 dims


@pytest.mark.parametrize("B,M,K,N", [(1, 128, 16, 128), (2, 200, 77, 150), (1, 129, 300, 385), (3, 256, 5, 128)])
def test_matmul_large_tile_kernel_vs_oracle(B, M, K, N):
    """M, N >= 128 take the 128x128 register-blocked kernel: ragged edges, K not a multiple of the K step,
    transposed (permuted) operands and a broadcast 2-D right operand -- host- and device-pointer entry points."""
    rng = np.random.default_rng(B * 1000 + M + K + N)
    a = rng.standard_normal((B, M, K)).astype(np.float32)
    b = rng.standard_normal((B, K, N)).astype(np.float32)
    w = rng.standard_normal((K, N)).astype(np.float32)
    at = np.ascontiguousarray(a.transpose(0, 2, 1))         # stored (B, K, M), used as its transpose
    tol = dict(atol=1e-5 * K + 2e-5, rtol=1e-5)
    want, want_w = C.matrix_multiply(a, b), C.matrix_multiply(a, w)
    DEV = fb.TensorBackend(fb.DeviceKernelOps)
    for o, mk in ((ops, T), (fb.DeviceKernelOps, lambda x: fb.tensor_from_numpy(x, backend=DEV))):
        close(o.matrix_multiply(mk(a), mk(b)).to_numpy(), want, **tol)
        close(o.matrix_multiply(mk(a), mk(w)).to_numpy(), want_w, **tol)
        close(o.matrix_multiply(mk(at).permute(0, 2, 1), mk(b)).to_numpy(), want, **tol)
