"""GPU parity of the fused softmax / layernorm kernels (softmax_kernel.so, layernorm_kernel.so)
against the oracle, the golden vectors made from the reference's composed ops, and -- when
oracle/_ref holds them -- the reference's own CUDA kernels compiled from /root/reference."""
import ctypes
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from oracle import attention_ref as R
from tests.gpu_util import golden, maxabs

pytestmark = pytest.mark.gpu
ops = fb.CudaKernelOps
T = fb.tensor_from_numpy


# kernel_tests/test_softmax_fw.py:14 atol=rtol=1e-3 ; we hold 1e-6 (fp32 round-off)
@pytest.mark.parametrize("path", golden("softmax_*.npz"), ids=os.path.basename)
def test_softmax_fw_bw_golden(path):
    z = np.load(path)
    inp = T(z["inp"])
    out = ops.attn_softmax_fw(inp, T(z["mask"]))
    assert out is inp  # in place, like the reference (cuda_kernel_ops.py:468)
    assert maxabs(out.to_numpy(), z["y"]) < 2e-6
    g = T(z["dy"])
    g2, _ = ops.attn_softmax_bw(g, T(z["y"]))
    assert maxabs(g2.to_numpy(), z["dx"]) < 2e-6


@pytest.mark.parametrize("B,H,F,Tn,future", [
    (2, 3, 5, 1, False), (2, 8, 17, 31, False), (1, 8, 64, 64, True), (3, 2, 40, 100, False),
    (2, 2, 33, 257, False), (1, 4, 128, 512, True), (2, 2, 9, 1000, False), (1, 2, 64, 1024, True),
    (1, 2, 16, 2048, False), (1, 1, 8, 4096, True), (1, 1, 4, 5000, False), (1, 1, 3, 16384, False),
    (1, 1, 3, 12345, False),
])
def test_softmax_fw_bw_shapes(B, H, F, Tn, future):
    rng = np.random.default_rng(B * 1000 + Tn)
    x = rng.uniform(-1, 1, (B, H, F, Tn)).astype(np.float32)
    valid = rng.integers(1, Tn + 1, B)
    mask = np.where(np.arange(Tn)[None, :] < valid[:, None], 0.0, -1e8).astype(np.float32)
    want = R.attn_softmax_fw(x, mask, mask_future=future)
    got = ops.attn_softmax_fw(T(x), T(mask), mask_future=future).to_numpy()
    assert maxabs(got, want) < 3e-6
    np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-4)
    dy = rng.uniform(-1, 1, x.shape).astype(np.float32)
    gg, _ = ops.attn_softmax_bw(T(dy), T(want.astype(np.float32)))
    assert maxabs(gg.to_numpy(), R.attn_softmax_bw(dy, want.astype(np.float32))) < 3e-6


def test_softmax_no_mask_and_device_api():
    lib = fb._lib.load("softmax_kernel")
    rng = np.random.default_rng(5)
    x = rng.uniform(-3, 3, (2, 2, 7, 96)).astype(np.float32)
    got = ops.attn_softmax_fw(T(x), None).to_numpy()
    assert maxabs(got, R.attn_softmax_fw(x)) < 3e-6
    # device-pointer entry point
    n = x.size
    p = lib.fa_malloc(n * 4)
    xs = np.ascontiguousarray(x)
    assert lib.fa_h2d(p, xs.ctypes.data_as(ctypes.c_void_p), n * 4) == 0
    assert lib.fa_attn_softmax_dev(p, None, 2, 2, 7, 96, 0, None) == 0
    out = np.empty_like(xs)
    assert lib.fa_d2h(out.ctypes.data_as(ctypes.c_void_p), p, n * 4) == 0
    lib.fa_free(p)
    assert maxabs(out, R.attn_softmax_fw(x)) < 3e-6
    # error path: too long a row is reported, not a crash
    assert lib.fa_attn_softmax_dev(p, None, 1, 1, 1, 20000, 0, None) == fb._lib.FA_ERR_UNSUPPORTED
    assert b"16384" in lib.fa_last_error()


# kernel_tests/test_layernorm_fw.py:22 atol 1e-2 rtol 1e-3; test_layernorm_bw.py:22 atol 1e-3 rtol 1e-2
@pytest.mark.parametrize("path", golden("layernorm_*.npz"), ids=os.path.basename)
def test_layernorm_golden(path):
    z = np.load(path)
    y, var, mean = ops.layernorm_fw(T(z["x"]), T(z["gamma"]), T(z["beta"]))
    assert maxabs(y.to_numpy(), z["y"]) < 2e-5
    assert maxabs(mean.to_numpy(), z["mean"]) < 1e-6
    assert maxabs(var.to_numpy(), z["var"] + 1e-8) < 1e-6
    dx, dg, db = ops.layernorm_bw(T(z["dy"]), T(z["x"]), T(z["gamma"]), T(z["beta"]), var, mean)
    assert dg.shape == (1, z["gamma"].shape[0]) and db.shape == (1, z["gamma"].shape[0])
    assert maxabs(dx.to_numpy(), z["dx"]) < 1e-4
    assert maxabs(dg.to_numpy().reshape(-1), z["dgamma"]) < 2e-4
    assert maxabs(db.to_numpy().reshape(-1), z["dbeta"]) < 2e-4


@pytest.mark.parametrize("rows,h", [(1, 4), (1024, 32), (7, 100), (300, 128), (65, 256), (33, 512), (129, 768),
                                    (17, 1024), (40, 2048), (9, 4096), (5, 8192), (3, 16384), (2000, 64)])
def test_layernorm_shapes(rows, h):
    rng = np.random.default_rng(rows * 7 + h)
    x = rng.uniform(-1, 1, (rows, h)).astype(np.float32) + 0.3
    g = rng.uniform(-1, 1, h).astype(np.float32)
    b = rng.uniform(-1, 1, h).astype(np.float32)
    dy = rng.uniform(-1, 1, (rows, h)).astype(np.float32)
    y, var, mean = ops.layernorm_fw(T(x), T(g), T(b))
    ye, ve, me = R.layernorm_fw(x, g, b)
    assert maxabs(y.to_numpy(), ye) < 3e-5
    assert maxabs(var.to_numpy(), ve) < 1e-5 and maxabs(mean.to_numpy(), me) < 1e-6
    dx, dg, db = ops.layernorm_bw(T(dy), T(x), T(g), T(b), var, mean)
    dxe, dge, dbe = R.layernorm_bw(dy, x, g, b, ve, me)
    assert maxabs(dx.to_numpy(), dxe) < 1e-4
    scale = max(1.0, np.sqrt(rows))
    assert maxabs(dg.to_numpy(), dge) < 2e-5 * scale * 4
    assert maxabs(db.to_numpy(), dbe) < 2e-5 * scale * 4


def test_layernorm_rejects_bad_hidden():
    x = np.zeros((4, 6), dtype=np.float32)
    with pytest.raises(fb.FlashAttnError, match="hidden_dim % 4"):
        ops.layernorm_fw(T(x), T(np.ones(6, np.float32)), T(np.zeros(6, np.float32)))


def test_autograd_nodes_softmax_layernorm():
    rng = np.random.default_rng(9)
    x = rng.uniform(-1, 1, (2, 2, 6, 24)).astype(np.float32)
    mask = np.zeros((2, 24), dtype=np.float32)
    mask[1, 20:] = -1e8
    xt = T(x, requires_grad=True)
    y = xt.attn_softmax(T(mask))
    dy = rng.uniform(-1, 1, x.shape).astype(np.float32)
    y.backward(T(dy))
    ye = R.attn_softmax_fw(x, mask)
    assert maxabs(y.to_numpy(), ye) < 3e-6
    assert maxabs(xt.grad.to_numpy(), R.attn_softmax_bw(dy, ye)) < 3e-6
    a = rng.uniform(-1, 1, (10, 64)).astype(np.float32)
    g, b = rng.uniform(-1, 1, 64).astype(np.float32), rng.uniform(-1, 1, 64).astype(np.float32)
    at, gt, bt = T(a, requires_grad=True), T(g, requires_grad=True), T(b, requires_grad=True)
    out = at.layernorm(gt, bt)
    da = rng.uniform(-1, 1, a.shape).astype(np.float32)
    out.backward(T(da))
    ye, ve, me = R.layernorm_fw(a, g, b)
    dxe, dge, dbe = R.layernorm_bw(da, a, g, b, ve, me)
    assert maxabs(at.grad.to_numpy(), dxe) < 1e-4
    assert maxabs(gt.grad.to_numpy(), dge.reshape(-1)) < 1e-4
    assert maxabs(bt.grad.to_numpy(), dbe.reshape(-1)) < 1e-4


# ------------------------------------------------------------------------------------------------------------------
# The reference's kernel_tests, run live: random shapes drawn like its TestDecorator (kernel_tests/test_utils.py:
# batch * seq <= 1024, seq <= 512, nhead 8, hidden 32; 5 draws), the fused op against the COMPOSED minitorch-style ops
# on the same backend, with the reference's own tolerances.
# ------------------------------------------------------------------------------------------------------------------
def _bs_sl(rng):
    seq = int(rng.integers(1, 513))
    return int(rng.integers(1, max(2, 1024 // seq + 1))), seq


@pytest.mark.parametrize("draw", range(5))
@pytest.mark.parametrize("storage", ["host", "device"])
def test_kernel_tests_softmax_recipe_live(draw, storage):
    """kernel_tests/test_softmax_fw.py:60-72 (softmax(inp + mask) baseline, 1e-3) and test_softmax_bw.py:48-52
    (soft * (grad - sum(grad * soft)), atol 1e-2 rtol 1e-3)."""
    backend = fb.TensorBackend(fb.DeviceKernelOps if storage == "device" else fb.CudaKernelOps)
    mk = lambda a, g=False: fb.tensor_from_numpy(np.asarray(a, np.float32), backend=backend, requires_grad=g)
    rng = np.random.default_rng(100 + draw)
    B, to_len = _bs_sl(rng)
    from_len, nhead = to_len, 8
    if B * nhead * from_len * to_len > 2 ** 24:
        from_len = max(1, 2 ** 24 // (B * nhead * to_len))
    inp = rng.uniform(-1, 1, (B, nhead, from_len, to_len)).astype(np.float32)
    valid = rng.integers(1, to_len + 1, B)
    mask = np.where(np.arange(to_len)[None, :] < valid[:, None], 0.0, -1e8).astype(np.float32)
    cust = mk(inp.copy()).attn_softmax(mk(mask.reshape(B, 1, 1, to_len)))
    base = fb.softmax(mk(inp) + mk(mask.reshape(B, 1, 1, to_len)), dim=3)
    np.testing.assert_allclose(cust.to_numpy(), base.to_numpy(), atol=1e-3, rtol=1e-3)
    grad = rng.uniform(-1, 1, inp.shape).astype(np.float32)
    soft = base.to_numpy()
    x = mk(inp.copy(), True)
    x.attn_softmax(mk(mask.reshape(B, 1, 1, to_len))).backward(mk(grad))
    want = soft * (grad - (grad * soft).sum(axis=3, keepdims=True))
    np.testing.assert_allclose(x.grad.to_numpy(), want, atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("draw", range(5))
@pytest.mark.parametrize("storage", ["host", "device"])
def test_kernel_tests_layernorm_recipe_live(draw, storage):
    """kernel_tests/test_layernorm_fw.py:50-69 (atol 1e-2 rtol 1e-3) and test_layernorm_bw.py:136-161
    (atol 1e-3 rtol 1e-2): composed mean / var / normalise / affine on the same backend as the baseline."""
    backend = fb.TensorBackend(fb.DeviceKernelOps if storage == "device" else fb.CudaKernelOps)
    mk = lambda a, g=False: fb.tensor_from_numpy(np.asarray(a, np.float32), backend=backend, requires_grad=g)
    rng = np.random.default_rng(200 + draw)
    b, s = _bs_sl(rng)
    rows, h = b * s, 32
    inp = rng.uniform(-1, 1, (rows, h)).astype(np.float32)
    gamma, beta = rng.uniform(-1, 1, h).astype(np.float32), rng.uniform(-1, 1, h).astype(np.float32)
    dy = rng.uniform(-1, 1, (rows, h)).astype(np.float32)
    outs = []
    for fused in (True, False):
        x, g_, b_ = mk(inp, True), mk(gamma, True), mk(beta, True)
        if fused:
            y = x.layernorm(g_, b_)
        else:
            mean = x.mean(dim=1).view(rows, 1)
            var = x.var(dim=1).view(rows, 1)
            y = g_ * ((x - mean) / ((var + 1e-8) ** 0.5)) + b_
        y.backward(mk(dy))
        outs.append((y.to_numpy(), x.grad.to_numpy(), g_.grad.to_numpy().reshape(h), b_.grad.to_numpy().reshape(h)))
    np.testing.assert_allclose(outs[0][0], outs[1][0], atol=1e-2, rtol=1e-3)
    for got, want in zip(outs[0][1:], outs[1][1:]):
        np.testing.assert_allclose(got, want, atol=1e-3 * max(1.0, float(np.abs(want).max())), rtol=1e-2)
