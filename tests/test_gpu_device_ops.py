"""DeviceKernelOps (device-resident plumbing): the CudaKernelOps surface with tensor storage in
HBM.  Same golden vectors and oracles as the host-pointer path: combine ops from the reference's FastOps, the
MultiHeadAttention golden (reference composed path + torch), the DecoderLM golden (reference's own run), and
bit-for-bit agreement with the host-pointer path where both run the same kernels."""
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from oracle import attention_ref as R
from oracle import combine_ref as C
from tests.gpu_util import golden, maxabs
from tests.test_gpu_combine import close
from tests.test_host_modules import decoder_loss, load_decoder

pytestmark = pytest.mark.gpu
DEV = fb.TensorBackend(fb.DeviceKernelOps)
ops = fb.DeviceKernelOps


def T(a, requires_grad=False):
    return fb.tensor_from_numpy(np.asarray(a, dtype=np.float32), backend=DEV, requires_grad=requires_grad)


@pytest.fixture(autouse=True)
def _fp32():
    ops.set_flash_mode("fp32")
    yield
    ops.set_flash_mode("fp32")


def test_storage_round_trip_and_views():
    a = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    t = T(a)
    assert isinstance(t._tensor._storage, fb.DeviceStorage)
    np.testing.assert_array_equal(t.to_numpy(), a)
    p = t.permute(2, 0, 1)
    np.testing.assert_array_equal(p.to_numpy(), a.transpose(2, 0, 1))
    c = p.contiguous()
    assert c._tensor.is_contiguous() and c._tensor._storage is not t._tensor._storage
    np.testing.assert_array_equal(c.view(4, 6).to_numpy(), a.transpose(2, 0, 1).reshape(4, 6))
    np.testing.assert_array_equal(t.zeros((5, 2)).to_numpy(), np.zeros((5, 2), np.float32))


def test_combine_golden_on_device_storage():
    g = np.load(golden("combine_ops.npz")[0])
    for name in C.UNARY:
        src = {"log": "pos", "inv": "nz"}.get(name, "a")
        close(ops.map(name)(T(g[src])).to_numpy(), g[f"map_{name}"])
    close(ops.map("neg")(T(g["a"]).permute(2, 0, 1)).to_numpy(), g["map_neg_perm"])
    for name in C.BINARY:
        a = {"log_back": "pos", "pow": "pos", "inv_back": "nz"}.get(name, "a")
        close(ops.zip(name)(T(g[a]), T(g["b"])).to_numpy(), g[f"zip_{name}"])
        if f"zipb_{name}" in g.files:
            close(ops.zip(name)(T(g[a]), T(g["brow"])).to_numpy(), g[f"zipb_{name}"])
    close(ops.zip("add")(T(g["a"]).permute(1, 0, 2), T(g["b"]).permute(1, 0, 2)).to_numpy(), g["zip_add_perm"])
    for dim in (0, 1, 2):
        close(ops.reduce("add", 0.0)(T(g["a"]), dim).to_numpy(), g[f"red_add_{dim}"], atol=5e-6)
        close(ops.reduce("mul", 1.0)(T(g["a"]), dim).to_numpy(), g[f"red_mul_{dim}"], rtol=1e-5)
        close(ops.reduce("max", -1e9)(T(g["a"]), dim).to_numpy(), g[f"red_max_{dim}"])
    close(ops.reduce("add", 0.0)(T(g["big"]), 1).to_numpy(), g["red_add_big"], atol=1e-4)
    A, Bm, W, Kt = (T(g[k]) for k in ("mm_A", "mm_B", "mm_W", "mm_Kt"))
    close(ops.matrix_multiply(A, Bm).to_numpy(), g["mm_batched"], atol=2e-5, rtol=1e-5)
    close(ops.matrix_multiply(A, W).to_numpy(), g["mm_bcast"], atol=2e-5, rtol=1e-5)
    close(ops.matrix_multiply(A, Kt.permute(0, 2, 1)).to_numpy(), g["mm_transposed"], atol=2e-5, rtol=1e-5)
    # 4-D @ 4-D (attention scores) and 2-D @ 2-D (Linear)
    rng = np.random.default_rng(2)
    q, k = rng.standard_normal((2, 3, 17, 8)).astype(np.float32), rng.standard_normal((2, 3, 17, 8)).astype(np.float32)
    close(ops.matrix_multiply(T(q), T(k).permute(0, 1, 3, 2)).to_numpy(), C.matrix_multiply(q, k.transpose(0, 1, 3, 2)),
          atol=2e-5, rtol=1e-5)
    x, w = rng.standard_normal((33, 70)).astype(np.float32), rng.standard_normal((70, 5)).astype(np.float32)
    close(ops.matrix_multiply(T(x), T(w)).to_numpy(), C.matrix_multiply(x, w), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("path", golden("attn_*.npz"), ids=os.path.basename)
def test_flash_golden_device_resident_equals_host_pointer_path(path):
    z = np.load(path)
    if "key_mask" in z.files:
        pytest.skip("the reference's flash_attention(q, k, v) surface has no mask argument")
    causal = bool(z["causal"])
    q, k, v = (T(z[n], requires_grad=True) for n in ("Q", "K", "V"))
    out = q.flash_attention_causal(k, v) if causal else q.flash_attention(k, v)
    assert maxabs(out.to_numpy(), z["O"]) < 1e-5
    out.backward(T(z["dO"]))
    for t, name in ((q, "dQ"), (k, "dK"), (v, "dV")):
        assert maxabs(t.grad.to_numpy(), z[name]) < 1e-5 * max(1.0, float(np.abs(z[name]).max())), name
    # same kernels behind the legacy host-pointer ABI: identical bits
    fb.CudaKernelOps.set_flash_mode("fp32")
    hq, hk, hv = (fb.tensor_from_numpy(z[n], requires_grad=True) for n in ("Q", "K", "V"))
    hout = hq.flash_attention_causal(hk, hv) if causal else hq.flash_attention(hk, hv)
    np.testing.assert_array_equal(out.to_numpy(), hout.to_numpy())


def test_flash_bf16_mode_on_device_tensors():
    rng = np.random.default_rng(12)
    Q, K, V, dO = (R.round_bf16(rng.standard_normal((2, 3, 300, 128)).astype(np.float32)) for _ in range(4))
    ops.set_flash_mode("bf16")
    q, k, v = (T(x, requires_grad=True) for x in (Q, K, V))
    out = q.flash_attention_causal(k, v)
    Oe, _, _ = R.attention_fwd(Q, K, V, causal=True)
    assert maxabs(out.to_numpy(), Oe) < 2e-2
    out.backward(T(dO))
    for t, want in zip((q, k, v), R.attention_bwd(Q, K, V, dO, causal=True)):
        assert maxabs(t.grad.to_numpy(), want) < 2e-2


@pytest.mark.parametrize("branch", ["flash", "fused", "composed"])
@pytest.mark.parametrize("path", golden("mha_cfg1_*.npz"), ids=os.path.basename)
def test_mha_module_device_resident(path, branch):
    z = np.load(path)
    layer = fb.MultiHeadAttention(z["X"].shape[-1], int(z["n_head"]), causal=bool(z["causal"]), p_dropout=0.0,
                                  bias=False, backend=DEV, use_flash_attention=branch == "flash",
                                  use_fused_kernel=branch == "fused")
    for lin, key in ((layer.q_projection, "Wq"), (layer.k_projection, "Wk"), (layer.v_projection, "Wv"),
                     (layer.out_projection, "Wo")):
        lin.weights.value = T(z[key], requires_grad=True)
    X = T(z["X"], requires_grad=True)
    Y = layer(X)
    for ref in ("Y_ref", "Y_torch"):
        np.testing.assert_allclose(Y.to_numpy(), z[ref], atol=1e-5, rtol=1e-5)
    Y.sum().backward()
    for ref in ("dX_ref", "dX_torch"):
        np.testing.assert_allclose(X.grad.to_numpy(), z[ref], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(layer.out_projection.weights.value.grad.to_numpy(), z["dWo_torch"], atol=2e-4, rtol=1e-5)
    np.testing.assert_allclose(layer.q_projection.weights.value.grad.to_numpy(), z["dWq_ref"], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("branch", ["composed", "flash", "fused"])
def test_decoder_lm_golden_device_resident(branch):
    z = np.load(golden("decoder_small.npz")[0])
    model, params = load_decoder(z, backend=DEV, use_flash_attention=branch == "flash", use_fused_kernel=branch == "fused")
    logits, total = decoder_loss(model, z, backend=DEV)
    tol = 1e-4 if branch == "fused" else 2e-5
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=tol * 10, rtol=tol)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < tol
    total.backward()
    for k in z.files:
        if k.startswith("g:"):
            got = params[k[2:]].value.grad.to_numpy()
            np.testing.assert_allclose(got, z[k], atol=tol * max(1.0, float(np.abs(z[k]).max())), rtol=10 * tol, err_msg=k)


# ------------------------------------------------------------------ fused lookup / loss kernels
@pytest.mark.parametrize("n,V,E", [(1, 7, 5), (77, 13, 64), (4992, 10000, 256), (300, 50, 1030)])
def test_embedding_gather_equals_one_hot_matmul(n, V, E):
    rng = np.random.default_rng(n + V)
    W = rng.standard_normal((V, E)).astype(np.float32)
    ids = rng.integers(0, V, (n,))
    dout = rng.standard_normal((n, E)).astype(np.float32)
    w, x = T(W, requires_grad=True), T(ids)
    out = fb.EmbeddingLookup.apply(x, w)
    np.testing.assert_array_equal(out.to_numpy(), W[ids])            # a gather moves bits
    out.backward(T(dout))
    want = np.zeros((V, E))
    np.add.at(want, ids, dout.astype(np.float64))
    np.testing.assert_allclose(w.grad.to_numpy(), want, atol=1e-5 * max(1.0, float(np.abs(want).max())), rtol=1e-5)
    g2 = ops.embedding_bw(x, T(dout), V).to_numpy()                  # ascending-token accumulation: deterministic
    np.testing.assert_array_equal(w.grad.to_numpy(), g2)
    if n * V <= 10 ** 6:                                             # and the formulation it replaces, on the device
        hot = fb.one_hot(x, V)
        np.testing.assert_array_equal(out.to_numpy(), (hot @ T(W)).to_numpy())


def test_embedding_out_of_range_id_gives_zero_row():
    W = np.arange(12, dtype=np.float32).reshape(3, 4)
    out = ops.embedding_fw(T(np.array([2.0, 5.0, -1.0, 0.0])), T(W)).to_numpy()
    np.testing.assert_array_equal(out, np.stack([W[2], np.zeros(4), np.zeros(4), W[0]]).astype(np.float32))


@pytest.mark.parametrize("n,C", [(1, 1), (5, 3), (257, 1000), (4992, 10000)])
def test_softmax_cross_entropy_kernel_vs_composed_and_fp64(n, C):
    rng = np.random.default_rng(C)
    x = (rng.standard_normal((n, C)) * 3).astype(np.float32)
    t = rng.integers(0, C, (n,))
    g = rng.standard_normal((n,)).astype(np.float32)
    lx = T(x, requires_grad=True)
    loss = fb.softmax_loss(lx, T(t), fused=True)
    x64 = x.astype(np.float64)
    mx = x64.max(axis=1)
    lse = mx + np.log(np.exp(x64 - mx[:, None]).sum(axis=1) + 1e-6)  # minitorch's log(x + EPS)
    np.testing.assert_allclose(loss.to_numpy(), lse - x64[np.arange(n), t], atol=2e-5, rtol=1e-5)
    loss.backward(T(g))
    p = np.exp(x64 - lse[:, None])
    p[np.arange(n), t] -= 1.0
    np.testing.assert_allclose(lx.grad.to_numpy(), g[:, None] * p, atol=2e-6, rtol=1e-4)
    if n * C <= 10 ** 6:       # the composed formulation (minitorch/nn.py:251-271) through the same device ops
        cx = T(x, requires_grad=True)
        closs = fb.softmax_loss(cx, T(t))
        np.testing.assert_allclose(loss.to_numpy(), closs.to_numpy(), atol=2e-5, rtol=1e-5)
        closs.backward(T(g))
        np.testing.assert_allclose(lx.grad.to_numpy(), cx.grad.to_numpy(), atol=5e-6, rtol=1e-4)


def test_decoder_lm_golden_with_fused_embedding_and_loss():
    z = np.load(golden("decoder_small.npz")[0])
    model, params = load_decoder(z, backend=DEV, use_flash_attention=True, use_fused_embedding=True)
    logits, total = decoder_loss(model, z, backend=DEV, fused_loss=True)
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=2e-4, rtol=2e-5)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < 2e-5
    total.backward()
    for k in z.files:
        if k.startswith("g:"):
            got = params[k[2:]].value.grad.to_numpy()
            np.testing.assert_allclose(got, z[k], atol=2e-5 * max(1.0, float(np.abs(z[k]).max())), rtol=2e-4, err_msg=k)


def test_fused_lookup_and_loss_through_host_pointer_ops_equal_device_ops():
    """CudaKernelOps.embedding_* / softmax_xent_* (host-pointer launch_* symbols) give the same bits as the
    device-resident ops, and the DecoderLM golden holds with them on reference-style host storage."""
    rng = np.random.default_rng(21)
    V, E, n, C = 50, 24, 77, 130
    W = rng.standard_normal((V, E)).astype(np.float32)
    ids = rng.integers(0, V, (7, 11))
    dout = rng.standard_normal((7, 11, E)).astype(np.float32)
    x = (rng.standard_normal((n, C)) * 2).astype(np.float32)
    t = rng.integers(0, C, (n,))
    g = rng.standard_normal((n,)).astype(np.float32)
    res = []
    for backend in (fb.default_backend(), DEV):
        mk = lambda a, r=False: fb.tensor_from_numpy(np.asarray(a, np.float32), backend=backend, requires_grad=r)
        w, lx = mk(W, True), mk(x, True)
        emb = fb.EmbeddingLookup.apply(mk(ids), w)
        emb.backward(mk(dout))
        loss = fb.softmax_loss(lx, mk(t), fused=True)
        loss.backward(mk(g))
        res.append((emb.to_numpy(), w.grad.to_numpy(), loss.to_numpy(), lx.grad.to_numpy()))
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(res[0][0], W[ids])
    z = np.load(golden("decoder_small.npz")[0])
    fb.CudaKernelOps.set_flash_mode("fp32")
    model, params = load_decoder(z, backend=fb.default_backend(), use_flash_attention=True, use_fused_embedding=True)
    logits, total = decoder_loss(model, z, backend=fb.default_backend(), fused_loss=True)
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=2e-4, rtol=2e-5)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < 2e-5
