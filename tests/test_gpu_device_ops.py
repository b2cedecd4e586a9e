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
