"""GPU parity of the flash-attention path in FA_MODE_FP32 (tolerance 1e-5, the bar of the
reference's tests/test_flash_attention.py:162-179) against golden vectors made from the
reference's composed attention, and against the fp64 oracle on shapes the reference's test
grid uses (head dims 4..1024, ragged N, generation-time N)."""
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from flashattn_b200 import device as dev
from oracle import attention_ref as R
from tests.gpu_util import golden, maxabs

pytestmark = pytest.mark.gpu
ops = fb.CudaKernelOps
T = fb.tensor_from_numpy
TOL = 1e-5


@pytest.fixture(autouse=True)
def _fp32_mode():
    ops.set_flash_mode("fp32")
    yield


@pytest.mark.parametrize("path", golden("attn_*.npz"), ids=os.path.basename)
def test_flash_golden_through_operator_surface(path):
    z = np.load(path)
    causal = bool(z["causal"])
    km = T(z["key_mask"]) if "key_mask" in z.files else None
    q, k, v = (T(z[n], requires_grad=True) for n in ("Q", "K", "V"))
    fw = ops.flash_attention_causal_fw if causal else ops.flash_attention_fw
    bw = ops.flash_attention_causal_bw if causal else ops.flash_attention_bw
    O, m, l = fw(q, k, v, key_mask=km)
    assert O.shape == z["Q"].shape and m.shape == z["Q"].shape[:3] and l.shape == z["Q"].shape[:3]
    assert maxabs(O.to_numpy(), z["O"]) < TOL
    dQ, dK, dV = bw(q, k, v, O, T(z["dO"]), m, l, key_mask=km)
    # gradients: 1e-5 relative to the gradient scale (the golden itself is an fp32 computation)
    for got, name in ((dQ, "dQ"), (dK, "dK"), (dV, "dV")):
        ref = z[name]
        assert maxabs(got.to_numpy(), ref) < TOL * max(1.0, float(np.abs(ref).max())), name
    # (m, l) are the row max of the scaled scores and sum exp(s - m)  (flashattention_kernel.cu:81-89)
    _, me, le = R.attention_fwd(z["Q"], z["K"], z["V"], causal=causal,
                                key_mask=z["key_mask"] if "key_mask" in z.files else None)
    assert maxabs(m.to_numpy(), me) < 1e-5 * max(1.0, float(np.abs(me).max()))
    np.testing.assert_allclose(l.to_numpy(), le, rtol=2e-5)


@pytest.mark.parametrize("causal", [False, True])
def test_autograd_function_with_permuted_inputs(causal):
    """q,k,v arrive as non-contiguous permuted views of (B,N,nh,d) storage, exactly what
    MultiHeadAttention.project_to_query_key_value produces (modules_transfomer.py:87-100)."""
    rng = np.random.default_rng(3)
    B, N, nh, d = 2, 50, 3, 16
    raw = [rng.standard_normal((B, N, nh, d)).astype(np.float32) for _ in range(3)]
    leaves = [T(r, requires_grad=True) for r in raw]
    q, k, v = (t.permute(0, 2, 1, 3) for t in leaves)
    out = q.flash_attention_causal(k, v) if causal else q.flash_attention(k, v)
    dO = rng.standard_normal((B, nh, N, d)).astype(np.float32)
    out.backward(T(dO))
    Qe, Ke, Ve = (r.transpose(0, 2, 1, 3) for r in raw)
    Oe, _, _ = R.attention_fwd(Qe, Ke, Ve, causal=causal)
    ge = R.attention_bwd(Qe, Ke, Ve, dO, causal=causal)
    assert maxabs(out.to_numpy(), Oe) < TOL
    for leaf, g in zip(leaves, ge):
        assert maxabs(leaf.grad.to_numpy(), g.transpose(0, 2, 1, 3)) < 2e-5


@pytest.mark.parametrize("B,H,N,d,causal", [
    (1, 1, 1, 4, True), (2, 2, 5, 8, False), (1, 2, 41, 32, True), (2, 1, 64, 4, True), (1, 2, 65, 16, False),
    (1, 1, 127, 64, True), (1, 2, 130, 128, False), (1, 1, 200, 192, True), (1, 1, 96, 256, False),
    (1, 1, 70, 320, True), (1, 1, 66, 512, False), (1, 1, 40, 1024, True), (2, 3, 300, 64, True),
    (1, 2, 1000, 32, False),
])
def test_device_api_fp32_shapes(B, H, N, d, causal):
    rng = np.random.default_rng(N * 31 + d)
    Q, K, V, dO = (rng.standard_normal((B, H, N, d)).astype(np.float32) for _ in range(4))
    dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(x) for x in (Q, K, V, dO))
    O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal)
    gq, gk, gv = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal)
    Oe, me, le = R.attention_fwd(Q, K, V, causal=causal)
    ge = R.attention_bwd(Q, K, V, dO, causal=causal)
    assert maxabs(O.to_numpy(), Oe) < TOL
    assert maxabs(m.to_numpy() + np.log(l.to_numpy()), me + np.log(le)) < 2e-5
    for got, want in zip((gq, gk, gv), ge):
        assert maxabs(got.to_numpy(), want) < TOL * max(1.0, float(np.abs(want).max()))


def test_kv_len_padding_and_strided_layout():
    rng = np.random.default_rng(11)
    B, H, N, d = 3, 2, 90, 32
    kv = np.array([90, 1, 47], dtype=np.int32)
    # (B,N,H,d) storage consumed in place through strides -- no .contiguous() copy
    raw = [rng.standard_normal((B, N, H, d)).astype(np.float32) for _ in range(4)]
    Q, K, V, dO = (r.transpose(0, 2, 1, 3) for r in raw)
    strides = (N * H * d, d, H * d)
    dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(r) for r in raw)
    dkv = dev.DeviceArray.from_numpy(kv)
    for causal in (False, True):
        O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal, kv_len=dkv, shape=(B, H, N, d), strides=strides)
        gq, gk, gv = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal, kv_len=dkv, shape=(B, H, N, d),
                                   strides=strides)
        Oe, _, _ = R.attention_fwd(Q, K, V, causal=causal, kv_len=kv)
        ge = R.attention_bwd(Q, K, V, dO, causal=causal, kv_len=kv)
        assert maxabs(O.to_numpy().transpose(0, 2, 1, 3), Oe) < TOL
        for got, want in zip((gq, gk, gv), ge):
            assert maxabs(got.to_numpy().transpose(0, 2, 1, 3), want) < 2e-5
        # padded keys receive exactly zero gradient
        assert not gk.to_numpy()[1, 1:].any() and not gv.to_numpy()[2, 47:].any()


def test_determinism_and_errors():
    rng = np.random.default_rng(2)
    Q, K, V, dO = (rng.standard_normal((1, 2, 257, 64)).astype(np.float32) for _ in range(4))
    d = [dev.DeviceArray.from_numpy(x) for x in (Q, K, V, dO)]
    runs = []
    for _ in range(2):
        O, m, l = dev.flash_fwd(d[0], d[1], d[2], causal=True)
        g = dev.flash_bwd(d[0], d[1], d[2], O, d[3], m, l, causal=True)
        runs.append([x.to_numpy() for x in (O,) + g])
    for a, b in zip(*runs):
        np.testing.assert_array_equal(a, b)  # no atomics in this mode: bitwise reproducible
    lib = fb._lib.load("flashattention_kernel")
    a = fb._lib.fa_attn_desc()
    a.B, a.H, a.N, a.d, a.dtype = 1, 1, 0, 8, 0
    assert lib.fa_flash_fwd_dev(a, None, None, None, None, None, None, None) == fb._lib.FA_ERR_INVALID
    assert b"bad shape" in lib.fa_last_error()
