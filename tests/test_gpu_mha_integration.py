"""MultiHeadAttention-level parity, mirroring the reference's tests/test_flash_attention.py
(:103-186 causal flash, :24-99 non-causal): same recipe (seeds 10, np.random.rand inputs, weights
copied from torch.nn.MultiheadAttention, bias=False, p_dropout=0), same 1e-5 tolerance, same
checked quantities (layer output, X.grad, out-projection weight grad, q/k/v weight grads).
The golden file holds the reference's composed minitorch path AND torch's answer for it
(tests/golden/make_golden.py).  Only the attention core runs on the GPU (the flash path under
test); the four bias-free Linear projections around it are plain numpy here because minitorch's
combine.so matmul is outside the hot path."""
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from tests.gpu_util import golden, maxabs

pytestmark = pytest.mark.gpu
T = fb.tensor_from_numpy


@pytest.mark.parametrize("path", golden("mha_cfg1_*.npz"), ids=os.path.basename)
def test_multihead_attention_flash_matches_reference_and_torch(path):
    z = np.load(path)
    fb.CudaKernelOps.set_flash_mode("fp32")
    X, Wq, Wk, Wv, Wo = (z[k].astype(np.float32) for k in ("X", "Wq", "Wk", "Wv", "Wo"))
    nh, causal = int(z["n_head"]), bool(z["causal"])
    B, N, E = X.shape
    d = E // nh
    x2 = X.reshape(B * N, E)
    # project_to_query_key_value (modules_transfomer.py:87-100): (B,N,nh,d) storage, permuted views
    leaves = [T((x2 @ W).reshape(B, N, nh, d), requires_grad=True) for W in (Wq, Wk, Wv)]
    q, k, v = (t.permute(0, 2, 1, 3) for t in leaves)
    attn = q.flash_attention_causal(k, v) if causal else q.flash_attention(k, v)
    # self_attention tail (:197-199) and out_projection (:227)
    A = attn.permute(0, 2, 1, 3).contiguous().view(B * N, E)
    Y = (A.to_numpy() @ Wo).reshape(B, N, E)
    for ref in ("Y_ref", "Y_torch"):
        np.testing.assert_allclose(Y, z[ref], atol=1e-5, rtol=1e-5)
    # result.sum().backward()  (:167)
    dY2 = np.ones((B * N, E), dtype=np.float32)
    dWo = A.to_numpy().T @ dY2
    A.backward(T(dY2 @ Wo.T))
    dq2, dk2, dv2 = (t.grad.to_numpy().reshape(B * N, E) for t in leaves)
    dX = (dq2 @ Wq.T + dk2 @ Wk.T + dv2 @ Wv.T).reshape(B, N, E)
    for ref in ("dX_ref", "dX_torch"):
        np.testing.assert_allclose(dX, z[ref], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(dWo, z["dWo_torch"], atol=2e-4, rtol=1e-5)
    for got, name in ((x2.T @ dq2, "dWq_ref"), (x2.T @ dk2, "dWk_ref"), (x2.T @ dv2, "dWv_ref")):
        np.testing.assert_allclose(got, z[name], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("N,E,nh,causal", [(2048, 64, 16, True), (2048, 1024, 2, True), (4096, 256, 4, True),
                                           (1024, 2048, 2, False)])
def test_reference_test_grid_points_against_fp64_oracle(N, E, nh, causal):
    """Grid points of tests/test_flash_attention.py:103-108 (head dims 4, 512, 64, 1024) that a 62 GB
    host cannot run through torch (SURVEY.md section 4 'Feasibility'): batch reduced to 2, oracle fp64."""
    from oracle import attention_ref as R
    fb.CudaKernelOps.set_flash_mode("fp32")
    rng = np.random.default_rng(N + E)
    B, d = 2, E // nh
    Q, K, V, dO = (rng.random((B, nh, N, d)).astype(np.float32) for _ in range(4))
    from flashattn_b200 import device as dev
    dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(x) for x in (Q, K, V, dO))
    O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal)
    g = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal)
    if N * N * B * nh * 8 < 6e9:
        Oe, _, _ = R.attention_fwd(Q, K, V, causal=causal)
        ge = R.attention_bwd(Q, K, V, dO, causal=causal)
        assert maxabs(O.to_numpy(), Oe) < 1e-5
        for got, want in zip(g, ge):
            assert maxabs(got.to_numpy(), want) < 1e-5 * max(1.0, float(np.abs(want).max()))
    else:  # oracle per (b,h) slice to bound memory
        Og = O.to_numpy()
        gg = [x.to_numpy() for x in g]
        for b, h in ((0, 0), (B - 1, nh - 1)):
            sl = (slice(b, b + 1), slice(h, h + 1))
            Oe, _, _ = R.attention_fwd(Q[sl], K[sl], V[sl], causal=causal)
            ge = R.attention_bwd(Q[sl], K[sl], V[sl], dO[sl], causal=causal)
            assert maxabs(Og[sl], Oe) < 1e-5
            for got, want in zip(gg, ge):
                assert maxabs(got[sl], want) < 1e-5 * max(1.0, float(np.abs(want).max()))
