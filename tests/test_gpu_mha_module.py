"""The MultiHeadAttention call site (minitorch/modules_transfomer.py:19-229) end to end on the GPU: projections
through combine.so's MatrixMultiply, attention core through each of the module's three branches -- flash
(the path under test), fused softmax kernel, composed matmul/softmax/matmul -- against the golden file made
from the reference's own composed minitorch path and torch.nn.MultiheadAttention
(tests/golden/make_golden.py; recipe and 1e-5 tolerance of tests/test_flash_attention.py:24-186)."""
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from tests.gpu_util import golden

pytestmark = pytest.mark.gpu
T = fb.tensor_from_numpy


def _build(z, **flags):
    nh, causal = int(z["n_head"]), bool(z["causal"])
    E = z["X"].shape[-1]
    layer = fb.MultiHeadAttention(E, nh, causal=causal, p_dropout=0.0, bias=False, **flags)
    for lin, key in ((layer.q_projection, "Wq"), (layer.k_projection, "Wk"), (layer.v_projection, "Wv"),
                     (layer.out_projection, "Wo")):
        lin.weights.value = T(z[key].astype(np.float32), requires_grad=True)
    return layer


@pytest.mark.parametrize("branch", ["flash", "fused", "composed"])
@pytest.mark.parametrize("path", golden("mha_cfg1_*.npz"), ids=os.path.basename)
def test_mha_module_three_branches(path, branch):
    z = np.load(path)
    fb.CudaKernelOps.set_flash_mode("fp32")
    layer = _build(z, use_flash_attention=branch == "flash", use_fused_kernel=branch == "fused")
    X = T(z["X"].astype(np.float32), requires_grad=True)
    Y = layer(X)
    assert Y.shape == z["X"].shape
    for ref in ("Y_ref", "Y_torch"):
        np.testing.assert_allclose(Y.to_numpy(), z[ref], atol=1e-5, rtol=1e-5)
    Y.sum().backward()                                      # result.sum().backward()  (:167)
    for ref in ("dX_ref", "dX_torch"):
        np.testing.assert_allclose(X.grad.to_numpy(), z[ref], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(layer.out_projection.weights.value.grad.to_numpy(), z["dWo_torch"], atol=2e-4, rtol=1e-5)
    for lin, name in ((layer.q_projection, "dWq_ref"), (layer.k_projection, "dWk_ref"), (layer.v_projection, "dWv_ref")):
        np.testing.assert_allclose(lin.weights.value.grad.to_numpy(), z[name], atol=2e-4, rtol=1e-4)


def test_mha_module_flash_equals_composed_with_bias_and_odd_length():
    """Same weights through the flash and the composed branch: N = 39 (config #2's sequence length), bias on."""
    rng = np.random.default_rng(8)
    B, N, E, nh = 3, 39, 64, 8
    np.random.seed(3)
    a = fb.MultiHeadAttention(E, nh, causal=True, p_dropout=0.0, bias=True, use_flash_attention=True)
    b = fb.MultiHeadAttention(E, nh, causal=True, p_dropout=0.0, bias=True)
    for la, lb in zip((a.q_projection, a.k_projection, a.v_projection, a.out_projection),
                      (b.q_projection, b.k_projection, b.v_projection, b.out_projection)):
        lb.weights.value = T(la.weights.value.to_numpy(), requires_grad=True)
        lb.bias.value = T(la.bias.value.to_numpy(), requires_grad=True)
    x = rng.standard_normal((B, N, E)).astype(np.float32)
    g = rng.standard_normal((B, N, E)).astype(np.float32)
    outs = []
    for layer in (a, b):
        X = T(x, requires_grad=True)
        Y = layer(X)
        Y.backward(T(g))
        outs.append((Y.to_numpy(), X.grad.to_numpy(), layer.q_projection.weights.value.grad.to_numpy(),
                     layer.v_projection.bias.value.grad.to_numpy()))
    for got, want in zip(*outs):
        np.testing.assert_allclose(got, want, atol=2e-5 * max(1.0, float(np.abs(want).max())), rtol=1e-5)
