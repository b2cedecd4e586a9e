"""Config #2's model family on the GPU libraries: DecoderLM (minitorch/modules_transfomer.py:339-453) with the
flash-attention path swapped in, every op through the C-ABI libraries (combine.so plumbing + the fused kernels).
(1) fixture-size golden made from the reference's own composed CPU run; (2) config #2's dimensions
(n_vocab 10000, n_embd 256, 8 heads, seq 39 after the label shift, causal) flash vs composed on the same weights."""
import numpy as np
import pytest

import flashattn_b200 as fb
from tests.gpu_util import golden
from tests.test_host_modules import decoder_loss, load_decoder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("branch", ["composed", "flash", "fused"])
def test_decoder_lm_golden_on_gpu(branch):
    z = np.load(golden("decoder_small.npz")[0])
    fb.CudaKernelOps.set_flash_mode("fp32")
    backend = fb.default_backend()
    model, params = load_decoder(z, backend=backend, use_flash_attention=branch == "flash",
                                 use_fused_kernel=branch == "fused")
    logits, total = decoder_loss(model, z, backend=backend)
    tol = 1e-4 if branch == "fused" else 2e-5   # fused: LayerNorm epsilon 1e-8 (kernel) instead of 1e-5
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=tol * 10, rtol=tol)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < tol
    total.backward()
    for k in z.files:
        if k.startswith("g:"):
            got = params[k[2:]].value.grad.to_numpy()
            np.testing.assert_allclose(got, z[k], atol=tol * max(1.0, float(np.abs(z[k]).max())), rtol=10 * tol,
                                       err_msg=k)


@pytest.mark.timeout(600)
def test_decoder_cfg2_flash_equals_composed():
    """BASELINE config #2 dimensions, batch cut to 8 to keep the composed arm quick: logits, loss and the
    first layer's q-projection gradient of the flash model against the composed model (fp32 mode)."""
    fb.CudaKernelOps.set_flash_mode("fp32")
    backend = fb.default_backend()
    n_vocab, n_embd, n_head, n_pos, B = 10000, 256, 8, 40, 8
    rng = np.random.default_rng(11111)
    ids = rng.integers(0, n_vocab, (B, n_pos))
    w = np.zeros((B, n_pos), np.float32)
    w[:, n_pos // 2:] = 1.0
    z = dict(input_ids=ids[:, :-1], labels=ids[:, 1:], label_token_weights=w[:, 1:])
    np.random.seed(5)
    kw = dict(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_positions=n_pos, p_dropout=0.0, ln_eps=1e-5,
              bias=True, backend=backend)
    flash = fb.DecoderLM(use_flash_attention=True, **kw)
    comp = fb.DecoderLM(use_flash_attention=False, **kw)
    pf, pc = dict(flash.named_parameters()), dict(comp.named_parameters())
    for name, prm in pf.items():
        pc[name].value = fb.tensor_from_numpy(prm.value.to_numpy(), backend=backend, requires_grad=True)
    out = []
    for model, params in ((flash, pf), (comp, pc)):
        logits, total = decoder_loss(model, z, backend=backend)
        total.backward()
        out.append((logits.to_numpy(), total.to_numpy().reshape(-1),
                    params["t_layer_1.attention.q_projection.weights"].value.grad.to_numpy(),
                    params["lm_head.bias"].value.grad.to_numpy()))
    for got, want in zip(*out):
        np.testing.assert_allclose(got, want, atol=2e-5 * max(1.0, float(np.abs(want).max())), rtol=1e-4)
