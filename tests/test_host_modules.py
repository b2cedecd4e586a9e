"""Host-side logic on CPU: HostTensor autograd and the MultiHeadAttention / DecoderLM call sites driven through
``OracleOps`` (tests/host_ops.py -- the oracle behind the CudaKernelOps surface), checked against golden vectors
made from the reference itself (tests/golden/make_golden.py).  The same modules run on the GPU libraries in
tests/test_gpu_mha_module.py and tests/test_gpu_decoder.py."""
import os

import numpy as np
import pytest

import flashattn_b200 as fb
from tests.gpu_util import golden
from tests.host_ops import OracleOps

BACKEND = fb.TensorBackend(OracleOps)


def T(a, requires_grad=False):
    return fb.tensor_from_numpy(np.asarray(a, dtype=np.float32), backend=BACKEND, requires_grad=requires_grad)


def load_decoder(z, **flags):
    n_vocab, n_embd, n_head, n_pos = (int(v) for v in z["cfg"])
    backend = flags.pop("backend", BACKEND)
    model = fb.DecoderLM(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_positions=n_pos, p_dropout=0.0,
                         ln_eps=1e-5, bias=True, backend=backend, **flags)
    params = dict(model.named_parameters())
    names = [k[2:] for k in z.files if k.startswith("p:")]
    if not flags.get("use_fused_kernel"):           # FusedLayerNorm holds plain tensors, not Parameters
        assert sorted(params) == sorted(names)      # same parameter tree as the reference's named_parameters()
    for name in (n for n in names if n in params):
        params[name].value = fb.tensor_from_numpy(z["p:" + name], backend=backend, requires_grad=True)
    return model, params


def decoder_loss(model, z, backend=BACKEND, fused_loss=False):
    mk = lambda a: fb.tensor_from_numpy(np.asarray(a, dtype=np.float32), backend=backend)
    logits = model(mk(z["input_ids"]))
    bs, l, c = logits.shape
    loss = fb.softmax_loss(logits.view(bs * l, c), mk(z["labels"]).view(bs * l), fused=fused_loss)
    w = mk(z["label_token_weights"]).view(bs * l)
    return logits, (loss * w).sum() / w.sum()


@pytest.mark.parametrize("branch", ["composed", "flash", "fused"])
def test_decoder_lm_matches_reference_golden(branch):
    z = np.load(golden("decoder_small.npz")[0])
    model, params = load_decoder(z, use_flash_attention=branch == "flash", use_fused_kernel=branch == "fused")
    logits, total = decoder_loss(model, z)
    # the fused branch swaps LayerNorm1d(eps=1e-5) for the fused kernel's fixed 1e-8 epsilon: same model up to that
    tol = 1e-4 if branch == "fused" else 2e-5
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=tol * 10, rtol=tol)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < tol
    total.backward()
    for k in z.files:
        if k.startswith("g:"):
            got = params[k[2:]].value.grad.to_numpy()
            np.testing.assert_allclose(got, z[k], atol=tol * max(1.0, float(np.abs(z[k]).max())), rtol=10 * tol,
                                       err_msg=k)


def test_decoder_lm_fused_embedding_and_loss_match_reference_golden():
    """Gather-kernel embeddings + single-kernel cross-entropy in place of the one-hot matmuls:
    same logits / loss / gradients as the reference's composed run."""
    z = np.load(golden("decoder_small.npz")[0])
    model, params = load_decoder(z, use_flash_attention=True, use_fused_embedding=True)
    logits, total = decoder_loss(model, z, fused_loss=True)
    np.testing.assert_allclose(logits.to_numpy(), z["logits"], atol=2e-4, rtol=2e-5)
    assert abs(float(total.to_numpy().reshape(-1)[0]) - float(z["loss"][0])) < 2e-5
    total.backward()
    for k in z.files:
        if k.startswith("g:"):
            got = params[k[2:]].value.grad.to_numpy()
            np.testing.assert_allclose(got, z[k], atol=2e-5 * max(1.0, float(np.abs(z[k]).max())), rtol=2e-4, err_msg=k)


@pytest.mark.parametrize("branch", ["flash", "fused", "composed"])
@pytest.mark.parametrize("path", golden("mha_cfg1_*.npz"), ids=os.path.basename)
def test_mha_module_host_logic(path, branch):
    z = np.load(path)
    layer = fb.MultiHeadAttention(z["X"].shape[-1], int(z["n_head"]), causal=bool(z["causal"]), p_dropout=0.0,
                                  bias=False, backend=BACKEND, use_flash_attention=branch == "flash",
                                  use_fused_kernel=branch == "fused")
    for lin, key in ((layer.q_projection, "Wq"), (layer.k_projection, "Wk"), (layer.v_projection, "Wv"),
                     (layer.out_projection, "Wo")):
        lin.weights.value = T(z[key], requires_grad=True)
    X = T(z["X"], requires_grad=True)
    Y = layer(X)
    np.testing.assert_allclose(Y.to_numpy(), z["Y_ref"], atol=1e-5, rtol=1e-5)
    Y.sum().backward()
    np.testing.assert_allclose(X.grad.to_numpy(), z["dX_ref"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(layer.q_projection.weights.value.grad.to_numpy(), z["dWq_ref"], atol=2e-4, rtol=1e-4)


def test_autograd_nodes_against_finite_differences():
    """GELU(x @ W + b).var / logsumexp / pow chain: analytic gradient vs central differences (fp64 oracle ops)."""
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal((5, 7)).astype(np.float32)
    W0 = rng.standard_normal((7, 3)).astype(np.float32)
    b0 = rng.standard_normal((3,)).astype(np.float32)

    def f(xv, Wv, bv):
        x, W, b = T(xv, True), T(Wv, True), T(bv, True)
        y = fb.GELU(x @ W + b)
        out = (fb.logsumexp(y, dim=1).view(5) * 0.5).sum() + (y.var(dim=0) + 1.0) .log().sum() - (y ** 2).mean()
        return out, (x, W, b)

    out, leaves = f(x0, W0, b0)
    out.backward()
    for arr, leaf in zip((x0, W0, b0), leaves):
        g = leaf.grad.to_numpy()
        for idx in [tuple(rng.integers(0, s) for s in arr.shape) for _ in range(4)]:
            e = 1e-2
            hi, lo = arr.copy(), arr.copy()
            hi[idx] += e
            lo[idx] -= e
            args_hi = [hi if a is arr else a for a in (x0, W0, b0)]
            args_lo = [lo if a is arr else a for a in (x0, W0, b0)]
            fd = (float(f(*args_hi)[0].to_numpy().reshape(-1)[0]) - float(f(*args_lo)[0].to_numpy().reshape(-1)[0])) / (2 * e)
            assert abs(fd - g[idx]) < 5e-3 * max(1.0, abs(fd)), (idx, fd, g[idx])


def test_generate_greedy_matches_stepwise_argmax_and_is_branch_independent():
    """generate() (project/run_machine_translation.py:300-325): flash and composed models with the same weights
    emit the same tokens, and each token is the arg-max of a fresh full-prefix forward."""
    z = np.load(golden("decoder_small.npz")[0])
    flash, _ = load_decoder(z, use_flash_attention=True)
    comp, _ = load_decoder(z)
    prompt = [int(t) for t in z["input_ids"][0, :5]]
    a = fb.generate(flash, prompt, model_max_length=12)
    b = fb.generate(comp, prompt, model_max_length=12)
    assert a == b and a[:5] == prompt and len(a) == 13
    logits = flash(T(np.asarray(a[:-1], dtype=np.float32).reshape(1, -1))).to_numpy()
    assert int(np.argmax(logits[0, -1])) == a[-1]


def test_generate_cached_host_logic_matches_full_prefix_loop():
    """Cache-aware generation (prefill once, then one position per token) through the oracle-backed
    ops: same greedy tokens as the reference-style full-prefix loop, per-step logits equal the last-position logits of a
    full forward, cache length bookkeeping, overflow is an error."""
    z = np.load(golden("decoder_small.npz")[0])
    model, _ = load_decoder(z, use_flash_attention=True)
    prompt = [int(t) for t in z["input_ids"][0, :5]]
    full = fb.generate(model, prompt, model_max_length=12)
    cached = fb.generate_cached(model, prompt, model_max_length=12)
    assert cached == full
    ops = model.backend.ops
    attn = model.t_layer_1.attention
    caches = [ops.kv_cache_new(1, attn.n_head, 9, attn.attn_hidden_dim) for _ in range(4)]
    lg = fb.decode_step(model, np.asarray(full[:6], dtype=np.float32).reshape(1, 6), caches).to_numpy()
    ref = model(T(np.asarray(full[:6], dtype=np.float32).reshape(1, 6))).to_numpy()
    np.testing.assert_allclose(lg, ref, atol=2e-5, rtol=1e-5)
    for t in range(6, 9):
        lg = fb.decode_step(model, np.asarray([[full[t]]], dtype=np.float32), caches).to_numpy()
        ref = model(T(np.asarray(full[:t + 1], dtype=np.float32).reshape(1, t + 1))).to_numpy()
        np.testing.assert_allclose(lg[0, 0], ref[0, t], atol=2e-5, rtol=1e-5)
    assert [c.len for c in caches] == [9] * 4
    import pytest
    with pytest.raises(ValueError):
        fb.decode_step(model, np.asarray([[full[9]]], dtype=np.float32), caches)


def test_module_tree_train_eval_and_dropout():
    """Module.train()/eval() reach every sub-module; Dropout is the identity at p = 0 or in eval mode and rescales
    by 1/(1-p) in training (minitorch/modules_basic.py:73-101)."""
    np.random.seed(0)
    layer = fb.TransformerLayer(16, 2, p_dropout=0.5, backend=BACKEND, use_flash_attention=True)
    layer.eval()
    assert not layer.attention.dropout.training and not layer.ff.dropout.training
    layer.train()
    assert layer.attention.dropout.training and layer.ff.training
    x = T(np.ones((4, 8)))
    d = fb.Dropout(0.5)
    d.eval()
    assert d(x) is x
    d.train()
    y = d(x).to_numpy()
    assert set(np.unique(y)) <= {0.0, 2.0} and 0 < (y == 0).sum() < y.size
    assert fb.Dropout(0.0)(x) is x


def test_linear_init_and_parameter_names():
    np.random.seed(1)
    lin = fb.Linear(64, 32, True, BACKEND)
    w = lin.weights.value.to_numpy()
    assert w.shape == (64, 32) and np.abs(w).max() <= (1 / 64) ** 0.5 + 1e-7       # uniform(+-1/sqrt(in))
    names = [n for n, _ in fb.DecoderLM(11, 16, 2, 8, backend=BACKEND).named_parameters()]
    assert "t_layer_3.attention.k_projection.bias" in names and "lm_head.weights" in names and len(names) == 70
    y = lin(T(np.ones((3, 64))))
    np.testing.assert_allclose(y.to_numpy(), np.ones((3, 64)) @ w + lin.bias.value.to_numpy(), atol=1e-5)


def test_transformer_layer_flash_equals_composed_and_fused():
    """Pre-LN block (minitorch/modules_transfomer.py:278-336): same weights through the three attention cores."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 10, 32)).astype(np.float32)
    g = rng.standard_normal((2, 10, 32)).astype(np.float32)
    np.random.seed(3)
    ref = fb.TransformerLayer(32, 4, p_dropout=0.0, ln_eps=1e-5, backend=BACKEND)
    outs = []
    for flags in ({}, {"use_flash_attention": True}):
        np.random.seed(3)
        layer = fb.TransformerLayer(32, 4, p_dropout=0.0, ln_eps=1e-5, backend=BACKEND, **flags)
        X = T(x, True)
        Y = layer(X)
        Y.backward(T(g))
        outs.append((Y.to_numpy(), X.grad.to_numpy(), layer.ff.linear_in.weights.value.grad.to_numpy()))
    for a, b in zip(*outs):
        np.testing.assert_allclose(a, b, atol=2e-5 * max(1.0, float(np.abs(b).max())), rtol=1e-5)
    assert ref is not None
