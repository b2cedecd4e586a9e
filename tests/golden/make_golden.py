#!/usr/bin/env python
"""Generate the golden vectors in this directory FROM THE REFERENCE ITSELF.

Run once in the build container (the only place /root/reference exists):

    NUMBA_DISABLE_CUDA=1 python tests/golden/make_golden.py

It imports the reference's minitorch from /root/reference (read-only, nothing is
copied into this repo) and drives its *composed* attention path
(minitorch/modules_transfomer.py:177-192 -> nn.softmax -> matmul) on the numba CPU
backend, exactly as BASELINE.md section 3 describes, with the two shims the survey
found necessary (fused-op stubs; >3-D matmul flatten).  It also evaluates the
reference tests' own oracle (torch.nn.MultiheadAttention, tests/test_flash_attention.py:40-70)
and the composed formulas kernel_tests/* use for softmax / layernorm.

Outputs (small .npz files, float32) are committed; the GPU box never needs the
reference.  tests/test_oracle.py pins oracle/attention_ref.py against them and
the -m gpu tests pin the CUDA path against them.
"""
import os
import sys

os.environ.setdefault("NUMBA_DISABLE_CUDA", "1")
REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import minitorch  # noqa: E402  (the reference's own package)
from minitorch.fast_ops import FastOps  # noqa: E402
from minitorch.tensor_ops import SimpleOps  # noqa: E402
from minitorch.tensor_functions import tensor_from_numpy  # noqa: E402

datatype = np.float32


class FastOpsX(FastOps):
    """FastOps + the fused-op stubs of SimpleOps (tensor_ops.py:240-269) + a matmul
    that flattens >3-D operands like CudaKernelOps.matrix_multiply
    (cuda_kernel_ops.py:357-369)."""

    attn_softmax_fw = SimpleOps.attn_softmax_fw
    attn_softmax_bw = SimpleOps.attn_softmax_bw
    layernorm_fw = SimpleOps.layernorm_fw
    layernorm_bw = SimpleOps.layernorm_bw
    flash_attention_fw = SimpleOps.flash_attention_fw
    flash_attention_bw = SimpleOps.flash_attention_bw
    flash_attention_causal_fw = SimpleOps.flash_attention_causal_fw
    flash_attention_causal_bw = SimpleOps.flash_attention_causal_bw

    @staticmethod
    def matrix_multiply(a, b):
        if len(a.shape) > 3:
            ls = list(a.shape[:-2])
            a3 = a.contiguous().view(int(np.prod(ls)), a.shape[-2], a.shape[-1])
            b3 = b.contiguous().view(int(np.prod(ls)), b.shape[-2], b.shape[-1])
            out = FastOps.matrix_multiply(a3, b3)
            return out.view(*ls, out.shape[-2], out.shape[-1])
        return FastOps.matrix_multiply(a, b)


backend = minitorch.TensorBackend(FastOpsX)


def mt(x, grad=True):
    return tensor_from_numpy(np.ascontiguousarray(x, dtype=datatype), backend=backend, requires_grad=grad)


def composed_attention(Q, K, V, dO, causal, key_mask=None):
    """The reference's composed branch (modules_transfomer.py:177-192) on (B,H,N,d)
    inputs, plus backward with an arbitrary upstream dO (sum(O*dO).backward())."""
    q, k, v = mt(Q), mt(K), mt(V)
    B, H, N, d = Q.shape
    kT = k.permute(0, 1, 3, 2)
    s = (q @ kT) / (d ** 0.5)
    if causal:
        mask = -np.finfo(datatype).max * np.triu(np.ones((B, H, N, N), dtype=datatype), 1)
        s = s + mt(mask)
    if key_mask is not None:
        km = np.broadcast_to(key_mask[:, None, None, :], (B, H, N, N))
        s = s + mt(km)
    o = minitorch.nn.softmax(s, dim=3) @ v
    (o * mt(dO, grad=False)).sum().backward()
    return (o.to_numpy(), q.grad.to_numpy(), k.grad.to_numpy(), v.grad.to_numpy())


def gen_attention():
    cases = {}
    rng = np.random.default_rng(0)
    specs = [
        ("cfg1", 2, 4, 64, 32, False, None),
        ("cfg1_causal", 2, 4, 64, 32, True, None),
        ("n39_causal", 3, 2, 39, 32, True, None),
        ("n100", 1, 2, 100, 16, False, None),
        ("n129_d64_causal", 1, 2, 129, 64, True, None),
        ("n5_d8", 2, 1, 5, 8, False, None),
        ("n1", 2, 2, 1, 32, True, None),
        ("pad_mask", 3, 2, 48, 32, False, "pad"),
        ("pad_mask_causal", 3, 2, 48, 32, True, "pad"),
        ("n160_d128", 1, 1, 160, 128, False, None),
    ]
    for name, B, H, N, d, causal, pad in specs:
        Q = rng.standard_normal((B, H, N, d)).astype(datatype)
        K = rng.standard_normal((B, H, N, d)).astype(datatype)
        V = rng.standard_normal((B, H, N, d)).astype(datatype)
        dO = rng.standard_normal((B, H, N, d)).astype(datatype)
        key_mask = None
        kv_len = None
        if pad:
            kv_len = np.array([N, 1, N // 2 + 3][:B], dtype=np.int32)
            key_mask = np.where(np.arange(N)[None, :] < kv_len[:, None], 0.0, -1e8).astype(datatype)
        O, dQ, dK, dV = composed_attention(Q, K, V, dO, causal, key_mask)
        d_ = dict(Q=Q, K=K, V=V, dO=dO, O=O, dQ=dQ, dK=dK, dV=dV,
                  causal=np.array(causal), B=B, H=H, N=N, d=d)
        if pad:
            d_["key_mask"] = key_mask
            d_["kv_len"] = kv_len
        cases[name] = d_
        print("attention", name, O.shape, float(np.abs(O).max()))
    for name, d_ in cases.items():
        np.savez_compressed(os.path.join(HERE, f"attn_{name}.npz"), **d_)


def gen_mha():
    """tests/test_flash_attention.py:24-99 recipe (seeds 10, rand inputs, torch MHA
    weights) at config #1, with the reference's composed path AND torch's answer."""
    for causal in (False, True):
        np.random.seed(10)
        torch.manual_seed(10)
        B, N, E, nh = 2, 64, 128, 4
        data = np.random.rand(B, N, E).astype(datatype)
        X = mt(data)
        X_ = torch.tensor(data, dtype=torch.float32, requires_grad=True)
        layer_ = torch.nn.MultiheadAttention(E, nh, bias=False, batch_first=True, dtype=torch.float32)
        layer = minitorch.MultiHeadAttention(E, nh, causal, 0.0, False, backend,
                                             use_fused_kernel=False, use_flash_attention=False)
        w_qkv = layer_.in_proj_weight.detach().numpy().T.copy()
        w_q_, w_k_, w_v_ = [w.copy() for w in np.split(w_qkv, 3, -1)]
        w_out_ = layer_.out_proj.weight.detach().numpy().T.copy()
        layer.q_projection.weights.value = mt(w_q_)
        layer.k_projection.weights.value = mt(w_k_)
        layer.v_projection.weights.value = mt(w_v_)
        layer.out_projection.weights.value = mt(w_out_)
        M = torch.triu(-float("inf") * torch.ones(N, N), 1) if causal else None
        result = layer(X)
        result_, _ = layer_(X_, X_, X_, attn_mask=M)
        result.sum().backward()
        result_.sum().backward()
        out = dict(
            X=data, Wq=w_q_, Wk=w_k_, Wv=w_v_, Wo=w_out_, n_head=nh, causal=np.array(causal),
            Y_ref=result.to_numpy(), Y_torch=result_.detach().numpy(),
            dX_ref=X.grad.to_numpy(), dX_torch=X_.grad.detach().numpy(),
            dWo_ref=layer.out_projection.weights.value.grad.to_numpy(),
            dWo_torch=layer_.out_proj.weight.grad.detach().numpy().T.copy(),
            dWq_ref=layer.q_projection.weights.value.grad.to_numpy(),
            dWk_ref=layer.k_projection.weights.value.grad.to_numpy(),
            dWv_ref=layer.v_projection.weights.value.grad.to_numpy(),
        )
        print("mha causal=%s |Y_ref - Y_torch| = %.3g  |dX| = %.3g" % (
            causal, np.abs(out["Y_ref"] - out["Y_torch"]).max(),
            np.abs(out["dX_ref"] - out["dX_torch"]).max()))
        np.savez_compressed(os.path.join(HERE, f"mha_cfg1_{'causal' if causal else 'full'}.npz"), **out)


def gen_softmax():
    """kernel_tests/test_softmax_fw.py:60-72 baseline: nn.softmax(inp + mask, dim=3);
    kernel_tests/test_softmax_bw.py:48-52 baseline formula."""
    rng = np.random.default_rng(3)
    for name, B, H, F, T in [("s1", 3, 8, 17, 40), ("s2", 2, 8, 5, 7), ("s3", 1, 8, 33, 130)]:
        inp = rng.uniform(-1, 1, (B, H, F, T)).astype(datatype)
        valid = rng.integers(1, T + 1, B)
        mask = np.where(np.arange(T)[None, :] < valid[:, None], 0.0, -1e8).astype(datatype)
        x = mt(inp) + mt(np.broadcast_to(mask[:, None, None, :], (B, 1, 1, T)).copy())
        y = minitorch.nn.softmax(x, dim=3)
        dy = rng.uniform(-1, 1, (B, H, F, T)).astype(datatype)
        yv = y.to_numpy()
        # test_softmax_bw.py baseline: soft_inp * (out_grad - sum(out_grad*soft_inp, dim=3))
        yt, dyt = mt(yv), mt(dy)
        tsum = (dyt * yt).sum(dim=3)
        dx = yt * (dyt - tsum)
        np.savez_compressed(os.path.join(HERE, f"softmax_{name}.npz"), inp=inp, mask=mask, y=yv, dy=dy,
                            dx=dx.to_numpy())
        print("softmax", name, yv.shape)


def gen_layernorm():
    """kernel_tests/test_layernorm_fw.py:50-69 and test_layernorm_bw.py:136-161 baselines."""
    rng = np.random.default_rng(4)
    for name, rows, h in [("l1", 96, 32), ("l2", 7, 256), ("l3", 33, 1028)]:
        x = rng.uniform(-1, 1, (rows, h)).astype(datatype)
        g = rng.uniform(-1, 1, (h,)).astype(datatype)
        b = rng.uniform(-1, 1, (h,)).astype(datatype)
        dy = rng.uniform(-1, 1, (rows, h)).astype(datatype)
        xt, gt, bt = mt(x), mt(g), mt(b)
        mean = xt.mean(dim=1)
        var = xt.var(dim=1)
        xhat = (xt - mean) / ((var + 1e-8) ** 0.5)       # kernel_tests/test_layernorm_fw.py baseline
        y = gt * xhat + bt
        (y * mt(dy, grad=False)).sum().backward()
        np.savez_compressed(os.path.join(HERE, f"layernorm_{name}.npz"), x=x, gamma=g, beta=b, dy=dy,
                            y=y.to_numpy(), mean=mean.to_numpy().reshape(rows),
                            var=var.to_numpy().reshape(rows),
                            dx=xt.grad.to_numpy(), dgamma=gt.grad.to_numpy().reshape(h),
                            dbeta=bt.grad.to_numpy().reshape(h))
        print("layernorm", name, x.shape)


def gen_combine():
    """The reference's own CPU implementation of the combine.so surface: FastOps.map / zip / reduce /
    matrix_multiply (minitorch/fast_ops.py:154-353) over its scalar operators (minitorch/operators.py),
    the same function table src/combine.cu:31-117 switches on (ids: cuda_kernel_ops.py:33-52)."""
    from minitorch import operators as ops
    rng = np.random.default_rng(5)
    out = {}
    a = rng.uniform(-2, 2, (3, 4, 5)).astype(datatype)
    a[0, 0, :3] = [0.0, 0.5, -0.5]
    b = rng.uniform(-2, 2, (3, 4, 5)).astype(datatype)
    b[1, 2, :] = a[1, 2, :]                      # exact ties for eq / is_close / lt
    b[2, 0, :] = a[2, 0, :] + 0.005
    pos = np.abs(a) + 0.1                        # domain of log / pow
    brow = rng.uniform(-2, 2, (1, 5)).astype(datatype)   # right-aligned broadcast operand
    nz = np.where(np.abs(a) < 0.1, datatype(0.3), a)     # inv(0) raises under numba (Python semantics)
    out.update(a=a, b=b, pos=pos, brow=brow, nz=nz)
    for name in ("id", "neg", "sigmoid", "relu", "exp", "tanh"):
        out[f"map_{name}"] = FastOps.map(getattr(ops, name))(mt(a, False)).to_numpy()
    out["map_inv"] = FastOps.map(ops.inv)(mt(nz, False)).to_numpy()
    out["map_log"] = FastOps.map(ops.log)(mt(pos, False)).to_numpy()
    # strided input view: map over a permuted tensor
    out["map_neg_perm"] = FastOps.map(ops.neg)(mt(a, False).permute(2, 0, 1)).to_numpy()
    out["zip_inv_back"] = FastOps.zip(ops.inv_back)(mt(nz, False), mt(b, False)).to_numpy()
    out["zipb_inv_back"] = FastOps.zip(ops.inv_back)(mt(nz, False), mt(brow, False)).to_numpy()
    for name in ("add", "mul", "lt", "eq", "relu_back", "is_close", "max"):
        out[f"zip_{name}"] = FastOps.zip(getattr(ops, name))(mt(a, False), mt(b, False)).to_numpy()
        out[f"zipb_{name}"] = FastOps.zip(getattr(ops, name))(mt(a, False), mt(brow, False)).to_numpy()
    out["zip_log_back"] = FastOps.zip(ops.log_back)(mt(pos, False), mt(b, False)).to_numpy()
    out["zip_pow"] = FastOps.zip(ops.pow)(mt(pos, False), mt(b, False)).to_numpy()
    out["zip_add_perm"] = FastOps.zip(ops.add)(mt(a, False).permute(1, 0, 2), mt(b, False).permute(1, 0, 2)).to_numpy()
    for dim in (0, 1, 2):
        out[f"red_add_{dim}"] = FastOps.reduce(ops.add, 0.0)(mt(a, False), dim).to_numpy()
        out[f"red_mul_{dim}"] = FastOps.reduce(ops.mul, 1.0)(mt(a, False), dim).to_numpy()
        out[f"red_max_{dim}"] = FastOps.reduce(ops.max, -1e9)(mt(a, False), dim).to_numpy()
    big = rng.uniform(-1, 1, (2, 700, 3)).astype(datatype)
    out["big"] = big
    out["red_add_big"] = FastOps.reduce(ops.add, 0.0)(mt(big, False), 1).to_numpy()
    # matmul: batched, batch-broadcast (2-D weight), transposed (permuted) operand, ragged tile sizes
    A = rng.standard_normal((4, 37, 70)).astype(datatype)
    Bm = rng.standard_normal((4, 70, 29)).astype(datatype)
    W = rng.standard_normal((70, 29)).astype(datatype)
    Kt = rng.standard_normal((4, 29, 70)).astype(datatype)
    out.update(mm_A=A, mm_B=Bm, mm_W=W, mm_Kt=Kt)
    out["mm_batched"] = FastOps.matrix_multiply(mt(A, False), mt(Bm, False)).to_numpy()
    out["mm_bcast"] = FastOps.matrix_multiply(mt(A, False), mt(W, False)).to_numpy()
    out["mm_transposed"] = FastOps.matrix_multiply(mt(A, False), mt(Kt, False).permute(0, 2, 1)).to_numpy()
    np.savez_compressed(os.path.join(HERE, "combine_ops.npz"), **out)
    print("combine", len(out), "arrays")


def gen_decoder():
    """Config #2's model family at fixture size: the reference's DecoderLM (minitorch/modules_transfomer.py:339-453)
    on its composed CPU path, with the MLE loss of project/run_machine_translation.py:164-192.  Flags are passed by
    KEYWORD (the reference's positional calls mis-route use_flash_attention).  Saves every parameter
    by its named_parameters() name, the logits, the loss and a few gradients."""
    np.random.seed(7)
    n_vocab, n_embd, n_head, n_pos, B = 97, 64, 8, 40, 3
    model = minitorch.DecoderLM(n_vocab=n_vocab, n_embd=n_embd, n_head=n_head, n_positions=n_pos, p_dropout=0.0,
                                ln_eps=1e-5, bias=True, backend=backend)
    rng = np.random.default_rng(11111)
    ids = rng.integers(0, n_vocab, (B, n_pos))
    weights = np.zeros((B, n_pos), dtype=datatype)
    weights[:, n_pos // 2:] = 1.0
    input_ids, labels, lw = ids[:, :-1], ids[:, 1:], weights[:, 1:]
    idx = mt(input_ids, grad=False)
    logits = model(idx=idx)
    bs, l, c = logits.shape
    loss = minitorch.nn.softmax_loss(logits=logits.view(bs * l, c), target=mt(labels, grad=False).view(bs * l))
    w = mt(lw, grad=False).view(bs * l)
    total = (loss * w).sum() / w.sum()
    total.backward()
    out = dict(input_ids=input_ids.astype(np.int64), labels=labels.astype(np.int64), label_token_weights=lw,
               logits=logits.to_numpy(), loss=np.array(total.to_numpy()).reshape(-1)[:1],
               cfg=np.array([n_vocab, n_embd, n_head, n_pos]))
    for name, prm in model.named_parameters():
        out["p:" + name] = prm.value.to_numpy()
        if prm.value.grad is not None and (name.endswith("q_projection.weights") or name.startswith("lm_head")
                                           or name.startswith("token_embeddings") or name.endswith("linear_in.bias")
                                           or name.endswith("out_projection.bias")):
            out["g:" + name] = prm.value.grad.to_numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_small.npz"), **out)
    print("decoder logits", logits.shape, "loss", float(out["loss"][0]), "arrays", len(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "combine":
        gen_combine()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decoder":
        gen_decoder()
        sys.exit(0)
    gen_attention()
    gen_mha()
    gen_softmax()
    gen_layernorm()
    gen_combine()
    gen_decoder()
