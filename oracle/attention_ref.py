"""CPU oracle for the fused-attention path (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``llmsys-project-flashattn_b200/``) never imports anything under ``oracle/``.

What it restates (citations relative to /root/reference):

* composed attention, the parity target named by BASELINE.json's north_star:
  ``minitorch/modules_transfomer.py:177-192`` --
  ``softmax((q @ kT) / sqrt(d) [+ causal mask], dim=3) @ v`` with the causal mask
  ``-finfo(float32).max * triu(ones, 1)`` (``modules_transfomer.py:63-71``) and
  the max-subtracted softmax of ``minitorch/nn.py:104-123`` (no epsilon).
* the analytic gradients that minitorch's autodiff produces for that graph
  (``minitorch/tensor_functions.py`` MatMul/Exp/Sum/Mul backward); written here in
  closed form: dV = P^T dO, dP = dO V^T, D = rowsum(dO*O), dS = P*(dP-D),
  dQ = c dS K, dK = c dS^T Q.
* the LightSeq-derived fused softmax (``src/softmax_kernel.cu:36-122,309-341``):
  additive key mask (B,to_len), optional future mask, ``1/(sum + 1e-8)``;
  backward ``dx = y * (dy - sum(y*dy))``.
* the fused layernorm (``src/layernorm_kernel.cu:37-98,193-368``): stored ``vars``
  already contain ``+1e-8`` and the backward adds ``1e-8`` again.

Pinning: tests/golden/*.npz were produced by ``tests/golden/make_golden.py`` which
imports the *reference itself* (minitorch on its numba CPU backend, and the
reference tests' own oracle ``torch.nn.MultiheadAttention``) in the build
container; ``tests/test_oracle.py`` checks every function below against those
vectors.  The reference's CUDA flash kernel does not compile so it
pins nothing.
"""
from __future__ import annotations

import numpy as np

F32_MAX = float(np.finfo(np.float32).max)
SOFTMAX_EPS = 1e-8  # src/softmax_kernel.cu:12
LN_EPS = 1e-8       # src/layernorm_kernel.cu:12


def _scores(Q, K, causal, key_mask, kv_len, dtype):
    """S = (Q K^T)/sqrt(d) + masks, in `dtype`.  Shapes (B,H,N,d) -> (B,H,N,N)."""
    Q = np.asarray(Q, dtype=dtype)
    K = np.asarray(K, dtype=dtype)
    B, H, N, d = Q.shape
    # modules_transfomer.py:181-190: matmul, then divide by d**0.5, then add mask
    S = np.matmul(Q, np.swapaxes(K, -1, -2)) / dtype(d ** 0.5)
    if causal:
        # modules_transfomer.py:63-71 (additive -finfo.max above the diagonal)
        S = S + dtype(-F32_MAX) * np.triu(np.ones((N, N), dtype=dtype), 1)
    if key_mask is not None:
        # src/softmax_kernel.cu:27-33: (B, to_len) additive, broadcast over heads/rows
        S = S + np.asarray(key_mask, dtype=dtype)[:, None, None, :]
    if kv_len is not None:
        kv_len = np.asarray(kv_len).reshape(B)
        pad = np.arange(N)[None, :] >= kv_len[:, None]          # (B, N)
        S = np.where(pad[:, None, None, :], dtype(-np.inf), S)
    return S


def attention_fwd(Q, K, V, causal=False, key_mask=None, kv_len=None, dtype=np.float64):
    """Return (O, m, l): O = softmax(S) V; m = rowmax(S); l = sum exp(S - m).

    (m, l) follow src/flashattention_kernel.cu:81-89 (row max of the *scaled*
    scores and the sum of exp(s - m)); LSE = m + log(l)."""
    S = _scores(Q, K, causal, key_mask, kv_len, dtype)
    m = S.max(axis=-1)
    E = np.exp(S - m[..., None])            # nn.py:121
    l = E.sum(axis=-1)                      # nn.py:122
    P = E / l[..., None]                    # nn.py:123
    O = np.matmul(P, np.asarray(V, dtype=dtype))
    return O, m, l


def attention_bwd(Q, K, V, dO, causal=False, key_mask=None, kv_len=None, dtype=np.float64):
    """Return (dQ, dK, dV) of sum(O * dO) w.r.t. Q, K, V for the composed graph."""
    Q = np.asarray(Q, dtype=dtype)
    K = np.asarray(K, dtype=dtype)
    V = np.asarray(V, dtype=dtype)
    dO = np.asarray(dO, dtype=dtype)
    d = Q.shape[-1]
    c = dtype(1.0) / dtype(d ** 0.5)
    S = _scores(Q, K, causal, key_mask, kv_len, dtype)
    m = S.max(axis=-1, keepdims=True)
    E = np.exp(S - m)
    P = E / E.sum(axis=-1, keepdims=True)
    dV = np.matmul(np.swapaxes(P, -1, -2), dO)
    dP = np.matmul(dO, np.swapaxes(V, -1, -2))
    D = (P * dP).sum(axis=-1, keepdims=True)
    dS = P * (dP - D)
    dQ = c * np.matmul(dS, K)
    dK = c * np.matmul(np.swapaxes(dS, -1, -2), Q)
    return dQ, dK, dV


def round_bf16(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what the bf16 mode feeds the MMA)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    out = r.astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if nan.any():
        out = out.copy()
        out[nan] = np.nan
    return out.reshape(x.shape)


def to_bf16_bits(x):
    """fp32 array -> uint16 array holding the bf16 bit patterns (RNE)."""
    return (round_bf16(x).view(np.uint32) >> 16).astype(np.uint16)


def from_bf16_bits(u):
    return (np.asarray(u, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


# --------------------------------------------------------------------------- #
# fused softmax (LightSeq semantics)
# --------------------------------------------------------------------------- #
def attn_softmax_fw(inp, attn_mask=None, mask_future=False, dtype=np.float64):
    """src/softmax_kernel.cu:36-122: y = exp(x+mask - max) / (sum + 1e-8).

    inp (B,H,from,to); attn_mask (B,to) additive or None; mask_future masks j > i."""
    x = np.asarray(inp, dtype=dtype)
    B, H, F, T = x.shape
    if attn_mask is not None:
        x = x + np.asarray(attn_mask, dtype=dtype).reshape(B, 1, 1, T)
    if mask_future:
        fut = np.arange(T)[None, :] > np.arange(F)[:, None]
        x = np.where(fut[None, None], dtype(-np.inf), x)
    mx = x.max(axis=-1, keepdims=True)
    e = np.exp(x - mx)
    return e / (e.sum(axis=-1, keepdims=True) + dtype(SOFTMAX_EPS))


def attn_softmax_bw(out_grad, soft_out, dtype=np.float64):
    """src/softmax_kernel.cu:309-341: dx = y * (dy - sum_j y*dy)."""
    dy = np.asarray(out_grad, dtype=dtype)
    y = np.asarray(soft_out, dtype=dtype)
    s = (dy * y).sum(axis=-1, keepdims=True)
    return y * (dy - s)


# --------------------------------------------------------------------------- #
# fused layernorm (LightSeq semantics incl. the double epsilon)
# --------------------------------------------------------------------------- #
def layernorm_fw(inp, gamma, beta, dtype=np.float64):
    """src/layernorm_kernel.cu:37-98. Returns (y, vars(+eps), means)."""
    x = np.asarray(inp, dtype=dtype)
    g = np.asarray(gamma, dtype=dtype).reshape(-1)
    b = np.asarray(beta, dtype=dtype).reshape(-1)
    mean = x.mean(axis=1)
    var = (x * x).mean(axis=1) - mean * mean + dtype(LN_EPS)      # :70
    y = g[None, :] * ((x - mean[:, None]) / np.sqrt(var)[:, None]) + b[None, :]
    return y, var, mean


def layernorm_bw(out_grad, inp, gamma, beta, var, mean, dtype=np.float64):
    """src/layernorm_kernel.cu:193-368. Returns (dx, dgamma(1,h), dbeta(1,h))."""
    dy = np.asarray(out_grad, dtype=dtype)
    x = np.asarray(inp, dtype=dtype)
    g = np.asarray(gamma, dtype=dtype).reshape(-1)
    var = np.asarray(var, dtype=dtype).reshape(-1)
    mean = np.asarray(mean, dtype=dtype).reshape(-1)
    h = x.shape[1]
    sd = np.sqrt(var + dtype(LN_EPS))[:, None]                    # :229, :310
    xhat = (x - mean[:, None]) / sd
    dbeta = dy.sum(axis=0, keepdims=True)
    dgamma = (dy * xhat).sum(axis=0, keepdims=True)
    dxhat = dy * g[None, :]
    s1 = dxhat.sum(axis=1, keepdims=True) / (h * sd)
    s2 = (dxhat * xhat).sum(axis=1, keepdims=True) / (h * sd)
    dx = dxhat / sd - s1 - xhat * s2                              # :348-362
    return dx, dgamma, dbeta


# --------------------------------------------------------------------------- #
# MultiHeadAttention (module-level restatement used by the MHA parity tests)
# --------------------------------------------------------------------------- #
def mha_fwd_bwd(X, Wq, Wk, Wv, Wo, n_head, causal, dY=None, dtype=np.float64):
    """minitorch/modules_transfomer.py:73-107,204-229 with bias=False, p_dropout=0.

    X (B,N,E); W* are (E_in,E_out) as minitorch's Linear stores them
    (modules_basic.py:107-150: y = x @ W).  Returns dict with Y, dX, dWq.. ."""
    X = np.asarray(X, dtype=dtype)
    Wq, Wk, Wv, Wo = (np.asarray(w, dtype=dtype) for w in (Wq, Wk, Wv, Wo))
    B, N, E = X.shape
    d = E // n_head
    x2 = X.reshape(B * N, E)

    def split(t):
        return t.reshape(B, N, n_head, d).transpose(0, 2, 1, 3)

    q, k, v = split(x2 @ Wq), split(x2 @ Wk), split(x2 @ Wv)
    O, _, _ = attention_fwd(q, k, v, causal=causal, dtype=dtype)
    A = O.transpose(0, 2, 1, 3).reshape(B * N, E)
    Y = (A @ Wo).reshape(B, N, E)
    out = {"Y": Y, "q": q, "k": k, "v": v, "O": O}
    if dY is None:
        dY = np.ones_like(Y)                 # result.sum().backward()
    dY2 = np.asarray(dY, dtype=dtype).reshape(B * N, E)
    dWo = A.T @ dY2
    dA = dY2 @ Wo.T
    dO = dA.reshape(B, N, n_head, d).transpose(0, 2, 1, 3)
    dq, dk, dv = attention_bwd(q, k, v, dO, causal=causal, dtype=dtype)

    def merge(t):
        return t.transpose(0, 2, 1, 3).reshape(B * N, E)

    dq2, dk2, dv2 = merge(dq), merge(dk), merge(dv)
    out.update(
        dWo=dWo, dWq=x2.T @ dq2, dWk=x2.T @ dk2, dWv=x2.T @ dv2,
        dX=(dq2 @ Wq.T + dk2 @ Wk.T + dv2 @ Wv.T).reshape(B, N, E), dO=dO,
    )
    return out
