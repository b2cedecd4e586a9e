"""CPU baseline "port": the reference's composed attention on its numba CPU backend, restated.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/attention_ref.py header for who may import it).

The reference's CPU path is minitorch's ``FastOps`` (minitorch/fast_ops.py:33-353,
``njit(parallel=True)`` map / zip / reduce / 3-D matmul) driving the composed attention of
minitorch/modules_transfomer.py:177-192 and minitorch/nn.py:104-123: every op is a separate
parallel pass that materialises a (B*H, N, N) fp32 buffer.  minitorch itself cannot travel to
the GPU box, so this module re-creates the same SEQUENCE OF PASSES with numba:

  forward : S = Q@K^T (matmul, one dot product per output, fast_ops.py:336-349 structure)
            S /= sqrt(d) (map) ; S += mask (zip, causal only) ; m = max(S) (reduce)
            E = exp(S - m) (zip + map) ; Z = sum(E) (reduce) ; P = E / Z (zip) ; O = P@V (matmul)
  backward: the passes minitorch's autodiff emits for that graph (matmul x4, the softmax
            quotient/exp/sum/max backward as zips, maps and reduces).

It uses contiguous indexing instead of FastOps' generic strided index arithmetic, so it is, if
anything, FASTER than the reference's own CPU path -- a conservative baseline.  Validated
against tests/golden (made by the real reference) in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
from numba import njit, prange

F32_MAX = np.float32(np.finfo(np.float32).max)


@njit(parallel=True, fastmath=True, cache=True)
def _bmm_nt(a, b, out):  # out[g,i,j] = sum_k a[g,i,k] * b[g,j,k]
    G, M, K = a.shape
    N = b.shape[1]
    for g in prange(G):
        for i in prange(M):
            for j in prange(N):
                acc = np.float32(0.0)
                for k in range(K):
                    acc += a[g, i, k] * b[g, j, k]
                out[g, i, j] = acc


@njit(parallel=True, fastmath=True, cache=True)
def _bmm_nn(a, b, out):  # out[g,i,j] = sum_k a[g,i,k] * b[g,k,j]
    G, M, K = a.shape
    N = b.shape[2]
    for g in prange(G):
        for i in prange(M):
            for j in prange(N):
                acc = np.float32(0.0)
                for k in range(K):
                    acc += a[g, i, k] * b[g, k, j]
                out[g, i, j] = acc


@njit(parallel=True, fastmath=True, cache=True)
def _bmm_tn(a, b, out):  # out[g,i,j] = sum_k a[g,k,i] * b[g,k,j]
    G, K, M = a.shape
    N = b.shape[2]
    for g in prange(G):
        for i in prange(M):
            for j in prange(N):
                acc = np.float32(0.0)
                for k in range(K):
                    acc += a[g, k, i] * b[g, k, j]
                out[g, i, j] = acc


@njit(parallel=True, cache=True)
def _map_scale(x, c, out):
    f, o = x.reshape(-1), out.reshape(-1)
    for i in prange(f.size):
        o[i] = f[i] * c


@njit(parallel=True, cache=True)
def _zip_add_causal(x, out):  # x + (-finfo.max) * triu(ones, 1)
    G, N, M = x.shape
    for g in prange(G):
        for i in range(N):
            for j in range(M):
                out[g, i, j] = x[g, i, j] + (-F32_MAX if j > i else np.float32(0.0))


@njit(parallel=True, cache=True)
def _reduce_max(x, out):
    G, N, M = x.shape
    for r in prange(G * N):
        g, i = r // N, r % N
        acc = x[g, i, 0]
        for j in range(1, M):
            acc = max(acc, x[g, i, j])
        out[g, i] = acc


@njit(parallel=True, cache=True)
def _zip_sub_bcast(x, v, out):
    G, N, M = x.shape
    for r in prange(G * N):
        g, i = r // N, r % N
        for j in range(M):
            out[g, i, j] = x[g, i, j] - v[g, i]


@njit(parallel=True, cache=True)
def _map_exp(x, out):
    f, o = x.reshape(-1), out.reshape(-1)
    for i in prange(f.size):
        o[i] = np.exp(f[i])


@njit(parallel=True, cache=True)
def _reduce_sum(x, out):
    G, N, M = x.shape
    for r in prange(G * N):
        g, i = r // N, r % N
        acc = np.float32(0.0)
        for j in range(M):
            acc += x[g, i, j]
        out[g, i] = acc


@njit(parallel=True, cache=True)
def _zip_div_bcast(x, v, out):
    G, N, M = x.shape
    for r in prange(G * N):
        g, i = r // N, r % N
        for j in range(M):
            out[g, i, j] = x[g, i, j] / v[g, i]


@njit(parallel=True, cache=True)
def _zip_mul(x, y, out):
    f, h, o = x.reshape(-1), y.reshape(-1), out.reshape(-1)
    for i in prange(f.size):
        o[i] = f[i] * h[i]


@njit(parallel=True, cache=True)
def _softmax_bw_tail(P, dP, rowdot, out):  # dS = P * (dP - rowdot)
    G, N, M = P.shape
    for r in prange(G * N):
        g, i = r // N, r % N
        for j in range(M):
            out[g, i, j] = P[g, i, j] * (dP[g, i, j] - rowdot[g, i])


def attention_fwd_bwd(Q, K, V, dO=None, causal=False):
    """Composed attention on (B,H,N,d) fp32 arrays; returns (O, dQ, dK, dV) (grads None if dO is None)."""
    B, H, N, d = Q.shape
    G = B * H
    q, k, v = (np.ascontiguousarray(x, dtype=np.float32).reshape(G, N, d) for x in (Q, K, V))
    S = np.empty((G, N, N), np.float32)
    _bmm_nt(q, k, S)
    T1 = np.empty_like(S)
    _map_scale(S, np.float32(1.0 / np.sqrt(d)), T1)
    if causal:
        T2 = np.empty_like(S)
        _zip_add_causal(T1, T2)
        T1 = T2
    mx = np.empty((G, N), np.float32)
    _reduce_max(T1, mx)
    T3 = np.empty_like(S)
    _zip_sub_bcast(T1, mx, T3)
    E = np.empty_like(S)
    _map_exp(T3, E)
    Z = np.empty((G, N), np.float32)
    _reduce_sum(E, Z)
    P = np.empty_like(S)
    _zip_div_bcast(E, Z, P)
    O = np.empty((G, N, d), np.float32)
    _bmm_nn(P, v, O)
    if dO is None:
        return O.reshape(B, H, N, d), None, None, None
    do = np.ascontiguousarray(dO, dtype=np.float32).reshape(G, N, d)
    dV = np.empty((G, N, d), np.float32)
    _bmm_tn(P, do, dV)
    dP = np.empty_like(S)
    _bmm_nt(do, v, dP)
    PdP = np.empty_like(S)
    _zip_mul(P, dP, PdP)
    rowdot = np.empty((G, N), np.float32)
    _reduce_sum(PdP, rowdot)
    dS = np.empty_like(S)
    _softmax_bw_tail(P, dP, rowdot, dS)
    dSs = np.empty_like(S)
    _map_scale(dS, np.float32(1.0 / np.sqrt(d)), dSs)
    dQ = np.empty((G, N, d), np.float32)
    _bmm_nn(dSs, k, dQ)
    dK = np.empty((G, N, d), np.float32)
    _bmm_tn(dSs, q, dK)
    shp = (B, H, N, d)
    return O.reshape(shp), dQ.reshape(shp), dK.reshape(shp), dV.reshape(shp)
