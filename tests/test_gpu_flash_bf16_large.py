"""bf16 tensor-core path at the LARGE ends of BASELINE's configs, where a full fp64 oracle is out of reach:
config #3's top end and config #5's sequence length (N = 8192, head_dim 64 and 128, causal and not), config #4 at
full size with a causal mask on top of its key padding, and a config-#5-shaped shard on every visible GPU.

Checks that need no O(N^2) oracle:
  * sampled QUERY rows against the fp64 composed formula (O, LSE and dQ of a row depend on that row only);
  * sampled KEY rows of dK / dV: column k of P needs every row's LSE, which is taken from the kernel's own (m, l)
    -- pinned by the sampled-row check above -- and D = rowsum(dO * O) from the kernel's O;
  * identities over EVERY element:  sum_k dV[k,:] = sum_q dO[q,:]  (non-causal: each row of P sums to 1), and
    sum dQ*Q = sum dK*K  per (batch, head) (both equal scale * sum dS*S);
  * exact zeros for padded keys.
Tolerance: 2e-2 max-abs (north_star, bf16) plus one bf16 ulp of the reference value.
"""
import numpy as np
import pytest

import flashattn_b200 as fb
from flashattn_b200 import device as dev
from oracle import attention_ref as R

pytestmark = pytest.mark.gpu
TOL = 2e-2
BF16_EPS = 2.0 ** -8


def _f64(x):
    return x.astype(np.float64)


def _check_slice(Q, K, V, dO, O, m, l, gq, gk, gv, causal, n_keys, rng, n_rows=14, n_cols=10):
    """All arguments are (N, d) / (N,) arrays of ONE (batch, head); n_keys = number of valid keys."""
    N, d = Q.shape
    sc = 1.0 / np.sqrt(d)
    q64, k64, v64, do64 = _f64(Q), _f64(K), _f64(V), _f64(dO)
    rows = np.unique(np.concatenate([[0, 127, 128, N - 1], rng.integers(0, N, n_rows)]))
    S = (q64[rows] @ k64[:n_keys].T) * sc
    if causal:
        S = np.where(np.arange(n_keys)[None, :] > rows[:, None], -np.inf, S)
    mx = S.max(axis=1, keepdims=True)
    P = np.exp(S - mx)
    lsum = P.sum(axis=1, keepdims=True)
    P /= lsum
    Oe = P @ v64[:n_keys]
    assert np.abs(O[rows] - Oe).max() < TOL
    lse_k = _f64(m) + np.log(_f64(l))
    assert np.abs(lse_k[rows] - (mx + np.log(lsum))[:, 0]).max() < 2e-3
    dS = P * (do64[rows] @ v64[:n_keys].T - (do64[rows] * Oe).sum(axis=1, keepdims=True))
    dQe = (dS @ k64[:n_keys]) * sc
    assert np.all(np.abs(gq[rows] - dQe) <= TOL + BF16_EPS * np.abs(dQe)), float(np.abs(gq[rows] - dQe).max())
    # key rows: column k of P from the kernel's LSE, D from the kernel's O
    cols = np.unique(np.concatenate([[0, min(127, n_keys - 1), n_keys - 1], rng.integers(0, n_keys, n_cols)]))
    Dv = (do64 * _f64(O)).sum(axis=1)
    Sc = (q64 @ k64[cols].T) * sc                       # (N, ncols)
    if causal:
        Sc = np.where(cols[None, :] > np.arange(N)[:, None], -np.inf, Sc)
    Pc = np.exp(Sc - lse_k[:, None])
    dVe = Pc.T @ do64
    dSc = Pc * (do64 @ v64[cols].T - Dv[:, None])
    dKe = (dSc.T @ q64) * sc
    assert np.all(np.abs(gv[cols] - dVe) <= TOL + BF16_EPS * np.abs(dVe)), float(np.abs(gv[cols] - dVe).max())
    assert np.all(np.abs(gk[cols] - dKe) <= TOL + BF16_EPS * np.abs(dKe)), float(np.abs(gk[cols] - dKe).max())


def _identities(Q, K, dO, gq, gk, gv, causal):
    """Over every element, per (batch, head)."""
    if not causal:
        sum_dv, sum_do = _f64(gv).sum(axis=2), _f64(dO).sum(axis=2)
        bound = BF16_EPS * (np.abs(_f64(gv)).sum(axis=2) + np.abs(_f64(dO)).sum(axis=2))
        assert np.all(np.abs(sum_dv - sum_do) <= bound)
    lhs, rhs = (_f64(gq) * _f64(Q)).sum(axis=(2, 3)), (_f64(gk) * _f64(K)).sum(axis=(2, 3))
    scale_ = np.abs(_f64(gq) * _f64(Q)).sum(axis=(2, 3)) + np.abs(_f64(gk) * _f64(K)).sum(axis=(2, 3))
    assert np.all(np.abs(lhs - rhs) <= 2 * BF16_EPS * scale_)


def _run(B, H, N, d, causal, kv, seed):
    rng = np.random.default_rng(seed)
    Q, K, V, dO = (R.round_bf16(rng.standard_normal((B, H, N, d), dtype=np.float32)) for _ in range(4))
    dq, dk, dv, ddo = (dev.DeviceArray.from_numpy(x, "bf16") for x in (Q, K, V, dO))
    dkv = dev.DeviceArray.from_numpy(kv) if kv is not None else None
    O, m, l = dev.flash_fwd(dq, dk, dv, causal=causal, kv_len=dkv)
    gq, gk, gv = dev.flash_bwd(dq, dk, dv, O, ddo, m, l, causal=causal, kv_len=dkv)
    return (Q, K, V, dO) + tuple(t.to_numpy() for t in (O, m, l, gq, gk, gv)), rng


@pytest.mark.timeout(600)
@pytest.mark.parametrize("d", [128, 64])
@pytest.mark.parametrize("causal", [False, True])
def test_bf16_n8192(d, causal):
    """Config #3's longest sequence / config #5's sequence length, both head dims, causal and not."""
    B, H, N = 1, 3, 8192
    (Q, K, V, dO, O, m, l, gq, gk, gv), rng = _run(B, H, N, d, causal, None, 100 + d + int(causal))
    for h in (0, H - 1):
        _check_slice(Q[0, h], K[0, h], V[0, h], dO[0, h], O[0, h], m[0, h], l[0, h], gq[0, h], gk[0, h], gv[0, h],
                     causal, N, rng)
    _identities(Q, K, dO, gq, gk, gv, causal)


@pytest.mark.timeout(900)
def test_bf16_full_cfg4_causal_with_padding():
    """BASELINE config #4 at full size (B=8, 32 heads, N=4096, d=128) with BOTH masks: causal + key padding."""
    B, H, N, d = 8, 32, 4096, 128
    kv = np.random.default_rng(45).integers(N // 2, N + 1, B).astype(np.int32)
    (Q, K, V, dO, O, m, l, gq, gk, gv), rng = _run(B, H, N, d, True, kv, 46)
    for b, h in ((0, 0), (5, 11), (B - 1, H - 1)):
        n = int(kv[b])
        _check_slice(Q[b, h], K[b, h], V[b, h], dO[b, h], O[b, h], m[b, h], l[b, h], gq[b, h], gk[b, h], gv[b, h],
                     True, n, rng, n_rows=10, n_cols=6)
        assert not gk[b, :, n:].any() and not gv[b, :, n:].any()      # padded keys: exactly zero gradients
    _identities(Q, K, dO, gq, gk, gv, True)


@pytest.mark.timeout(900)
def test_cfg5_shaped_shard_on_every_gpu():
    """Config #5 geometry (32 heads, N=8192, d=128) as a small batch shard on EVERY visible device of this process
    (fa_set_device): the per-device state of the library (pools, streams) must follow the current device."""
    lib = fb._lib.load("flashattention_kernel")
    ndev = int(lib.fa_device_count())
    if ndev < 2:
        pytest.skip("needs at least two GPUs in one process")
    try:
        for g in range(min(ndev, 8)):
            fb._lib.check(lib, lib.fa_set_device(g))
            B, H, N, d = 1, 32, 8192, 128
            (Q, K, V, dO, O, m, l, gq, gk, gv), rng = _run(B, H, N, d, bool(g & 1), None, 500 + g)
            h = (7 * g) % H
            _check_slice(Q[0, h], K[0, h], V[0, h], dO[0, h], O[0, h], m[0, h], l[0, h], gq[0, h], gk[0, h], gv[0, h],
                         bool(g & 1), N, rng, n_rows=8, n_cols=6)
    finally:
        lib.fa_set_device(0)
