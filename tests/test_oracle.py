"""Pin oracle/attention_ref.py against the golden vectors produced by the reference
itself (tests/golden/make_golden.py: reference minitorch composed path on its numba
CPU backend + torch.nn.MultiheadAttention, the reference tests' own oracle)."""
import glob
import os

import numpy as np
import pytest

from oracle import attention_ref as R

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(pattern):
    files = sorted(glob.glob(os.path.join(G, pattern)))
    assert files, pattern
    return files


@pytest.mark.parametrize("path", _load("attn_*.npz"), ids=os.path.basename)
def test_attention_oracle_matches_reference_composed(path):
    z = np.load(path)
    causal = bool(z["causal"])
    km = z["key_mask"] if "key_mask" in z.files else None
    O, m, l = R.attention_fwd(z["Q"], z["K"], z["V"], causal=causal, key_mask=km)
    dQ, dK, dV = R.attention_bwd(z["Q"], z["K"], z["V"], z["dO"], causal=causal, key_mask=km)
    # reference path is fp32 numba; oracle is fp64 -> tolerance = fp32 round-off of the reference
    np.testing.assert_allclose(O, z["O"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(dQ, z["dQ"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(dK, z["dK"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(dV, z["dV"], atol=2e-5, rtol=1e-4)
    # kv_len form of the padding mask is the same function
    if "kv_len" in z.files:
        O2, _, _ = R.attention_fwd(z["Q"], z["K"], z["V"], causal=causal, kv_len=z["kv_len"])
        np.testing.assert_allclose(O2, O, atol=1e-12)
    # (m, l) reproduce the softmax normaliser
    S = R._scores(z["Q"], z["K"], causal, km, None, np.float64)
    np.testing.assert_allclose(m + np.log(l), np.log(np.exp(S - S.max(-1, keepdims=True)).sum(-1)) + S.max(-1),
                               atol=1e-12)


@pytest.mark.parametrize("path", _load("mha_*.npz"), ids=os.path.basename)
def test_mha_oracle_matches_reference_and_torch(path):
    z = np.load(path)
    out = R.mha_fwd_bwd(z["X"], z["Wq"], z["Wk"], z["Wv"], z["Wo"], int(z["n_head"]), bool(z["causal"]))
    for ref in ("Y_ref", "Y_torch"):
        np.testing.assert_allclose(out["Y"], z[ref], atol=1e-5, rtol=1e-5)
    for ref in ("dX_ref", "dX_torch"):
        np.testing.assert_allclose(out["dX"], z[ref], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(out["dWo"], z["dWo_ref"], atol=2e-4, rtol=1e-5)
    np.testing.assert_allclose(out["dWo"], z["dWo_torch"], atol=2e-4, rtol=1e-5)
    for w in ("dWq", "dWk", "dWv"):
        np.testing.assert_allclose(out[w], z[w + "_ref"], atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("path", _load("softmax_*.npz"), ids=os.path.basename)
def test_softmax_oracle(path):
    z = np.load(path)
    y = R.attn_softmax_fw(z["inp"], z["mask"])
    # kernel_tests/test_softmax_fw.py:14 tolerance is 1e-3; the +1e-8 epsilon differs from nn.softmax by <1e-8
    np.testing.assert_allclose(y, z["y"], atol=1e-6, rtol=1e-5)
    dx = R.attn_softmax_bw(z["dy"], z["y"])
    np.testing.assert_allclose(dx, z["dx"], atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("path", _load("layernorm_*.npz"), ids=os.path.basename)
def test_layernorm_oracle(path):
    z = np.load(path)
    y, var, mean = R.layernorm_fw(z["x"], z["gamma"], z["beta"])
    np.testing.assert_allclose(y, z["y"], atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(mean, z["mean"], atol=1e-6)
    np.testing.assert_allclose(var, z["var"] + 1e-8, atol=1e-6)
    dx, dg, db = R.layernorm_bw(z["dy"], z["x"], z["gamma"], z["beta"], var, mean)
    # kernel_tests/test_layernorm_bw.py:22 tolerance atol 1e-3 rtol 1e-2 (double epsilon is far below that)
    np.testing.assert_allclose(dx, z["dx"], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(dg.reshape(-1), z["dgamma"], atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(db.reshape(-1), z["dbeta"], atol=1e-4, rtol=1e-4)


def test_bf16_rounding_helpers():
    x = np.array([1.0, 1.00390625, 1.0078125, -3.14159, 1e-30, 65504.0], dtype=np.float32)
    r = R.round_bf16(x)
    assert r[0] == 1.0 and r[2] == 1.0078125
    assert r[1] == 1.0  # tie -> even
    np.testing.assert_array_equal(R.from_bf16_bits(R.to_bf16_bits(x)), r)


@pytest.mark.parametrize("name", ["attn_cfg1.npz", "attn_cfg1_causal.npz", "attn_n39_causal.npz", "attn_n100.npz"])
def test_numba_port_matches_reference_composed(name):
    """oracle/numba_composed.py (the CPU-baseline port) against vectors made by the real reference."""
    from oracle import numba_composed as NC
    z = np.load(os.path.join(G, name))
    O, dQ, dK, dV = NC.attention_fwd_bwd(z["Q"], z["K"], z["V"], z["dO"], causal=bool(z["causal"]))
    np.testing.assert_allclose(O, z["O"], atol=3e-6, rtol=1e-5)
    np.testing.assert_allclose(dQ, z["dQ"], atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(dK, z["dK"], atol=3e-5, rtol=1e-4)
    np.testing.assert_allclose(dV, z["dV"], atol=3e-5, rtol=1e-4)


def test_combine_oracle_matches_reference_fastops():
    """oracle/combine_ref.py against tests/golden/combine_ops.npz (the reference's own FastOps
    map/zip/reduce/matrix_multiply outputs, generated by tests/golden/make_golden.py)."""
    from oracle import combine_ref as C
    g = np.load(os.path.join(G, "combine_ops.npz"))
    tol = dict(rtol=2e-6, atol=2e-6)
    for name in C.UNARY:
        src = {"log": "pos", "inv": "nz"}.get(name, "a")
        np.testing.assert_allclose(C.tensor_map(C.FN_IDS[name], g[src]), g[f"map_{name}"], **tol)
    np.testing.assert_allclose(C.tensor_map(4, np.transpose(g["a"], (2, 0, 1))), g["map_neg_perm"], **tol)
    for name in C.BINARY:
        a = {"log_back": "pos", "pow": "pos", "inv_back": "nz"}.get(name, "a")
        np.testing.assert_allclose(C.tensor_zip(C.FN_IDS[name], g[a], g["b"]), g[f"zip_{name}"], **tol)
        if f"zipb_{name}" in g.files:
            np.testing.assert_allclose(C.tensor_zip(C.FN_IDS[name], g[a], g["brow"]), g[f"zipb_{name}"], **tol)
    for dim in (0, 1, 2):
        np.testing.assert_allclose(C.tensor_reduce(1, g["a"], dim, 0.0), g[f"red_add_{dim}"], rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(C.tensor_reduce(2, g["a"], dim, 1.0), g[f"red_mul_{dim}"], rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(C.tensor_reduce(16, g["a"], dim, -1e9), g[f"red_max_{dim}"], **tol)
    np.testing.assert_allclose(C.tensor_reduce(1, g["big"], 1, 0.0), g["red_add_big"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(C.matrix_multiply(g["mm_A"], g["mm_B"]), g["mm_batched"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(C.matrix_multiply(g["mm_A"], g["mm_W"]), g["mm_bcast"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(C.matrix_multiply(g["mm_A"], np.swapaxes(g["mm_Kt"], 1, 2)), g["mm_transposed"],
                               rtol=1e-5, atol=2e-5)
