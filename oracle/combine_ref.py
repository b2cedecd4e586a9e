"""CPU oracle for the map / zip / reduce / matmul plumbing row (TEST INFRASTRUCTURE, not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Restates, in numpy (float64 internally, cast to float32 at the end):
* the function table of ``/root/reference/src/combine.cu:31-117`` (ids from
  ``minitorch/cuda_kernel_ops.py:33-52``; scalar definitions ``minitorch/operators.py:12-150``);
* ``tensorMap`` / ``tensorZip`` right-aligned broadcasting (``src/combine.cu:119-146,201-311``,
  ``minitorch/tensor_data.py`` ``broadcast_index``);
* ``tensorReduce``: fold along one dimension starting from ``reduce_value`` (``src/combine.cu:255-311``);
* ``MatrixMultiply``: batched (B,m,n) @ (B,n,p) with a batch of one broadcast (``src/combine.cu:148-199``).

Pinning: ``tests/golden/combine_ops.npz`` was produced by ``tests/golden/make_golden.py`` from the
reference's own ``FastOps`` map/zip/reduce/matrix_multiply (``minitorch/fast_ops.py:154-353``) run on
its scalar ``operators``; ``tests/test_oracle.py`` checks every function here against it.
"""
from __future__ import annotations

import numpy as np

EPS = 1e-6  # minitorch/operators.py:107

FN_IDS = {"add": 1, "mul": 2, "id": 3, "neg": 4, "lt": 5, "eq": 6, "sigmoid": 7, "relu": 8, "relu_back": 9,
          "log": 10, "log_back": 11, "exp": 12, "inv": 13, "inv_back": 14, "is_close": 15, "max": 16, "pow": 17,
          "tanh": 18}
UNARY = ("id", "neg", "sigmoid", "relu", "log", "exp", "inv", "tanh")
BINARY = ("add", "mul", "lt", "eq", "relu_back", "log_back", "inv_back", "is_close", "max", "pow")


def apply_fn(fn_id: int, x, y=0.0):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    with np.errstate(all="ignore"):
        if fn_id == 1: return x + y
        if fn_id == 2: return x * y
        if fn_id == 3: return x + 0 * y
        if fn_id == 4: return -x + 0 * y
        if fn_id == 5: return (x < y).astype(np.float64)
        if fn_id == 6: return (x == y).astype(np.float64)
        if fn_id == 7: return np.where(x >= 0, 1.0 / (1.0 + np.exp(-x)), np.exp(x) / (1.0 + np.exp(x))) + 0 * y
        if fn_id == 8: return np.maximum(x, 0.0) + 0 * y
        if fn_id == 9: return np.where(x > 0, y, 0.0)
        if fn_id == 10: return np.log(x + EPS) + 0 * y
        if fn_id == 11: return y / (x + EPS)
        if fn_id == 12: return np.exp(x) + 0 * y
        if fn_id == 13: return 1.0 / x + 0 * y
        if fn_id == 14: return -(1.0 / (x * x)) * y
        if fn_id == 15: return ((x - y < 1e-2) & (y - x < 1e-2)).astype(np.float64)
        if fn_id == 16: return np.where(x > y, x, y)
        if fn_id == 17: return np.power(x, y)
        if fn_id == 18: return np.tanh(x) + 0 * y
    return x + y  # src/combine.cu:113 default


def tensor_map(fn_id: int, a, out_shape=None):
    r = apply_fn(fn_id, a)
    if out_shape is not None:
        r = np.broadcast_to(r, out_shape)
    return np.ascontiguousarray(r, dtype=np.float32)


def tensor_zip(fn_id: int, a, b):
    return np.ascontiguousarray(apply_fn(fn_id, a, b), dtype=np.float32)


def tensor_reduce(fn_id: int, a, dim: int, start: float):
    a = np.asarray(a, dtype=np.float64)
    acc = np.full(a.shape[:dim] + (1,) + a.shape[dim + 1:], float(start), dtype=np.float64)
    for s in range(a.shape[dim]):
        acc = apply_fn(fn_id, acc, np.take(a, [s], axis=dim))
    return acc.astype(np.float32)


def matrix_multiply(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.matmul(a, b).astype(np.float32)
