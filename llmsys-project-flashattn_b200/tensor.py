"""Stand-alone host tensor + the four fused autograd Functions of the reference.

The reference's minitorch package cannot travel with this repo, so the parity tests and
the benchmark need a minimal carrier that speaks the same protocol ``CudaKernelOps``
expects from ``minitorch.Tensor`` (minitorch/tensor.py:72-436): host fp32 storage behind
``_tensor._storage`` with shape/strides, ``contiguous``/``permute``/``view``/``zeros``,
``Tensor.make`` and a ``backend`` whose attributes are the fused ops
(minitorch/tensor_ops.py:97-104).  The autograd nodes mirror
minitorch/tensor_functions.py:435-516 (Attn_Softmax, LayerNorm, FlashAttention,
FlashAttentionCausal) with the Attn_Softmax.backward unpacking bug fixed
(SURVEY.md 2.4).  Layout ops here are pure index bookkeeping on host memory; all
arithmetic goes through the CUDA libraries.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .cuda_kernel_ops import CudaKernelOps

datatype = np.float32


class TensorBackend:
    """Binder with the reference's attribute names (minitorch/tensor_ops.py:97-104)."""

    def __init__(self, ops=CudaKernelOps):
        self.ops = ops
        self.cuda = getattr(ops, "cuda", False)
        # storage lives in HBM (DeviceKernelOps) instead of host numpy memory (CudaKernelOps, like the reference)
        self.device = getattr(ops, "device_resident", False)
        if hasattr(ops, "map"):   # map / zip / reduce / matmul over combine.so, same attribute names as the reference
            for name in ("neg", "sigmoid", "relu", "log", "exp", "id", "inv", "tanh"):
                setattr(self, name + "_map", ops.map(name))
            self.id_cmap = self.id_map
            for name in ("add", "mul", "lt", "eq", "is_close", "relu_back", "log_back", "inv_back"):
                setattr(self, name + "_zip", ops.zip(name))
            self.pow_scalar_zip = ops.zip("pow")
            self.max_zip = ops.zip("max")
            self.add_reduce = ops.reduce("add", 0.0)
            self.mul_reduce = ops.reduce("mul", 1.0)
            self.max_reduce = ops.reduce("max", -1e9)
            self.matrix_multiply = ops.matrix_multiply
        for name in ("attn_softmax_fw", "attn_softmax_bw", "layernorm_fw", "layernorm_bw", "flash_attention_fw",
                     "flash_attention_bw", "flash_attention_causal_fw", "flash_attention_causal_bw"):
            setattr(self, name, getattr(ops, name))
        # optional fused lookup / loss ops; callers fall back to the one-hot formulation without them
        for name in ("embedding_fw", "embedding_bw", "softmax_xent_fw", "softmax_xent_bw"):
            if hasattr(ops, name):
                setattr(self, name, getattr(ops, name))


class _Data:
    """Strided view over a flat fp32 array (the slice of minitorch.TensorData the ops touch)."""

    def __init__(self, storage: np.ndarray, shape: Sequence[int], strides: Optional[Sequence[int]] = None):
        self._storage = storage
        self.shape = tuple(int(s) for s in shape)
        if strides is None:
            st, acc = [], 1
            for s in reversed(self.shape):
                st.append(acc)
                acc *= s
            strides = tuple(reversed(st))
        self.strides = tuple(int(s) for s in strides)

    @property
    def _shape(self) -> np.ndarray:      # minitorch.TensorData attribute names; built once (read-only by convention)
        a = self.__dict__.get("_shape_i32")
        if a is None:
            a = self.__dict__["_shape_i32"] = np.ascontiguousarray(np.array(self.shape, dtype=np.int32).reshape(-1))
        return a

    @property
    def _strides(self) -> np.ndarray:
        a = self.__dict__.get("_strides_i32")
        if a is None:
            a = self.__dict__["_strides_i32"] = np.ascontiguousarray(np.array(self.strides, dtype=np.int32).reshape(-1))
        return a

    def is_contiguous(self) -> bool:
        exp = 1
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != exp:
                return False
            exp *= s
        return True

    def view_array(self) -> np.ndarray:
        return np.lib.stride_tricks.as_strided(self._storage, self.shape, tuple(s * 4 for s in self.strides))


class Context:
    def __init__(self):
        self.saved_values: Tuple = ()

    def save_for_backward(self, *values):
        self.saved_values = values


class Function:
    @classmethod
    def apply(cls, *inputs: "HostTensor") -> "HostTensor":
        ctx = Context()
        out = cls.forward(ctx, *[t.detach() if isinstance(t, HostTensor) else t for t in inputs])
        if any(isinstance(t, HostTensor) and t.requires_grad() for t in inputs):
            out._node = (cls, ctx, inputs)
        return out


class HostTensor:
    def __init__(self, data: _Data, backend: Optional[TensorBackend] = None):
        self._tensor = data
        self.backend = backend if backend is not None else default_backend()
        self.f = self.backend
        self.grad: Optional[HostTensor] = None
        self._requires_grad = False
        self._node = None

    # ---- construction -------------------------------------------------------------
    @staticmethod
    def make(storage, shape, strides=None, backend=None) -> "HostTensor":
        backend = backend if backend is not None else default_backend()
        if backend.device:
            st = backend.ops.to_storage(storage)
        else:
            st = np.ascontiguousarray(storage, dtype=datatype).reshape(-1)
        return HostTensor(_Data(st, shape, strides), backend)

    def zeros(self, shape=None) -> "HostTensor":
        shape = self.shape if shape is None else tuple(shape)
        n = int(np.prod(shape)) if len(shape) else 1
        if self.backend.device:
            return HostTensor(_Data(self.backend.ops.zeros_storage(n), shape), self.backend)
        return HostTensor.make(np.zeros(n, dtype=datatype), shape, backend=self.backend)

    def detach(self) -> "HostTensor":
        return HostTensor(self._tensor, self.backend)

    def requires_grad_(self, x: bool = True) -> None:
        self._requires_grad = bool(x)

    def requires_grad(self) -> bool:
        return self._requires_grad or self._node is not None

    # ---- layout -------------------------------------------------------------------
    @property
    def shape(self):
        return self._tensor.shape

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def contiguous(self) -> "HostTensor":
        if self._tensor.is_contiguous():
            return self
        if self.backend.device:
            out = self.f.id_map(self.detach())      # strided read -> fresh contiguous buffer, on the device
        else:
            arr = np.ascontiguousarray(self._tensor.view_array())
            out = HostTensor.make(arr.reshape(-1), self.shape, backend=self.backend)
        if self.requires_grad():
            out._node = (_Contiguous, None, (self,))
        return out

    def permute(self, *order) -> "HostTensor":
        d = self._tensor
        out = HostTensor(_Data(d._storage, [d.shape[i] for i in order], [d.strides[i] for i in order]), self.backend)
        if self.requires_grad():
            out._node = (_Permute, tuple(order), (self,))
        return out

    def view(self, *shape) -> "HostTensor":
        assert self._tensor.is_contiguous(), "view needs a contiguous tensor (minitorch/tensor_functions.py View)"
        out = HostTensor(_Data(self._tensor._storage, shape), self.backend)
        if self.requires_grad():
            out._node = (_View, self.shape, (self,))
        return out

    def to_numpy(self) -> np.ndarray:
        if self.backend.device:
            host = self.backend.ops.storage_to_numpy(self._tensor._storage)
            return np.array(np.lib.stride_tricks.as_strided(host, self.shape, tuple(s * 4 for s in self._tensor.strides)),
                            dtype=datatype)
        return np.array(self._tensor.view_array(), dtype=datatype)

    # ---- fused ops (minitorch/tensor.py:424-436) ------------------------------------
    def attn_softmax(self, mask: "HostTensor") -> "HostTensor":
        return Attn_Softmax.apply(self, mask)

    def layernorm(self, gamma: "HostTensor", beta: "HostTensor") -> "HostTensor":
        return LayerNorm.apply(self, gamma, beta)

    def flash_attention(self, k: "HostTensor", v: "HostTensor") -> "HostTensor":
        return FlashAttention.apply(self, k, v)

    def flash_attention_causal(self, k: "HostTensor", v: "HostTensor") -> "HostTensor":
        return FlashAttentionCausal.apply(self, k, v)

    # ---- arithmetic the MultiHeadAttention call site needs (minitorch/tensor.py:196-260), all through
    # combine.so's map / zip / reduce / MatrixMultiply ------------------------------------------------
    def _lift(self, b) -> "HostTensor":
        """Python scalar -> (1,) tensor.  Constants are cached per backend (on device storage every new scalar
        would otherwise be its own H2D copy); they are never written to."""
        if isinstance(b, HostTensor):
            return b
        cache = self.backend.__dict__.setdefault("_scalar_cache", {})
        key = float(b)
        t = cache.get(key)
        if t is None:
            if len(cache) > 256:
                cache.clear()
            t = cache[key] = tensor_from_numpy(np.array([b], dtype=datatype), backend=self.backend)
        return t

    def __matmul__(self, b: "HostTensor") -> "HostTensor":
        return MatMul.apply(self, b)

    def __add__(self, b) -> "HostTensor":
        return Add.apply(self, self._lift(b))

    def __mul__(self, b) -> "HostTensor":
        return Mul.apply(self, self._lift(b))

    def __truediv__(self, b) -> "HostTensor":
        return Mul.apply(self, Inv.apply(self._lift(b)))

    def __neg__(self) -> "HostTensor":
        return Neg.apply(self)

    def __sub__(self, b) -> "HostTensor":
        return Add.apply(self, Neg.apply(self._lift(b)))

    def __radd__(self, b) -> "HostTensor":
        return self + b

    def __rmul__(self, b) -> "HostTensor":
        return self * b

    def __pow__(self, b) -> "HostTensor":
        return PowerScalar.apply(self, self._lift(b))

    def exp(self) -> "HostTensor":
        return Exp.apply(self)

    def log(self) -> "HostTensor":
        return Log.apply(self)

    def tanh(self) -> "HostTensor":
        return Tanh.apply(self)

    def mean(self, dim: Optional[int] = None) -> "HostTensor":
        return self.sum(dim) / (self.shape[dim] if dim is not None else self.size)

    def var(self, dim: int) -> "HostTensor":
        """Biased variance, keeps the reduced dim (minitorch/tensor.py:248-260)."""
        diff = (self - self.sum(dim) / self.shape[dim]) ** 2
        return diff.sum(dim) / self.shape[dim]

    def sum(self, dim: Optional[int] = None) -> "HostTensor":
        if dim is None:
            return Sum.apply(self.contiguous().view(self.size), 0)
        return Sum.apply(self, dim)

    # ---- autodiff -----------------------------------------------------------------
    def backward(self, grad_output: Optional["HostTensor"] = None) -> None:
        """Reverse-mode sweep over the recorded nodes (minitorch/autodiff.py:130-163)."""
        order, seen = [], set()

        def visit(t):
            if id(t) in seen:
                return
            seen.add(id(t))
            if t._node is not None:
                for inp in t._node[2]:
                    if isinstance(inp, HostTensor):
                        visit(inp)
            order.append(t)

        visit(self)
        if grad_output is None:
            assert self.size == 1, "backward() without a gradient needs a scalar (minitorch/tensor.py:394-407)"
            grad_output = tensor_from_numpy(np.ones(self.shape, dtype=datatype), backend=self.backend)
        grads = {id(self): grad_output}
        for t in reversed(order):
            g = grads.pop(id(t), None)
            if g is None:
                continue
            if t._node is None:
                if t._requires_grad:
                    t.grad = g if t.grad is None else _add(t.grad, g)
                continue
            fn, ctx, inputs = t._node
            inputs = tuple(i for i in inputs if isinstance(i, HostTensor))
            outs = fn.backward(ctx, g)
            if not isinstance(outs, tuple):
                outs = (outs,)
            for inp, gi in zip(inputs, outs):
                if gi is None or not inp.requires_grad():
                    continue
                grads[id(inp)] = gi if id(inp) not in grads else _add(grads[id(inp)], gi)


def _add(a: HostTensor, b: HostTensor) -> HostTensor:
    if a.backend.device:
        return a.f.add_zip(a, b)
    return tensor_from_numpy(a.to_numpy() + b.to_numpy(), backend=a.backend)


class _Contiguous:
    @staticmethod
    def backward(ctx, g):
        return g


class _Permute:
    @staticmethod
    def backward(order, g):
        inv = [0] * len(order)
        for i, o in enumerate(order):
            inv[o] = i
        return g.permute(*inv)


class _View:
    @staticmethod
    def backward(orig_shape, g):
        return g.contiguous().view(*orig_shape)


def _unbroadcast(g: HostTensor, shape) -> HostTensor:
    """Sum a gradient back to the shape of a broadcast operand (minitorch/tensor.py expand :293-326)."""
    shape = tuple(shape)
    if g.shape == shape:
        return g
    pad = len(g.shape) - len(shape)
    for dim, n in enumerate(g.shape):
        if dim < pad or shape[dim - pad] == 1 and n != 1:
            g = g.f.add_reduce(g, dim)
    return g.contiguous().view(*shape)


class QKVPart(Function):
    """One of the three outputs of the fused Q/K/V projection: the forward of all three is ONE GEMM
    (ops.qkv_projection, run by the first part and shared through `call`); each part is its own autograd node whose
    backward is the Linear backward of its weight (dx = g @ W^T, dW = x^T @ g -- tensor-core GEMMs in bf16 GEMM mode)."""

    @staticmethod
    def forward(ctx, x2, w, call, j):
        ctx.save_for_backward(x2, w)
        return call[j]

    @staticmethod
    def backward(ctx, g):
        x2, w = ctx.saved_values
        g = g.contiguous()
        return g.f.matrix_multiply(g, w.permute(1, 0)), g.f.matrix_multiply(x2.permute(1, 0), g)


def fused_qkv(x2, wq, wk, wv):
    """(q2, k2, v2) = x2 @ wq, x2 @ wk, x2 @ wv through one fused GEMM, or None when the backend cannot do it."""
    ops = x2.backend.ops
    if not hasattr(ops, "qkv_projection") or getattr(ops, "gemm_mode", "fp32") != "bf16":
        return None
    call = ops.qkv_projection(x2.detach(), wq.detach(), wk.detach(), wv.detach())
    if call is None:
        return None
    return tuple(QKVPart.apply(x2, w, call, j) for j, w in enumerate((wq, wk, wv)))


class MatMul(Function):
    """minitorch/tensor_functions.py:413-432."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a.f.matrix_multiply(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_values

        def tr(t):
            order = list(range(len(t.shape)))
            order[-2], order[-1] = order[-1], order[-2]
            return t.permute(*order)

        return (_unbroadcast(g.f.matrix_multiply(g, tr(b)), a.shape), _unbroadcast(g.f.matrix_multiply(tr(a), g), b.shape))


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a.shape, b.shape)
        return a.f.add_zip(a, b)

    @staticmethod
    def backward(ctx, g):
        sa, sb = ctx.saved_values
        return _unbroadcast(g, sa), _unbroadcast(g, sb)


class Mul(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a.f.mul_zip(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_values
        return _unbroadcast(g.f.mul_zip(g, b), a.shape), _unbroadcast(g.f.mul_zip(g, a), b.shape)


class Neg(Function):
    @staticmethod
    def forward(ctx, a):
        return a.f.neg_map(a)

    @staticmethod
    def backward(ctx, g):
        return g.f.neg_map(g)


class Inv(Function):
    @staticmethod
    def forward(ctx, a):
        ctx.save_for_backward(a)
        return a.f.inv_map(a)

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_values
        return g.f.inv_back_zip(a, g)


class Exp(Function):
    @staticmethod
    def forward(ctx, a):
        out = a.f.exp_map(a)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_values
        return g.f.mul_zip(g, out)


class Sum(Function):
    @staticmethod
    def forward(ctx, a, dim):
        ctx.save_for_backward(a.shape)
        return a.f.add_reduce(a, int(dim))

    @staticmethod
    def backward(ctx, g):
        (shape,) = ctx.saved_values
        return g.f.add_zip(g, g.zeros(shape))   # broadcast the reduced gradient back over `dim`


class Max(Function):
    """Row max with the reference's argmax gradient (minitorch/nn.py:81-97)."""

    @staticmethod
    def forward(ctx, a, dim):
        out = a.f.max_reduce(a, int(dim))
        ctx.save_for_backward(a, out)
        return out

    @staticmethod
    def backward(ctx, g):
        a, out = ctx.saved_values
        return g.f.mul_zip(g.f.eq_zip(out, a), g)


class PowerScalar(Function):
    """a ** scalar (minitorch/tensor_functions.py:133-183)."""

    @staticmethod
    def forward(ctx, a, scalar):
        ctx.save_for_backward(a, scalar)
        return a.f.pow_scalar_zip(a, scalar)

    @staticmethod
    def backward(ctx, g):
        a, scalar = ctx.saved_values
        f = g.f
        return f.mul_zip(g, f.mul_zip(scalar, f.pow_scalar_zip(a, f.add_zip(scalar, scalar._lift(-1.0))))), None


class Tanh(Function):
    @staticmethod
    def forward(ctx, a):
        out = a.f.tanh_map(a)
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_values
        f = g.f
        return f.mul_zip(g, f.add_zip(f.neg_map(f.mul_zip(out, out)), out._lift(1.0)))


class Log(Function):
    @staticmethod
    def forward(ctx, a):
        ctx.save_for_backward(a)
        return a.f.log_map(a)

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_values
        return g.f.log_back_zip(a, g)


def softmax(x: HostTensor, dim: int) -> HostTensor:
    """Composed softmax, op for op as minitorch/nn.py:104-123 (max-subtracted, no epsilon)."""
    e = (x - Max.apply(x, dim)).exp()
    return e / e.sum(dim)


def logsumexp(x: HostTensor, dim: int) -> HostTensor:
    """minitorch/nn.py:229-246 (keeps the reduced dim)."""
    mx = Max.apply(x, dim)
    return mx + (x - mx).exp().sum(dim).log()


def one_hot(x: HostTensor, num_classes: int) -> HostTensor:
    """minitorch/nn.py:212-222 (an np.eye row lookup there).  Host storage: the rows are written directly;
    device storage: eq_zip of the indices against a class-id row, so the (n, C) matrix is born in HBM."""
    if x.backend.device:
        n = x.size
        idx = x.contiguous().view(n, 1)
        classes = tensor_from_numpy(np.arange(num_classes, dtype=datatype).reshape(1, num_classes), backend=x.backend)
        return x.f.eq_zip(idx, classes).view(*x.shape, num_classes)
    idx = x.to_numpy().astype(np.int64).reshape(-1)
    hot = np.zeros((idx.size, num_classes), dtype=datatype)
    hot[np.arange(idx.size), idx] = 1.0
    return tensor_from_numpy(hot.reshape(*x.shape, num_classes), backend=x.backend)


class EmbeddingLookup(Function):
    """weights[ids] as one gather kernel; same values and gradient as one_hot(ids) @ weights
    (minitorch/modules_basic.py:55-71) without materialising the (tokens, vocab) matrix."""

    @staticmethod
    def forward(ctx, ids, weights):
        ctx.save_for_backward(ids, weights.shape[0])
        return ids.f.embedding_fw(ids, weights)

    @staticmethod
    def backward(ctx, g):
        ids, V = ctx.saved_values
        return None, g.f.embedding_bw(ids, g, V)


class SoftmaxCrossEntropy(Function):
    """logsumexp(logits) - logits[target] per row in one kernel each way (minitorch/nn.py:251-271 composed)."""

    @staticmethod
    def forward(ctx, logits, target):
        loss, lse = logits.f.softmax_xent_fw(logits, target)
        ctx.save_for_backward(logits, target, lse)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, target, lse = ctx.saved_values
        return g.f.softmax_xent_bw(g, logits, target, lse), None


def softmax_loss(logits: HostTensor, target: HostTensor, fused: bool = False) -> HostTensor:
    """Cross entropy with reduction=None (minitorch/nn.py:251-271): (minibatch, C), (minibatch,) -> (minibatch,).
    `fused=True` uses the single-kernel version when the backend has one."""
    if fused and hasattr(logits.f, "softmax_xent_fw"):
        return SoftmaxCrossEntropy.apply(logits, target)
    result = logsumexp(logits, dim=1) - (logits * one_hot(target, logits.shape[1])).sum(dim=1)
    return result.view(logits.shape[0])


def GELU(x: HostTensor) -> HostTensor:
    """tanh-approximated GELU, op for op as minitorch/nn.py:205-209."""
    return 0.5 * x * (1 + (float(np.sqrt(2 / np.pi)) * (x + 0.044715 * (x ** 3))).tanh())


class Attn_Softmax(Function):
    @staticmethod
    def forward(ctx, inp, mask):
        out = inp.f.attn_softmax_fw(inp, mask)
        ctx.save_for_backward(out, mask)   # the softmax OUTPUT is what the backward needs
        return out

    @staticmethod
    def backward(ctx, out_grad):
        soft, _mask = ctx.saved_values
        # the bw kernel works in place: hand it a private contiguous copy of the incoming gradient
        if out_grad.backend.device:
            g = out_grad.f.id_map(out_grad)
        else:
            g = tensor_from_numpy(out_grad.to_numpy(), backend=out_grad.backend)
        g, _ = out_grad.f.attn_softmax_bw(g, soft)
        return g, None


class LayerNorm(Function):
    @staticmethod
    def forward(ctx, inp, gamma, beta):
        ln_res, var, means = inp.f.layernorm_fw(inp, gamma, beta)
        ctx.save_for_backward(inp, gamma, beta, var, means)
        return ln_res

    @staticmethod
    def backward(ctx, out_grad):
        inp, gamma, beta, var, means = ctx.saved_values
        dx, dg, db = out_grad.f.layernorm_bw(out_grad, inp, gamma, beta, var, means)
        return dx, dg.view(gamma.shape[0]), db.view(beta.shape[0])


class FlashAttention(Function):
    @staticmethod
    def forward(ctx, Q, K, V):
        O, m, l = Q.f.flash_attention_fw(Q, K, V)
        ctx.save_for_backward(Q, K, V, O, m, l)
        return O

    @staticmethod
    def backward(ctx, out_grad):
        Q, K, V, O, m, l = ctx.saved_values
        return out_grad.f.flash_attention_bw(Q, K, V, O, out_grad, m, l)


class FlashAttentionCausal(Function):
    @staticmethod
    def forward(ctx, Q, K, V):
        O, m, l = Q.f.flash_attention_causal_fw(Q, K, V)
        ctx.save_for_backward(Q, K, V, O, m, l)
        return O

    @staticmethod
    def backward(ctx, out_grad):
        Q, K, V, O, m, l = ctx.saved_values
        return out_grad.f.flash_attention_causal_bw(Q, K, V, O, out_grad, m, l)


_default_backend = None


def default_backend() -> TensorBackend:
    global _default_backend
    if _default_backend is None:
        _default_backend = TensorBackend(CudaKernelOps)
    return _default_backend


def tensor_from_numpy(arr, backend: Optional[TensorBackend] = None, requires_grad: bool = False) -> HostTensor:
    """Same name/meaning as minitorch/tensor_functions.py:632."""
    arr = np.ascontiguousarray(arr, dtype=datatype)
    t = HostTensor.make(arr.reshape(-1).copy(), arr.shape, backend=backend)
    t.requires_grad_(requires_grad)
    return t
