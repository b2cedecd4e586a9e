"""``CudaKernelOps`` -- the operator surface of the reference's minitorch/cuda_kernel_ops.py
for the fused ops, backed by the B200 libraries.

Same static-method names, argument order and return tuples as the reference
(minitorch/cuda_kernel_ops.py: attn_softmax_fw :440, attn_softmax_bw :471, layernorm_fw
:498, layernorm_bw :543, flash_attention_fw :606, flash_attention_bw :677,
flash_attention_causal_fw :761, flash_attention_causal_bw :811), so
``TensorBackend(CudaKernelOps)`` (minitorch/tensor_ops.py:97-104) binds them unchanged.
Differences, all deliberate:
  * no pycuda / torch imports (the reference only used them to fetch a stream handle);
  * every operand is made contiguous (the reference forgot O and dO, :745,:749);
  * a failing launch raises ``FlashAttnError`` instead of exit()/silent return;
  * optional ``key_mask`` (additive (B,N) padding mask) on the flash entry points.
The tensors are duck-typed: anything with the minitorch ``Tensor`` protocol
(``shape``, ``contiguous()``, ``_tensor._storage`` fp32 1-D numpy, ``zeros(shape)``,
``backend``, ``Tensor.make``) works -- the reference's own ``minitorch.Tensor`` or the
stand-alone ``HostTensor`` in tensor.py.
"""
from __future__ import annotations

import numpy as np

from . import _lib

datatype = np.float32


def _storage(t) -> np.ndarray:
    return t._tensor._storage


def _make_like(ref, arr: np.ndarray):
    """New tensor of ref's class/backend holding `arr` (fp32, C-order)."""
    arr = np.ascontiguousarray(arr, dtype=datatype)
    out = type(ref).make(arr.reshape(-1), tuple(arr.shape), backend=ref.backend)
    if hasattr(out, "requires_grad_"):
        out.requires_grad_(True)  # the reference creates m with requires_grad=True (:624-628)
    return out


class _PinnedPool:
    """fp32 numpy arrays over page-locked host memory (``fa_malloc_host``), recycled by size.

    The outputs of the flash entry points (O, m, l, dQ, dK, dV) are allocated here when they are large: the library's
    D2H copies / widening threads then write into memory that is already mapped and pinned instead of first-touching a
    fresh ``numpy.zeros`` block page by page, and a later call that reads them (the backward reads O) skips staging.
    A block returns to the free list when the last numpy view of it dies."""

    MIN_BYTES = 1 << 20
    MAX_CACHED = 8 << 30

    def __init__(self):
        self.free = {}
        self.cached = 0

    def _release(self, ptr, nbytes):
        if self.cached + nbytes <= self.MAX_CACHED:
            self.free.setdefault(nbytes, []).append(ptr)
            self.cached += nbytes
        else:
            try:
                _lib.load("flashattention_kernel").fa_free_host(ptr)
            except Exception:
                pass

    def empty(self, count: int):
        import ctypes
        import weakref
        nbytes = int(count) * 4
        if nbytes < self.MIN_BYTES:
            return None
        lst = self.free.get(nbytes)
        if lst:
            ptr = lst.pop()
            self.cached -= nbytes
        else:
            lib = _lib.load("flashattention_kernel")
            ptr = lib.fa_malloc_host(nbytes)
            if not ptr and self.free:       # pinned memory is exhausted: give the cached blocks back and retry
                for nb, ptrs in self.free.items():
                    for q in ptrs:
                        lib.fa_free_host(q)
                self.free, self.cached = {}, 0
                ptr = lib.fa_malloc_host(nbytes)
            if not ptr:
                return None
        buf = (ctypes.c_float * int(count)).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return np.frombuffer(buf, dtype=datatype)


_pinned = _PinnedPool()


def _out_like(ref, shape):
    """Output tensor of ref's class/backend; large ones live in pinned memory and are NOT zero-filled (the library
    overwrites every element of its outputs)."""
    count = int(np.prod(shape))
    arr = _pinned.empty(count)
    if arr is None:
        return ref.zeros(tuple(shape))
    return type(ref).make(arr, tuple(shape), backend=ref.backend)


def _mask_ptr(key_mask, B, N):
    if key_mask is None:
        return None, None
    km = np.ascontiguousarray(key_mask.to_numpy() if hasattr(key_mask, "to_numpy") else key_mask, dtype=datatype)
    km = km.reshape(B, N)
    return km, km.ctypes.data_as(_lib.c_void_p)


# function ids of the combine.so switch (minitorch/cuda_kernel_ops.py:33-52, src/combine.cu:11-28).  Keys are
# the operator NAMES so that the reference's ``minitorch.operators`` callables (looked up through
# ``__name__``), plain strings and raw ids all resolve.
fn_map = {"add": 1, "mul": 2, "id": 3, "neg": 4, "lt": 5, "eq": 6, "sigmoid": 7, "relu": 8, "relu_back": 9,
          "log": 10, "log_back": 11, "exp": 12, "inv": 13, "inv_back": 14, "is_close": 15, "max": 16, "pow": 17,
          "tanh": 18}


def _fn_id(fn) -> int:
    if isinstance(fn, int):
        return fn
    name = fn if isinstance(fn, str) else getattr(fn, "__name__", None)
    if name not in fn_map:
        raise KeyError(f"{fn!r} is not one of the 18 functions combine.so implements")  # reference: KeyError too
    return fn_map[name]


def _i32(seq) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(seq, dtype=np.int32).reshape(-1))


def _layout(t):
    d = t._tensor
    shape = getattr(d, "_shape", None)
    strides = getattr(d, "_strides", None)
    return _i32(d.shape if shape is None else shape), _i32(d.strides if strides is None else strides)


def shape_broadcast(s1, s2):
    """minitorch/tensor_data.py shape_broadcast: right-aligned numpy-style union."""
    n = max(len(s1), len(s2))
    a, b = (1,) * (n - len(s1)) + tuple(s1), (1,) * (n - len(s2)) + tuple(s2)
    out = []
    for x, y in zip(a, b):
        if x != y and x != 1 and y != 1:
            raise IndexError(f"cannot broadcast {tuple(s1)} with {tuple(s2)}")
        out.append(max(x, y))
    return tuple(out)


def _size(t) -> int:
    return int(np.prod(t.shape)) if len(t.shape) else 1


class CudaKernelOps:
    cuda = True

    # ------------------------------------------------------------------ map / zip / reduce / matmul
    # (minitorch/cuda_kernel_ops.py:58-437 over combine.so)
    @staticmethod
    def map(fn):
        fn_id = _fn_id(fn)

        def ret(a, out=None):
            lib = _lib.load("combine")
            if out is None:
                out = a.zeros(a.shape)
            osh, ost = _layout(out)
            ash, ast = _layout(a)
            if len(ash) < len(osh):      # `out` given with more dims: right-align like broadcast_index
                pad = len(osh) - len(ash)
                ash = _i32([1] * pad + list(ash))
                ast = _i32([0] * pad + list(ast))
            lib.tensorMap(_storage(out), osh, ost, _size(out), _storage(a), ash, ast, _size(a), len(osh), fn_id)
            _lib.check(lib)
            return out

        return ret

    @staticmethod
    def zip(fn):
        fn_id = _fn_id(fn)

        def ret(a, b):
            lib = _lib.load("combine")
            out = a.zeros(shape_broadcast(a.shape, b.shape))
            osh, ost = _layout(out)
            ash, ast = _layout(a)
            bsh, bst = _layout(b)
            lib.tensorZip(_storage(out), osh, ost, _size(out), len(osh), _storage(a), ash, ast, _size(a), len(ash),
                          _storage(b), bsh, bst, _size(b), len(bsh), fn_id)
            _lib.check(lib)
            return out

        return ret

    @staticmethod
    def reduce(fn, start: float = 0.0):
        fn_id = _fn_id(fn)

        def ret(a, dim: int):
            lib = _lib.load("combine")
            out_shape = list(a.shape)
            out_shape[dim] = 1
            out = a.zeros(tuple(out_shape))
            osh, ost = _layout(out)
            ash, ast = _layout(a)
            lib.tensorReduce(_storage(out), osh, ost, _size(out), _storage(a), ash, ast, int(dim), float(start),
                             len(ash), fn_id)
            _lib.check(lib)
            return out

        return ret

    @staticmethod
    def matrix_multiply(a, b):
        """Batched matmul with the reference's shape handling (:343-437): 2-D operands become a batch of
        one (broadcast against the other side), >3-D operands are flattened over their leading dims."""
        lib = _lib.load("combine")
        both_2d = 0
        if len(a.shape) == 2:
            a = a.contiguous().view(1, a.shape[0], a.shape[1])
            both_2d += 1
        if len(b.shape) == 2:
            b = b.contiguous().view(1, b.shape[0], b.shape[1])
            both_2d += 1
        both_2d = both_2d == 2
        ls = list(shape_broadcast(a.shape[:-2], b.shape[:-2])) + [a.shape[-2], b.shape[-1]]
        if a.shape[-1] != b.shape[-2]:
            raise AssertionError(f"matmul inner dims differ: {a.shape} @ {b.shape}")
        out = a.zeros(tuple(ls))
        more_3d = len(ls) > 3
        if more_3d:
            out = out.view(int(np.prod(ls[:-2])), ls[-2], ls[-1])
        if len(a.shape) > 3:
            a = a.contiguous().view(int(np.prod(a.shape[:-2])), a.shape[-2], a.shape[-1])
        if len(b.shape) > 3:
            b = b.contiguous().view(int(np.prod(b.shape[:-2])), b.shape[-2], b.shape[-1])
        osh, ost = _layout(out)
        ash, ast = _layout(a)
        bsh, bst = _layout(b)
        lib.MatrixMultiply(_storage(out), osh, ost, _storage(a), ash, ast, _storage(b), bsh, bst, out.shape[0],
                           a.shape[1], b.shape[2])
        _lib.check(lib)
        if both_2d:
            out = out.view(out.shape[1], out.shape[2])
        if more_3d:
            out = out.view(*ls)
        return out

    # ------------------------------------------------------------------ flash attention
    @staticmethod
    def _flash_fw(Q, K, V, causal: bool, key_mask=None):
        lib = _lib.load("flashattention_kernel")
        B, nh, N, d = Q.shape
        O, l, m = _out_like(Q, (B, nh, N, d)), _out_like(Q, (B, nh, N)), _out_like(Q, (B, nh, N))
        if hasattr(m, "requires_grad_"):
            m.requires_grad_(True)  # the reference creates m with requires_grad=True (:624-628)
        Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
        keep, kptr = _mask_ptr(key_mask, B, N)
        if key_mask is None and not causal:
            lib.launch_flashattention_forward(_storage(Q), _storage(K), _storage(V), _storage(O), _storage(l),
                                              _storage(m), B, nh, N, d)
        elif key_mask is None:
            lib.launch_flashattention_forward_causal(_storage(Q), _storage(K), _storage(V), _storage(O),
                                                     _storage(l), _storage(m), B, nh, N, d)
        else:
            lib.launch_flashattention_forward_masked(_storage(Q), _storage(K), _storage(V), _storage(O),
                                                     _storage(l), _storage(m), kptr, int(causal), B, nh, N, d)
        _lib.check(lib)
        return O, m, l

    @staticmethod
    def _flash_bw(Q, K, V, O, dO, m, l, causal: bool, key_mask=None):
        lib = _lib.load("flashattention_kernel")
        B, nh, N, d = Q.shape
        dQ, dK, dV = (_out_like(Q, (B, nh, N, d)) for _ in range(3))
        Q, K, V, O, dO, m, l = (t.contiguous() for t in (Q, K, V, O, dO, m, l))
        keep, kptr = _mask_ptr(key_mask, B, N)
        args = (_storage(Q), _storage(K), _storage(V), _storage(O), _storage(dQ), _storage(dK), _storage(dV),
                _storage(dO), _storage(l), _storage(m))
        if key_mask is None and not causal:
            lib.launch_flashattention_backward(*args, B, nh, N, d)
        elif key_mask is None:
            lib.launch_flashattention_backward_causal(*args, B, nh, N, d)
        else:
            lib.launch_flashattention_backward_masked(*args, kptr, int(causal), B, nh, N, d)
        _lib.check(lib)
        return dQ, dK, dV

    @staticmethod
    def flash_attention_fw(Q, K, V, key_mask=None):
        return CudaKernelOps._flash_fw(Q, K, V, False, key_mask)

    @staticmethod
    def flash_attention_bw(Q, K, V, O, dO, m, l, key_mask=None):
        return CudaKernelOps._flash_bw(Q, K, V, O, dO, m, l, False, key_mask)

    @staticmethod
    def flash_attention_causal_fw(Q, K, V, key_mask=None):
        return CudaKernelOps._flash_fw(Q, K, V, True, key_mask)

    @staticmethod
    def flash_attention_causal_bw(Q, K, V, O, dO, m, l, key_mask=None):
        return CudaKernelOps._flash_bw(Q, K, V, O, dO, m, l, True, key_mask)

    # ------------------------------------------------------------------ fused softmax
    @staticmethod
    def attn_softmax_fw(inp, mask, mask_future: bool = False):
        """In place, like the reference (:440-468): returns `inp` itself."""
        lib = _lib.load("softmax_kernel")
        batch_size, nhead, from_len, to_len = inp.shape
        mptr = None
        if mask is not None:
            ms = _storage(mask.contiguous())
            if ms.size < batch_size * to_len:  # e.g. a (1,1,T,T) dummy: broadcast the first row
                ms = np.ascontiguousarray(np.broadcast_to(ms[:to_len], (batch_size, to_len))).reshape(-1)
            mptr = ms.ctypes.data_as(_lib.c_void_p)
        lib.launch_attn_softmax(_storage(inp), mptr, batch_size, nhead, from_len, to_len, bool(mask_future), None)
        _lib.check(lib)
        return inp

    @staticmethod
    def attn_softmax_bw(out_grad, soft_inp):
        lib = _lib.load("softmax_kernel")
        rows = out_grad.shape[0] * out_grad.shape[1] * out_grad.shape[2]
        softmax_len = soft_inp.shape[3]
        lib.launch_attn_softmax_bw(_storage(out_grad), _storage(soft_inp), rows, softmax_len, None)
        _lib.check(lib)
        return out_grad, soft_inp

    # ------------------------------------------------------------------ fused layernorm
    @staticmethod
    def layernorm_fw(inp, gamma, beta):
        lib = _lib.load("layernorm_kernel")
        batch_size, hidden_dim = inp.shape
        ln_res = inp.zeros(inp.shape)
        var = inp.zeros((batch_size,))
        means = inp.zeros((batch_size,))
        inp, gamma, beta = inp.contiguous(), gamma.contiguous(), beta.contiguous()
        lib.launch_layernorm(_storage(ln_res), _storage(var), _storage(means), _storage(inp), _storage(gamma),
                             _storage(beta), batch_size, hidden_dim, None)
        _lib.check(lib)
        return ln_res, var, means

    @staticmethod
    def layernorm_bw(out_grad, inp, gamma, beta, var, mean):
        lib = _lib.load("layernorm_kernel")
        batch_size, hidden_dim = inp.shape
        gamma_grad = gamma.zeros((1, gamma.shape[0]))
        beta_grad = beta.zeros((1, beta.shape[0]))
        inp_grad = inp.zeros(inp.shape)
        out_grad, inp, gamma, beta, var, mean = (t.contiguous() for t in (out_grad, inp, gamma, beta, var, mean))
        lib.launch_layernorm_bw(_storage(gamma_grad), _storage(beta_grad), _storage(inp_grad), _storage(out_grad),
                                _storage(inp), _storage(gamma), _storage(beta), _storage(var), _storage(mean),
                                batch_size, hidden_dim, None, None)
        _lib.check(lib)
        return inp_grad, gamma_grad, beta_grad

    # ------------------------------------------------------------------ embedding lookup / cross-entropy
    @staticmethod
    def embedding_fw(ids, weights):
        lib = _lib.load("combine")
        ids, weights = ids.contiguous(), weights.contiguous()
        V, E = weights.shape
        out = weights.zeros(tuple(ids.shape) + (E,))
        lib.launch_embedding_fw(_storage(out), _storage(ids), _storage(weights), _size(ids), V, E)
        _lib.check(lib)
        return out

    @staticmethod
    def embedding_bw(ids, out_grad, num_embeddings: int):
        lib = _lib.load("combine")
        ids, out_grad = ids.contiguous(), out_grad.contiguous()
        E = out_grad.shape[-1]
        dW = out_grad.zeros((num_embeddings, E))
        lib.launch_embedding_bw(_storage(dW), _storage(ids), _storage(out_grad), _size(ids), num_embeddings, E)
        _lib.check(lib)
        return dW

    @staticmethod
    def softmax_xent_fw(logits, target):
        lib = _lib.load("combine")
        logits, target = logits.contiguous(), target.contiguous()
        n, C = logits.shape
        loss, lse = logits.zeros((n,)), logits.zeros((n,))
        lib.launch_softmax_xent_fw(_storage(loss), _storage(lse), _storage(logits), _storage(target), n, C)
        _lib.check(lib)
        return loss, lse

    @staticmethod
    def softmax_xent_bw(out_grad, logits, target, lse):
        lib = _lib.load("combine")
        out_grad, logits, target, lse = (t.contiguous() for t in (out_grad, logits, target, lse))
        n, C = logits.shape
        dx = logits.zeros((n, C))
        lib.launch_softmax_xent_bw(_storage(dx), _storage(out_grad), _storage(logits), _storage(target), _storage(lse), n, C)
        _lib.check(lib)
        return dx

    # ------------------------------------------------------------------ mode switch
    @staticmethod
    def set_flash_mode(mode: str) -> None:
        """'fp32' (default, <=1e-5) or 'bf16' (tcgen05 tensor cores, <=2e-2)."""
        lib = _lib.load("flashattention_kernel")
        lib.fa_set_mode({"fp32": _lib.FA_MODE_FP32, "bf16": _lib.FA_MODE_BF16}[mode])

    @staticmethod
    def get_flash_mode() -> str:
        lib = _lib.load("flashattention_kernel")
        return "bf16" if lib.fa_get_mode() == _lib.FA_MODE_BF16 else "fp32"
