"""``DeviceKernelOps`` -- the same operator surface as ``CudaKernelOps`` with tensor storage RESIDENT IN HBM.

SURVEY.md 8(f)-1: in the reference every op is a host-pointer call (cudaMalloc + H2D + kernel + D2H + cudaFree,
minitorch/cuda_kernel_ops.py:60-437 / src/combine.cu:315-580), so once attention is fused the PCIe round trips of
the plumbing dominate a training step.  Here a tensor's ``_tensor._storage`` is a ``DeviceStorage`` (fp32, from
the stream-ordered pool), every op calls the ``*_dev`` entry point of the C ABI on the default stream, and data
crosses PCIe only in ``tensor_from_numpy`` / ``to_numpy``.  Bind it with ``TensorBackend(DeviceKernelOps)``; the
autograd nodes and the MultiHeadAttention / DecoderLM modules are unchanged.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .cuda_kernel_ops import _fn_id, _i32, shape_broadcast

datatype = np.float32


def _fa():
    return _lib.load("flashattention_kernel")


class DeviceStorage:
    """Flat fp32 device buffer; freed (stream-ordered) on garbage collection."""

    __slots__ = ("ptr", "size")

    def __init__(self, size: int):
        self.size = int(size)
        lib = _fa()
        self.ptr = lib.fa_malloc_async(max(self.size, 1) * 4, None)
        if not self.ptr:
            _lib.check(lib)
            raise MemoryError(f"fa_malloc_async({self.size * 4}) failed")

    @classmethod
    def from_numpy(cls, arr) -> "DeviceStorage":
        host = np.ascontiguousarray(arr, dtype=datatype).reshape(-1)
        out = cls(host.size)
        if host.size:
            lib = _fa()
            _lib.check(lib, lib.fa_h2d(out.ptr, host.ctypes.data_as(ctypes.c_void_p), host.size * 4))
        return out

    def to_numpy(self) -> np.ndarray:
        host = np.empty(self.size, dtype=datatype)
        if self.size:
            lib = _fa()
            _lib.check(lib, lib.fa_d2h(host.ctypes.data_as(ctypes.c_void_p), self.ptr, self.size * 4))
        return host

    def __len__(self):
        return self.size

    def __del__(self):
        try:
            if self.ptr:
                _fa().fa_free_async(self.ptr, None)
                self.ptr = None
        except Exception:
            pass


def _st(t) -> DeviceStorage:
    return t._tensor._storage


def _layout(t):
    d = t._tensor
    return d._shape, d._strides        # cached int32 arrays (tensor.py::_Data)


def _size(t) -> int:
    return int(np.prod(t.shape)) if len(t.shape) else 1


def _bf16_copy(t):
    """bf16 image of a contiguous fp32 tensor (device scratch, 2 bytes per element)."""
    lib = _fa()
    n = _size(t)
    buf = DeviceStorage((n + 1) // 2)
    _lib.check(lib, lib.fa_cast_f32_to_bf16_dev(_st(t).ptr, buf.ptr, n, None))
    return buf


class KVCache:
    """Keys and values of the positions decoded so far, resident in HBM: fp32 (B, nh, capacity, d) each."""

    def __init__(self, B: int, nh: int, capacity: int, d: int):
        self.B, self.nh, self.capacity, self.d = int(B), int(nh), int(capacity), int(d)
        n = self.B * self.nh * self.capacity * self.d
        self.k, self.v = DeviceStorage(n), DeviceStorage(n)
        self.len = 0


class DeviceKernelOps:
    cuda = True
    device_resident = True
    flash_mode = "fp32"       # "bf16": tcgen05 tensor-core kernels for head_dim 64 / 128 (operands rounded on device)
    gemm_mode = "fp32"        # "bf16": 2-D matmuls (Linear fwd / bwd, lm_head) on the tcgen05 GEMM, fp32 accumulation

    # ---- storage hooks used by HostTensor ------------------------------------------------------------------
    @staticmethod
    def to_storage(x):
        return x if isinstance(x, DeviceStorage) else DeviceStorage.from_numpy(x)

    @staticmethod
    def zeros_storage(n: int):
        st = DeviceStorage(n)
        lib = _fa()
        _lib.check(lib, lib.fa_memset_async(st.ptr, 0, max(int(n), 1) * 4, None))
        return st

    @staticmethod
    def storage_to_numpy(st) -> np.ndarray:
        return st.to_numpy()

    # ---- map / zip / reduce / matmul -------------------------------------------------------------------------
    @staticmethod
    def map(fn):
        fn_id = _fn_id(fn)

        def ret(a, out=None):
            lib = _lib.load("combine")
            if out is None:
                out = a.zeros(a.shape)
            osh, ost = _layout(out)
            ash, ast = _layout(a)
            _lib.check(lib, lib.fa_map_dev(_st(out).ptr, osh, ost, len(osh), _st(a).ptr, ash, ast, len(ash), fn_id, None))
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
            _lib.check(lib, lib.fa_zip_dev(_st(out).ptr, osh, ost, len(osh), _st(a).ptr, ash, ast, len(ash), _st(b).ptr,
                                           bsh, bst, len(bsh), fn_id, None))
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
            _lib.check(lib, lib.fa_reduce_dev(_st(out).ptr, osh, ost, _st(a).ptr, ash, ast, len(ash), int(dim),
                                              float(start), fn_id, None))
            return out

        return ret

    @staticmethod
    def matrix_multiply(a, b):
        """Same shape handling as CudaKernelOps.matrix_multiply (minitorch/cuda_kernel_ops.py:343-437); operands
        are used in place through their strides (transposed views need no copy)."""
        lib = _lib.load("combine")
        if a.shape[-1] != b.shape[-2]:
            raise AssertionError(f"matmul inner dims differ: {a.shape} @ {b.shape}")
        lead = shape_broadcast(a.shape[:-2], b.shape[:-2])
        ls = list(lead) + [a.shape[-2], b.shape[-1]]

        def as3(t):
            if len(t.shape) == 2:
                return (1,) + tuple(t.shape), (0,) + tuple(t._tensor.strides), t
            if len(t.shape) == 3:
                return tuple(t.shape), tuple(t._tensor.strides), t
            t = t.contiguous()
            nb = int(np.prod(t.shape[:-2]))
            return (nb, t.shape[-2], t.shape[-1]), (t.shape[-2] * t.shape[-1], t.shape[-1], 1), t

        ash, ast, a = as3(a)
        bsh, bst, b = as3(b)
        nb = int(np.prod(lead)) if lead else 1
        if ash[0] not in (1, nb) or bsh[0] not in (1, nb):
            raise IndexError(f"cannot broadcast matmul batches {a.shape} @ {b.shape}")
        out = a.zeros(tuple(ls))
        m, p = ls[-2], ls[-1]
        if DeviceKernelOps.gemm_mode == "bf16" and nb == 1 and DeviceKernelOps._gemm_bf16(out, a, ast, b, bst, m, p, ash[2]):
            return out
        lib_rc = lib.fa_matmul_dev(_st(out).ptr, _i32((nb, m, p)), _i32((m * p, p, 1)), _st(a).ptr, _i32(ash), _i32(ast),
                                   _st(b).ptr, _i32(bsh), _i32(bst), None)
        _lib.check(lib, lib_rc)
        return out

    @staticmethod
    def _gemm_bf16(out, a, ast, b, bst, m, p, k) -> bool:
        """out (m, p) fp32 = a (m, k) @ b (k, p) on the tcgen05 GEMM (operands rounded to bf16 on the device, fp32
        accumulation).  Transposed VIEWS are consumed in place (the kernel takes either memory order per operand), so
        the backward GEMMs dx = dy @ W^T and dW = x^T @ dy of a Linear layer need no copies.  Returns False when the
        layout does not fit (caller falls back to the fp32 CUDA-core matmul)."""
        if len(_st(a)) != m * k or len(_st(b)) != k * p or m * p * k < (1 << 16):
            return False
        am, ak = ast[-2], ast[-1]
        bk, bn = bst[-2], bst[-1]
        if ak == 1 and am >= k:
            a_mn, lda = 0, am
        elif am == 1 and ak >= m:
            a_mn, lda = 1, ak
        else:
            return False
        if bn == 1 and bk >= p:
            b_mn, ldb = 1, bk
        elif bk == 1 and bn >= k:
            b_mn, ldb = 0, bn
        else:
            return False
        if lda % 8 or ldb % 8:
            return False
        lib = _lib.load("combine")
        a16, b16 = DeviceStorage((m * k + 1) // 2), DeviceStorage((k * p + 1) // 2)
        fl = _fa()
        _lib.check(fl, fl.fa_cast_f32_to_bf16_dev(_st(a).ptr, a16.ptr, m * k, None))
        _lib.check(fl, fl.fa_cast_f32_to_bf16_dev(_st(b).ptr, b16.ptr, k * p, None))
        _lib.check(lib, lib.fa_gemm_bf16_dev(_st(out).ptr, 0, p, a16.ptr, a_mn, lda, b16.ptr, b_mn, ldb, m, p, k, None))
        return True

    @staticmethod
    def qkv_projection(x2, wq, wk, wv):
        """Fused Q/K/V projection: x2 (M, E) times the concatenated (E, 3E) weight in ONE tcgen05
        GEMM whose three column blocks land in separate (M, E) fp32 buffers -- (B, N, nh, d) storage, which the flash
        kernels consume in place through their strides.  Returns None when the shapes do not fit (caller uses three
        matmuls)."""
        M, E = x2.shape
        if E % 32 or any(tuple(w.shape) != (E, E) for w in (wq, wk, wv)) or len(_st(x2)) != M * E or \
                x2._tensor.strides != (E, 1) or any(w._tensor.strides != (E, 1) or len(_st(w)) != E * E for w in (wq, wk, wv)):
            return None
        lib, fl = _lib.load("combine"), _fa()
        cat = DeviceStorage(3 * E * E)                       # [E][3E] fp32: three strided column-block copies
        osh, ost = _i32((E, E)), _i32((3 * E, 1))
        for j, w in enumerate((wq, wk, wv)):
            ish, ist = _layout(w)
            _lib.check(lib, lib.fa_map_dev(cat.ptr + j * E * 4, osh, ost, 2, _st(w).ptr, ish, ist, 2, _fn_id("id"), None))
        w16, x16 = DeviceStorage((3 * E * E + 1) // 2), DeviceStorage((M * E + 1) // 2)
        _lib.check(fl, fl.fa_cast_f32_to_bf16_dev(cat.ptr, w16.ptr, 3 * E * E, None))
        _lib.check(fl, fl.fa_cast_f32_to_bf16_dev(_st(x2).ptr, x16.ptr, M * E, None))
        outs = [x2.zeros((M, E)) for _ in range(3)]
        _lib.check(lib, lib.fa_qkv_proj_bf16_dev(_st(outs[0]).ptr, _st(outs[1]).ptr, _st(outs[2]).ptr, 0, x16.ptr, w16.ptr,
                                                 M, E, None))
        return outs

    # ---- flash attention ---------------------------------------------------------------------------------------
    @staticmethod
    def _desc(B, nh, N, d, causal, bf16, strides=None):
        a = _lib.fa_attn_desc()
        a.B, a.H, a.N, a.d = B, nh, N, d
        a.dtype = _lib.FA_DTYPE_BF16 if bf16 else _lib.FA_DTYPE_F32
        a.causal = int(bool(causal))
        if strides is not None:
            a.stride_b, a.stride_h, a.stride_n = strides
        return a

    @staticmethod
    def _shared_layout(tensors):
        """(stride_b, stride_h, stride_n) when every tensor is the SAME dense, unit-inner-stride view -- e.g. the
        permuted (B, nh, N, d) views of (B, N, nh, d) storage that project_to_query_key_value hands over
        (modules_transfomer.py:87-100).  The kernels then consume them in place (TMA descriptors / index math carry
        the strides): no .contiguous() copies on either side of the attention core."""
        first = tensors[0]._tensor
        st = first.strides
        if st[3] != 1 or len(_st(tensors[0])) != _size(tensors[0]):
            return None
        dense = sorted(zip(st, first.shape), reverse=True)       # a permutation of a packed buffer?
        acc = 1
        for stride, extent in reversed(dense):
            if extent != 1 and stride != acc:
                return None
            acc *= extent
        for t in tensors[1:]:
            if t._tensor.strides != st or t.shape != tensors[0].shape or len(_st(t)) != _size(t):
                return None
        if any(x % 8 for x in st[:3]):                           # TMA wants 16-byte multiples
            return None
        return st[0], st[1], st[2]

    @staticmethod
    def _like(t, strides4):
        """Uninitialised-but-zeroed tensor with t's shape laid out with the given strides."""
        from .tensor import HostTensor, _Data
        st = DeviceKernelOps.zeros_storage(_size(t))
        return HostTensor(_Data(st, t.shape, strides4), t.backend)

    @staticmethod
    def _flash_fw(Q, K, V, causal):
        lib = _fa()
        B, nh, N, d = Q.shape
        lay = DeviceKernelOps._shared_layout((Q, K, V))
        if lay is None:
            Q, K, V = Q.contiguous(), K.contiguous(), V.contiguous()
            O = Q.zeros((B, nh, N, d))
        else:
            O = DeviceKernelOps._like(Q, Q._tensor.strides)
        m, l = Q.zeros((B, nh, N)), Q.zeros((B, nh, N))
        bf16 = DeviceKernelOps.flash_mode == "bf16" and d % 8 == 0 and 8 <= d <= 128
        a = DeviceKernelOps._desc(B, nh, N, d, causal, bf16, lay)
        if bf16:
            q, k, v = (_bf16_copy(t) for t in (Q, K, V))
            o = DeviceStorage((_size(O) + 1) // 2)
            _lib.check(lib, lib.fa_flash_fwd_dev(ctypes.byref(a), q.ptr, k.ptr, v.ptr, o.ptr, _st(m).ptr, _st(l).ptr, None))
            _lib.check(lib, lib.fa_cast_bf16_to_f32_dev(o.ptr, _st(O).ptr, _size(O), None))
        else:
            _lib.check(lib, lib.fa_flash_fwd_dev(ctypes.byref(a), _st(Q).ptr, _st(K).ptr, _st(V).ptr, _st(O).ptr,
                                                 _st(m).ptr, _st(l).ptr, None))
        return O, m, l

    @staticmethod
    def _flash_bw(Q, K, V, O, dO, m, l, causal):
        lib = _fa()
        B, nh, N, d = Q.shape
        m, l = m.contiguous(), l.contiguous()
        lay = DeviceKernelOps._shared_layout((Q, K, V, O, dO))
        if lay is None:
            Q, K, V, O, dO = (t.contiguous() for t in (Q, K, V, O, dO))
            grads = tuple(Q.zeros((B, nh, N, d)) for _ in range(3))
        else:
            grads = tuple(DeviceKernelOps._like(Q, Q._tensor.strides) for _ in range(3))
        bf16 = DeviceKernelOps.flash_mode == "bf16" and d % 8 == 0 and 8 <= d <= 128
        a = DeviceKernelOps._desc(B, nh, N, d, causal, bf16, lay)
        if bf16:
            q, k, v, o, do = (_bf16_copy(t) for t in (Q, K, V, O, dO))
            g16 = [DeviceStorage((_size(Q) + 1) // 2) for _ in range(3)]
            _lib.check(lib, lib.fa_flash_bwd_dev(ctypes.byref(a), q.ptr, k.ptr, v.ptr, o.ptr, do.ptr, _st(m).ptr,
                                                 _st(l).ptr, g16[0].ptr, g16[1].ptr, g16[2].ptr, None))
            for src, dst in zip(g16, grads):
                _lib.check(lib, lib.fa_cast_bf16_to_f32_dev(src.ptr, _st(dst).ptr, _size(dst), None))
        else:
            _lib.check(lib, lib.fa_flash_bwd_dev(ctypes.byref(a), _st(Q).ptr, _st(K).ptr, _st(V).ptr, _st(O).ptr,
                                                 _st(dO).ptr, _st(m).ptr, _st(l).ptr, _st(grads[0]).ptr,
                                                 _st(grads[1]).ptr, _st(grads[2]).ptr, None))
        return grads

    @staticmethod
    def flash_attention_fw(Q, K, V):
        return DeviceKernelOps._flash_fw(Q, K, V, False)

    @staticmethod
    def flash_attention_bw(Q, K, V, O, dO, m, l):
        return DeviceKernelOps._flash_bw(Q, K, V, O, dO, m, l, False)

    @staticmethod
    def flash_attention_causal_fw(Q, K, V):
        return DeviceKernelOps._flash_fw(Q, K, V, True)

    @staticmethod
    def flash_attention_causal_bw(Q, K, V, O, dO, m, l):
        return DeviceKernelOps._flash_bw(Q, K, V, O, dO, m, l, True)

    # ---- decode shapes: KV cache + split-KV attention -------------------------------------------
    @staticmethod
    def kv_cache_new(B: int, nh: int, capacity: int, d: int) -> "KVCache":
        return KVCache(B, nh, capacity, d)

    @staticmethod
    def kv_cache_append(cache: "KVCache", k, v) -> None:
        """Copy the (B, nh, n_new, d) keys / values (any strides, e.g. the permuted views of
        project_to_query_key_value) into positions [len, len + n_new) of the cache."""
        lib = _lib.load("combine")
        B, nh, n_new, d = k.shape
        if (B, nh, d) != (cache.B, cache.nh, cache.d) or cache.len + n_new > cache.capacity:
            raise ValueError(f"kv_cache_append: {k.shape} does not fit cache {cache.B, cache.nh, cache.capacity, cache.d} "
                             f"at position {cache.len}")
        osh = _i32((B, nh, n_new, d))
        ost = _i32((nh * cache.capacity * d, cache.capacity * d, d, 1))
        for t, buf in ((k, cache.k), (v, cache.v)):
            ish, ist = _layout(t)
            _lib.check(lib, lib.fa_map_dev(buf.ptr + cache.len * d * 4, osh, ost, 4, _st(t).ptr, ish, ist, 4,
                                           _fn_id("id"), None))
        cache.len += n_new

    @staticmethod
    def flash_decode(q, cache: "KVCache"):
        """softmax(q K^T / sqrt(d)) V for ONE query token per (batch, head) against the cached positions:
        q (B, nh, 1, d) -> (B, nh, 1, d), fp32.  Causality is implicit (the query is the newest position)."""
        lib = _fa()
        B, nh, one, d = q.shape
        assert one == 1 and (B, nh, d) == (cache.B, cache.nh, cache.d)
        q = q.contiguous()
        out = q.zeros((B, nh, 1, d))
        a = _lib.fa_decode_desc()
        a.B, a.H, a.d, a.L, a.L_cap, a.dtype = B, nh, d, cache.len, cache.capacity, _lib.FA_DTYPE_F32
        _lib.check(lib, lib.fa_flash_decode_dev(ctypes.byref(a), _st(q).ptr, cache.k.ptr, cache.v.ptr, _st(out).ptr,
                                                None, None))
        return out

    # ---- fused softmax / layernorm -----------------------------------------------------------------------------
    @staticmethod
    def attn_softmax_fw(inp, mask, mask_future: bool = False):
        """In place on a contiguous `inp` (returns it), like the reference (:440-468)."""
        lib = _lib.load("softmax_kernel")
        B, nhead, from_len, to_len = inp.shape
        assert inp._tensor.is_contiguous(), "attn_softmax_fw works in place on a contiguous tensor"
        mptr = None
        if mask is not None:
            mask = mask.contiguous()
            if _size(mask) < B * to_len:
                raise ValueError("attn_softmax_fw: mask must hold (batch, to_len) additive values")
            mptr = _st(mask).ptr
        _lib.check(lib, lib.fa_attn_softmax_dev(_st(inp).ptr, mptr, B, nhead, from_len, to_len, int(bool(mask_future)),
                                                None))
        return inp

    @staticmethod
    def attn_softmax_bw(out_grad, soft_inp):
        lib = _lib.load("softmax_kernel")
        rows = out_grad.shape[0] * out_grad.shape[1] * out_grad.shape[2]
        assert out_grad._tensor.is_contiguous()
        soft_inp = soft_inp.contiguous()
        _lib.check(lib, lib.fa_attn_softmax_bw_dev(_st(out_grad).ptr, _st(soft_inp).ptr, rows, soft_inp.shape[3], None))
        return out_grad, soft_inp

    @staticmethod
    def layernorm_fw(inp, gamma, beta):
        lib = _lib.load("layernorm_kernel")
        rows, hidden = inp.shape
        ln_res, var, means = inp.zeros(inp.shape), inp.zeros((rows,)), inp.zeros((rows,))
        inp, gamma, beta = inp.contiguous(), gamma.contiguous(), beta.contiguous()
        _lib.check(lib, lib.fa_layernorm_dev(_st(ln_res).ptr, _st(var).ptr, _st(means).ptr, _st(inp).ptr, _st(gamma).ptr,
                                             _st(beta).ptr, rows, hidden, None))
        return ln_res, var, means

    @staticmethod
    def layernorm_bw(out_grad, inp, gamma, beta, var, mean):
        lib = _lib.load("layernorm_kernel")
        rows, hidden = inp.shape
        gamma_grad, beta_grad = gamma.zeros((1, gamma.shape[0])), beta.zeros((1, beta.shape[0]))
        inp_grad = inp.zeros(inp.shape)
        out_grad, inp, gamma, beta, var, mean = (t.contiguous() for t in (out_grad, inp, gamma, beta, var, mean))
        _lib.check(lib, lib.fa_layernorm_bw_dev(_st(gamma_grad).ptr, _st(beta_grad).ptr, _st(inp_grad).ptr,
                                                _st(out_grad).ptr, _st(inp).ptr, _st(gamma).ptr, _st(beta).ptr,
                                                _st(var).ptr, _st(mean).ptr, rows, hidden, None))
        return inp_grad, gamma_grad, beta_grad

    # ---- embedding lookup / softmax cross-entropy without one-hot matmuls --------------------
    @staticmethod
    def embedding_fw(ids, weights):
        """(*,) ids, (V, E) weights -> (*, E): rows of `weights` (== one_hot(ids) @ weights)."""
        lib = _lib.load("combine")
        ids, weights = ids.contiguous(), weights.contiguous()
        V, E = weights.shape
        out = weights.zeros(tuple(ids.shape) + (E,))
        _lib.check(lib, lib.fa_embedding_fw_dev(_st(out).ptr, _st(ids).ptr, _st(weights).ptr, _size(ids), V, E, None))
        return out

    @staticmethod
    def embedding_bw(ids, out_grad, num_embeddings: int):
        """Gradient of the table: (V, E), token rows summed per id in ascending token order (deterministic)."""
        lib = _lib.load("combine")
        ids, out_grad = ids.contiguous(), out_grad.contiguous()
        E = out_grad.shape[-1]
        dW = out_grad.zeros((num_embeddings, E))
        _lib.check(lib, lib.fa_embedding_bw_dev(_st(dW).ptr, _st(ids).ptr, _st(out_grad).ptr, _size(ids), num_embeddings, E,
                                                None))
        return dW

    @staticmethod
    def softmax_xent_fw(logits, target):
        """(n, C) logits, (n,) targets -> (loss (n,), lse (n,))."""
        lib = _lib.load("combine")
        logits, target = logits.contiguous(), target.contiguous()
        n, C = logits.shape
        loss, lse = logits.zeros((n,)), logits.zeros((n,))
        _lib.check(lib, lib.fa_softmax_xent_fw_dev(_st(loss).ptr, _st(lse).ptr, _st(logits).ptr, _st(target).ptr, n, C, None))
        return loss, lse

    @staticmethod
    def softmax_xent_bw(out_grad, logits, target, lse):
        lib = _lib.load("combine")
        out_grad, logits, target, lse = (t.contiguous() for t in (out_grad, logits, target, lse))
        n, C = logits.shape
        dx = logits.zeros((n, C))
        _lib.check(lib, lib.fa_softmax_xent_bw_dev(_st(dx).ptr, _st(out_grad).ptr, _st(logits).ptr, _st(target).ptr,
                                                   _st(lse).ptr, n, C, None))
        return dx

    @staticmethod
    def set_flash_mode(mode: str) -> None:
        assert mode in ("fp32", "bf16")
        DeviceKernelOps.flash_mode = mode

    @staticmethod
    def set_gemm_mode(mode: str) -> None:
        """'fp32' (default: fp32 CUDA-core matmul, <= 1e-5) or 'bf16' (tcgen05 GEMM, bf16 operands / fp32 accumulate)."""
        if mode not in ("fp32", "bf16"):
            raise ValueError(mode)
        DeviceKernelOps.gemm_mode = mode

    @staticmethod
    def get_flash_mode() -> str:
        return DeviceKernelOps.flash_mode
