"""``OracleOps``: a CPU stand-in for ``CudaKernelOps`` built on the oracle (TEST INFRASTRUCTURE).

Same static-method surface (map / zip / reduce / matrix_multiply and the eight fused ops), numpy arithmetic
from ``oracle/``.  It lets the ``-m "not gpu"`` suite drive the host-side logic -- HostTensor autograd, the
MultiHeadAttention / TransformerLayer / DecoderLM call sites -- against the golden vectors without a GPU.
The product never imports this module.
"""
import numpy as np

from flashattn_b200.cuda_kernel_ops import _fn_id, shape_broadcast
from oracle import attention_ref as R
from oracle import combine_ref as C

f32 = np.float32


def _new(ref, arr):
    arr = np.array(arr, dtype=f32, order="C")     # always a fresh writable buffer
    return type(ref).make(arr.reshape(-1), tuple(arr.shape), backend=ref.backend)


class OracleOps:
    cuda = False

    @staticmethod
    def map(fn):
        fid = _fn_id(fn)
        return lambda a, out=None: _new(a, C.tensor_map(fid, a.to_numpy()))

    @staticmethod
    def zip(fn):
        fid = _fn_id(fn)

        def ret(a, b):
            shape = shape_broadcast(a.shape, b.shape)
            return _new(a, np.broadcast_to(C.tensor_zip(fid, a.to_numpy(), b.to_numpy()), shape))
        return ret

    @staticmethod
    def reduce(fn, start=0.0):
        fid = _fn_id(fn)
        return lambda a, dim: _new(a, C.tensor_reduce(fid, a.to_numpy(), int(dim), start))

    @staticmethod
    def matrix_multiply(a, b):
        return _new(a, C.matrix_multiply(a.to_numpy(), b.to_numpy()))

    @staticmethod
    def _fw(Q, K, V, causal):
        O, m, l = R.attention_fwd(Q.to_numpy(), K.to_numpy(), V.to_numpy(), causal=causal)
        return _new(Q, O), _new(Q, m), _new(Q, l)

    @staticmethod
    def _bw(Q, K, V, O, dO, m, l, causal):
        g = R.attention_bwd(Q.to_numpy(), K.to_numpy(), V.to_numpy(), dO.to_numpy(), causal=causal)
        return tuple(_new(Q, x) for x in g)

    flash_attention_fw = staticmethod(lambda Q, K, V: OracleOps._fw(Q, K, V, False))
    flash_attention_causal_fw = staticmethod(lambda Q, K, V: OracleOps._fw(Q, K, V, True))
    flash_attention_bw = staticmethod(lambda Q, K, V, O, dO, m, l: OracleOps._bw(Q, K, V, O, dO, m, l, False))
    flash_attention_causal_bw = staticmethod(lambda Q, K, V, O, dO, m, l: OracleOps._bw(Q, K, V, O, dO, m, l, True))

    @staticmethod
    def attn_softmax_fw(inp, mask, mask_future=False):
        y = R.attn_softmax_fw(inp.to_numpy(), None if mask is None else mask.to_numpy(), mask_future)
        inp._tensor._storage[:] = np.ascontiguousarray(y, dtype=f32).reshape(-1)   # in place, like the kernel
        return inp

    @staticmethod
    def attn_softmax_bw(out_grad, soft_inp):
        return _new(out_grad, R.attn_softmax_bw(out_grad.to_numpy(), soft_inp.to_numpy())), soft_inp

    @staticmethod
    def layernorm_fw(inp, gamma, beta):
        y, var, mean = R.layernorm_fw(inp.to_numpy(), gamma.to_numpy(), beta.to_numpy())
        return _new(inp, y), _new(inp, var), _new(inp, mean)

    @staticmethod
    def layernorm_bw(out_grad, inp, gamma, beta, var, mean):
        dx, dg, db = R.layernorm_bw(out_grad.to_numpy(), inp.to_numpy(), gamma.to_numpy(), beta.to_numpy(),
                                    var.to_numpy(), mean.to_numpy())
        return _new(inp, dx), _new(inp, np.reshape(dg, (1, -1))), _new(inp, np.reshape(db, (1, -1)))

    # fused lookup / loss ops: numpy restatement of the one-hot formulations they replace
    @staticmethod
    def embedding_fw(ids, weights):
        return _new(weights, weights.to_numpy()[ids.to_numpy().astype(np.int64)])

    @staticmethod
    def embedding_bw(ids, out_grad, num_embeddings):
        idx = ids.to_numpy().astype(np.int64).reshape(-1)
        g = out_grad.to_numpy().astype(np.float64).reshape(idx.size, -1)
        dW = np.zeros((num_embeddings, g.shape[1]))
        np.add.at(dW, idx, g)
        return _new(out_grad, dW)

    @staticmethod
    def softmax_xent_fw(logits, target):
        x = logits.to_numpy().astype(np.float64)
        t = target.to_numpy().astype(np.int64)
        mx = x.max(axis=1)
        lse = mx + np.log(np.exp(x - mx[:, None]).sum(axis=1) + C.EPS)      # minitorch's log adds EPS
        return _new(logits, lse - x[np.arange(len(t)), t]), _new(logits, lse)

    @staticmethod
    def softmax_xent_bw(out_grad, logits, target, lse):
        x = logits.to_numpy().astype(np.float64)
        t = target.to_numpy().astype(np.int64)
        p = np.exp(x - lse.to_numpy().astype(np.float64)[:, None])
        p[np.arange(len(t)), t] -= 1.0
        return _new(logits, out_grad.to_numpy().astype(np.float64)[:, None] * p)

    # decode shapes: numpy KV cache + the decode-attention formula, so that the cache bookkeeping of
    # MultiHeadAttention.forward_cached / decode_step / generate_cached is testable without a GPU
    class _Cache:
        def __init__(self, B, nh, capacity, d):
            self.B, self.nh, self.capacity, self.d = B, nh, capacity, d
            self.k = np.zeros((B, nh, capacity, d), dtype=f32)
            self.v = np.zeros((B, nh, capacity, d), dtype=f32)
            self.len = 0

    @staticmethod
    def kv_cache_new(B, nh, capacity, d):
        return OracleOps._Cache(B, nh, capacity, d)

    @staticmethod
    def kv_cache_append(cache, k, v):
        n = k.shape[2]
        if cache.len + n > cache.capacity:
            raise ValueError("cache overflow")
        cache.k[:, :, cache.len:cache.len + n] = k.to_numpy()
        cache.v[:, :, cache.len:cache.len + n] = v.to_numpy()
        cache.len += n

    @staticmethod
    def flash_decode(q, cache):
        Q = q.to_numpy().astype(np.float64)                       # (B, nh, 1, d)
        K = cache.k[:, :, :cache.len].astype(np.float64)
        V = cache.v[:, :, :cache.len].astype(np.float64)
        s = np.einsum("bhqd,bhkd->bhqk", Q, K) / np.sqrt(Q.shape[-1])
        p = np.exp(s - s.max(-1, keepdims=True))
        p /= p.sum(-1, keepdims=True)
        return _new(q, np.einsum("bhqk,bhkd->bhqd", p, V))
