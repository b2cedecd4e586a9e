"""The call site of the hot path: ``MultiHeadAttention`` with the reference's constructor flags and method
names (minitorch/modules_transfomer.py:19-229), over the stand-alone ``HostTensor`` carrier.

Three attention cores behind one module, selected exactly as the reference does (:137-192):
  * ``use_flash_attention=True``  -> ``q.flash_attention[_causal](k, v)``      (the path this repo replaces)
  * ``use_fused_kernel=True``     -> ``(q @ kT / sqrt(d)).attn_softmax(mask) @ v``   (companion softmax kernel)
  * neither                       -> ``softmax(q @ kT / sqrt(d) (+ causal mask), dim=3) @ v``  (the composed
                                     path that north_star names as the parity target)
Every arithmetic op (projections included) runs on the GPU through the C-ABI libraries; there is no numpy
arithmetic here.  ``Linear`` follows minitorch/modules_basic.py:104-155 (weights (in, out), uniform
+-1/sqrt(in) init), ``Dropout`` :73-101.
"""
from __future__ import annotations

import numpy as np

from .tensor import (GELU, EmbeddingLookup, HostTensor, TensorBackend, default_backend, fused_qkv, one_hot, softmax,
                     tensor_from_numpy)

datatype = np.float32


class Parameter:
    def __init__(self, value: HostTensor):
        self.value = value
        self.value.requires_grad_(True)


class Module:
    def __init__(self):
        self.training = True

    def named_parameters(self):
        """(dotted name, Parameter) pairs, named like minitorch/module.py:48-67."""
        out = []
        for k, v in self.__dict__.items():
            if isinstance(v, Parameter):
                out.append((k, v))
            elif isinstance(v, Module):
                out.extend((f"{k}.{n}", p) for n, p in v.named_parameters())
        return out

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def train(self):
        self.training = True
        for v in self.__dict__.values():
            if isinstance(v, Module):
                v.train()

    def eval(self):
        self.training = False
        for v in self.__dict__.values():
            if isinstance(v, Module):
                v.eval()

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)


class Linear(Module):
    def __init__(self, in_size: int, out_size: int, bias: bool, backend: TensorBackend = None):
        super().__init__()
        bound = (1.0 / in_size) ** 0.5
        self.out_size = out_size
        self.weights = Parameter(tensor_from_numpy(np.random.uniform(-bound, bound, (in_size, out_size)),
                                                   backend=backend, requires_grad=True))
        self.bias = Parameter(tensor_from_numpy(np.random.uniform(-bound, bound, (out_size,)), backend=backend,
                                                requires_grad=True)) if bias else None

    def forward(self, x: HostTensor) -> HostTensor:
        batch, in_size = x.shape
        out = x.view(batch, in_size) @ self.weights.value
        if self.bias is not None:
            out = out + self.bias.value
        return out


class Dropout(Module):
    def __init__(self, p_dropout: float = 0.1):
        super().__init__()
        self.p_dropout = p_dropout

    def forward(self, x: HostTensor) -> HostTensor:
        if self.p_dropout == 0 or not self.training:
            return x
        mask = tensor_from_numpy(np.random.binomial(1, 1 - self.p_dropout, x.shape), backend=x.backend)
        return (x * mask) / (1 - self.p_dropout)


class MultiHeadAttention(Module):
    def __init__(self, n_embd: int, n_head: int, causal: bool = False, p_dropout: float = 0.1, bias: bool = True,
                 backend: TensorBackend = None, use_fused_kernel: bool = False, use_flash_attention: bool = False):
        super().__init__()
        self.backend = backend if backend is not None else default_backend()
        self.n_embd = n_embd
        self.n_head = n_head
        self.causal = causal
        self.attn_hidden_dim = n_embd // n_head
        self.q_projection = Linear(n_embd, n_embd, bias, self.backend)
        self.k_projection = Linear(n_embd, n_embd, bias, self.backend)
        self.v_projection = Linear(n_embd, n_embd, bias, self.backend)
        self.out_projection = Linear(n_embd, n_embd, bias, self.backend)
        self.dropout = Dropout(p_dropout)
        self.use_fused_kernel = use_fused_kernel
        self.use_flash_attention = use_flash_attention

    def create_causal_mask(self, bs, nh, seq_len):
        """(bs, nh, T, T) additive mask, -finfo(f32).max above the diagonal (:63-71)."""
        key = (bs, nh, seq_len)
        if getattr(self, "_mask_key", None) != key:     # constant per shape: build (and upload) it once
            mask = -np.finfo(datatype).max * np.triu(np.ones((bs, nh, seq_len, seq_len), dtype=datatype), 1)
            self._mask, self._mask_key = tensor_from_numpy(mask, backend=self.backend), key
        return self._mask

    def project_to_query_key_value(self, x: HostTensor):
        """q, k, v: permuted NON-contiguous (B, nh, N, d) views of (B, N, nh, d) storage; kT (B, nh, d, N) (:73-107)."""
        batch_size, seq_len, n_embd = x.shape
        x2 = x.contiguous().view(batch_size * seq_len, n_embd)
        split = (batch_size, seq_len, self.n_head, self.attn_hidden_dim)
        fused = None
        if self.q_projection.bias is None and self.k_projection.bias is None and self.v_projection.bias is None:
            # one GEMM for all three projections when the backend has a tensor-core GEMM mode
            fused = fused_qkv(x2, self.q_projection.weights.value, self.k_projection.weights.value,
                              self.v_projection.weights.value)
        if fused is not None:
            q2, k2, v2 = fused
        else:
            q2, k2, v2 = self.q_projection(x2), self.k_projection(x2), self.v_projection(x2)
        q = q2.view(*split).permute(0, 2, 1, 3)
        k = k2.view(*split)
        kT = k.permute(0, 2, 3, 1)
        k = k.permute(0, 2, 1, 3)
        v = v2.view(*split).permute(0, 2, 1, 3)
        return q, k, kT, v

    def self_attention(self, q: HostTensor, kT: HostTensor, v: HostTensor) -> HostTensor:
        """`kT` is K itself (B, nh, N, d) on the flash branch, K transposed otherwise (:110-199)."""
        batch_size, num_head, queries_len, q_dim = q.shape
        k_dim = kT.shape[3] if self.use_flash_attention else kT.shape[2]
        assert q_dim == k_dim == v.shape[3]
        scale = self.attn_hidden_dim ** 0.5
        if self.use_fused_kernel:
            if self.causal:
                mask = self.create_causal_mask(batch_size, num_head, queries_len)
                result = ((q @ kT) / scale + mask).attn_softmax(None) @ v
            else:
                result = ((q @ kT) / scale).attn_softmax(None) @ v
        elif self.use_flash_attention:
            result = q.flash_attention_causal(kT, v) if self.causal else q.flash_attention(kT, v)
        elif self.causal:
            result = softmax((q @ kT) / scale + self.create_causal_mask(batch_size, num_head, queries_len), dim=3) @ v
        else:
            result = softmax((q @ kT) / scale, dim=3) @ v
        return result.permute(0, 2, 1, 3).contiguous().view(batch_size, queries_len, self.n_embd)

    def forward(self, x: HostTensor) -> HostTensor:
        batch_size, seq_len, n_embd = x.shape
        q, k, kT, v = self.project_to_query_key_value(x)
        if self.use_flash_attention:
            if self.n_embd / self.n_head > 2048:   # the reference's guard (:219-221)
                print("Please reduce n_embd or increase n_head")
                return None
            attn = self.self_attention(q, k, v)
        else:
            attn = self.self_attention(q, kT, v)
        return self.out_projection(attn.view(batch_size * seq_len, n_embd)).view(batch_size, seq_len, n_embd)


    def forward_cached(self, x: HostTensor, cache) -> HostTensor:
        """Decode-time forward (the reference has no cache): `x` holds only the NEW positions
        (B, n_new, E); their keys / values are appended to `cache` and the queries attend to everything cached.
        n_new == 1 runs the split-KV decode kernel; a longer block is the prefill of an EMPTY cache and goes
        through the module's regular (causal) attention core."""
        batch_size, n_new, n_embd = x.shape
        ops = self.backend.ops
        q, k, kT, v = self.project_to_query_key_value(x)
        ops.kv_cache_append(cache, k, v)
        if n_new == 1:
            attn = ops.flash_decode(q, cache).view(batch_size, 1, n_embd)
        else:
            assert cache.len == n_new and self.causal, "prefill expects an empty cache and a causal module"
            attn = self.self_attention(q, k if self.use_flash_attention else kT, v)
        return self.out_projection(attn.view(batch_size * n_new, n_embd)).view(batch_size, n_new, n_embd)


class Embedding(Module):
    """One-hot @ weights, exactly the reference's formulation (minitorch/modules_basic.py:29-71)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, backend: TensorBackend = None, fused: bool = False):
        super().__init__()
        self.fused = fused        # gather kernel instead of one_hot @ weights (same values)
        self.backend = backend
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.weights = Parameter(tensor_from_numpy(np.random.normal(0, 1, (num_embeddings, embedding_dim)),
                                                   backend=backend, requires_grad=True))

    def forward(self, x: HostTensor) -> HostTensor:
        bs, seq_len = x.shape
        if self.fused and hasattr(x.f, "embedding_fw"):
            return EmbeddingLookup.apply(x, self.weights.value)
        hot = one_hot(x, self.num_embeddings).view(bs * seq_len, self.num_embeddings)
        return (hot @ self.weights.value).view(bs, seq_len, self.embedding_dim)


class LayerNorm1d(Module):
    """Composed layer norm.  Like the reference (minitorch/modules_basic.py:158-199) the forward returns the
    normalised input WITHOUT applying weights / bias (they exist as parameters and receive no gradient)."""

    def __init__(self, dim: int, eps: float, backend: TensorBackend = None):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.weights = Parameter(tensor_from_numpy(np.ones((dim,)), backend=backend))
        self.bias = Parameter(tensor_from_numpy(np.zeros((dim,)), backend=backend))

    def forward(self, x: HostTensor) -> HostTensor:
        mean = x.mean(dim=1)
        var = x.var(dim=1)
        return (x - mean) / (var + self.eps) ** 0.5


class FusedLayerNorm(Module):
    """x.layernorm(gamma, beta) through layernorm_kernel.so (minitorch/modules_basic.py:202-210)."""

    def __init__(self, n_embd: int, backend: TensorBackend = None):
        super().__init__()
        self.n_embd = n_embd
        self.gamma = tensor_from_numpy(np.ones((n_embd,)), backend=backend)
        self.beta = tensor_from_numpy(np.zeros((n_embd,)), backend=backend)

    def forward(self, x: HostTensor) -> HostTensor:
        return x.layernorm(self.gamma, self.beta)


class FeedForward(Module):
    """Linear -> GELU -> Linear -> dropout (minitorch/modules_transfomer.py:232-275)."""

    def __init__(self, n_embd: int, middle_dim: int = 256, p_dropout: float = 0.1, bias: bool = True,
                 backend: TensorBackend = None):
        super().__init__()
        self.linear_in = Linear(n_embd, middle_dim, bias=bias, backend=backend)
        self.linear_out = Linear(middle_dim, n_embd, bias=bias, backend=backend)
        self.dropout = Dropout(p_dropout)

    def forward(self, x: HostTensor) -> HostTensor:
        batch_size, seq_len, n_embd = x.shape
        x = GELU(self.linear_in(x.contiguous().view(batch_size * seq_len, n_embd)))
        return self.dropout(self.linear_out(x)).view(batch_size, seq_len, n_embd)


class TransformerLayer(Module):
    """Pre-LN block (minitorch/modules_transfomer.py:278-336).  The flags reach MultiHeadAttention by KEYWORD:
    the reference's positional call hands `use_flash_attention` to `use_fused_kernel`."""

    def __init__(self, n_embd: int, n_head: int, p_dropout: float = 0.1, ln_eps: float = 1e-8, bias: bool = True,
                 backend: TensorBackend = None, use_fused_kernel: bool = False, use_flash_attention: bool = False):
        super().__init__()
        self.attention = MultiHeadAttention(n_embd, n_head, causal=True, p_dropout=p_dropout, bias=bias,
                                            backend=backend, use_fused_kernel=use_fused_kernel,
                                            use_flash_attention=use_flash_attention)
        self.ff = FeedForward(n_embd, 256, p_dropout, bias, backend)
        self.use_fused_kernel = use_fused_kernel
        if not use_fused_kernel:
            self.ln_1 = LayerNorm1d(n_embd, ln_eps, backend)
            self.ln_2 = LayerNorm1d(n_embd, ln_eps, backend)
        else:
            self.ln_1 = FusedLayerNorm(n_embd, backend)
            self.ln_2 = FusedLayerNorm(n_embd, backend)

    def forward(self, x: HostTensor) -> HostTensor:
        batch_size, seq_len, x_dim = x.shape
        a = self.ln_1(x.contiguous().view(batch_size * seq_len, x_dim)).view(batch_size, seq_len, x_dim)
        a = self.attention(a) + x
        y = self.ln_2(a.contiguous().view(batch_size * seq_len, x_dim)).view(batch_size, seq_len, x_dim)
        return self.ff(y) + a

    def forward_cached(self, x: HostTensor, cache) -> HostTensor:
        batch_size, seq_len, x_dim = x.shape
        a = self.ln_1(x.contiguous().view(batch_size * seq_len, x_dim)).view(batch_size, seq_len, x_dim)
        a = self.attention.forward_cached(a, cache) + x
        y = self.ln_2(a.contiguous().view(batch_size * seq_len, x_dim)).view(batch_size, seq_len, x_dim)
        return self.ff(y) + a


class DecoderLM(Module):
    """Decoder-only Pre-LN transformer with four layers (minitorch/modules_transfomer.py:339-453); config #2 of
    BASELINE.json is DecoderLM(n_vocab=10000, n_embd=256, n_head=8, n_positions=40) with flash attention."""

    def __init__(self, n_vocab: int, n_embd: int, n_head: int, n_positions: int, p_dropout: float = 0.1,
                 ln_eps: float = 1e-5, bias: bool = True, backend: TensorBackend = None,
                 use_fused_kernel: bool = False, use_flash_attention: bool = False, use_fused_embedding: bool = False):
        super().__init__()
        self.backend = backend if backend is not None else default_backend()
        self.n_embd = n_embd
        self.n_vocab = n_vocab
        # use_fused_embedding (an extension, default off): gather kernel instead of the one-hot matmul
        self.token_embeddings = Embedding(n_vocab, n_embd, self.backend, fused=use_fused_embedding)
        self.position_embeddings = Embedding(n_vocab, n_embd, self.backend, fused=use_fused_embedding)   # (sic) n_vocab rows
        kw = dict(p_dropout=p_dropout, ln_eps=ln_eps, bias=bias, backend=self.backend,
                  use_fused_kernel=use_fused_kernel, use_flash_attention=use_flash_attention)
        self.t_layer_1 = TransformerLayer(n_embd, n_head, **kw)
        self.t_layer_2 = TransformerLayer(n_embd, n_head, **kw)
        self.t_layer_3 = TransformerLayer(n_embd, n_head, **kw)
        self.t_layer_4 = TransformerLayer(n_embd, n_head, **kw)
        self.dropout = Dropout(p_dropout)
        self.lm_head = Linear(n_embd, n_vocab, bias, self.backend)
        self.use_fused_kernel = use_fused_kernel
        self.ln = FusedLayerNorm(n_embd, self.backend) if use_fused_kernel else LayerNorm1d(n_embd, ln_eps, self.backend)

    def forward(self, idx: HostTensor) -> HostTensor:
        batch_size, seq_len = idx.shape
        position_id = tensor_from_numpy(np.arange(seq_len, dtype=datatype).reshape(1, seq_len), backend=self.backend)
        x = self.token_embeddings(idx) + self.position_embeddings(position_id).view(1, seq_len, self.n_embd)
        for layer in (self.t_layer_1, self.t_layer_2, self.t_layer_3, self.t_layer_4):
            x = layer(x)
        x = self.ln(x.contiguous().view(batch_size * seq_len, self.n_embd))
        return self.lm_head(x).view(batch_size, seq_len, self.n_vocab)


def _layers(model):
    return (model.t_layer_1, model.t_layer_2, model.t_layer_3, model.t_layer_4)


def decode_step(model: DecoderLM, new_ids, caches) -> HostTensor:
    """Logits (B, n_new, n_vocab) of the NEW positions only; `caches` (one KVCache per layer) hold the rest."""
    ids = np.asarray(new_ids, dtype=datatype)
    batch_size, n_new = ids.shape
    start = caches[0].len
    idx = tensor_from_numpy(ids, backend=model.backend)
    pos = tensor_from_numpy(np.arange(start, start + n_new, dtype=datatype).reshape(1, n_new), backend=model.backend)
    x = model.token_embeddings(idx) + model.position_embeddings(pos).view(1, n_new, model.n_embd)
    for layer, cache in zip(_layers(model), caches):
        x = layer.forward_cached(x, cache)
    x = model.ln(x.contiguous().view(batch_size * n_new, model.n_embd))
    return model.lm_head(x).view(batch_size, n_new, model.n_vocab)


def generate_cached(model: DecoderLM, token_ids, model_max_length: int, eos_id: int = -1):
    """Greedy decoding with a KV cache: the prompt is run once (prefill), then every new token costs ONE position
    through the layers and one split-KV decode-attention call per layer, instead of the reference's full-prefix
    re-run per token (project/run_machine_translation.py:299-325).  Same tokens as generate() (tests pin that).
    Needs a device-resident backend (DeviceKernelOps: kv_cache_new / kv_cache_append / flash_decode)."""
    ops = model.backend.ops
    if not hasattr(ops, "flash_decode"):
        raise RuntimeError("generate_cached needs a backend with a KV cache (TensorBackend(DeviceKernelOps))")
    model.eval()
    ids = [int(t) for t in token_ids]
    attn = model.t_layer_1.attention
    caches = [ops.kv_cache_new(1, attn.n_head, model_max_length + 1, attn.attn_hidden_dim) for _ in _layers(model)]
    new = list(ids)
    while len(ids) <= model_max_length:
        logits = decode_step(model, np.asarray(new, dtype=datatype).reshape(1, len(new)), caches)
        gen_id = int(np.argmax(logits.to_numpy()[0, len(new) - 1, :]))
        if gen_id == eos_id:
            break
        ids.append(gen_id)
        new = [gen_id]
    return ids


def generate(model: DecoderLM, token_ids, model_max_length: int, eos_id: int = -1):
    """Greedy decoding exactly as the reference does it (project/run_machine_translation.py:300-325): the model is
    re-run on the WHOLE prefix for every new token (no KV cache) and the arg-max of the last position is appended,
    until `eos_id` or `model_max_length` tokens.  Returns the full id list.  Attention here is the decode regime of
    SURVEY.md 8(f)-3 (batch 1, N <= model_max_length): launch-latency bound, so what matters is that nothing
    crosses PCIe per op -- use a device-resident backend."""
    model.eval()
    ids = [int(t) for t in token_ids]
    while len(ids) <= model_max_length:
        x = tensor_from_numpy(np.asarray(ids, dtype=datatype).reshape(1, len(ids)), backend=model.backend)
        logits = model(x)
        gen_id = int(np.argmax(logits.to_numpy()[0, len(ids) - 1, :]))
        if gen_id == eos_id:
            break
        ids.append(gen_id)
    return ids
