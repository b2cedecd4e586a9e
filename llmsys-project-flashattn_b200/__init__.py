"""B200-native fused attention path behind minitorch's CudaKernelOps operator surface.

Layout:
  csrc/                CUDA sources (tcgen05/TMA attention, fp32 attention, softmax, layernorm) + C ABI
  compile_cuda.sh      builds minitorch/cuda_kernels/{flashattention,softmax,layernorm}_kernel.so for sm_100a
  cuda_kernel_ops.py   CudaKernelOps: the reference's fused-op entry points over ctypes
  tensor.py            stand-alone host tensor + FlashAttention / Attn_Softmax / LayerNorm autograd nodes
  modules_transformer.py  MultiHeadAttention / Linear / Dropout: the reference's call site of the path
  device_ops.py        DeviceKernelOps: the same operator surface with tensor storage resident in HBM
  device.py            device-resident buffers + the *_dev entry points (bench / sharded runs)
"""
from . import _lib, device, sharding
from ._lib import FlashAttnError
from .cuda_kernel_ops import CudaKernelOps
from .device_ops import DeviceKernelOps, DeviceStorage, KVCache
from .tensor import (Attn_Softmax, EmbeddingLookup, SoftmaxCrossEntropy, FlashAttention, FlashAttentionCausal, HostTensor, LayerNorm, TensorBackend,
                     default_backend, logsumexp, one_hot, softmax, softmax_loss, GELU,
                     tensor_from_numpy)
from . import modules_transformer
from .modules_transformer import (DecoderLM, Dropout, Embedding, FeedForward, FusedLayerNorm, LayerNorm1d, Linear,
                                  MultiHeadAttention, TransformerLayer, decode_step, generate, generate_cached)

__all__ = ["CudaKernelOps", "DeviceKernelOps", "DeviceStorage", "TensorBackend", "HostTensor", "tensor_from_numpy", "default_backend", "FlashAttention",
           "FlashAttentionCausal", "Attn_Softmax", "LayerNorm", "FlashAttnError", "_lib", "device", "sharding", "softmax", "modules_transformer",
           "MultiHeadAttention", "Linear", "Dropout", "DecoderLM", "TransformerLayer", "FeedForward", "Embedding",
           "LayerNorm1d", "FusedLayerNorm", "softmax_loss", "logsumexp", "one_hot", "GELU", "EmbeddingLookup", "SoftmaxCrossEntropy", "generate", "generate_cached", "decode_step", "KVCache"]
