#!/usr/bin/env python
"""CPU arm of bench.py: the REFERENCE's own composed attention (minitorch/modules_transfomer.py:177-192 ->
minitorch/nn.py:104-123 softmax -> matmul) on the reference's own numba CPU backend (minitorch/fast_ops.py FastOps),
imported from the overlay tree baseline/_ref (built by baseline/build_overlay.sh; unmodified reference Python).

Run as a subprocess with NUMBA_DISABLE_CUDA=1 so that minitorch/nn.py:56-61 selects FastOps.  Two shims are needed
because the reference never wires FastOps to the fused-op surface: the eight fused-op stubs of
SimpleOps (tensor_ops.py:240-269) and a matmul that flattens >3-D operands like CudaKernelOps.matrix_multiply
(cuda_kernel_ops.py:357-369).  Prints one JSON line: seconds per fwd+bwd step, cores, sample.
"""
import json
import os
import sys
import time

os.environ.setdefault("NUMBA_DISABLE_CUDA", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
OVERLAY = os.path.join(HERE, "_ref")
sys.path.insert(0, OVERLAY)

import numpy as np  # noqa: E402


def main():
    B, H, N, d, steps, warmup = (int(x) for x in sys.argv[1:7])
    import numba
    import minitorch
    from minitorch.fast_ops import FastOps
    from minitorch.tensor_functions import tensor_from_numpy
    from minitorch.tensor_ops import SimpleOps

    class FastOpsX(FastOps):
        attn_softmax_fw = SimpleOps.attn_softmax_fw
        attn_softmax_bw = SimpleOps.attn_softmax_bw
        layernorm_fw = SimpleOps.layernorm_fw
        layernorm_bw = SimpleOps.layernorm_bw
        flash_attention_fw = SimpleOps.flash_attention_fw
        flash_attention_bw = SimpleOps.flash_attention_bw
        flash_attention_causal_fw = SimpleOps.flash_attention_causal_fw
        flash_attention_causal_bw = SimpleOps.flash_attention_causal_bw

        @staticmethod
        def matrix_multiply(a, b):
            if len(a.shape) > 3:
                ls = list(a.shape[:-2])
                a3 = a.contiguous().view(int(np.prod(ls)), a.shape[-2], a.shape[-1])
                b3 = b.contiguous().view(int(np.prod(ls)), b.shape[-2], b.shape[-1])
                out = FastOps.matrix_multiply(a3, b3)
                return out.view(*ls, out.shape[-2], out.shape[-1])
            return FastOps.matrix_multiply(a, b)

    backend = minitorch.TensorBackend(FastOpsX)
    rng = np.random.default_rng(0)
    Q, K, V, dO = (rng.standard_normal((B, H, N, d)).astype(np.float32) for _ in range(4))

    def mt(x, grad=True):
        return tensor_from_numpy(np.ascontiguousarray(x), backend=backend, requires_grad=grad)

    def step():
        q, k, v = mt(Q), mt(K), mt(V)
        s = (q @ k.permute(0, 1, 3, 2)) / (d ** 0.5)
        o = minitorch.nn.softmax(s, dim=3) @ v
        (o * mt(dO, grad=False)).sum().backward()
        return float(q.grad.to_numpy().reshape(-1)[0])

    for _ in range(max(1, warmup)):     # includes the numba JIT of every kernel on the path
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    print(json.dumps(dict(seconds_per_step=dt, cores=int(numba.get_num_threads()), B=B, H=H, N=N, d=d, steps=steps)))


if __name__ == "__main__":
    main()
