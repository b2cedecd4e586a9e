#!/usr/bin/env python
"""Error budget of one grid point of the reference's tests/test_flash_attention.py:103-186 (run from the overlay):
the same recipe evaluated three ways -- the reference's minitorch MHA with flash attention on this repo's libraries,
torch.nn.MultiheadAttention in fp32 (the test's oracle) and in fp64 (taken as the truth) -- to see whose rounding error
a 1e-5 violation is.   usage: overlay_error_budget.py nh n_embd N [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OVERLAY = os.path.join(ROOT, "baseline", "_ref")
os.chdir(OVERLAY)
sys.path.insert(0, OVERLAY)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import minitorch  # noqa: E402
from minitorch.cuda_kernel_ops import CudaKernelOps  # noqa: E402

nh, n_embd, N = (int(x) for x in sys.argv[1:4])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
backend = minitorch.TensorBackend(CudaKernelOps)
np.random.seed(10)
torch.manual_seed(10)
data = np.random.rand(B, N, n_embd)
X = minitorch.tensor_from_numpy(data, backend, True)
layer_ = torch.nn.MultiheadAttention(n_embd, nh, 0.0, bias=False, batch_first=True, dtype=torch.float32)
layer = minitorch.MultiHeadAttention(n_embd, nh, True, 0.0, bias=False, backend=backend, use_fused_kernel=False,
                                     use_flash_attention=True)
w_qkv = layer_.in_proj_weight.detach().numpy().T.copy()
w_q_, w_k_, w_v_ = [w.copy() for w in np.split(w_qkv, 3, -1)]
w_out_ = layer_.out_proj.weight.detach().numpy().T.copy()
for proj, w in ((layer.q_projection, w_q_), (layer.k_projection, w_k_), (layer.v_projection, w_v_),
                (layer.out_projection, w_out_)):
    proj.weights.value = minitorch.tensor_from_numpy(w, backend=backend, requires_grad=True)
M = torch.triu(-float("inf") * torch.ones(N, N), 1)
res = layer(X)
res.sum().backward()
ours = dict(Y=res.to_numpy(), dX=X.grad.to_numpy(), dWo=layer.out_projection.weights.value.grad.to_numpy())


def run_torch(dtype):
    x = torch.tensor(data, dtype=dtype, requires_grad=True)
    lay = torch.nn.MultiheadAttention(n_embd, nh, 0.0, bias=False, batch_first=True, dtype=dtype)
    with torch.no_grad():
        lay.in_proj_weight.copy_(layer_.in_proj_weight.to(dtype))
        lay.out_proj.weight.copy_(layer_.out_proj.weight.to(dtype))
    y, _ = lay(x, x, x, attn_mask=M.to(dtype))
    y.sum().backward()
    return dict(Y=y.detach().numpy(), dX=x.grad.numpy(), dWo=lay.out_proj.weight.grad.numpy().T)


t32 = run_torch(torch.float32)
t64 = run_torch(torch.float64)
for k in ("Y", "dX", "dWo"):
    tol = 1e-5 + 1e-5 * np.abs(t32[k])
    viol = int((np.abs(ours[k] - t32[k]) > tol).sum())
    eo = np.abs(ours[k] - t64[k])
    et = np.abs(t32[k].astype(np.float64) - t64[k])
    print(f"{k:4s} |ours-torch32| max {np.abs(ours[k] - t32[k]).max():.3e} violations {viol}/{ours[k].size};  "
          f"|ours-fp64| max {eo.max():.3e} rms {np.sqrt((eo ** 2).mean()):.3e};  "
          f"|torch32-fp64| max {et.max():.3e} rms {np.sqrt((et ** 2).mean()):.3e};  max|x| {np.abs(t64[k]).max():.3g}",
          flush=True)
