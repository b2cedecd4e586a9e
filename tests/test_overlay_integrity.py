"""CPU checks around the reference overlay (baseline/_ref, built by baseline/build_overlay.sh):

* where /root/reference exists (the build container), the overlay's python files are byte-identical to the reference's,
  except the ONE recorded patch (Attn_Softmax.backward in minitorch/tensor_functions.py) and the added pycuda stub /
  cuda_kernels directory -- i.e. the GPU acceptance test really runs the reference's own code;
* the reference's own ctypes binding (minitorch/cuda_kernel_ops.py:26-29) loads this repo's four libraries (symbol
  resolution only; no compute without a GPU);
* the log classifier of tools/run_overlay_tests.py tells a tolerance tail from a hard failure.
"""
import difflib
import filecmp
import os
import subprocess
import sys

import pytest

from tools import run_overlay_tests as R

REF = "/root/reference"
needs_overlay = pytest.mark.skipif(not os.path.isdir(os.path.join(R.OVERLAY, "minitorch")),
                                   reason="overlay not built (baseline/build_overlay.sh)")


@needs_overlay
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "minitorch")), reason="reference tree not present on this box")
def test_overlay_is_the_reference_plus_the_recorded_patch():
    changed = []
    for sub in ("minitorch", "tests", "kernel_tests", "project"):
        for root, _, files in os.walk(os.path.join(REF, sub)):
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(root, f)
                dst = os.path.join(R.OVERLAY, os.path.relpath(src, REF))
                assert os.path.exists(dst), dst
                if not filecmp.cmp(src, dst, shallow=False):
                    changed.append(os.path.relpath(src, REF))
    assert filecmp.cmp(os.path.join(REF, "test_utils.py"), os.path.join(R.OVERLAY, "test_utils.py"), shallow=False)
    assert changed == ["minitorch/tensor_functions.py"], changed
    a = open(os.path.join(REF, "minitorch/tensor_functions.py")).read().splitlines()
    b = open(os.path.join(R.OVERLAY, "minitorch/tensor_functions.py")).read().splitlines()
    diff = [ln for ln in difflib.unified_diff(a, b, lineterm="", n=0) if ln[:1] in "+-" and ln[:3] not in ("+++", "---")]
    removed = [ln[1:].strip() for ln in diff if ln[0] == "-"]
    added = [ln[1:].strip() for ln in diff if ln[0] == "+"]
    assert sorted(removed) == sorted(['print("INSIDE FORWARD?")', "(inp,) = ctx.saved_values", "print(inp)",
                                      'print("INSIDE BACKWARD?")']), removed
    assert len(added) == 1 and added[0].startswith("inp, mask = ctx.saved_values"), added


@needs_overlay
def test_reference_binding_loads_the_libraries():
    """`from minitorch.cuda_kernel_ops import CudaKernelOps` in the overlay = ctypes.CDLL of the four .so + symbol lookups
    the reference makes at call time (checked here by name)."""
    code = ("import minitorch\n"
            "from minitorch import cuda_kernel_ops as c\n"
            "for lib, names in ((c.lib, ['tensorMap', 'tensorZip', 'tensorReduce', 'MatrixMultiply']),\n"
            "                   (c.lib_softmax, ['launch_attn_softmax', 'launch_attn_softmax_bw']),\n"
            "                   (c.lib_layernorm, ['launch_layernorm', 'launch_layernorm_bw']),\n"
            "                   (c.lib_flashattention, ['launch_flashattention_forward', 'launch_flashattention_backward',\n"
            "                                           'launch_flashattention_forward_causal',\n"
            "                                           'launch_flashattention_backward_causal'])):\n"
            "    for n in names:\n"
            "        getattr(lib, n)\n"
            "print('bound')\n")
    env = dict(os.environ, NUMBA_DISABLE_CUDA="1", PYTHONPATH=R.OVERLAY)
    p = subprocess.run([sys.executable, "-c", code], cwd=R.OVERLAY, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "bound" in p.stdout, p.stderr[-2000:]


def test_failure_classifier():
    log = """..F.F
=================================== FAILURES ===================================
_ test_multihead_attention_flash_attention_is_causal[CudaKernelOps-0.0-2-256-2048-64] _
E       Mismatched elements: 5 / 33554432 (1.49e-05%)
E       Max absolute difference among violations: 1.1211261e-05
tests/test_flash_attention.py:170: AssertionError
_ test_multihead_attention_flash_attention_is_causal[CudaKernelOps-0.0-4-2048-2048-64] _
E       Mismatched elements: 2048 / 4194304 (0.0488%)
E       Max absolute difference among violations: 4.720688e-05
tests/test_flash_attention.py:174: AssertionError
=========================== short test summary info ============================
"""
    f = R.parse_failures(log)
    a = f["test_multihead_attention_flash_attention_is_causal[CudaKernelOps-0.0-2-256-2048-64]"]
    b = f["test_multihead_attention_flash_attention_is_causal[CudaKernelOps-0.0-4-2048-2048-64]"]
    assert a["tail"] and a["check"] == "X.grad" and a["mismatched"] == 5
    assert not b["tail"] and b["check"] == "W_out.grad"


def test_grid_selection_follows_host_memory(monkeypatch):
    monkeypatch.setattr(R, "mem_available", lambda: 64 << 30)
    pts = R.causal_grid("full")
    assert (16, 64, 4096) not in pts and (2, 64, 2048) in pts
    assert all(64 * nh * N * N * 16 <= 0.6 * (64 << 30) for nh, e, N in pts)
    assert R.causal_grid("quick") == [(nh, e, 2048) for e in (64, 256, 512) for nh in (2, 4)]
