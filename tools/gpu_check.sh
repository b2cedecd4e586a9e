#!/usr/bin/env bash
# One GPU-box pass over the test groups, each in its own process (a faulting kernel kills
# only its own CUDA context).  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n 25 gpurun_out/$name.log; }
run companions python -m pytest tests/test_gpu_companions.py -q -m gpu
run vs_ref_kernels python -m pytest tests/test_gpu_vs_reference_kernels.py -q -m gpu
run flash_fp32 python -m pytest tests/test_gpu_flash_fp32.py -q -m gpu
run diag_p0 python tools/diag_fwd.py 0
run diag_p1 python tools/diag_fwd.py 1
run flash_bf16 python -m pytest tests/test_gpu_flash_bf16.py -q -m gpu
