#!/usr/bin/env bash
# One GPU-box pass over the test groups, each in its own process (a faulting kernel kills
# only its own CUDA context).  Logs land in gpurun_out/.   usage: gpu_check.sh [group ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? ($name)"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
groups="${@:-companions vs_ref_kernels flash_fp32 diag_fwd diag_bwd flash_bf16}"
for g in $groups; do
  case $g in
    companions) run companions python -m pytest tests/test_gpu_companions.py -q -m gpu ;;
    vs_ref_kernels) run vs_ref_kernels python -m pytest tests/test_gpu_vs_reference_kernels.py -q -m gpu ;;
    flash_fp32) run flash_fp32 python -m pytest tests/test_gpu_flash_fp32.py -q -m gpu ;;
    diag_fwd) run diag_fwd python tools/diag_fwd.py ;;
    diag_bwd) run diag_bwd python tools/diag_bwd.py ;;
    flash_bf16) run flash_bf16 python -m pytest tests/test_gpu_flash_bf16.py -q -m gpu ;;
    combine) run combine python -m pytest tests/test_gpu_combine.py -q -m gpu ;;
    mha) run mha python -m pytest tests/test_gpu_mha_integration.py -q -m gpu ;;
    extra) run extra python tools/bench_extra.py --out gpurun_out/extra.json ;;
    extra_attn) run extra_attn python tools/bench_extra.py --what sweep,cfg4 --out gpurun_out/extra_attn.json ;;
    extra_comp) run extra_comp python tools/bench_extra.py --what companions --out gpurun_out/extra_comp.json ;;
    alltests) run alltests python -m pytest tests -x -q -m gpu ;;
    bench) run bench python bench.py --steps 10 --warmup 3 ;;
    benchq) run benchq python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu ;;
    smoke) run smoke python -c "import __graft_entry__ as g; g.smoke()" ;;
  esac
done
