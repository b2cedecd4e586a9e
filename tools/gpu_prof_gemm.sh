#!/usr/bin/env bash
# plain run first, then one full ncu capture of the tcgen05 GEMM (8192^3) and of the decode kernel
mkdir -p gpurun_out
CMD="python tools/gemm_bench.py"
timeout 300 $CMD > gpurun_out/plain_gemm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/r02_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"; tail -n 2 gpurun_out/ncu_gemm.log
