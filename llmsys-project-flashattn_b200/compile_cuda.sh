#!/usr/bin/env bash
# Build the four C-ABI shared libraries for sm_100a (B200), in-tree.
# Mirrors the reference's compile_cuda.sh (same output names under minitorch/cuda_kernels/
# so the reference's cuda_kernel_ops.py:26-29 can load them unchanged), but with an explicit
# arch: tcgen05 / TMEM / TMA only exist for sm_100a.
#   usage: bash compile_cuda.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/minitorch/cuda_kernels"
mkdir -p "$OUT"
NVCC="${NVCC:-nvcc}"
FLAGS=(-std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC "$@")
# host-side fp32 <-> bf16 conversion of the legacy ABI's staging threads (plain C++, AVX2 chosen at run time)
"${CXX:-g++}" -O3 -std=c++17 -fPIC -c "$HERE/csrc/host_convert.cpp" -o "$OUT/host_convert.o"
pids=()
for name in flashattention_kernel softmax_kernel layernorm_kernel combine; do
  extra=()
  [ "$name" = flashattention_kernel ] && extra=("$OUT/host_convert.o")
  "$NVCC" "${FLAGS[@]}" -o "$OUT/$name.so" "$HERE/csrc/$name.cu" "${extra[@]}" &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
[ $rc -eq 0 ] && echo "built: $(ls "$OUT"/*.so | tr '\n' ' ')"
exit $rc
