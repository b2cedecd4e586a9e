#!/usr/bin/env bash
# usage: build_variant.sh <name> [extra nvcc flags...]  -- builds build/<name>/flashattention_kernel.so
# (alternative builds for tools/sweep_variants.sh and the FA_TRACE timeline tools)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
name=$1; shift
mkdir -p "$HERE/build/$name"
g++ -O3 -std=c++17 -fPIC -c "$HERE/llmsys-project-flashattn_b200/csrc/host_convert.cpp" -o "$HERE/build/$name/host_convert.o"
nvcc -std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --shared -Xcompiler -fPIC "$@" \
  -o "$HERE/build/$name/flashattention_kernel.so" "$HERE/llmsys-project-flashattn_b200/csrc/flashattention_kernel.cu" "$HERE/build/$name/host_convert.o"
echo "built build/$name"
