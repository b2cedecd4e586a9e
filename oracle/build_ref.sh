#!/usr/bin/env bash
# Compile the REFERENCE's own CUDA sources for the companion kernels (softmax, layernorm)
# from where they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the
# GPU box with the snapshot).  Nothing is copied into the repo.  The reference's flash kernel
# (src/flashattention_kernel.cu) is NOT built: it does not compile as committed
# (undeclared max_sram_size :362/:771, tile_idx_Q :476 -- SURVEY.md 2.3).
# These libraries are CHECKERS for tests/ only; the product never loads them.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "reference not present at $REF; keeping any prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT"
for name in softmax_kernel layernorm_kernel; do
  nvcc -O2 -gencode arch=compute_100,code=sm_100 --shared -Xcompiler -fPIC -I "$REF/src" \
       -o "$OUT/ref_$name.so" "$REF/src/$name.cu"
done
echo "built: $(ls "$OUT"/*.so | tr '\n' ' ')"
