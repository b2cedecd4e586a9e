#!/usr/bin/env bash
# Build the REFERENCE OVERLAY: the reference's own Python (minitorch/, tests/, kernel_tests/, test_utils.py,
# project/) copied from /root/reference into the git-ignored baseline/_ref/, with THIS repo's four C-ABI
# libraries dropped where the reference's minitorch/cuda_kernel_ops.py:26-29 loads them
# (cwd-relative minitorch/cuda_kernels/*.so).  Running the reference's tests from that directory is the
# north-star acceptance test: the reference's own binding + the reference's own tests + the new kernels.
#
# Nothing of the reference enters the git history (baseline/_ref/ is in .gitignore); the directory is NOT in
# .gpurunignore, so it travels to the GPU box like the built .so files.  /root/reference only exists in the
# build container: on the GPU box this script only refreshes the .so files of an existing overlay.
#
# Deviations from the reference tree (all recorded here and in INTEGRATION.md):
#   1. a stub `pycuda` package (pycuda/__init__.py, autoinit.py, driver.py, gpuarray.py -- empty): pycuda is not
#      installed and only the dead matrix_multiply_cublas (cuda_kernel_ops.py:243-338) would use it;
#   2. minitorch/tensor_functions.py Attn_Softmax.backward (:444-450): `(inp,) = ctx.saved_values` -> two values
#      were saved (:440), so the reference's own line raises ValueError; fixed to `inp, mask = ctx.saved_values`
#      and the two debug print() calls (one stringifies the whole tensor) dropped.  Without this
#      kernel_tests/test_softmax_bw.py cannot pass on ANY backend.
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="$HERE/_ref"
KDIR="$ROOT/llmsys-project-flashattn_b200/minitorch/cuda_kernels"

if [ -d "$REF/minitorch" ]; then
  rm -rf "$OUT"
  mkdir -p "$OUT"
  cp -r "$REF/minitorch" "$REF/tests" "$REF/kernel_tests" "$REF/project" "$OUT/"
  cp "$REF/test_utils.py" "$OUT/"
  chmod -R u+w "$OUT"
  find "$OUT" -name __pycache__ -type d -prune -exec rm -rf {} +
  mkdir -p "$OUT/pycuda"
  for f in __init__ autoinit driver gpuarray; do
    echo "# stub: pycuda is not installed; the reference only needs the import to succeed" > "$OUT/pycuda/$f.py"
  done
  python - "$OUT/minitorch/tensor_functions.py" <<'PY'
import re, sys
p = sys.argv[1]
s = open(p).read()
old = "        (inp,) = ctx.saved_values\n        print(inp)\n        print(\"INSIDE BACKWARD?\")\n"
new = "        inp, mask = ctx.saved_values  # overlay fix: two values are saved by forward()\n"
assert s.count(old) == 1, "Attn_Softmax.backward no longer matches the recorded patch"
s = s.replace(old, new)
s = s.replace("        print(\"INSIDE FORWARD?\")\n", "")
open(p, "w").write(s)
PY
  echo "overlay: copied reference python from $REF"
elif [ ! -d "$OUT/minitorch" ]; then
  echo "overlay: no reference at $REF and no existing overlay at $OUT -- nothing to do"
  exit 0
fi
mkdir -p "$OUT/minitorch/cuda_kernels"
for name in combine softmax_kernel layernorm_kernel flashattention_kernel; do
  [ -f "$KDIR/$name.so" ] || { echo "overlay: $KDIR/$name.so missing; run compile_cuda.sh first" >&2; exit 1; }
  cp "$KDIR/$name.so" "$OUT/minitorch/cuda_kernels/$name.so"
done
echo "overlay ready: $OUT (libraries: $(ls "$OUT/minitorch/cuda_kernels" | tr '\n' ' '))"
