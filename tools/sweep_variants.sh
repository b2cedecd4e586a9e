#!/usr/bin/env bash
# usage: sweep_variants.sh dir1 dir2 ...   -- quick bench of alternative builds of flashattention_kernel.so
for d in "$@"; do
  FLASHATTN_B200_KERNEL_DIR=$PWD/$d timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        k = json.loads(line)['kernels']; print('$d', 'fwd_ms %.3f (%.0f TF)  bwd_ms %.3f (%.0f TF)' % (k['fwd_ms'], k['fwd_tflops'], k['bwd_ms'], k['bwd_tflops']))
    elif 'rror' in line: print(line.strip()[:200])
"
done
