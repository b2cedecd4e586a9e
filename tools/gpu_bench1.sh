#!/usr/bin/env bash
# bench + profiling pass (1 GPU): plain bench, then launch list and one full ncu capture of the forward kernel.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_flash_bf16.py -q -m gpu > gpurun_out/flash_bf16.log 2>&1; echo "bf16 tests rc=$?"; tail -n 5 gpurun_out/flash_bf16.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.log 2>&1; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_full.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwd_kernel -s 3 -c 1 -o gpurun_out/prof_fwd -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -n 3 gpurun_out/ncu2.log
