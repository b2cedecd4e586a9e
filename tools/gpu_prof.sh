#!/usr/bin/env bash
# usage: gpu_prof.sh <kernel-regex> <out-name> [skip]   -- plain run first, then launch list + one full ncu capture
mkdir -p gpurun_out
K=$1; OUT=$2; SKIP=${3:-3}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras"
timeout 600 $CMD > gpurun_out/plain_$OUT.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$OUT.csv $CMD > gpurun_out/ncul_$OUT.log 2>&1
echo "ncu launches rc=$?"
timeout 600 $CMD > gpurun_out/plain2_$OUT.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/$OUT -f $CMD > gpurun_out/ncu_$OUT.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_$OUT.log
