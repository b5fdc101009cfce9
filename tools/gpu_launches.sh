#!/bin/bash
# ncu launch list only (per-kernel device time; cold-cache, serialised: compare shares)
mkdir -p gpurun_out
CMD="python tools/profile_step.py --workload ${1:-8m} --steps 3 --warmup 2"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/plain.log
