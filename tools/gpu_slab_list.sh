#!/bin/bash
# per-kernel device times of the slab step at per-rank size (virtual ranks on one GPU), ncu launch list
mkdir -p gpurun_out
CMD="python tools/profile_slab.py --ranks ${1:-8} --steps 2 --warmup 2"
timeout 600 $CMD > gpurun_out/slab_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/slab_launches.csv $CMD > gpurun_out/slab_ncu.log 2>&1
tail -2 gpurun_out/slab_plain.log
