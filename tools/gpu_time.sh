#!/bin/bash
# timing only: per-kernel-group event timings at 8M (and optionally 1M)
mkdir -p gpurun_out
SPHSM_GROUPS=1 timeout 300 python tools/profile_step.py --workload ${1:-8m} --steps 10 --warmup 3 2>&1 | tee gpurun_out/time.log
