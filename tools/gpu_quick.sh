#!/bin/bash
# quick loop: GPU tests + per-kernel-group event timings at 8M (no ncu)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|^FAILED|^E  .*(Error|err )" gpurun_out/pytest.log | head -20
SPHSM_GROUPS=1 timeout 300 python tools/profile_step.py --workload 8m --steps 10 --warmup 3 2>&1 | tee gpurun_out/quick.log
