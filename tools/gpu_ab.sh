#!/bin/bash
# GPU tests, then the kernel-path A/B (bit-identity + timing)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -8 gpurun_out/pytest.log
timeout 900 python tools/kernel_ab.py --workload 8m --steps 10 --skip-check --variants gen4,gen6_t128_s2 > gpurun_out/kernel_ab.log 2>&1; echo "ab exit $?" >> gpurun_out/kernel_ab.log
cat gpurun_out/kernel_ab.log | tail -10
