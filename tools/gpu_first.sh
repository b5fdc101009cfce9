#!/bin/bash
# first GPU contact: smoke, tests, short bench (each under its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 600 python bench.py --workload 8m --steps 20 --warmup 3 > gpurun_out/bench_8m.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_8m.log
tail -3 gpurun_out/bench_8m.log
