#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale_parity.py tests/test_gpu_slabs.py -m gpu -q --timeout 600 > gpurun_out/pytest_b1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_b1.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest_b1.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench_n1.log; tail -5 gpurun_out/bench_n1.err
timeout 300 python bench.py --steps 100 --warmup 3 --workload 32m --no-cpu-baseline > gpurun_out/bench_32m_n1.log 2>&1; echo "bench32 exit $?"; tail -c 1500 gpurun_out/bench_32m_n1.log
