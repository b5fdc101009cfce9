#!/bin/bash
# round-end style validation on one GPU: smoke, GPU tests, default bench (own arm + reference arm), then the ncu launch
# list and the full captures of the two neighbour passes (each after its command has run once without ncu)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_8m.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_8m.log
tail -2 gpurun_out/bench_8m.log | cut -c1-600
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_ref.log | cut -c1-400
bash tools/gpu_profile.sh notest
