#!/bin/bash
# new tests, then a full ncu capture of the generation-6 neighbour passes (one launch each) after a plain run of the same command
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_scale_parity.py -m gpu -q --timeout 600 -k "midrun or staged or scale" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
CMD="python tools/profile_step.py --workload 8m --steps 3 --warmup 2"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
tail -1 gpurun_out/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass_b -s 2 -c 1 -f -o gpurun_out/prof_pass_b $CMD > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass_a -s 2 -c 1 -f -o gpurun_out/prof_pass_a $CMD > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/*.ncu-rep
