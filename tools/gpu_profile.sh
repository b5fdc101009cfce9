#!/bin/bash
# tests, then ncu launch list + full captures of the two neighbour passes (plain run first, as the recipe requires)
mkdir -p gpurun_out
if [ "$1" != "notest" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "passed|failed|^FAILED|^E  " gpurun_out/pytest.log | head -40
fi
CMD="python tools/profile_step.py --workload 8m --steps 3 --warmup 2"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/plain.log
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass_b -s 2 -c 1 -f -o gpurun_out/prof_pass_b $CMD > gpurun_out/ncu_b.log 2>&1
timeout 300 $CMD > gpurun_out/plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass_a -s 2 -c 1 -f -o gpurun_out/prof_pass_a $CMD > gpurun_out/ncu_a.log 2>&1
ls -la gpurun_out/
