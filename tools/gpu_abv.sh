#!/bin/bash
mkdir -p gpurun_out
for k in base A8 B7 A10 base A8 B7 A10; do
  if [ $k = base ]; then unset SPHSM_LIB_PATH; else export SPHSM_LIB_PATH=$PWD/sph_sm_monodomain_b200/libsphsm_b200_$k.so; fi
  echo "variant=$k"; timeout -s KILL 300 python tools/kernel_ab.py --workload 8m --steps 10 --skip-check --variants gen4 2>&1 | grep "\[time\]"
done
