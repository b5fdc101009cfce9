#!/bin/bash
# the drop-in class behind the headless replay of main.cpp's protocol on the reference's own 5211-particle set: 500 steps, stimulation
# off at half time, Get_Paticles() every frame (what display_points does), report line of cpp:785-792
mkdir -p gpurun_out
python - <<'PY'
import numpy as np
g = np.load("tests/golden/cfg2_5211.npz")
g["positions"].astype("<f4").tofile("gpurun_out/cfg2.xyz")
PY
python -m sph_sm_monodomain_b200.build > /dev/null 2>&1
for mode in "" "--staged"; do
  ./sph_sm_monodomain_b200/sphsm_headless --xyz gpurun_out/cfg2.xyz --steps 500 $mode > gpurun_out/headless_cfg2$mode.log 2>&1
  echo "mode '$mode':"; grep ";" gpurun_out/headless_cfg2$mode.log | tail -1; grep headless gpurun_out/headless_cfg2$mode.log | tail -1
done
