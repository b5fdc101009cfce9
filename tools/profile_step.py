#!/usr/bin/env python
"""Short, deterministic driver for ncu: build the benchmark lattice, run a few fused steps through the C-ABI.
    python tools/profile_step.py [--workload 8m] [--steps 3] [--warmup 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sph_sm_monodomain_b200 import Sim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="8m")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
wl = bench.parse_workload(a.workload)
pos, world, fixed, stim = bench.workload_inputs(wl)
sim = Sim(capacity=len(pos), world=world, diagnostics=False)
sim.Init_Fluid(pos)
sim.set_masks(fixed, stim)
if wl["quadratic"]:
    sim.flip_quadratic()
sim.Animation(a.warmup)
sim.sync()
sim.Animation(a.steps)
sim.sync()
print("ms/step", sim.last_step_ms() / a.steps, "launches", sim.launch_count())
if os.environ.get("SPHSM_GROUPS"):
    for k, v in sim.profile_step(5).items():
        print(f"  {k:45s} {v * 1e3:9.1f} us")
