#!/usr/bin/env python
"""The slab step at per-rank size on ONE GPU: R virtual ranks (sphsm_step_group) over the benchmark lattice, a few steps.
Run under `ncu --metrics gpu__time_duration.sum` to list what one rank's kernels cost at 1/R of the workload.
    python tools/profile_slab.py [--workload 8m] [--ranks 8] [--steps 3] [--warmup 2]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sph_sm_monodomain_b200 import LocalGroup, Sim, slabs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="8m")
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
a = ap.parse_args()
wl = bench.parse_workload(a.workload)
pos, world, fixed, stim = bench.workload_inputs(wl)
axis = slabs.slab_axis_for(world)
npl = slabs.num_planes(world, axis)
parts = slabs.partition_planes(slabs.plane_histogram(pos, axis, npl), a.ranks)
sims = []
for r in range(a.ranks):
    s = Sim(capacity=len(pos), world=world, diagnostics=False, slab_axis=axis)
    s.Init_Fluid(pos)
    s.set_masks(fixed, stim)
    if wl["quadratic"]:
        s.flip_quadratic()
    sims.append(s)
grp = LocalGroup(sims)
for s, (lo, hi) in zip(sims, parts):
    s.set_slab(lo, hi)
grp.step(a.warmup)
for s in sims:
    s.sync()
t0 = time.perf_counter()
grp.step(a.steps)
for s in sims:
    s.sync()
dt = time.perf_counter() - t0
print("virtual ranks", a.ranks, "wall ms/step (all ranks serialised on one GPU)", 1e3 * dt / a.steps,
      "owned", [s.comm_info()["own_end"] - s.comm_info()["own_begin"] for s in sims])
