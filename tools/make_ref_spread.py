#!/usr/bin/env python
"""Measure the REFERENCE's own build-to-build trajectory spread: the genuine class compiled with the flags its
Makefile ships (-Ofast) against the IEEE-strict -O2 build that generated tests/golden/*.npz.  The long-trajectory
parity test bounds the CUDA path's deviation by a small multiple of this spread (the dynamics amplify rounding
differences: stiff pressure term, rho ~ 30x rest density).  Writes tests/golden/ref_spread.json.  Runs only where
oracle/_ref exists (this container)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import CpuSim  # noqa: E402
from tests.common import CONFIGS, advance_to, load_golden, setup_from_golden  # noqa: E402

out = {}
for name, cfg in CONFIGS.items():
    if cfg["quadratic"]:
        continue
    g, kw = load_golden(name)
    sim = CpuSim("ref_ofast", **kw)
    setup_from_golden(sim, g, False)
    done, rows = 0, {}
    for cp in [int(c) for c in g["checkpoints"]]:
        done = advance_to(sim, g, done, cp)
        p = sim.particles()
        dpos = np.abs(p["pos"].astype(np.float64) - g[f"step{cp}.pos"])
        dvm = np.abs(p["Vm"].astype(np.float64) - g[f"step{cp}.Vm"])
        rows[str(cp)] = {"pos_max": float(dpos.max()), "pos_mean": float(dpos.mean()), "vm_max": float(dvm.max()), "vm_mean": float(dvm.mean())}
        print(name, cp, rows[str(cp)], flush=True)
    out[name] = rows
with open(os.path.join(ROOT, "tests", "golden", "ref_spread.json"), "w") as fh:
    json.dump({"what": "genuine reference, g++ -Ofast (its Makefile's flags) vs g++ -O2 -ffp-contract=off (the golden vectors)", "spread": out}, fh, indent=1)
