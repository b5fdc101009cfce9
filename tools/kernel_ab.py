#!/usr/bin/env python
"""A/B of the neighbour-pass kernel paths on one GPU, through the C-ABI:

  1. bit-identity: every path (gathered generation 4, staged generation 6 with 64 / 128 targets per block, 2 / 4 candidates
     per iteration, staging off) must leave every field of every particle bit-identical after a few steps, on a jittered
     lattice (staged blocks), on the reference's dense mesh with the warp path off (blocks that do not fit: gathered inside the
     generation-6 kernel) and on a sparse cloud (key ranges too long to stage);
  2. timing: per-kernel-group CUDA-event times of each path on the benchmark workload.

    python tools/kernel_ab.py [--workload 8m] [--steps 10] [--skip-check]
Writes gpurun_out/kernel_ab.json.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sph_sm_monodomain_b200 import Sim, inputs  # noqa: E402
from sph_sm_monodomain_b200.sim import tune  # noqa: E402

VARIANTS = {
    "gen4": dict(**{"pass": 4}),
    "gen6_t128_s2": dict(**{"pass": 6}, stage6=1, t6=128, b_step6=2),
    "gen6_t128_s4": dict(**{"pass": 6}, stage6=1, t6=128, b_step6=4),
    "gen6_t64_s2": dict(**{"pass": 6}, stage6=1, t6=64, b_step6=2),
    "gen6_t64_s4": dict(**{"pass": 6}, stage6=1, t6=64, b_step6=4),
    "gen6_unstaged": dict(**{"pass": 6}, stage6=0, t6=128, b_step6=2),
}


def apply(v):
    for k, val in VARIANTS[v].items():
        tune(k, val)


def run_case(name, pos, world, fixed, stim, quadratic, steps, warp_path=1):
    out = {}
    ref = None
    tune("warp_path", warp_path)
    for v in VARIANTS:
        apply(v)
        sim = Sim(capacity=len(pos), world=world, diagnostics=True)
        sim.Init_Fluid(pos)
        sim.set_masks(fixed, stim)
        if quadratic:
            sim.flip_quadratic()
        sim.Animation(steps)
        p = sim.particles()
        sim.close()
        if ref is None:
            ref = p
            out[v] = "reference"
        else:
            bad = [f for f in p.dtype.names if f != "_pad" and p[f].tobytes() != ref[f].tobytes()]
            out[v] = "bit-identical" if not bad else {f: float(np.abs(p[f].astype(np.float64) - ref[f]).max()) for f in bad}
        print(f"[check] {name:28s} {v:16s} {out[v]}", flush=True)
    tune("warp_path", 1)
    return out


def checks():
    res = {}
    # jittered lattice, linear and quadratic (staged blocks; rows of 1, 2 and 4 lattice lines)
    pos, world = inputs.lattice(96, 40, 40, jitter=0.05)
    fixed, stim = inputs.lattice_masks(pos, 96, 8)
    stimv = np.where(stim, np.float32(300.0), np.float32(0.0)).astype(np.float32)
    res["jitter_96x40x40_linear"] = run_case("jitter_96x40x40_linear", pos, world, fixed.astype(np.uint8), stimv, False, 4)
    res["jitter_96x40x40_quadratic"] = run_case("jitter_96x40x40_quadratic", pos, world, fixed.astype(np.uint8), stimv, True, 4)
    # the reference's dense mesh through the thread-per-particle kernels (warp path off): blocks too full to stage
    g = np.load(os.path.join(bench.ROOT, "tests", "golden", "cfg2_5211_wave.npz"))
    res["cfg2_wave_dense"] = run_case("cfg2_wave_dense", np.ascontiguousarray(g["positions"], np.float32), (1.5, 1.5, 1.5),
                                      g["init.fixed"].astype(np.uint8), g["init.stim"].astype(np.float32), False, 4, warp_path=0)
    # a sparse random cloud: key ranges of a block span thousands of cells
    rng = np.random.default_rng(7)
    cloud = (rng.random((20000, 3), dtype=np.float32) * np.float32(1.45) + np.float32(0.02)).astype(np.float32)
    res["sparse_cloud_20k"] = run_case("sparse_cloud_20k", cloud, (1.5, 1.5, 1.5), np.zeros(20000, np.uint8),
                                       np.where(cloud[:, 0] < 0.3, np.float32(300), np.float32(0)).astype(np.float32), False, 3)
    return res


def timing(workload, steps, variants):
    wl = bench.parse_workload(workload)
    pos, world, fixed, stim = bench.workload_inputs(wl)
    res = {}
    for v in variants:
        apply(v)
        sim = Sim(capacity=len(pos), world=world, diagnostics=False)
        sim.Init_Fluid(pos)
        sim.set_masks(fixed, stim)
        if wl["quadratic"]:
            sim.flip_quadratic()
        sim.Animation(5)
        sim.sync()
        sim.Animation(steps)
        sim.sync()
        ms = sim.last_step_ms() / steps
        groups = sim.profile_step(5)
        sim.close()
        res[v] = dict(ms_per_step=ms, groups_us={k: round(val * 1e3, 1) for k, val in groups.items() if val > 0})
        print(f"[time] {workload} {v:16s} {ms:.4f} ms/step  " + "  ".join(f"{k.split('(')[0]}={val * 1e3:.0f}" for k, val in groups.items() if val > 0),
              flush=True)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="8m")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--variants", default=",".join(VARIANTS))
    a = ap.parse_args()
    out = {}
    if not a.skip_check:
        out["checks"] = checks()
    out["timing"] = {a.workload: timing(a.workload, a.steps, a.variants.split(","))}
    os.makedirs(os.path.join(bench.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(bench.ROOT, "gpurun_out", "kernel_ab.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    ok = all(v in ("reference", "bit-identical") for c in out.get("checks", {}).values() for v in c.values())
    print("ALL BIT-IDENTICAL" if ok else "MISMATCH")
