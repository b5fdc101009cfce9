"""Shared replay logic: drive any simulation object that follows the reference class's method names
(oracle.CpuSim on the CPU, sph_sm_monodomain_b200.Sim on the GPU) through a golden configuration."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STATE = ("pos", "vel", "dens", "pres", "Vm", "Inter_Vm", "Iion", "w", "stim")
STAGE_OUT = {
    2: ("predicted_vel", "goal", "corrected_vel"),
    3: ("inter_vel",),
    4: ("dens", "pres"),
    5: ("Iion", "w"),
    6: ("acc", "Inter_Vm"),
    7: ("vel", "pos", "Vm"),
}

# name -> (quadratic, kwargs for the simulation ctor)
CONFIGS = {
    "cfg1_4944": dict(quadratic=False),
    "cfg2_5211": dict(quadratic=False),
    "cfg2_5211_wave": dict(quadratic=False),
    "cube_4913": dict(quadratic=False),
    "cube_4913_quadratic": dict(quadratic=True),
    "cfg1_4944_quadratic": dict(quadratic=True),
    "lattice_24x10x12": dict(quadratic=False),
}


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    kw = {}
    if name.startswith("lattice"):
        world = np.load(os.path.join(GOLDEN, name + ".world.npy"))
        kw = dict(capacity=len(g["positions"]), world=tuple(float(w) for w in world))
    return g, kw


def setup_from_golden(sim, g, quadratic):
    """Initial state exactly as tools/make_golden.py set it: positions, then the fixed / stim masks."""
    sim.Init_Fluid(g["positions"])
    sim.set_fields(fixed=g["init.fixed"], stim=g["init.stim"])
    if quadratic:
        assert sim.flip_quadratic()


def advance_to(sim, g, done, target):
    off = int(g["stim_off_before"])
    while done < target:
        if off >= 0 and done == off:
            sim.turnOffStim()
        nxt = target
        if off >= 0 and done < off < target:
            nxt = off
        sim.Animation(nxt - done)
        done = nxt
    return done


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def rel_err(x, ref, scale=None):
    """max |x-ref| / max(|ref|, field scale) — the tolerance metric of SURVEY.md §8c."""
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    if scale is None:
        scale = np.abs(ref).max() if ref.size else 1.0
    den = np.maximum(np.abs(ref), max(float(scale), 1e-30))
    return float((np.abs(x - ref) / den).max()) if ref.size else 0.0
