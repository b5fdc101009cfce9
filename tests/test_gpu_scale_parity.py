"""GPU: value parity on the configurations the performance numbers are quoted on (BASELINE.json configs[2] / configs[3]):
whole fused steps of the production kernels on the 1M lattice (linear shape matching), its seeded-jitter variant, and the 8M
lattice (quadratic), every field against the CPU oracle run with double moment accumulation — the oracle of record for the
summation order at this size (SURVEY.md §7 hard part 3: the reference's sequential float sums carry ~1e-4 of their own
rounding at N >= 1M).  Metric and tolerance classes as tests/test_gpu_parity.py; achieved errors go to the parity report.
"""
import numpy as np
import pytest

from oracle import CpuSim
from tests.common import STAGE_OUT
from tests.test_gpu_parity import TOL, assert_close, field_scale

pytestmark = pytest.mark.gpu

# Steps after the first may contain DISCRETE events that rounding decides: a particle that lands within an ulp of a wall is
# reflected (cpp:620-646: velocity sign flip) in one implementation and not in the other — on the jittered 1M lattice exactly one
# particle does at step 2 (y = 8.887e-4 - 8.8875e-4; the fast path integrates with one FMA, the reference with a multiply and an
# add).  It and the neighbours it then pushes are legitimate outliers of a trajectory comparison, so later steps bound the
# NUMBER of particles beyond the tolerance (ten per million) instead of the maximum; step 1 has no such freedom.
MAX_OUTLIER_FRACTION = 1e-5


def per_particle_err(name, got, ref, params):
    s = field_scale(name, ref, params)
    g, r = got.astype(np.float64).reshape(len(got), -1), ref.astype(np.float64).reshape(len(ref), -1)
    den = np.maximum(np.abs(r), s if s > 0 else 1e-30)
    return (np.abs(g - r) / den).max(axis=1)

CASES = {
    # name: (dims, quadratic, jitter, steps to compare at)
    "1m_linear": ((100, 100, 100), False, 0.0, (1, 3)),
    "1m_linear_jitter": ((100, 100, 100), False, 0.05, (1, 3)),
    "8m_quadratic": ((512, 125, 125), True, 0.0, (1,)),
}


def acc_bound(got, want, consts, jitter, quantile=1.0):
    """Absolute bound on |acc - acc_ref|.  On a lattice near rest the acceleration is a cancelling sum, and its viscosity part
    multiplies velocity DIFFERENCES between neighbours by V * mu * Visco(r) / rho (cpp:559, 568) — ~3 per neighbour here — so it
    amplifies whatever (within-tolerance) difference the intermediate velocities carry (their own origin: the goal positions'
    last-bit rounding times alpha / dt = 97).  The bound is what those measured velocity differences imply, plus TOL of the
    field's infinity norm: acc must not differ by MORE than its inputs explain."""
    h, s = float(consts["kernel"]), float(consts["kernel"]) * 0.9
    r_min = s * (1.0 - 2.0 * jitter)
    rho_min = float(want["dens"].min())
    coupling = 8 * (float(want["mass"].max()) / rho_min) * float(consts["mu"]) * float(consts["Spiky_constant"]) * (h - r_min) / rho_min
    div = np.abs(got["inter_vel"].astype(np.float64) - want["inter_vel"]).max(axis=1)
    div = float(div.max() if quantile >= 1.0 else np.quantile(div, quantile))
    return TOL * float(np.abs(want["acc"]).max()) + 2.0 * coupling * div


def tol_for(field, quadratic, step):
    """Step 1: 1e-5 (2e-5 on the quadratic goal: the truncated 9x9 Jacobi inverse amplifies the moment rounding, Q7) on every
    field the north_star names and on the velocity-like ones scaled as in test_gpu_parity.field_scale.  Later steps: the stiff
    pressure term amplifies rounding differences, same classes as test_fused_step_vs_golden."""
    base = 2 * TOL if (quadratic and field in ("goal", "corrected_vel", "inter_vel", "vel", "acc", "pos")) else TOL
    return base if step == 1 else 20 * base


@pytest.mark.parametrize("case", list(CASES))
def test_fused_steps_at_scale(case, parity_record):
    from sph_sm_monodomain_b200 import Sim, inputs

    dims, quadratic, jitter, steps = CASES[case]
    pos, world = inputs.lattice(*dims, jitter=jitter)
    fixed, stim = inputs.lattice_masks(pos, dims[0], 8)
    stim = np.where(stim, np.float32(300.0), np.float32(0.0)).astype(np.float32)
    n = len(pos)
    ora = CpuSim("port", capacity=n, world=world, moments_in_double=True)
    ora.Init_Fluid(pos)
    ora.set_fields(fixed=fixed.astype(np.uint8), stim=stim)
    sim = Sim(capacity=n, world=world, diagnostics=True)
    sim.Init_Fluid(pos)
    sim.set_masks(fixed, stim)
    if quadratic:
        assert ora.flip_quadratic() and sim.flip_quadratic()
    c = ora.constants()
    params = {"alpha": float(c["alpha"]), "Time_Delta": float(c["Time_Delta"]), "pos_scale": float(np.abs(pos).max())}
    done = 0
    for target in steps:
        ora.Animation(target - done)
        sim.Animation(target - done)
        done = target
        got, want = sim.particles(), ora.particles()
        assert np.array_equal(got["fixed"], want["fixed"]) and np.array_equal(got["stim"], want["stim"])
        for st in range(2, 8):
            for f in STAGE_OUT[st]:
                tol = tol_for(f, quadratic, target)
                if target > 1:  # (see MAX_OUTLIER_FRACTION)
                    e = per_particle_err(f, got[f], want[f], params) if f != "acc" else np.abs(got["acc"].astype(np.float64) - want["acc"]).max(axis=1)
                    lim = tol if f != "acc" else acc_bound(got, want, c, jitter, quantile=1.0 - MAX_OUTLIER_FRACTION)
                    frac = float((e > lim).mean())
                    typical = float(np.quantile(e, 1.0 - MAX_OUTLIER_FRACTION))
                    parity_record("fused_steps_at_scale", f"{case}/step{target}", f + " (all but 1e-5 of the particles)", typical, lim)
                    parity_record("fused_steps_at_scale", f"{case}/step{target}", f + " (fraction of particles beyond the tolerance)", frac, MAX_OUTLIER_FRACTION)
                    assert frac <= MAX_OUTLIER_FRACTION, (f, frac, float(e.max()))
                    continue
                if f == "acc":
                    bound = acc_bound(got, want, c, jitter)
                    dacc = float(np.abs(got["acc"].astype(np.float64) - want["acc"]).max())
                    parity_record("fused_steps_at_scale", f"{case}/step{target}", "acc (absolute; bound = what the velocity differences imply)", dacc, bound)
                    assert dacc <= bound, (dacc, bound)
                    continue
                err = assert_close(f, got[f], want[f], params, tol=tol)
                parity_record("fused_steps_at_scale", f"{case}/step{target}", f, err, tol)
    ora.close()
    sim.close()
