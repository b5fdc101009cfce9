"""CPU: the C restatement (oracle/sphsm_oracle.c) against the committed golden vectors that
tools/make_golden.py generated from the genuine reference — bit-exact on every field."""
import numpy as np
import pytest

from oracle import CpuSim
from tests.common import CONFIGS, STAGE_OUT, STATE, advance_to, bits_equal, load_golden, setup_from_golden


@pytest.fixture(scope="module")
def kats(golden_dir):
    return dict(np.load(golden_dir + "/kats.npz"))


def test_constants_bit_exact(kats):
    sim = CpuSim("port")
    c = sim.constants()
    for name, val in zip(kats["const_names"], kats["const_values"]):
        assert np.float32(c[str(name)]).tobytes() == np.float32(val).tobytes(), name
    # SURVEY.md §3.5 hex values
    assert float(c["Time_Delta"]).hex() == "0x1.9398dc0000000p-9"
    assert float(c["Poly6_constant"]).hex() == "0x1.5bdf8c0000000p+42"
    assert float(c["Spiky_constant"]).hex() == "0x1.a0e1b80000000p+31"
    assert float(c["B_spline_constant"]).hex() == "0x1.36d97a0000000p+12"


def test_smoothing_kernels_bit_exact(kats):
    sim = CpuSim("port")
    assert bits_equal(np.asarray([sim.Poly6(x) for x in kats["r2"]], np.float32), kats["poly6"])
    assert bits_equal(np.asarray([sim.Spiky(x) for x in kats["r"]], np.float32), kats["spiky"])
    assert bits_equal(np.asarray([sim.Visco(x) for x in kats["r"]], np.float32), kats["visco"])
    assert bits_equal(np.asarray([sim.B_spline_2(x) for x in kats["r"]], np.float32), kats["bspline2"])
    assert sim.Poly6(0.0) == pytest.approx(24479.4004, rel=1e-7)  # SURVEY.md §8c known answer


def test_cell_hash_bit_exact(kats):
    sim = CpuSim("port")
    got = np.asarray([sim.cell_hash(*p) for p in kats["hash_points"]], np.int32)
    assert np.array_equal(got, kats["hash_values"])
    assert (got == -1).any() and (got >= 0).any()


def test_small_matrices_bit_exact(kats):
    sim = CpuSim("port")
    for m, r in zip(kats["m3"], kats["m3_polar"]):
        assert bits_equal(sim.polar3(m), r)
    for m, ok, inv in zip(kats["m3"], kats["m3_inv_ok"], kats["m3_inv"]):
        o, i = sim.invert3(m)
        assert o == bool(ok) and bits_equal(i, inv)
    assert not kats["m3_inv_ok"][2]  # singular matrix: invert() is a no-op
    for m, inv in zip(kats["m9"], kats["m9_inv"]):
        assert bits_equal(sim.invert9(m), inv)


def test_isolated_particle_density():
    """Double self term (Q1): an isolated particle has rho = 2 * 0.2 * Poly6(0) = 9791.76."""
    sim = CpuSim("port")
    sim.Init_Fluid(np.asarray([[0.5, 0.5, 0.5], [1.0, 1.0, 1.0]], np.float32))
    sim.stage("Find_neighbors")
    sim.stage("Compute_Density_SingPressure")
    p = sim.particles()
    assert p["dens"][0] == pytest.approx(9791.76, rel=1e-6)
    assert np.signbit(p["pres"][0]) and p["pres"][0] == 0.0  # stim == 0 -> pres = -0.0f (Q2)


# trajectory lengths: the long ones dominate the CPU suite (~35 s per 1000 steps at N~5k)
@pytest.mark.parametrize("name", list(CONFIGS))
def test_replay_golden_bit_exact(name):
    g, kw = load_golden(name)
    quadratic = CONFIGS[name]["quadratic"]
    sim = CpuSim("port", **kw)
    setup_from_golden(sim, g, quadratic)
    # step 1, stage by stage
    sim.stage(1)
    start, idx = sim.cells_csr()
    occ = np.flatnonzero(np.diff(start) > 0)
    assert np.array_equal(occ, g["s1.occupied_cells"])
    assert np.array_equal(np.diff(start)[occ], g["s1.occupied_count"])
    assert np.array_equal(idx, g["s1.bucket_items"])
    for st in range(2, 8):
        sim.stage(st)
        p = sim.particles()
        for f in STAGE_OUT[st]:
            assert bits_equal(p[f], g[f"s1.stage{st}.{f}"]), (st, f)
    done = 1
    for cp in g["checkpoints"]:
        done = advance_to(sim, g, done, int(cp))
        p = sim.particles()
        for f in STATE:
            assert bits_equal(p[f], g[f"step{cp}.{f}"]), (int(cp), f)


def test_known_answers_cfg1():
    """SURVEY.md §8c survey-time known answers for config 1 (genuine reference, -O2)."""
    g, _ = load_golden("cfg1_4944")
    assert g["init.fixed"].sum() == 1675 and (g["init.stim"] > 0).all()
    assert np.allclose(g["step1.pos"][0], [0.481091917, 0.561482131, 0.491041154], rtol=0, atol=1e-9)
    assert g["step1.dens"][0] == pytest.approx(15356.4863, rel=1e-7)
    assert g["step1.pres"][0] == pytest.approx(7122.24316, rel=1e-7)
    assert np.all(g["step1.Vm"] == np.float32(0.0711150765))
    assert g["step1.pres"][4943] == 15000.0
    assert g["step1000.Vm"].mean() == -200.0
