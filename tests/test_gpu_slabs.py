"""GPU (one device): the multi-GPU slab step run as VIRTUAL RANKS (sphsm_comm_init_local / sphsm_step_group — the same
phases the NCCL path runs, collectives as device copies) against the single-GPU step on the same input.

With the canonical in-cell order everywhere (ascending original index, params.reserved[1]) every neighbour sum visits its
candidates in the same order on both sides, so halo / migrant handling is bit-identical; only the shape-matching moment
sums are combined in a different order (per-rank partial sums added on the host), which can move the goal positions by an
ulp.  Bound: 2e-6 of the field scale after 25 steps.  The production default orders only the planes next to a slab face
(that is all the exchange needs); interior cells then sum in arrival order and the run differs from the single-GPU one by
summation-order rounding: bound 1e-5.  Ownership must partition the particles in both modes."""
import numpy as np
import pytest

from sph_sm_monodomain_b200 import inputs, slabs
from tests.common import rel_err

pytestmark = pytest.mark.gpu


def make(Sim, pos, world, fixed, stim, quadratic, **kw):
    p_kw = dict(capacity=len(pos), world=world, diagnostics=False, slab_axis=0)
    p_kw.update(kw)
    s = Sim(**p_kw)
    s.Init_Fluid(pos)
    s.set_masks(fixed.astype(np.uint8), np.where(stim, np.float32(300), np.float32(0)).astype(np.float32))
    if quadratic:
        s.flip_quadratic()
    return s


def setup_case(dims, jitter):
    pos, world = inputs.lattice(*dims, jitter=jitter)
    fixed, stim = inputs.lattice_masks(pos, dims[0], 4)
    return pos, world, fixed, stim


def canonical_everywhere(sim):
    p = sim.get_params()
    p.reserved[1] = 1  # in-cell order = ascending original index in EVERY cell
    sim._ck(sim.lib.sphsm_set_params(sim.h, p))


@pytest.mark.parametrize("canonical", [True, False], ids=["canonical", "default"])
@pytest.mark.parametrize("nranks,quadratic,jitter", [(2, False, 0.0), (3, True, 0.05), (4, False, 0.05)])
def test_virtual_ranks_match_single_gpu(nranks, quadratic, jitter, canonical):
    from sph_sm_monodomain_b200 import LocalGroup, Sim

    dims = (40, 9, 8)
    steps = 25
    pos, world, fixed, stim = setup_case(dims, jitter)
    n = len(pos)
    single = make(Sim, pos, world, fixed, stim, quadratic)
    canonical_everywhere(single)
    single.Animation(steps)
    ids1, xyz1 = single.download_owned()
    ref = np.empty((n, 3), np.float32)
    ref[ids1] = xyz1
    assert len(ids1) == n

    npl = slabs.num_planes(world, 0)
    parts = slabs.partition_planes(slabs.plane_histogram(pos, 0, npl), nranks)
    sims = [make(Sim, pos, world, fixed, stim, quadratic) for _ in range(nranks)]
    grp = LocalGroup(sims)
    for s, (lo, hi) in zip(sims, parts):
        if canonical:
            canonical_everywhere(s)  # default: only the planes on either side of a slab face are put in canonical order
        s.set_slab(lo, hi)
    own0 = [s.comm_info() for s in sims]
    assert sum(i["own_end"] - i["own_begin"] for i in own0) == n
    grp.step(steps)
    got, owner = grp.gather_positions(n)
    assert (owner >= 0).all(), "some particle is owned by no rank"
    # ownership follows the plane each particle had at the last sort: at most one plane off after the final integration
    pl = slabs.plane_of(got, 0)
    for r, (lo, hi) in enumerate(parts):
        assert ((pl[owner == r] >= lo - 1) & (pl[owner == r] < hi + 1)).all()
    err = rel_err(got, ref)
    assert err <= (2e-6 if canonical else 1e-5), err
    infos = [s.comm_info() for s in sims]
    for r, i in enumerate(infos):  # interior ranks carry two halo planes, end ranks one
        halos = (i["own_begin"] > 0) + (i["n_local"] > i["own_end"])
        assert halos == (r > 0) + (r < nranks - 1)


def test_migration_across_faces_and_mask_updates():
    """Particles pushed across slab faces (strong stimulation, many steps) change owner; stim_off and per-step set_masks keep
    working on the distributed state; the slab run stays on the single-GPU trajectory."""
    from sph_sm_monodomain_b200 import LocalGroup, Sim

    dims = (40, 9, 8)
    pos, world, fixed, stim = setup_case(dims, 0.0)
    n = len(pos)
    npl = slabs.num_planes(world, 0)
    parts = slabs.partition_planes(slabs.plane_histogram(pos, 0, npl), 2)
    single = make(Sim, pos, world, fixed, stim, False)
    p = single.get_params()
    p.reserved[1] = 1
    single._ck(single.lib.sphsm_set_params(single.h, p))
    sims = [make(Sim, pos, world, fixed, stim, False) for _ in range(2)]
    grp = LocalGroup(sims)
    for s, (lo, hi) in zip(sims, parts):
        s.set_slab(lo, hi)
    owner0 = grp.gather_positions(n)[1]
    stim2 = np.where(pos[:, 0] > pos[:, 0].mean(), np.float32(300), np.float32(0)).astype(np.float32)
    for k in range(6):
        single.Animation(20)
        grp.step(20)
        if k == 2:
            single.turnOffStim()
            for s in sims:
                s.turnOffStim()
        if k == 3:
            single.set_masks(None, stim2)
            for s in sims:
                s.set_masks(None, stim2)
    got, owner = grp.gather_positions(n)
    ids1, xyz1 = single.download_owned()
    ref = np.empty((n, 3), np.float32)
    ref[ids1] = xyz1
    assert (owner >= 0).all()
    assert (owner != owner0).sum() > 0, "no particle migrated: the test does not exercise migration"
    assert rel_err(got, ref) <= 5e-4  # 120 steps: trajectory bound (rounding grows along the stiff dynamics)


def test_slab_errors_are_reported():
    from sph_sm_monodomain_b200 import LocalGroup, Sim, SphsmError

    pos, world, fixed, stim = setup_case((20, 6, 6), 0.0)
    a = make(Sim, pos, world, fixed, stim, False)
    with pytest.raises(SphsmError):
        a.set_slab(0, 5)  # no communicator yet
    b = make(Sim, pos, world, fixed, stim, False, slab_axis=-1)
    with pytest.raises(SphsmError):
        LocalGroup([b])  # reference key order has no slab axis
    c = make(Sim, pos, world, fixed, stim, False)
    grp = LocalGroup([c])
    with pytest.raises(SphsmError):
        grp.step(1)  # slab not applied
    c.set_slab(0, slabs.num_planes(world, 0))
    grp.step(3)  # one rank, no neighbours: plain step through the phase program
    assert c.comm_info()["own_end"] - c.comm_info()["own_begin"] == len(pos)


def test_exchange1_messages_follow_the_face_populations():
    """sphsm_tune("x1_dynamic", 1): exchange-1 messages start at the full halo capacity, shrink to the lagged face population + margin after three unpacked
    exchanges, are sized identically by the two sides of every face, go back to full capacity when the slab is set again, and leave every
    bit of the result where the fixed-capacity messages put it."""
    from sph_sm_monodomain_b200 import LocalGroup, Sim
    from sph_sm_monodomain_b200.sim import tune

    dims = (40, 9, 8)
    pos, world, fixed, stim = setup_case(dims, 0.05)
    n = len(pos)
    npl = slabs.num_planes(world, 0)
    parts = slabs.partition_planes(slabs.plane_histogram(pos, 0, npl), 3)

    def run(dynamic):
        tune("x1_dynamic", dynamic)
        try:
            sims = [make(Sim, pos, world, fixed, stim, True) for _ in range(3)]
            grp = LocalGroup(sims)
            for s, (lo, hi) in zip(sims, parts):
                s.set_slab(lo, hi)
            cap = sims[0].comm_info()["halo_capacity"]
            sizes = []
            for _ in range(8):
                grp.step(1)
                sizes.append([s.x1_sizes() for s in sims])
            got, _ = grp.gather_positions(n)
            # a new slab (collective) voids the history: full-size messages again
            for s, (lo, hi) in zip(sims, parts):
                s.set_slab(lo, hi)
            grp.step(1)
            after = [s.x1_sizes() for s in sims]
            for s in sims:
                s.close()
            return cap, sizes, after, got
        finally:
            tune("x1_dynamic", 0)  # the default

    cap, sizes, after, got = run(1)
    assert all(v == cap for per in sizes[:3] for s in per for v in s), sizes[:3]
    for per in sizes[3:]:
        for r in range(3):
            to_l, to_r, from_l, from_r = per[r]
            if r > 0:
                assert to_l < cap and from_l < cap and to_l % 256 == 0
                assert from_l == per[r - 1][1] and to_l == per[r - 1][3]  # both sides of the face agree
            if r < 2:
                assert to_r < cap and from_r < cap
    assert all(v == cap for s in after for v in s), after
    cap0, sizes0, _, got0 = run(0)
    assert cap0 == cap and all(v == cap for per in sizes0 for s in per for v in s)
    assert got.tobytes() == got0.tobytes()
