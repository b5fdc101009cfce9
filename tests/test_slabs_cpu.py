"""CPU: host-side logic of the multi-GPU path — slab axis, plane histogram, balanced partition — and the N>1 plumbing
(world_size-2 gloo): every rank histograms its share of the particles, the counts are all-reduced, and all ranks derive
the same slabs that cover every particle exactly once."""
import os
import socket

import numpy as np
import pytest

from sph_sm_monodomain_b200 import inputs, slabs


def test_axis_planes_and_plane_of_match_device_rule():
    pos, world = inputs.lattice(30, 6, 5)
    axis = slabs.slab_axis_for(world)
    assert axis == 0
    npl = slabs.num_planes(world, axis)
    assert npl == int(np.ceil(np.float32(world[0]) / np.float32(0.04)))
    pl = slabs.plane_of(pos, axis)
    assert pl.min() >= 0 and pl.max() < npl
    # float32 division + truncation, not float64: a coordinate sitting on a cell face must land where the device puts it
    x = np.float32(0.04) * np.float32(7)
    assert slabs.plane_of(np.array([[x, 0, 0]], np.float32), 0)[0] == int(np.float32(x) / np.float32(0.04))


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_partition_covers_all_planes_and_balances(nranks):
    pos, world = inputs.lattice(64, 9, 9)
    npl = slabs.num_planes(world, 0)
    hist = slabs.plane_histogram(pos, 0, npl)
    assert hist.sum() == len(pos)
    parts = slabs.partition_planes(hist, nranks)
    assert parts[0][0] == 0 and parts[-1][1] == npl
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(hi > lo for lo, hi in parts)
    counts = [int(hist[lo:hi].sum()) for lo, hi in parts]
    assert sum(counts) == len(pos)
    assert max(counts) - min(counts) <= 2 * hist.max()  # plane granularity


def test_partition_degenerate_inputs():
    with pytest.raises(ValueError):
        slabs.partition_planes(np.ones(3, np.int64), 4)
    assert slabs.partition_planes(np.array([0, 0, 10, 0]), 2) in ([(0, 2), (2, 4)], [(0, 3), (3, 4)])
    assert slabs.partition_planes(np.zeros(5, np.int64), 5) == [(k, k + 1) for k in range(5)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        pos, world = inputs.lattice(48, 7, 6)
        axis = slabs.slab_axis_for(world)
        npl = slabs.num_planes(world, axis)
        mine = pos[rank::world_size]  # each rank sees only its share of the input
        hist = torch.from_numpy(slabs.plane_histogram(mine, axis, npl))
        dist.all_reduce(hist)
        parts = slabs.partition_planes(hist.numpy(), world_size)
        lo, hi = parts[rank]
        pl = slabs.plane_of(pos, axis)
        owned = int(((pl >= lo) & (pl < hi)).sum())
        tot = torch.tensor([owned])
        dist.all_reduce(tot)
        # the 128-byte id the NCCL bootstrap would broadcast from rank 0 travels the same way
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident[:] = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        out.put((rank, parts, owned, int(tot[0]), len(pos), bytes(ident.numpy().tobytes())))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_ranks_agree_on_slabs():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, parts0, own0, tot0, n0, id0), (r1, parts1, own1, tot1, n1, id1) = res
    assert parts0 == parts1 and tot0 == tot1 == n0 == n1 == own0 + own1
    assert abs(own0 - own1) <= 2 * 7 * 6 * 2
    assert id0 == id1 == bytes(range(128))
