#!/usr/bin/env python
"""NCCL slab path vs the single-GPU step on the same input (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/mg_parity.py [--dims 96x24x24] [--steps 30] [--quadratic]

Every rank steps its slab through sphsm_step (ncclSend/ncclRecv halos + migrants, moment ncclAllReduce); rank 0 also runs
the whole set on its own GPU with the canonical in-cell order and compares the gathered positions.  Prints one JSON line
and exits non-zero when the deviation exceeds --tol (default 1e-5 of the field scale; with --canonical every cell on every
rank sums in ascending-id order and the result is expected to be bit-identical)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sph_sm_monodomain_b200 import Sim, inputs, slabs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dims", default="96x24x24")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--quadratic", action="store_true")
ap.add_argument("--jitter", type=float, default=0.05)
ap.add_argument("--tol", type=float, default=1e-5)
ap.add_argument("--canonical", action="store_true", help="ascending-id order in every cell on all ranks (bit-level comparison)")
ap.add_argument("--expect-overflow", type=int, default=0,
                help="halo capacity too small on purpose: every rank must get SphsmError within a few steps, none may hang")
a = ap.parse_args()
rank, world_size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims = tuple(int(v) for v in a.dims.split("x"))
pos, world = inputs.lattice(*dims, jitter=a.jitter)
fixed, stim = inputs.lattice_masks(pos, dims[0], 4)
fixed = fixed.astype(np.uint8)
stimv = np.where(stim, np.float32(300), np.float32(0)).astype(np.float32)
n = len(pos)
axis = slabs.slab_axis_for(world)
npl = slabs.num_planes(world, axis)
parts = slabs.partition_planes(slabs.plane_histogram(pos, axis, npl), world_size)


def make(**kw):
    s = Sim(capacity=n, world=world, device=local, diagnostics=False, slab_axis=axis, **kw)
    s.Init_Fluid(pos)
    s.set_masks(fixed, stimv)
    if a.quadratic:
        s.flip_quadratic()
    return s


if a.expect_overflow:
    # error path: rank-local halo overflow -> reported by EVERY rank at the end of the following step, no rank left in a collective
    from sph_sm_monodomain_b200 import SphsmError

    sim = make(halo_capacity=a.expect_overflow)
    ident = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    sim.comm_init(world_size, rank, bytes(ident.cpu().numpy().tobytes()))
    sim.set_slab(*parts[rank])
    failed_at, msg = -1, ""
    for k in range(6):
        try:
            sim.Animation(1)
            sim.sync()
        except SphsmError as e:
            failed_at, msg = k, str(e)
            break
    t = torch.tensor([failed_at], device=f"cuda:{local}")
    got = [torch.zeros_like(t) for _ in range(world_size)]
    dist.all_gather(got, t)
    steps_failed = [int(x.item()) for x in got]
    ok = all(s >= 0 for s in steps_failed) and len(set(steps_failed)) == 1
    if rank == 0:
        print(json.dumps({"mg_error_path": "ok" if ok else "FAIL", "ranks": world_size, "failed_at_step": steps_failed, "message_rank0": msg[:160]}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

sim = make()
if a.canonical:
    pp = sim.get_params()
    pp.reserved[1] = 1
    sim._ck(sim.lib.sphsm_set_params(sim.h, pp))
ident = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
if rank == 0:
    ident.copy_(torch.frombuffer(bytearray(Sim.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(ident, 0)
sim.comm_init(world_size, rank, bytes(ident.cpu().numpy().tobytes()))
sim.set_slab(*parts[rank])
sim.Animation(a.steps)
sim.sync()
ids, xyz = sim.download_owned()
# gather (padded) on the GPU
cnt = torch.tensor([len(ids)], device=f"cuda:{local}")
cnts = [torch.zeros_like(cnt) for _ in range(world_size)]
dist.all_gather(cnts, cnt)
cap = int(max(c.item() for c in cnts))
pad_ids = torch.full((cap,), -1, dtype=torch.int32, device=f"cuda:{local}")
pad_xyz = torch.zeros((cap, 3), dtype=torch.float32, device=f"cuda:{local}")
pad_ids[: len(ids)] = torch.from_numpy(ids).to(pad_ids.device)
pad_xyz[: len(ids)] = torch.from_numpy(xyz).to(pad_xyz.device)
all_ids = [torch.zeros_like(pad_ids) for _ in range(world_size)]
all_xyz = [torch.zeros_like(pad_xyz) for _ in range(world_size)]
dist.all_gather(all_ids, pad_ids)
dist.all_gather(all_xyz, pad_xyz)
rc = 0
if rank == 0:
    got = np.full((n, 3), np.nan, np.float32)
    seen = np.zeros(n, np.int32)
    for i, x in zip(all_ids, all_xyz):
        i, x = i.cpu().numpy(), x.cpu().numpy()
        m = i >= 0
        got[i[m]] = x[m]
        seen[i[m]] += 1
    single = make()
    p = single.get_params()
    p.reserved[1] = 1
    single._ck(single.lib.sphsm_set_params(single.h, p))
    single.Animation(a.steps)
    i1, x1 = single.download_owned()
    ref = np.empty((n, 3), np.float32)
    ref[i1] = x1
    err = float(np.abs(got.astype(np.float64) - ref).max() / np.abs(ref).max()) if (seen == 1).all() else float("nan")
    ok = bool((seen == 1).all() and err <= a.tol)
    print(json.dumps({"mg_parity": "ok" if ok else "FAIL", "ranks": world_size, "particles": n, "steps": a.steps, "quadratic": a.quadratic,
                      "max_rel_dev_vs_single_gpu": err, "owned_per_rank": [int(c.item()) for c in cnts],
                      "every_particle_owned_once": bool((seen == 1).all()), "bit_identical": bool(np.array_equal(got, ref)),
                      "push_exchange": sim.push_exchange(), "push_allreduce": sim.push_allreduce()}), flush=True)
    rc = 0 if ok else 1
dist.barrier()
dist.destroy_process_group()
sys.exit(rc)
