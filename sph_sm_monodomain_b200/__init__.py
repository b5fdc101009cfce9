"""sph_sm_monodomain_b200 — B200 (sm_100a) implementation of the SPH_SM_monodomain per-timestep particle pipeline.

* ``csrc/``      hand-written CUDA kernels + the extern "C" layer (include/sphsm_b200.h) -> libsphsm_b200.so
* ``sim.Sim``    host-side mirror of the reference class API over that C-ABI (the C++ drop-in is include/SPH_SM_monodomain.h)
* ``slabs``     host-side slab partition for the multi-GPU path
* ``inputs``     the reference's CSV rule / lattice generator and the synthetic benchmark lattices
* ``build``      in-tree nvcc build

Importing the package never touches CUDA; constructing a ``Sim`` loads the library and fails loudly if it is
missing or no device is usable (there is no CPU fallback).
"""
from . import inputs, slabs  # noqa: F401
from .sim import PARTICLE_DTYPE, STAGES, LocalGroup, Sim, SphsmError, default_params  # noqa: F401

__all__ = ["Sim", "LocalGroup", "slabs", "SphsmError", "default_params", "inputs", "PARTICLE_DTYPE", "STAGES"]
