"""Host-side slab decomposition for the multi-GPU path: which axis to cut, the cell plane of a position
(exactly as the device computes it), per-plane particle counts and the balanced partition into contiguous plane ranges.
Pure numpy; the counts are combined across ranks by the caller (torch.distributed all_reduce in bench.py)."""
from __future__ import annotations

import numpy as np

KERNEL_H = np.float32(0.04)


def slab_axis_for(world) -> int:
    """Cut along the longest axis (ties -> the lowest axis index)."""
    return int(np.argmax(np.asarray(world, dtype=np.float64)))


def num_planes(world, axis: int, h=KERNEL_H) -> int:
    """Grid_Size along the axis, ceil(World / Cell_Size) in float like the reference ctor (cpp:32-35)."""
    return int(np.ceil(np.float32(world[axis]) / np.float32(h)))


def plane_of(pos: np.ndarray, axis: int, h=KERNEL_H) -> np.ndarray:
    """Cell plane of every position: float32 division by Cell_Size and C truncation (cpp:127-134)."""
    return (np.asarray(pos, np.float32)[:, axis] / np.float32(h)).astype(np.int64)


def plane_histogram(pos: np.ndarray, axis: int, nplanes: int, h=KERNEL_H) -> np.ndarray:
    pl = plane_of(pos, axis, h)
    pl = pl[(pl >= 0) & (pl < nplanes)]
    return np.bincount(pl, minlength=nplanes).astype(np.int64)


def partition_planes(hist: np.ndarray, nranks: int):
    """Contiguous plane ranges [(lo, hi)] covering [0, len(hist)), every rank at least one plane, particle counts as
    even as plane granularity allows (cut where the prefix sum crosses k/nranks of the total)."""
    hist = np.asarray(hist, np.int64)
    nplanes = len(hist)
    if nranks < 1 or nranks > nplanes:
        raise ValueError(f"cannot cut {nplanes} planes into {nranks} slabs")
    prefix = np.concatenate([[0], np.cumsum(hist)])
    total = int(prefix[-1])
    cuts = [0]
    for k in range(1, nranks):
        target = total * k / nranks
        c = int(np.searchsorted(prefix, target, side="left"))
        if c > 0 and abs(prefix[c - 1] - target) <= abs(prefix[min(c, nplanes)] - target):
            c -= 1
        c = max(c, cuts[-1] + 1)             # at least one plane per rank so far
        c = min(c, nplanes - (nranks - k))   # and one for each rank still to come
        cuts.append(c)
    cuts.append(nplanes)
    return [(cuts[k], cuts[k + 1]) for k in range(nranks)]
