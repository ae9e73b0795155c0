"""Multi-GPU host logic for the dense-grid path: slab sharding along the slowest-varying volume
axis (x in the reference layout flat=(i*Ry+j)*Rz+k, main.py:364), the halo exchange marching cubes needs, and the mesh
merge. One process per GPU; torch.distributed carries the single exchange step (an all-gather of boundary planes).
The arithmetic stays in the CUDA library: this module only moves planes and renumbers vertex ids.

Seam rule: a vertex belongs to the rank that owns the lower voxel of its grid edge. A slab's faces may reference
vertices of the first plane of the next slab; the kernel numbers those right after the slab's own vertices (in the
next rank's canonical order), so global id = base[rank+1] + (local id - n_owned) and no welding is needed.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

HALO_LO = 2   # planes below a slab needed by the Sobel/trilinear normal stencil (recon_util.py:9-48 incl. the +0.5 voxel quirk)
HALO_HI = 3   # planes above: 1 for the cells, +2 for vertex ids of the next rank's first plane and the normal stencil


def slab_range(rx: int, world: int, rank: int) -> Tuple[int, int]:
    """Planes [start, end) of `rank`: contiguous, sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(rx, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def halo_planes(rx: int, start: int, end: int) -> Tuple[int, int]:
    """(lo, hi) halo widths actually available for the slab [start, end) of a volume with rx planes."""
    return min(HALO_LO, start), min(HALO_HI, rx - end)


def merge_meshes(parts: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]]):
    """Concatenate per-rank (verts, faces, normals) into the single-volume mesh. Face indices >= the rank's own vertex
    count refer to the next rank's first vertices (see module docstring)."""
    counts = [p[0].shape[0] for p in parts]
    base = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    faces = []
    for r, (v, f, n) in enumerate(parts):
        f = f.astype(np.int64)
        own = counts[r]
        g = np.where(f < own, f + base[r], f - own + base[min(r + 1, len(parts))])
        faces.append(g)
    verts = np.concatenate([p[0] for p in parts], 0)
    normals = np.concatenate([p[2] for p in parts], 0) if parts[0][2] is not None else None
    return verts, np.concatenate(faces, 0).astype(np.int32), normals


def exchange_halo(slab, rank: int, world: int, rx: int):
    """The path's one exchange step: a single all-gather of each rank's boundary planes (first HALO_HI, last HALO_LO),
    after which every rank holds [lo halo | own planes | hi halo] (NCCL over NVLink on GPUs, gloo in the CPU tests).
    `slab` is this rank's (nx, Ry, Rz) block of the volume; nx >= HALO_HI is required."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return slab
    nx = slab.shape[0]
    if nx < HALO_HI:
        raise ValueError('slab of %d planes is thinner than the halo (%d)' % (nx, HALO_HI))
    send = torch.cat([slab[:HALO_HI], slab[nx - HALO_LO:]], 0).contiguous()          # (HALO_HI + HALO_LO, Ry, Rz)
    gathered = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(gathered, send)
    start, end = slab_range(rx, world, rank)
    lo, hi = halo_planes(rx, start, end)
    parts = []
    if lo:
        parts.append(gathered[rank - 1][HALO_HI + (HALO_LO - lo):])                 # previous rank's last `lo` planes
    parts.append(slab)
    if hi:
        parts.append(gathered[rank + 1][:hi])                                       # next rank's first `hi` planes
    return torch.cat(parts, 0).contiguous()
