"""Marching-cubes case tables, generated (not transcribed) so they are watertight by construction.

The reference calls skimage.measure.marching_cubes (utils/recon_util.py:64; scikit_image==0.17.2, Lewiner),
whose source is not vendored. This module builds the 256-case table for a classic marching-cubes with a
CONSISTENT face-ambiguity rule, so neighbouring cells always agree on a shared face and the surface is closed:

* corner c of a cell has offset (c&1, (c>>1)&1, (c>>2)&1) in (x,y,z); case bit c is set when value[c] > iso
  (the 'inside' test of skimage: value > level);
* the 12 cell edges are (axis, corner_a) -> corner_b = corner_a | (1<<axis), numbered axis*4 + rank (EDGES below);
* on every cube face the cut edges are paired; an ambiguous face (4 cut edges, diagonal corners inside) always
  separates the INSIDE corners (pairs the two edges adjacent to each inside corner) -- a rule that depends only on
  the face's own corner states, hence watertight across cells;
* the segments close into loops, each loop is fan-triangulated and wound so that the right-hand-rule normal points
  from inside (value > iso) to outside, i.e. along -gradient. (The reference then reverses skimage's face winding,
  recon_util.py:69; avatarcap_b200.mesh documents the final convention.)

`python -m avatarcap_b200.mc_tables` regenerates csrc/mc_tables.inc for the CUDA kernels.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

CORNER_OFFSETS = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.int32)

# edge id = axis*4 + rank; endpoints (a, b) with b = a | (1<<axis)
EDGES: List[Tuple[int, int]] = []
for _axis in range(3):
    for _a in range(8):
        if not (_a >> _axis) & 1:
            EDGES.append((_a, _a | (1 << _axis)))
assert len(EDGES) == 12
EDGE_AXIS = np.array([e // 4 for e in range(12)], dtype=np.int32)
# owner voxel offset of each edge = offset of its lower corner
EDGE_OWNER_OFFSET = np.array([CORNER_OFFSETS[a] for a, _ in EDGES], dtype=np.int32)

# the 6 faces as cyclic corner quadruples
_FACES = []
for _axis in range(3):
    for _side in (0, 1):
        u, v = [ax for ax in range(3) if ax != _axis]
        base = _side << _axis
        _FACES.append([base, base | (1 << u), base | (1 << u) | (1 << v), base | (1 << v)])


def _edge_id(a: int, b: int) -> int:
    lo, hi = min(a, b), max(a, b)
    return EDGES.index((lo, hi))


def _build_case(case: int) -> List[Tuple[int, int, int]]:
    inside = [(case >> c) & 1 for c in range(8)]
    adj = {e: [] for e in range(12)}
    for quad in _FACES:
        cut = []   # (edge id, position i: edge between quad[i] and quad[i+1])
        for i in range(4):
            a, b = quad[i], quad[(i + 1) % 4]
            if inside[a] != inside[b]:
                cut.append((_edge_id(a, b), i))
        if len(cut) == 2:
            adj[cut[0][0]].append(cut[1][0]); adj[cut[1][0]].append(cut[0][0])
        elif len(cut) == 4:
            # ambiguous: pair the two edges adjacent to each inside corner
            for i in range(4):
                if inside[quad[i]]:
                    e_prev = _edge_id(quad[(i - 1) % 4], quad[i]); e_next = _edge_id(quad[i], quad[(i + 1) % 4])
                    adj[e_prev].append(e_next); adj[e_next].append(e_prev)
        else:
            assert len(cut) == 0
    mid = {e: 0.5 * (CORNER_OFFSETS[a] + CORNER_OFFSETS[b]).astype(np.float64) for e, (a, b) in enumerate(EDGES)}
    tris: List[Tuple[int, int, int]] = []
    seen = set()
    for e0 in range(12):
        if e0 in seen or not adj[e0]:
            continue
        assert len(adj[e0]) == 2
        loop = [e0]; seen.add(e0); prev, cur = e0, adj[e0][0]
        while cur != e0:
            loop.append(cur); seen.add(cur)
            n0, n1 = adj[cur]
            nxt = n1 if n0 == prev else n0
            prev, cur = cur, nxt
        # orientation: Newell normal vs the local inside direction
        pts = np.array([mid[e] for e in loop])
        nrm = np.zeros(3)
        for i in range(len(loop)):
            p, q = pts[i], pts[(i + 1) % len(loop)]
            nrm += np.cross(p, q)
        g = np.zeros(3)
        for e in loop:
            a, b = EDGES[e]
            ia, ib = (a, b) if inside[a] else (b, a)
            g += (CORNER_OFFSETS[ia] - CORNER_OFFSETS[ib]).astype(np.float64)
        d = float(nrm @ g)
        assert abs(d) > 1e-9, (case, loop)
        if d > 0:                       # normal currently points towards the inside -> reverse
            loop = loop[::-1]
        for i in range(1, len(loop) - 1):
            tris.append((loop[0], loop[i], loop[i + 1]))
    return tris


def build_tables():
    """-> (ntri (256,) uint8, tri (256, MAXT*3) int8 padded with -1, edge_mask (256,) uint16)."""
    cases = [_build_case(c) for c in range(256)]
    maxt = max(len(t) for t in cases)
    ntri = np.array([len(t) for t in cases], dtype=np.uint8)
    tri = -np.ones((256, maxt * 3), dtype=np.int8)
    mask = np.zeros(256, dtype=np.uint16)
    for c, t in enumerate(cases):
        for i, (a, b, d) in enumerate(t):
            tri[c, 3 * i: 3 * i + 3] = (a, b, d)
            mask[c] |= (1 << a) | (1 << b) | (1 << d)
    return ntri, tri, mask


NTRI, TRI, EDGE_MASK = build_tables()
MAX_TRI = TRI.shape[1] // 3


def emit_cuda_include(path: str) -> None:
    """Tables live in GLOBAL memory and are staged into shared memory by each CTA: they are indexed by the per-thread case number,
    and a divergent index serialises __constant__ accesses (one address per cycle per warp)."""
    a = np.array(EDGES, dtype=np.int32)
    assert [int(x) for x in EDGE_AXIS] == [e >> 2 for e in range(12)]
    corner_nibbles = sum(int(c) << (4 * e) for e, c in enumerate(a[:, 0]))
    lines = ['// GENERATED by `python -m avatarcap_b200.mc_tables` -- do not edit.',
             '// Marching-cubes case tables (see avatarcap_b200/mc_tables.py for the construction).',
             '#define AVC_MC_MAX_TRI %d' % MAX_TRI,
             '__device__ const unsigned char g_mc_ntri[256] = {%s};' % ','.join(str(int(x)) for x in NTRI),
             '// triangle list per case: 3 edge numbers per triangle, -1 padded to 16 bytes (one uint4 per case)',
             '__device__ __align__(16) const signed char g_mc_tri[256][16] = {']
    for c in range(256):
        lines.append('  {%s},' % ','.join(str(int(x)) for x in list(TRI[c]) + [-1] * (16 - MAX_TRI * 3)))
    lines.append('};')
    lines.append('// edge e: axis = e >> 2; lower corner offset (dx | dy<<1 | dz<<2) = nibble e of this constant')
    lines.append('#define AVC_MC_EDGE_CORNER_NIBBLES 0x%xull' % corner_nibbles)
    txt = '\n'.join(lines) + '\n'
    if os.path.exists(path) and open(path).read() == txt:
        return                      # unchanged: keep the timestamp so that make does not rebuild everything
    with open(path, 'w') as f:
        f.write(txt)


if __name__ == '__main__':
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc', 'mc_tables.inc')
    emit_cuda_include(out)
    print('wrote', out, 'max triangles per cell =', MAX_TRI, 'total tris over cases =', int(NTRI.sum()))
