"""ORACLE (test infrastructure only) -- CPU restatement of utils/recon_util.py (mesh extraction + normals).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

PARITY UNPINNED at one boundary: recon_util.py:64 calls skimage.measure.marching_cubes (scikit_image==0.17.2,
method 'lewiner'), a third-party Cython routine that is neither vendored under /root/reference nor installed here.
Its published behaviour is restated (inside <=> value > level; one vertex per sign-changing grid edge, placed by
linear interpolation; spacing scales index coordinates) with a classic marching cubes traced per cell HERE, independently of
the product's generated case table (the two are compared case by case and mesh by mesh through `surface_signature`, which
is blind to the fan each side chose). Lewiner's MC33 topology can differ in ambiguous cells (which face diagonal a
face-ambiguous pattern takes: `face_rule`; and, in a few interior-ambiguous cases, an extra cell-centre vertex); vertex
positions on edges are identical. tests/test_oracle_vs_golden.py bounds that gap on the body SDF.
Everything else in recon_util.py (lines 9-63, 65-70) is restated verbatim and pinned by tests/golden/mesh_golden.npz.

Vertex / face ORDER is this repo's own canonical order (skimage's order is unspecified):
  vertices ascending in (owner voxel linear index (i*Ry+j)*Rz+k, axis x<y<z); faces ascending in (cell linear
  index, triangle number in the case table).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------------------------------
# Independent per-cell tracer. Nothing below imports the product's case table (avatarcap_b200/mc_tables.py): corners, edges and
# faces are enumerated from coordinates here, the iso-contour of every cube face is drawn from the face's own corner signs, the
# contour segments are chained into closed loops and each loop is fan-triangulated. The product's table is checked AGAINST this
# (tests/test_oracle_vs_golden.py::test_product_case_table_matches_the_independent_tracer), not the other way round.
_CORNERS = [(x, y, z) for z in (0, 1) for y in (0, 1) for x in (0, 1)]            # corner index c = x | y << 1 | z << 2


def _cell_edges():
    """12 edges as (owner corner offset (3,), axis); the edge runs from `owner` to owner + e_axis. Ordered by (axis, z, y, x of the
    owner) -- only used as an internal key, the mesh identifies an edge by (owner voxel, axis)."""
    out = []
    for axis in range(3):
        for c in _CORNERS:
            if c[axis] == 0:
                out.append((c, axis))
    return out


_EDGES = _cell_edges()
_EDGE_KEY = {(c, a): i for i, (c, a) in enumerate(_EDGES)}


def _edge_between(p, q):
    axis = [i for i in range(3) if p[i] != q[i]]
    assert len(axis) == 1
    lo = p if p[axis[0]] == 0 else q
    return _EDGE_KEY[(lo, axis[0])]


def _face_cycles():
    """the 6 cube faces as cyclic corner quadruples"""
    faces = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in (0, 1):
            quad = []
            for du, dv in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]; c[axis] = side; c[u] = du; c[v] = dv
                quad.append(tuple(c))
            faces.append(quad)
    return faces


_FACE_CYCLES = _face_cycles()


def case_loops(case: int, face_rule: str = 'separate'):
    """Closed, oriented iso-contour loops of one sign pattern: list of loops, each a list of cell-edge numbers (into _EDGES).
    face_rule decides the face-ambiguous patterns (4 sign changes around a face): 'separate' cuts off each INSIDE corner,
    'join' cuts off each OUTSIDE corner (the two ways a bilinear face interpolant can resolve, i.e. what an asymptotic decider /
    MC33 chooses between from the VALUES; a rule that only looks at the face's signs is consistent across cells either way).
    Orientation: right-hand normal from inside (value > iso) to outside."""
    ins = {c: (case >> (c[0] | c[1] << 1 | c[2] << 2)) & 1 for c in _CORNERS}
    nxt = {}
    def link(e1, e2):
        nxt.setdefault(e1, []).append(e2); nxt.setdefault(e2, []).append(e1)
    for quad in _FACE_CYCLES:
        cuts = [i for i in range(4) if ins[quad[i]] != ins[quad[(i + 1) % 4]]]
        if len(cuts) == 2:
            link(_edge_between(quad[cuts[0]], quad[(cuts[0] + 1) % 4]), _edge_between(quad[cuts[1]], quad[(cuts[1] + 1) % 4]))
        elif len(cuts) == 4:
            want = 1 if face_rule == 'separate' else 0
            for i in range(4):
                if ins[quad[i]] == want:                      # cut this corner off: the two face edges that meet in it
                    link(_edge_between(quad[i - 1], quad[i]), _edge_between(quad[i], quad[(i + 1) % 4]))
        else:
            assert not cuts
    loops, done = [], set()
    for start in sorted(nxt):
        if start in done:
            continue
        assert len(nxt[start]) == 2
        loop, prev, cur = [start], None, start
        while True:
            done.add(cur)
            a, b = nxt[cur]
            step = a if a != prev else b
            if len(loop) > 1 and step == start:
                break
            if a == b:                                         # cannot happen on a cube (an edge lies on two distinct faces)
                raise AssertionError
            prev, cur = cur, step
            loop.append(cur)
        # orientation by the Newell normal against the summed inside -> outside directions of the loop's edges
        mids = np.array([np.array(_EDGES[e][0], float) + 0.5 * np.eye(3)[_EDGES[e][1]] for e in loop])
        nrm = sum(np.cross(mids[i], mids[(i + 1) % len(loop)]) for i in range(len(loop)))
        out_dir = np.zeros(3)
        for e in loop:
            c, ax = _EDGES[e]
            d = np.eye(3)[ax]
            out_dir += -d if ins[c] == 0 else d               # from the inside end point towards the outside one
        if float(nrm @ out_dir) < 0:
            loop = loop[::-1]
        loops.append(loop)
    return loops


def _tracer_tables(face_rule: str):
    """per case: triangle count and (tri, corner) -> (owner dx, dy, dz, axis), from the loops (fan triangulation)"""
    ntri = np.zeros(256, np.int64)
    tab = -np.ones((256, 5, 3, 4), np.int64)
    for case in range(256):
        t = 0
        for loop in case_loops(case, face_rule):
            for i in range(1, len(loop) - 1):
                for k, e in enumerate((loop[0], loop[i], loop[i + 1])):
                    c, ax = _EDGES[e]
                    tab[case, t, k] = (c[0], c[1], c[2], ax)
                t += 1
        ntri[case] = t
    return ntri, tab


_TRACER = {}


def marching_cubes(vol: np.ndarray, level: float, spacing=(1.0, 1.0, 1.0), face_rule: str = 'separate', return_owner: bool = False):
    """Restates skimage.measure.marching_cubes(volume, level, spacing=...)[:2] (see module docstring).
    -> verts (V,3) float32 in index*spacing coordinates, faces (F,3) int32 (normal = -gradient, i.e. 'descent').
    return_owner: also (owner voxel linear index of every vertex, axis of every vertex, cell linear index of every face)."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    X, Y, Z = vol.shape
    if not (vol.min() <= level <= vol.max()):
        raise ValueError('Surface level must be within volume data range.')     # skimage behaviour
    if face_rule not in _TRACER:
        _TRACER[face_rule] = _tracer_tables(face_rule)
    NT, TAB = _TRACER[face_rule]
    inside = vol > np.float32(level)
    nvox = X * Y * Z
    # --- vertices: one per sign-changing edge, owned by the lower voxel -------------------------------
    cut = np.zeros((nvox, 3), dtype=bool)
    cx = np.zeros((X, Y, Z), bool); cx[:-1] = inside[:-1] != inside[1:]
    cy = np.zeros((X, Y, Z), bool); cy[:, :-1] = inside[:, :-1] != inside[:, 1:]
    cz = np.zeros((X, Y, Z), bool); cz[:, :, :-1] = inside[:, :, :-1] != inside[:, :, 1:]
    cut[:, 0] = cx.reshape(-1); cut[:, 1] = cy.reshape(-1); cut[:, 2] = cz.reshape(-1)
    flat_cut = cut.reshape(-1)
    vid = np.cumsum(flat_cut) - 1                       # vertex id of (voxel, axis) in canonical order
    keys = np.nonzero(flat_cut)[0]
    vox = keys // 3; axis = keys % 3
    i = vox // (Y * Z); j = (vox // Z) % Y; k = vox % Z
    base = np.stack([i, j, k], 1).astype(np.float32)
    strides = np.array([Y * Z, Z, 1])
    fv = vol.reshape(-1)
    va = fv[vox]; vb = fv[vox + strides[axis]]
    t = (np.float32(level) - va) / (vb - va)             # linear interpolation along the edge, float32
    verts = base.copy()
    verts[np.arange(len(keys)), axis] += t
    verts = (verts * np.asarray(spacing, dtype=np.float32)).astype(np.float32)
    # --- faces -------------------------------------------------------------------------------------------
    ci = np.zeros((X - 1, Y - 1, Z - 1), dtype=np.int32)
    for c, (dx, dy, dz) in enumerate(_CORNERS):
        ci |= inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int32) << c
    ci = ci.reshape(-1)
    active = np.nonzero(NT[ci] > 0)[0]
    cc = ci[active]
    ii = active // ((Y - 1) * (Z - 1)); jj = (active // (Z - 1)) % (Y - 1); kk = active % (Z - 1)
    nt = NT[cc]
    first = np.cumsum(nt) - nt
    total = int(nt.sum())
    cell_of = np.repeat(np.arange(len(active)), nt)
    tnum = np.arange(total) - first[cell_of]
    faces = np.empty((total, 3), dtype=np.int32)
    for corner in range(3):
        o = TAB[cc[cell_of], tnum, corner]               # (F, 4): owner offset + axis
        ovox = ((ii[cell_of] + o[:, 0]) * Y + (jj[cell_of] + o[:, 1])) * Z + (kk[cell_of] + o[:, 2])
        faces[:, corner] = vid[ovox * 3 + o[:, 3]]
    if return_owner:
        cell_lin = (ii[cell_of] * Y + jj[cell_of]) * Z + kk[cell_of]      # lowest-corner voxel of the face's cell
        return verts, faces, vox, axis, cell_lin
    return verts, faces


def surface_signature(faces: np.ndarray, cell_of_face: np.ndarray) -> np.ndarray:
    """Triangulation-independent description of a marching-cubes surface: per cell, the DIRECTED boundary edges of the cell's
    triangle patch (edges used by two triangles of the same cell cancel). Two meshes over the same vertices are the same surface
    cell by cell (same loops, same orientation), whatever fan each table chose, iff their signatures are equal. -> sorted (n,3)
    int64 rows (cell, a, b)."""
    f = np.asarray(faces, np.int64); c = np.asarray(cell_of_face, np.int64)
    e = np.concatenate([np.stack([c, f[:, 0], f[:, 1]], 1), np.stack([c, f[:, 1], f[:, 2]], 1), np.stack([c, f[:, 2], f[:, 0]], 1)], 0)
    V = int(f.max()) + 1 if len(f) else 1
    key = (e[:, 0] * V + e[:, 1]) * V + e[:, 2]
    rkey = (e[:, 0] * V + e[:, 2]) * V + e[:, 1]
    keep = ~np.isin(key, rkey)                           # an interior edge of the patch appears once in each direction
    out = e[keep]
    return out[np.lexsort((out[:, 2], out[:, 1], out[:, 0]))]


def cells_of_faces(vol: np.ndarray, level: float) -> np.ndarray:
    """cell (lowest-corner voxel linear index) of every face of ANY marching cubes that emits its faces in ascending cell order
    with the triangle count the loops dictate (sum of len(loop) - 2): lets the tests segment the product's face list per cell."""
    vol = np.ascontiguousarray(vol, np.float32)
    return marching_cubes(vol, level, return_owner=True)[4]


def extract_normal_volume(vol: np.ndarray, voxel_size: np.ndarray) -> np.ndarray:
    """recon_util.py:9-29: three 3x3x3 Sobel cross-correlations with zero padding, / (16*2*voxel) -> (X,Y,Z,3)."""
    vol = vol.astype(np.float32)
    X, Y, Z = vol.shape
    pad = np.zeros((X + 2, Y + 2, Z + 2), dtype=np.float32)
    pad[1:-1, 1:-1, 1:-1] = vol
    sm = np.array([1, 2, 1], dtype=np.float32)
    out = np.zeros((X, Y, Z, 3), dtype=np.float32)
    dims = (X, Y, Z)
    for ax in range(3):
        o0, o1 = [a for a in range(3) if a != ax]
        acc = np.zeros((X, Y, Z), dtype=np.float32)
        for a in range(3):
            for b in range(3):
                plus = [None] * 3; minus = [None] * 3
                plus[ax] = slice(2, 2 + dims[ax]); minus[ax] = slice(0, dims[ax])
                plus[o0] = minus[o0] = slice(a, a + dims[o0])
                plus[o1] = minus[o1] = slice(b, b + dims[o1])
                acc += (sm[a] * sm[b]) * (pad[tuple(plus)] - pad[tuple(minus)])
        out[..., ax] = acc / np.float32(16 * 2 * voxel_size[ax])
    return out


def _trilinear_border_np(vol_c: np.ndarray, g: np.ndarray) -> np.ndarray:
    """vol_c (D,H,W,C); g (N,3) normalised with g[:,0]->W, g[:,1]->H, g[:,2]->D (grid_sample convention)."""
    D, H, W, C = vol_c.shape
    def un(x, n):
        return np.clip(((x + 1) / 2) * (n - 1), 0, n - 1).astype(np.float32)
    ix = un(g[:, 0], W); iy = un(g[:, 1], H); iz = un(g[:, 2], D)
    x0 = np.floor(ix); y0 = np.floor(iy); z0 = np.floor(iz)
    tx = ix - x0; ty = iy - y0; tz = iz - z0
    x0 = x0.astype(int); y0 = y0.astype(int); z0 = z0.astype(int)
    out = np.zeros((g.shape[0], C), dtype=np.float32)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi = x0 + dx; yi = y0 + dy; zi = z0 + dz
                ok = (xi <= W - 1) & (yi <= H - 1) & (zi <= D - 1)
                w = ((tx if dx else 1 - tx) * (ty if dy else 1 - ty) * (tz if dz else 1 - tz) * ok).astype(np.float32)
                out += vol_c[np.minimum(zi, D - 1), np.minimum(yi, H - 1), np.minimum(xi, W - 1)] * w[:, None]
    return out


def extract_normal_from_volume(vol: np.ndarray, voxel_size: np.ndarray, pts_grid: np.ndarray) -> np.ndarray:
    """recon_util.py:32-48. pts_grid (N,3) normalised volume coords in (x,y,z) order; the reference reorders to
    [2,1,0] because grid_sample's x indexes the LAST volume axis. No epsilon in the normalisation (:46-47)."""
    nv = extract_normal_volume(vol, voxel_size)                # (X,Y,Z,3) == (D,H,W,C)
    n = _trilinear_border_np(nv, pts_grid[:, [2, 1, 0]].astype(np.float32))
    return (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)


def same_surface(faces_a, faces_b, cells) -> bool:
    """faces_a / faces_b (F,3) over the SAME vertex numbering, both in ascending cell order with `cells[f]` the cell of face f:
    True when they describe the same oriented surface cell by cell (see surface_signature), whatever triangulation of the loops."""
    fa = np.asarray(faces_a); fb = np.asarray(faces_b)
    return fa.shape == fb.shape and np.array_equal(surface_signature(fa, cells), surface_signature(fb, cells))


def recon_mesh(occ_volume: np.ndarray, volume_res, bounds: np.ndarray, iso_value: float = 0.5, return_cells: bool = False):
    """recon_util.recon_mesh (recon_util.py:51-70). -> vertices (V,3) f32, faces (F,3) i32, normals (V,3) f32
    (+ the cell of every face with return_cells, for same_surface)."""
    vol = np.asarray(occ_volume, dtype=np.float32).reshape(volume_res)
    bounds = np.asarray(bounds, dtype=np.float32)
    volume_len = bounds[1] - bounds[0]                                         # :60
    voxel_size = volume_len / np.array(volume_res, dtype=np.float32)           # :61
    vertices, faces, _, _, cells = marching_cubes(vol, iso_value, spacing=voxel_size, return_owner=True)       # :64
    vertices = vertices + bounds[0] + 0.5 * voxel_size                         # :65
    vertices_grid = 2 * (vertices - bounds[0]) / volume_len - 1.0              # :66
    normals = extract_normal_from_volume(vol, voxel_size, vertices_grid)       # :67
    normals = -normals                                                         # :68
    faces = faces[:, [2, 1, 0]]                                                # :69
    if return_cells:
        return vertices.astype(np.float32), faces.astype(np.int32), normals.astype(np.float32), cells
    return vertices.astype(np.float32), faces.astype(np.int32), normals.astype(np.float32)


def chamfer(a: np.ndarray, b: np.ndarray) -> float:
    """Symmetric mean nearest-neighbour distance (metres) between two vertex sets."""
    from scipy.spatial import cKDTree
    da, _ = cKDTree(b).query(a); db, _ = cKDTree(a).query(b)
    return float(0.5 * (da.mean() + db.mean()))


def contains_points(verts: np.ndarray, faces: np.ndarray, pts: np.ndarray, chunk: int = 2048) -> np.ndarray:
    """Restates trimesh.Trimesh.contains (ray-casting parity; dataset/avatarcap_dataset.py:120-123; trimesh is a third-party
    dependency that is not installed here -> PARITY UNPINNED, pinned by analytic shapes in the tests): a point is inside a closed
    mesh iff a ray along +z crosses the surface an odd number of times. float64, top-left rule on the projected triangles."""
    v = np.asarray(verts, np.float64); f = np.asarray(faces, np.int64); p = np.asarray(pts, np.float64)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    area = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])
    ok = area != 0
    a, b, c, area = a[ok], b[ok], c[ok], area[ok]
    sgn = np.sign(area)
    out = np.zeros(len(p), bool)

    def edge(pa, pb, px, py):
        e = sgn[None] * ((pb[None, :, 0] - pa[None, :, 0]) * (py[:, None] - pa[None, :, 1]) - (pb[None, :, 1] - pa[None, :, 1]) * (px[:, None] - pa[None, :, 0]))
        dx = (sgn * (pb[:, 0] - pa[:, 0]))[None]; dy = (sgn * (pb[:, 1] - pa[:, 1]))[None]
        return (e > 0) | ((e == 0) & (((dy == 0) & (dx < 0)) | (dy < 0)))

    for s in range(0, len(p), chunk):
        q = p[s:s + chunk]; px, py, pz = q[:, 0], q[:, 1], q[:, 2]
        ins = edge(a, b, px, py) & edge(b, c, px, py) & edge(c, a, px, py)
        w0 = ((b[None, :, 0] - px[:, None]) * (c[None, :, 1] - py[:, None]) - (b[None, :, 1] - py[:, None]) * (c[None, :, 0] - px[:, None])) / area[None]
        w1 = ((c[None, :, 0] - px[:, None]) * (a[None, :, 1] - py[:, None]) - (c[None, :, 1] - py[:, None]) * (a[None, :, 0] - px[:, None])) / area[None]
        zc = w0 * a[None, :, 2] + w1 * b[None, :, 2] + (1 - w0 - w1) * c[None, :, 2]
        out[s:s + chunk] = ((ins & (zc > pz[:, None])).sum(1) % 2) == 1
    return out
