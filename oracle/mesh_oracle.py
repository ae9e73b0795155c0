"""ORACLE (test infrastructure only) -- CPU restatement of utils/recon_util.py (mesh extraction + normals).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

PARITY UNPINNED at one boundary: recon_util.py:64 calls skimage.measure.marching_cubes (scikit_image==0.17.2,
method 'lewiner'), a third-party Cython routine that is neither vendored under /root/reference nor installed here.
Its published behaviour is restated (inside <=> value > level; one vertex per sign-changing grid edge, placed by
linear interpolation; spacing scales index coordinates) with a classic marching-cubes whose case table resolves
ambiguous faces consistently (avatarcap_b200/mc_tables.py). Lewiner's MC33 topology can differ from it in ambiguous
cells (face count, and rarely an extra cell-centre vertex); vertex positions on edges are identical.
Everything else in recon_util.py (lines 9-63, 65-70) is restated verbatim and pinned by tests/golden/mesh_golden.npz.

Vertex / face ORDER is this repo's own canonical order (skimage's order is unspecified):
  vertices ascending in (owner voxel linear index (i*Ry+j)*Rz+k, axis x<y<z); faces ascending in (cell linear
  index, triangle number in the case table).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from avatarcap_b200.mc_tables import CORNER_OFFSETS, EDGES, EDGE_AXIS, EDGE_OWNER_OFFSET, NTRI, TRI


def marching_cubes(vol: np.ndarray, level: float, spacing=(1.0, 1.0, 1.0)) -> Tuple[np.ndarray, np.ndarray]:
    """Restates skimage.measure.marching_cubes(volume, level, spacing=...)[:2] (see module docstring).
    -> verts (V,3) float32 in index*spacing coordinates, faces (F,3) int32 (normal = -gradient, i.e. 'descent')."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    X, Y, Z = vol.shape
    if not (vol.min() <= level <= vol.max()):
        raise ValueError('Surface level must be within volume data range.')     # skimage behaviour
    inside = vol > np.float32(level)
    nvox = X * Y * Z
    # --- vertices: one per sign-changing edge, owned by the lower voxel -------------------------------
    cut = np.zeros((nvox, 3), dtype=bool)
    v3 = inside
    cx = np.zeros((X, Y, Z), bool); cx[:-1] = v3[:-1] != v3[1:]
    cy = np.zeros((X, Y, Z), bool); cy[:, :-1] = v3[:, :-1] != v3[:, 1:]
    cz = np.zeros((X, Y, Z), bool); cz[:, :, :-1] = v3[:, :, :-1] != v3[:, :, 1:]
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
    for c in range(8):
        dx, dy, dz = CORNER_OFFSETS[c]
        ci |= inside[dx:X - 1 + dx, dy:Y - 1 + dy, dz:Z - 1 + dz].astype(np.int32) << c
    ci = ci.reshape(-1)
    active = np.nonzero(NTRI[ci] > 0)[0]
    cc = ci[active]
    ii = active // ((Y - 1) * (Z - 1)); jj = (active // (Z - 1)) % (Y - 1); kk = active % (Z - 1)
    nt = NTRI[cc].astype(np.int64)
    # expand (cell, t) pairs in order
    first = np.cumsum(nt) - nt
    total = int(nt.sum())
    cell_of = np.repeat(np.arange(len(active)), nt)
    tnum = np.arange(total) - first[cell_of]
    faces = np.empty((total, 3), dtype=np.int32)
    for corner in range(3):
        e = TRI[cc[cell_of], 3 * tnum + corner].astype(np.int64)
        own = EDGE_OWNER_OFFSET[e]
        ovox = ((ii[cell_of] + own[:, 0]) * Y + (jj[cell_of] + own[:, 1])) * Z + (kk[cell_of] + own[:, 2])
        faces[:, corner] = vid[ovox * 3 + EDGE_AXIS[e]]
    return verts, faces


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


def recon_mesh(occ_volume: np.ndarray, volume_res, bounds: np.ndarray, iso_value: float = 0.5):
    """recon_util.recon_mesh (recon_util.py:51-70). -> vertices (V,3) f32, faces (F,3) i32, normals (V,3) f32."""
    vol = np.asarray(occ_volume, dtype=np.float32).reshape(volume_res)
    bounds = np.asarray(bounds, dtype=np.float32)
    volume_len = bounds[1] - bounds[0]                                         # :60
    voxel_size = volume_len / np.array(volume_res, dtype=np.float32)           # :61
    vertices, faces = marching_cubes(vol, iso_value, spacing=voxel_size)       # :64
    vertices = vertices + bounds[0] + 0.5 * voxel_size                         # :65
    vertices_grid = 2 * (vertices - bounds[0]) / volume_len - 1.0              # :66
    normals = extract_normal_from_volume(vol, voxel_size, vertices_grid)       # :67
    normals = -normals                                                         # :68
    faces = faces[:, [2, 1, 0]]                                                # :69
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
