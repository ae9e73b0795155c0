"""CPU: the oracle (oracle/*.py) against golden vectors produced by the reference's own modules
(tests/golden/gen_golden.py). This is what pins the oracle (the reference ships no tests)."""
import numpy as np
import pytest
import torch

from oracle import field_oracle as fo
from oracle import mesh_oracle as mo
from avatarcap_b200 import synth
from helpers import golden_scene, load_golden, maxabs


@pytest.fixture(scope='module')
def scene():
    return golden_scene()


def test_feature_maps_reproducible(scene):
    g = load_golden('avatar_golden.npz')
    assert abs(float(scene['pose_map'].astype(np.float64).sum()) - float(g['fmap_sum'])) < 1e-6
    r = load_golden('recon_golden.npz')
    assert abs(float(scene['image_map'].astype(np.float64).sum()) - float(r['fmap_sum'])) < 1e-6


def test_positional_encoding_bit_exact():
    g = load_golden('avatar_golden.npz')
    pe = fo.embed(torch.from_numpy(g['pts']), 10).numpy()
    assert pe.shape == (g['pts'].shape[0], 63)
    assert np.array_equal(pe, g['pe'])
    assert np.array_equal(fo.embed(torch.from_numpy(g['pts']), 0).numpy(), g['pts'])    # multires 0 -> identity


def test_occupancy_query(scene):
    g = load_golden('avatar_golden.npz')
    o = fo.occupancy_query(scene['avatar_sd'], g['pts'], scene['pose_map'], g['center'], with_texture=True)
    assert maxabs(o['nonrigid_offset'], g['nonrigid_offset']) < 5e-7
    assert maxabs(o['nonrigid_offset'], g['warp_query']) < 5e-7
    assert maxabs(o['cano_pts_ov'], g['cano_pts_ov']) < 5e-5         # fp32 rounding through 2^9 PE frequencies
    assert maxabs(o['rgb'], g['rgb']) < 1e-6
    assert maxabs(o['alpha'], g['alpha']) < 5e-5 * max(1.0, float(np.abs(g['alpha']).max()))   # density head is O(100)
    assert np.ptp(g['cano_pts_ov']) > 1.0                            # the field is non-degenerate (O(1) like a trained SDF)


def test_occupancy_query_f64_gap(scene):
    """fp32 reference vs fp64 oracle: the inherent fp32 noise that the 1e-4 budget has to absorb."""
    g = load_golden('avatar_golden.npz')
    o = fo.occupancy_query(scene['avatar_sd'], g['pts'], scene['pose_map'], g['center'], dtype=torch.float64)
    assert maxabs(o['cano_pts_ov'], g['cano_pts_ov']) < 3e-5


def test_recon_infer(scene):
    g = load_golden('recon_golden.npz')
    ov = fo.recon_infer(scene['recon_sd'], g['pts'], scene['image_map'], g['center'])
    assert g['ov'].shape == (1, ov.shape[0])          # the reference returns (1,N) for B=1 (arch_recon.py:74)
    assert maxabs(ov, g['ov'][0]) < 2e-6
    assert g['ov'].min() < 0.3 and g['ov'].max() > 0.7


def test_lbs(scene):
    g = load_golden('lbs_golden.npz'); fr = scene['frame']
    lbs = fo.calculate_lbs(g['verts'], fr['cano_smpl_v'], fr['smpl_skinning_weights'])
    assert maxabs(lbs, g['lbs']) < 1e-6
    assert np.allclose(lbs.sum(1), 1.0, atol=1e-5)
    live, mats = fo.skinning(g['verts'], lbs, fr['cano2live_jnt_mats'])
    assert maxabs(live, g['live']) < 2e-6 and maxabs(mats, g['mats']) < 2e-6
    assert maxabs(fo.skinning_normal(g['normals'], lbs, fr['cano2live_jnt_mats']), g['live_normals']) < 2e-6


@pytest.mark.parametrize('space', ['posed', 'cano', 'temp'])
def test_geotex_forward(scene, space):
    g = load_golden('avatar_golden.npz'); fr = scene['frame']
    wvol = synth.blend_weight_volume(fr)
    w = g['fwd_wpts_live'] if space == 'posed' else g['fwd_wpts_cano']
    o = fo.geotex_forward(scene['avatar_sd'], w, g['fwd_dists'], fr, scene['pose_map'], wvol, space)
    assert maxabs(o['nonrigid_offset'], g['fwd_%s_off' % space]) < 1e-6
    assert maxabs(o['occ'], g['fwd_%s_occ' % space]) < 1e-4
    assert maxabs(o['raw'], g['fwd_%s_raw' % space]) < 1e-5
    after = g['fwd_%s_wpts_after' % space]
    if space == 'cano':      # reference quirk: in-place += offsets on the caller's tensor
        assert maxabs(after, w + g['fwd_cano_off']) < 1e-6
    else:
        assert np.array_equal(after, w)


def test_sobel_normals_and_grid():
    g = load_golden('mesh_golden.npz')
    nv = mo.extract_normal_volume(g['vol'], g['voxel'])
    assert maxabs(nv, g['normal_volume']) < 2e-5 * float(np.abs(g['normal_volume']).max())
    n = mo.extract_normal_from_volume(g['vol'], g['voxel'], g['grid_pts'])
    assert maxabs(n, g['normals']) < 2e-6
    assert np.array_equal(fo.generate_volume_points(g['bounds'], (7, 9, 5)), g['vol_pts_7_9_5'])
    gp = fo.generate_volume_points(g['bounds'], (64, 33, 128))
    assert np.array_equal(gp[::97], g['vol_pts_64_33_128_sub'])
    assert maxabs(synth.volume_points(g['bounds'], (64, 33, 128)), gp) < 2.5e-7


def test_marching_cubes_properties():
    """The MC restatement is unpinned against skimage (not installed); pin it by surface properties instead."""
    R = 48
    ax = np.linspace(-1, 1, R)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    vol = (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)          # positive inside (reference convention)
    h = 2.0 / (R - 1)
    v, f = mo.marching_cubes(vol, 0.0, (h, h, h))
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0).astype(np.int64)
    key = e[:, 0] * len(v) + e[:, 1]; rkey = e[:, 1] * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key) and np.isin(rkey, key).all()       # closed, consistently oriented manifold
    assert len(v) - len(e) // 2 + len(f) == 2                                  # Euler characteristic of a sphere
    p = v - 1.0
    tri = p[f]; nrm = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert ((nrm * tri.mean(1)).sum(1) > 0).all()                             # normals along -gradient ('descent')
    assert abs(0.5 * np.linalg.norm(nrm, axis=1).sum() - 4 * np.pi * 0.36) < 0.02
    assert np.abs(np.linalg.norm(p, axis=1) - 0.6).max() < 2e-3
    # vertex count == number of sign-changing grid edges
    ins = vol > 0
    n_edges = (ins[1:] != ins[:-1]).sum() + (ins[:, 1:] != ins[:, :-1]).sum() + (ins[:, :, 1:] != ins[:, :, :-1]).sum()
    assert len(v) == n_edges
    with pytest.raises(ValueError):
        mo.marching_cubes(vol, 5.0)


def test_marching_cubes_watertight_on_noise():
    rs = np.random.RandomState(3)
    vol = rs.normal(0, 1, (14, 11, 9)).astype(np.float32)
    pad = -5 * np.ones((16, 13, 11), np.float32); pad[1:-1, 1:-1, 1:-1] = vol      # closed: everything outside is 'outside'
    v, f = mo.marching_cubes(pad, 0.0)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0).astype(np.int64)
    key = e[:, 0] * len(v) + e[:, 1]; rkey = e[:, 1] * len(v) + e[:, 0]
    assert len(np.unique(key)) == len(key) and np.isin(rkey, key).all()
    assert f.min() == 0 and f.max() == len(v) - 1


def test_recon_mesh_conventions():
    """recon_util.py:61-69: voxel = len/res, +0.5 voxel shift, negated normals, reversed winding."""
    res = (20, 24, 12)
    bounds = np.array([[-0.5, -0.6, -0.3], [0.5, 0.6, 0.3]], np.float32)
    ii, jj, kk = np.meshgrid(*[np.arange(r) for r in res], indexing='ij')
    c = np.array(res) / 2.0 - 0.5
    vol = (4.0 - np.sqrt((ii - c[0]) ** 2 + (jj - c[1]) ** 2 + (kk - c[2]) ** 2)).astype(np.float32)
    v, f, n = mo.recon_mesh(vol, res, bounds, 0.0)
    voxel = (bounds[1] - bounds[0]) / np.array(res, np.float32)
    centre = bounds[0] + (c + 0.5) * voxel
    r_idx = np.linalg.norm((v - centre) / voxel, axis=1)
    assert np.abs(r_idx - 4.0).max() < 0.08
    assert ((n * (v - centre)).sum(1) > 0).all()                  # returned normals point outward (field is positive inside)
    tri = v[f]; fn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert ((fn * (tri.mean(1) - centre)).sum(1) < 0).all()        # faces[:, [2,1,0]] reverses the 'descent' winding
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)


def test_nerf_render_vertex_colours(scene):
    """NerfRenderer.render + raw2outputs driven as main.py:464-478 (texture template integrated along -normal)."""
    g = load_golden('nerf_golden.npz'); fr = scene['frame']
    wvol = synth.blend_weight_volume(fr)
    R = len(g['verts'])
    o = fo.nerf_render(scene['avatar_sd'], g['verts'] + g['normals'], -g['normals'], np.full(R, 0.95, np.float32), np.full(R, 1.05, np.float32),
                       np.ones(R, np.float32), fr, scene['pose_map'], wvol, 'cano', 0.02, 0.05)
    assert np.array_equal(o['near'], g['near_after']) and np.array_equal(o['far'], g['far_after'])    # depth-1 rays: near 0.98, far 1.05
    assert maxabs(o['raw'], g['raw']) < 2e-5
    assert maxabs(o['rgb_map'], g['rgb_map']) < 2e-5 and maxabs(o['acc_map'], g['acc_map']) < 2e-5 and maxabs(o['depth_map'], g['depth_map']) < 2e-5
    assert g['acc_map'].max() > 0.9 and (g['acc_map'] > 0.05).mean() > 0.05                           # the compositing is exercised


def test_contains_points_oracle_on_analytic_shapes():
    """trimesh.contains (avatarcap_dataset.py:120-123) is a third-party routine that is not installed: the ray-parity restatement is
    pinned by shapes whose inside is known in closed form -- spheres (any orientation of the faces), a box with grid points exactly on
    its faces' projections, nested and disjoint components."""
    from helpers import uv_sphere
    from oracle import field_oracle as fo
    bounds = np.array([[-0.9, -1.0, -0.35], [0.95, 0.9, 0.3]], np.float32)
    res = (30, 34, 18)
    pts = fo.generate_volume_points(bounds, res)
    c1, r1, c2, r2 = np.array([0.1, -0.2, 0.0]), 0.27, np.array([-0.45, 0.5, 0.05]), 0.2
    v1, f1 = uv_sphere(c1, r1); v2, f2 = uv_sphere(c2, r2, 16, 20)
    verts = np.concatenate([v1, v2], 0); faces = np.concatenate([f1, f2[:, ::-1] + len(v1)], 0)     # second sphere wound the other way
    ins = mo.contains_points(verts, faces, pts)
    d1 = np.linalg.norm(pts - c1, axis=1); d2 = np.linalg.norm(pts - c2, axis=1)
    assert ins[(d1 < 0.93 * r1) | (d2 < 0.9 * r2)].all() and not ins[(d1 > 1.01 * r1) & (d2 > 1.01 * r2)].any() and ins.sum() > 50
    # a hollow shell: points inside the inner sphere are outside the solid (two crossings)
    vi, fi = uv_sphere(c1, 0.12)
    shell = mo.contains_points(np.concatenate([v1, vi], 0), np.concatenate([f1, fi + len(v1)], 0), pts)
    assert not shell[d1 < 0.1].any() and shell[(d1 > 0.14) & (d1 < 0.24)].all()
    # axis-aligned box whose faces pass exactly through grid columns: every column is counted once (top-left rule), half-open in x, y
    gx = np.unique(pts[:, 0]); gy = np.unique(pts[:, 1])
    x0, x1, y0, y1, z0, z1 = gx[5], gx[12], gy[7], gy[20], -0.2, 0.17
    bv = np.array([[x, y, z] for x in (x0, x1) for y in (y0, y1) for z in (z0, z1)], np.float64)
    bf = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
    box = mo.contains_points(bv, bf, pts)
    strictly = (pts[:, 0] > x0) & (pts[:, 0] < x1) & (pts[:, 1] > y0) & (pts[:, 1] < y1) & (pts[:, 2] > z0) & (pts[:, 2] < z1)
    closed = (pts[:, 0] >= x0) & (pts[:, 0] <= x1) & (pts[:, 1] >= y0) & (pts[:, 1] <= y1) & (pts[:, 2] > z0) & (pts[:, 2] < z1)
    assert box[strictly].all() and not box[~closed].any()
    on_edge = closed & ~strictly
    cols = box[on_edge]
    assert 0 < cols.sum() < cols.size                     # boundary columns belong to exactly one side, not to both or neither


def test_product_case_table_matches_the_independent_tracer():
    """The product's generated table (avatarcap_b200/mc_tables.py -> csrc/mc_tables.inc) against the oracle's own per-cell tracer, which
    shares no code or table with it: for each of the 256 sign patterns the triangles must bound exactly the tracer's oriented loops
    (directed boundary edges equal, triangle count = sum(len(loop) - 2)) -- a table error cannot hide behind a shared import."""
    from avatarcap_b200 import mc_tables as T
    for case in range(256):
        loops = mo.case_loops(case)
        want = {(mo._EDGES[lp[i]], mo._EDGES[lp[(i + 1) % len(lp)]]) for lp in loops for i in range(len(lp))}

        def pe(e):
            a, _ = T.EDGES[e]
            return (tuple(int(x) for x in T.CORNER_OFFSETS[a]), int(T.EDGE_AXIS[e]))
        d = []
        for t in range(int(T.NTRI[case])):
            a, b, c = (int(x) for x in T.TRI[case, 3 * t:3 * t + 3])
            d += [(pe(a), pe(b)), (pe(b), pe(c)), (pe(c), pe(a))]
        ds = set(d)
        assert len(ds) == len(d)                                          # no directed edge twice
        assert {x for x in d if (x[1], x[0]) not in ds} == want, case
        assert int(T.NTRI[case]) == sum(len(lp) - 2 for lp in loops), case
    # the same on a mesh: the table-driven numpy restatement of the kernel's face pass == the tracer's mesh, as surfaces
    rs = np.random.RandomState(5)
    vol = rs.normal(0, 1, (19, 17, 13)).astype(np.float32)
    v, f, vox, axis, cells = mo.marching_cubes(vol, 0.0, return_owner=True)
    X, Y, Z = vol.shape
    ins = vol > 0
    vid = -np.ones((X * Y * Z, 3), np.int64); vid[vox, axis] = np.arange(len(v))
    faces = []
    for cell in np.unique(cells):
        i, j, k = cell // (Y * Z), (cell // Z) % Y, cell % Z
        case = sum(int(ins[i + (c & 1), j + ((c >> 1) & 1), k + ((c >> 2) & 1)]) << c for c in range(8))
        for t in range(int(T.NTRI[case])):
            tri = []
            for e in T.TRI[case, 3 * t:3 * t + 3]:
                o = T.EDGE_OWNER_OFFSET[e]
                tri.append(vid[((i + o[0]) * Y + (j + o[1])) * Z + (k + o[2]), T.EDGE_AXIS[e]])
            faces.append(tri)
    assert mo.same_surface(np.array(faces), f, cells)


def test_skimage_lewiner_when_available():
    """recon_util.py:64 calls skimage.measure.marching_cubes (Lewiner). When a box has scikit-image (this image does not), pin the
    restatement against the real thing: same vertex set (up to order), and a face count that differs only through ambiguous cells."""
    skm = pytest.importorskip('skimage.measure')
    rs = np.random.RandomState(2)
    ax = np.linspace(-1, 1, 40)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    vol = (0.55 - np.sqrt(x * x + 1.3 * y * y + 0.8 * z * z) + 0.02 * rs.normal(0, 1, x.shape)).astype(np.float32)
    v, f = mo.marching_cubes(vol, 0.0)
    mc = getattr(skm, 'marching_cubes', None) or skm.marching_cubes_lewiner
    sv, sf = mc(vol, 0.0)[:2]
    assert abs(len(sv) - len(v)) <= 1e-3 * len(v) + _ambiguous_cells(vol, 0.0)[0]
    assert mo.chamfer(sv.astype(np.float32), v) < 1e-3


def _ambiguous_cells(vol, level):
    """(# cells whose sign pattern MC33 / Lewiner may triangulate differently from a sign-only table, # surface cells):
    face-ambiguous patterns (a face with diagonal corners inside) and the interior-ambiguous pattern (two inside -- or two outside --
    corners on a body diagonal, MC33 case 4 and its supersets are already face-ambiguous except that one)."""
    quads = [(0, 1, 3, 2), (4, 5, 7, 6), (0, 1, 5, 4), (2, 3, 7, 6), (0, 2, 6, 4), (1, 3, 7, 5)]      # corner c = x | y << 1 | z << 2
    amb = np.zeros(256, bool)
    for c in range(256):
        b = [(c >> k) & 1 for k in range(8)]
        face = any(b[p] == b[r] and b[q] == b[s] and b[p] != b[q] for p, q, r, s in quads)
        ones = [k for k in range(8) if b[k]]; zeros = [k for k in range(8) if not b[k]]
        diag = any(len(g) == 2 and g[0] ^ g[1] == 7 for g in (ones, zeros))
        amb[c] = face or diag
    ins = vol > level
    res = vol.shape
    case = np.zeros(tuple(r - 1 for r in res), np.int32)
    for k in range(8):
        dx, dy, dz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        case |= ins[dx:res[0] - 1 + dx, dy:res[1] - 1 + dy, dz:res[2] - 1 + dz].astype(np.int32) << k
    active = (case > 0) & (case < 255)
    return int(amb[case[active]].sum()), int(active.sum())


def test_lewiner_gap_is_bounded_on_the_body_sdf_256():
    """What skimage's Lewiner (MC33 topology) can do differently from the sign-only table, on BASELINE's 256^3 body volume
    (north_star: vertex count and Chamfer within 1e-3):
      * vertices on grid edges are the same in every variant (one per sign-changing edge, linear interpolation);
      * a face-ambiguous cell may take the other face diagonal: different triangles over the SAME vertices -- both resolutions are
        enumerated here (`face_rule` 'separate' / 'join') and give identical vertex sets, hence Chamfer 0 between them;
      * MC33 may add ONE cell-centre vertex in an ambiguous cell (its cases 7.3, 10.x, 12.x, 13.x): at most one extra vertex per
        ambiguous cell, less than a voxel diagonal away from the cell's other vertices.
    So |dV| / V <= n_ambiguous / V and the symmetric Chamfer distance <= 0.5 * (n_ambiguous / V) * voxel diagonal."""
    body = synth.SynthBody(); fr = synth.make_frame(body)
    res = (256, 256, 256)
    pts = synth.volume_points(fr['cano_bounds'], res)
    vol = np.empty(len(pts), np.float32)
    for s0 in range(0, len(pts), 1 << 21):                          # chunks bound the float64 temporaries of the analytic SDF
        vol[s0:s0 + (1 << 21)] = synth.body_sdf(pts[s0:s0 + (1 << 21)], synth.cano_pose())
    del pts
    vol = vol.reshape(res)
    n_amb, n_cells = _ambiguous_cells(vol, 0.0)
    ins = vol > 0
    V = int((ins[1:] != ins[:-1]).sum() + (ins[:, 1:] != ins[:, :-1]).sum() + (ins[:, :, 1:] != ins[:, :, :-1]).sum())
    voxel = (fr['cano_bounds'][1] - fr['cano_bounds'][0]) / np.array(res, np.float32)
    diag = float(np.linalg.norm(voxel))
    print('256^3 body SDF: %d vertices, %d surface cells, %d ambiguous (%.4f%%); bound on |dV|/V %.2e, on Chamfer %.2e m'
          % (V, n_cells, n_amb, 100.0 * n_amb / n_cells, n_amb / V, 0.5 * n_amb / V * diag))
    assert V > 100_000
    assert n_amb / V < 1e-3
    assert 0.5 * (n_amb / V) * diag < 1e-3
    # both face resolutions, explicitly, on a sub-volume that contains ambiguous cells: same vertices, different triangles
    sub = vol[96:160, 64:192, :]
    va, fa = mo.marching_cubes(sub, 0.0, face_rule='separate')
    vb, fb = mo.marching_cubes(sub, 0.0, face_rule='join')
    assert np.array_equal(va, vb) and mo.chamfer(va, vb) == 0.0
    if _ambiguous_cells(sub, 0.0)[0]:
        assert len(fa) != len(fb) or not np.array_equal(fa, fb)


def test_marching_cubes_ambiguous_cells_are_rare():
    """quantifies the one unpinned boundary (skimage's Lewiner topology): only cells with a face-ambiguous sign pattern can be
    triangulated differently, and on a smooth body SDF they are a fraction of a per cent of the surface cells"""
    quads = [(0, 1, 3, 2), (4, 5, 7, 6), (0, 1, 5, 4), (2, 3, 7, 6), (0, 2, 6, 4), (1, 3, 7, 5)]      # corner c = x | y << 1 | z << 2
    amb = np.zeros(256, bool)
    for c in range(256):
        b = [(c >> k) & 1 for k in range(8)]
        amb[c] = any(b[p] == b[r] and b[q] == b[s] and b[p] != b[q] for p, q, r, s in quads)
    body = synth.SynthBody(); fr = synth.make_frame(body)
    res = (64, 64, 64)
    vol = synth.body_sdf(synth.volume_points(fr['cano_bounds'], res), synth.cano_pose()).reshape(res)
    ins = vol > 0
    case = np.zeros(tuple(r - 1 for r in res), np.int32)
    for k in range(8):
        dx, dy, dz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        case |= ins[dx:res[0] - 1 + dx, dy:res[1] - 1 + dy, dz:res[2] - 1 + dz].astype(np.int32) << k
    active = (case > 0) & (case < 255)
    frac = float(amb[case[active]].mean())
    assert active.sum() > 5000 and frac < 5e-3, frac
    # and the triangulation the repo uses is watertight there too (consistent face rule): every edge of the mesh is shared by two faces
    v, f = mo.marching_cubes(vol, 0.0)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0), axis=1)
    u, counts = np.unique(e, axis=0, return_counts=True)
    assert set(np.unique(counts)) <= {1, 2}
    # open edges only where the body itself leaves the volume (the head touches the +y bound): both end points on a boundary plane
    open_v = v[u[counts == 1].reshape(-1)]
    on_bound = ((open_v == 0) | (open_v == np.array(res, np.float32) - 1)).any(axis=1)
    assert on_bound.all() and (counts == 1).sum() < 50
