"""GPU (-m gpu): the CUDA path, called through the C ABI (avatarcap_b200.engine -> libavatarcap_b200.so), against the
CPU oracle on the same seeded inputs and against the golden vectors made by the reference's own modules.

Tolerances (BASELINE.json north_star): occupancy within 1e-4 abs; vertex count and Chamfer within 1e-3.
Integer / index outputs (faces, KNN indices, scatter) are compared bit-exactly."""
import numpy as np
import pytest
import torch

from helpers import golden_scene, tpose_scene, load_golden, maxabs
from avatarcap_b200 import synth

pytestmark = pytest.mark.gpu

IMPLS = ['simt', 'tc2']


@pytest.fixture(scope='module')
def eng():
    from avatarcap_b200.engine import Engine
    e = Engine()
    yield e
    e.close()


@pytest.fixture(scope='module')
def scene(eng):
    s = golden_scene()
    eng.load_avatar(s['avatar_sd']); eng.load_recon(s['recon_sd'])
    return s


def _impl_ok(eng, impl):
    if impl in ('tc', 'tc2') and not eng.has_tensor_core_path:
        pytest.fail('tensor-core path not available on this build/device: the product path must exist on the B200')


@pytest.mark.parametrize('impl', IMPLS)
def test_avatar_vs_golden(eng, scene, impl):
    _impl_ok(eng, impl)
    g = load_golden('avatar_golden.npz')
    eng.set_pose_feature_map(scene['pose_map'])
    o = eng.eval_occupancy(g['pts'], g['center'], want_offsets=True, want_texture=True, impl=impl)
    torch.cuda.synchronize()
    tol_occ = 1e-4
    assert maxabs(o['occ'].cpu().numpy(), g['cano_pts_ov'][:, 0]) < tol_occ
    assert maxabs(o['off'].cpu().numpy(), g['nonrigid_offset']) < 2e-6
    assert maxabs(o['rgb'].cpu().numpy(), g['rgb']) < 1e-5
    assert maxabs(o['alpha'].cpu().numpy(), g['alpha'][:, 0]) < 1e-4 * max(1.0, float(np.abs(g['alpha']).max()))   # density is O(100)
    off = eng.eval_warp(g['pts'], g['center'], impl=impl)
    assert maxabs(off.cpu().numpy(), g['warp_query']) < 2e-6
    q = g['pts'] + g['nonrigid_offset']
    rgb, alpha, occ = eng.eval_template(q, impl=impl)
    assert maxabs(occ.cpu().numpy(), g['occ'][:, 0]) < tol_occ
    assert maxabs(rgb.cpu().numpy(), g['rgb']) < 1e-5


@pytest.mark.parametrize('impl', IMPLS)
def test_avatar_occupancy_type_and_errors(eng, scene, impl):
    _impl_ok(eng, impl)
    g = load_golden('avatar_golden.npz')
    eng.set_pose_feature_map(scene['pose_map'])
    a = eng.eval_occupancy(g['pts'][:300], g['center'], if_type='sdf', impl=impl)['occ']
    b = eng.eval_occupancy(g['pts'][:300], g['center'], if_type='occupancy', impl=impl)['occ']
    assert maxabs(torch.sigmoid(a).cpu().numpy(), b.cpu().numpy()) < 1e-6      # arch_avatar.py:77-80
    with pytest.raises(ValueError):
        eng.eval_occupancy(g['pts'][:4], g['center'], if_type='bogus', impl=impl)  # config.py:22 / arch_avatar.py:82


@pytest.mark.parametrize('impl', IMPLS)
@pytest.mark.parametrize('n', [0, 1, 63, 64, 65, 127, 129, 1000])
def test_ragged_sizes(eng, scene, impl, n):
    """empty and ragged point lists: tile tails must not change any value"""
    _impl_ok(eng, impl)
    g = load_golden('avatar_golden.npz')
    eng.set_pose_feature_map(scene['pose_map']); eng.set_image_feature_map(scene['image_map'])
    o = eng.eval_occupancy(g['pts'][:n], g['center'], impl=impl)
    assert o['occ'].shape == (n,) and o['off'].shape == (n, 3)
    if n:
        assert maxabs(o['occ'].cpu().numpy(), g['cano_pts_ov'][:n, 0]) < 1e-4
    r = eng.eval_recon(g['pts'][:n], g['center'], impl=impl)
    assert r.shape == (n,)
    if n:
        assert maxabs(r.cpu().numpy(), load_golden('recon_golden.npz')['ov'][0, :n]) < 1e-4


@pytest.mark.parametrize('impl', IMPLS)
def test_recon_vs_golden(eng, scene, impl):
    _impl_ok(eng, impl)
    g = load_golden('recon_golden.npz')
    eng.set_image_feature_map(scene['image_map'])
    ov = eng.eval_recon(g['pts'], g['center'], impl=impl)
    assert maxabs(ov.cpu().numpy(), g['ov'][0]) < 1e-4


def test_api_mirror_shapes(eng, scene):
    """the reference-facing functions keep the reference's dict keys / shapes (arch_avatar.py:379-381, arch_recon.py:74)"""
    from avatarcap_b200 import api
    g = load_golden('avatar_golden.npz')
    dev = eng.device
    batch = {'cano_pts': torch.from_numpy(g['pts'])[None].to(dev), 'cano_smpl_center': torch.from_numpy(g['center'])[None].to(dev)}
    fmap = torch.from_numpy(scene['pose_map'])[None].to(dev)
    out = api.occupancy_query(eng, batch, fmap)
    assert set(out) == {'cano_pts_ov', 'nonrigid_offset'}
    assert out['cano_pts_ov'].shape == (1, len(g['pts']), 1) and out['nonrigid_offset'].shape == (1, len(g['pts']), 3)
    assert maxabs(out['cano_pts_ov'][0].cpu().numpy(), g['cano_pts_ov']) < 1e-4
    ov = api.recon_infer(eng, batch, torch.from_numpy(scene['image_map'])[None].to(dev))
    assert ov.shape == (1, len(g['pts']))
    rgb, alpha, occ = api.template_forward(eng, batch['cano_pts'])
    assert rgb.shape == (1, len(g['pts']), 3) and alpha.shape == (1, len(g['pts']), 1) and occ.shape == (1, len(g['pts']), 1)


def test_state_errors():
    from avatarcap_b200.engine import Engine
    from avatarcap_b200._lib import AvcError
    e = Engine()
    with pytest.raises(AvcError):
        e.eval_occupancy(np.zeros((4, 3), np.float32), [0, 0, 0])      # weights not loaded
    e.load_avatar(synth.avatar_state_dict())
    with pytest.raises(AvcError):
        e.eval_occupancy(np.zeros((4, 3), np.float32), [0, 0, 0])      # feature map not set
    bad = synth.avatar_state_dict(); bad['warping_field.mlp.conv1.weight'] = np.zeros((256, 99, 1), np.float32)
    with pytest.raises((ValueError, AssertionError)):
        e.load_avatar(bad)
    e.close()


# --------------------------------------------------------------------------------------------- LBS / KNN
def test_knn_exact(eng, scene):
    from oracle import field_oracle as fo
    fr = scene['frame']
    rs = np.random.RandomState(5)
    q = (fr['cano_smpl_v'][rs.randint(0, synth.N_VERTS, 5000)] + rs.normal(0, 0.05, (5000, 3))).astype(np.float32)
    for K in (1, 4):
        d2, idx = eng.knn(q, fr['cano_smpl_v'], K)
        rd, ri = fo.knn_points(torch.from_numpy(q), torch.from_numpy(fr['cano_smpl_v']), K)
        assert np.array_equal(idx.cpu().numpy(), ri.numpy())
        assert np.array_equal(d2.cpu().numpy(), rd.numpy())          # same summation order, no FMA contraction


def test_lbs_vs_golden(eng, scene):
    g = load_golden('lbs_golden.npz'); fr = scene['frame']
    lbs = eng.lbs_weights(g['verts'], fr['cano_smpl_v'], fr['smpl_skinning_weights'])
    assert maxabs(lbs.cpu().numpy(), g['lbs']) < 2e-6
    live, mats = eng.skin_points(g['verts'], lbs, fr['cano2live_jnt_mats'], True)
    assert maxabs(live.cpu().numpy(), g['live']) < 3e-6 and maxabs(mats.cpu().numpy(), g['mats']) < 3e-6
    ln = eng.skin_normals(g['normals'], lbs, fr['cano2live_jnt_mats'])
    assert maxabs(ln.cpu().numpy(), g['live_normals']) < 3e-6
    v2, n2 = eng.skin_mesh(g['verts'], g['normals'], fr['cano_smpl_v'], fr['smpl_skinning_weights'], fr['cano2live_jnt_mats'])
    assert maxabs(v2.cpu().numpy(), g['live']) < 3e-6 and maxabs(n2.cpu().numpy(), g['live_normals']) < 3e-6
    from avatarcap_b200 import api
    su = api.SmplUtil(fr['smpl_skinning_weights'], eng)
    with pytest.raises(ValueError):
        su.calculate_lbs(torch.from_numpy(g['verts'])[None])        # smpl_util.py:30-31
    su.set_cano_smpl_vertices(torch.from_numpy(fr['cano_smpl_v']))
    assert su.calculate_lbs(torch.from_numpy(g['verts'])[None]).shape == (1, len(g['verts']), 24)


@pytest.mark.parametrize('space', ['posed', 'cano', 'temp'])
def test_geotex_forward_vs_golden(eng, scene, space):
    from avatarcap_b200 import api
    g = load_golden('avatar_golden.npz'); fr = scene['frame']; dev = eng.device
    wvol = torch.from_numpy(synth.blend_weight_volume(fr)).to(dev)
    batch = {k: torch.from_numpy(fr[k])[None].to(dev) for k in ('cano_smpl_center', 'cano_bounds', 'cano2live_jnt_mats', 'live_smpl_v')}
    w = torch.from_numpy((g['fwd_wpts_live'] if space == 'posed' else g['fwd_wpts_cano']).copy())[None].to(dev)
    o = api.geotex_forward(eng, w, torch.from_numpy(g['fwd_dists'])[None].to(dev), batch, torch.from_numpy(scene['pose_map'])[None].to(dev),
                           torch.from_numpy(fr['smpl_skinning_weights']).to(dev), torch.from_numpy(fr['cano_smpl_v']).to(dev), wvol, space)
    assert set(o) == {'raw', 'occ', 'nonrigid_offset'}
    assert maxabs(o['nonrigid_offset'][0].cpu().numpy(), g['fwd_%s_off' % space]) < 3e-6
    assert maxabs(o['occ'][0].cpu().numpy(), g['fwd_%s_occ' % space]) < 2e-4      # posed: + inverse-LBS rounding through the 2^9 PE
    assert maxabs(o['raw'][0].cpu().numpy(), g['fwd_%s_raw' % space]) < 1e-4
    assert maxabs(w[0].cpu().numpy(), g['fwd_%s_wpts_after' % space]) < 3e-6      # 'cano' mutates the input in place


# --------------------------------------------------------------------------------------------- grid / scatter / mesh
def test_grid_and_scatter(eng):
    from oracle import field_oracle as fo
    g = load_golden('mesh_golden.npz')
    for res in ((7, 9, 5), (64, 33, 128)):
        pts = eng.make_grid(g['bounds'], res).cpu().numpy()
        ref = fo.generate_volume_points(g['bounds'], res)
        assert maxabs(pts, ref) < 2.5e-7
    assert np.array_equal(eng.make_grid(g['bounds'], (7, 9, 5)).cpu().numpy(), g['vol_pts_7_9_5'])
    slab = eng.make_grid(g['bounds'], (64, 33, 128), 10, 7).cpu().numpy()
    assert np.array_equal(slab, eng.make_grid(g['bounds'], (64, 33, 128)).cpu().numpy()[10 * 33 * 128:17 * 33 * 128])
    rs = np.random.RandomState(2)
    for n in (1, 1000, 1024 * 3 + 17):
        flag = rs.rand(n) < 0.3
        vals = rs.normal(0, 1, int(flag.sum())).astype(np.float32); fill = rs.normal(0, 1, int((~flag).sum())).astype(np.float32)
        out = eng.scatter_fill(torch.from_numpy(flag), torch.from_numpy(vals), torch.from_numpy(fill)).cpu().numpy()
        assert np.array_equal(out, fo.scatter_fill(flag, vals, fill, (n,)))


def _mesh_case(eng, vol, bounds, iso):
    from oracle import mesh_oracle as mo
    res = vol.shape
    v, f, n = eng.extract_mesh(torch.from_numpy(vol), bounds, iso)
    rv, rf, rn, cells = mo.recon_mesh(vol, res, bounds, iso, return_cells=True)
    assert v.shape[0] == rv.shape[0] and f.shape[0] == rf.shape[0]               # exact counts
    # exact topology, cell by cell, against the oracle's INDEPENDENT tracer (blind to the fan each side chose inside a loop)
    assert mo.same_surface(f.cpu().numpy(), rf, cells)
    assert maxabs(v.cpu().numpy(), rv) < 1e-6
    good = np.linalg.norm(rn, axis=1) > 0.5
    assert maxabs(n.cpu().numpy()[good], rn[good]) < 2e-4
    return v, f, n


def test_mesh_sphere_and_noise(eng):
    bounds = np.array([[-0.9, -1.0, -0.35], [0.95, 0.9, 0.3]], np.float32)
    for res in ((32, 32, 32), (40, 24, 18), (2, 2, 2), (5, 3, 2)):
        ii, jj, kk = np.meshgrid(*[np.arange(r) for r in res], indexing='ij')
        c = np.array(res) / 2.0 - 0.3
        vol = (min(res) / 3.0 - np.sqrt((ii - c[0]) ** 2 + (jj - c[1]) ** 2 + (kk - c[2]) ** 2)).astype(np.float32)
        _mesh_case(eng, vol, bounds, 0.0)
    rs = np.random.RandomState(11)
    vol = rs.normal(0, 1, (33, 29, 31)).astype(np.float32)
    _mesh_case(eng, vol, bounds, 0.0)
    _mesh_case(eng, (1 / (1 + np.exp(-vol))).astype(np.float32), bounds, 0.5)   # recon iso (recon_util.py:51 default)
    nv, nf = eng.mc_count(torch.from_numpy(vol), 100.0)
    assert (nv, nf) == (0, 0)
    from avatarcap_b200 import api
    with pytest.raises(ValueError):
        api.recon_mesh(eng, torch.from_numpy(vol).cuda(), vol.shape, bounds, 100.0)


def test_mc_emit_counted_reuse_and_fallback(eng):
    """avc_mc_emit_counted: reuses the preceding count's scan, and recounts when anything intervened or an argument differs"""
    import ctypes as C
    from avatarcap_b200 import _lib
    rs = np.random.RandomState(12)
    bounds = np.array([[-1, -1, -0.4], [1, 1, 0.4]], np.float32)
    vol = torch.from_numpy(rs.normal(0, 1, (40, 36, 20)).astype(np.float32)).to(eng.device)
    vol2 = torch.from_numpy(rs.normal(0, 1, (40, 36, 20)).astype(np.float32)).to(eng.device)
    ref = [t.clone() for t in eng.extract_mesh(vol, bounds, 0.0)]                       # count + emit_counted
    ref2 = [t.clone() for t in eng.extract_mesh(vol2, bounds, 0.25)]

    def emit(fn, v_, iso, nv, nf):
        verts = torch.empty((nv, 3), device=eng.device); faces = torch.empty((nf, 3), device=eng.device, dtype=torch.int32); nrm = torch.empty_like(verts)
        eng._check(fn(eng._h, C.c_void_p(v_.data_ptr()), _lib.i3(v_.shape), _lib.f6(bounds.reshape(6)), float(iso), 0, 0, 0, v_.shape[0],
                      C.c_void_p(verts.data_ptr()), C.c_void_p(nrm.data_ptr()), C.c_void_p(faces.data_ptr()), nv, nf, eng._stream()))
        return verts, faces, nrm
    nv, nf = eng.mc_count(vol, 0.0)
    eng.rasterize(np.zeros((3, 3), np.float32), None, None, np.identity(4, np.float32), 8, 8)       # overwrites the scratch buffer
    for a, b in zip(emit(eng.lib.avc_mc_emit_counted, vol, 0.0, nv, nf), ref):
        assert torch.equal(a, b)
    nv, nf = eng.mc_count(vol, 0.0)                                                                 # count for vol, emit for vol2 / other iso
    nv2, nf2 = ref2[0].shape[0], ref2[1].shape[0]
    for a, b in zip(emit(eng.lib.avc_mc_emit_counted, vol2, 0.25, nv2, nf2), ref2):
        assert torch.equal(a, b)
    for a, b in zip(emit(eng.lib.avc_mc_emit, vol, 0.0, nv, nf), ref):
        assert torch.equal(a, b)


def test_mc_extract_capacity_bounded_async(eng):
    """avc_mc_extract: no host round trip, device-side counts; too small a capacity sets the overflow flag, keeps the counts exact
    and fills the first cap entries -- never a silent truncation. Many chunks (every block sums the chunk records before it) and both iso values."""
    rs = np.random.RandomState(21)
    bounds = np.array([[-1, -1, -0.4], [1, 1, 0.4]], np.float32)
    from oracle import mesh_oracle as mo
    for shape, iso in (((96, 80, 64), 0.0), ((50, 33, 27), 0.3)):
        vol_np = rs.normal(0, 1, shape).astype(np.float32)
        vol = torch.from_numpy(vol_np).to(eng.device)
        rv, rf, rn, cells = mo.recon_mesh(vol_np, shape, bounds, iso, return_cells=True)
        nv, nf = rv.shape[0], rf.shape[0]
        assert eng.mc_count(vol, iso) == (nv, nf)
        v, f, n, counts = eng.extract_mesh_async(vol, bounds, iso, nv + 100, nf + 7)
        c = counts.tolist()
        assert c[0] == nv and c[1] == nf and c[2] == nv and c[3] == 0
        assert mo.same_surface(f[:nf].cpu().numpy(), rf, cells) and maxabs(v[:nv].cpu().numpy(), rv) < 1e-6
        for cap_v, cap_f, flags in ((nv // 2, nf + 1, 1), (nv, nf // 3, 2), (0, 0, 3)):
            v2, f2, n2, c2 = eng.extract_mesh_async(vol, bounds, iso, cap_v, cap_f)
            c2 = c2.tolist()
            assert c2[:3] == [nv, nf, nv] and c2[3] == flags
            assert torch.equal(v2[:min(cap_v, nv)], v[:min(cap_v, nv)]) and torch.equal(f2[:min(cap_f, nf)], f[:min(cap_f, nf)])
        # the front end retries an overflow with the exact sizes and remembers them
        eng._keep['mc_hint'] = (16, 16)
        v3, f3, n3 = eng.extract_mesh(vol, bounds, iso)
        assert torch.equal(v3, v[:nv]) and torch.equal(f3, f[:nf]) and torch.equal(n3, n[:nv])
        for _ in range(3):                                   # repeated runs are deterministic (ticket order does not leak into the result)
            v4, f4, n4 = eng.extract_mesh(vol, bounds, iso)
            assert torch.equal(v4, v3) and torch.equal(f4, f3)


def test_mesh_normals_vs_reference_golden(eng):
    """Sobel + trilinear normals against the reference's own conv3d / grid_sample output (mesh_golden.npz)."""
    from oracle import mesh_oracle as mo
    g = load_golden('mesh_golden.npz')
    vol = g['vol']; res = vol.shape
    bounds = np.stack([g['bounds'][0], g['bounds'][0] + g['voxel'] * np.array(res, np.float32)], 0)
    v, f, n = eng.extract_mesh(torch.from_numpy(vol), bounds, 0.0)
    grid = 2 * (v.cpu().numpy() - bounds[0]) / (bounds[1] - bounds[0]) - 1.0
    ref = -mo.extract_normal_from_volume(vol, g['voxel'], grid.astype(np.float32))
    assert maxabs(n.cpu().numpy(), ref) < 5e-4


@pytest.mark.parametrize('res', [(37, 20, 22), (37, 20, 24)])
def test_mesh_slabs_equal_whole(eng, res):
    """multi-GPU seam rule: slabs along x with (2,3) halo planes reproduce the single-volume mesh exactly
    (Rz = 22: scalar classification; Rz = 24: the float4 path) -- and both equal the CPU oracle"""
    import os
    from avatarcap_b200 import shard
    from oracle import mesh_oracle as mo
    rs = np.random.RandomState(4)
    ii, jj, kk = np.meshgrid(*[np.arange(r) for r in res], indexing='ij')
    vol = (9.0 - np.sqrt((ii - 18) ** 2 + (jj - 10) ** 2 + (kk - 11) ** 2) + 0.8 * rs.normal(0, 1, res)).astype(np.float32)
    bounds = np.array([[-0.9, -1.0, -0.35], [0.95, 0.9, 0.3]], np.float32)
    v, f, n = eng.extract_mesh(torch.from_numpy(vol), bounds, 0.0)
    rv, rf, rn, cells = mo.recon_mesh(vol, res, bounds, 0.0, return_cells=True)
    assert mo.same_surface(f.cpu().numpy(), rf, cells) and maxabs(v.cpu().numpy(), rv) < 1e-6
    # the A/B knob selects the scalar classification path: same mesh bit for bit
    os.environ['AVC_MC_SCALAR'] = '1'
    try:
        v2, f2, n2 = eng.extract_mesh(torch.from_numpy(vol), bounds, 0.0)
    finally:
        del os.environ['AVC_MC_SCALAR']
    assert torch.equal(v2, v) and torch.equal(f2, f) and torch.equal(n2, n)
    for world in (2, 3, 5):
        parts = []
        for r in range(world):
            s, e = shard.slab_range(res[0], world, r)
            lo, hi = shard.halo_planes(res[0], s, e)
            sub = torch.from_numpy(vol[s - lo:e + hi].copy())
            pv, pf, pn = eng.extract_mesh(sub, bounds, 0.0, halo_lo=lo, halo_hi=hi, x_origin=s - lo, gres_x=res[0])
            parts.append((pv.cpu().numpy(), pf.cpu().numpy(), pn.cpu().numpy()))
        mv, mf, mn = shard.merge_meshes(parts)
        assert np.array_equal(mf, f.cpu().numpy())
        assert np.array_equal(mv, v.cpu().numpy())
        assert maxabs(mn, n.cpu().numpy()) < 1e-6


# --------------------------------------------------------------------------------------------- BASELINE configs
_ORACLE_CACHE = {}


def _cached(key, fn):
    """the CPU oracle is by far the slowest part of these tests: evaluate it once per configuration, not once per implementation"""
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = fn()
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize('impl', IMPLS)
def test_config1_dense_64(eng, impl):
    """BASELINE config 1: T-pose body, 64^3 dense grid, occupancy vs the CPU oracle (restated reference)."""
    _impl_ok(eng, impl)
    from oracle import field_oracle as fo
    s = tpose_scene(64)
    eng.load_avatar(s['avatar_sd'])
    eng.set_pose_feature_map(s['pose_map'])
    fr = s['frame']
    pts = eng.make_grid(fr['cano_bounds'], (64, 64, 64))
    o = eng.eval_occupancy(pts, fr['cano_smpl_center'], impl=impl)
    ref = _cached('config1', lambda: fo.occupancy_query(s['avatar_sd'], pts.cpu().numpy(), s['pose_map'], fr['cano_smpl_center']))
    err = maxabs(o['occ'].cpu().numpy(), ref['cano_pts_ov'][:, 0])
    print('config1 %s: occ max-abs err %.3g (range %.3g..%.3g), off err %.3g' % (
        impl, err, ref['cano_pts_ov'].min(), ref['cano_pts_ov'].max(), maxabs(o['off'].cpu().numpy(), ref['nonrigid_offset'])))
    assert err < 1e-4
    assert maxabs(o['off'].cpu().numpy(), ref['nonrigid_offset']) < 2e-6
    # the split entry points (WarpingField.query, DoubleTNet.forward) over MANY tiles per CTA must reproduce the fused query
    off = eng.eval_warp(pts, fr['cano_smpl_center'], impl=impl)
    assert torch.equal(off, o['off'])
    rgb, alpha, occ = eng.eval_template(pts + off, impl=impl)
    assert maxabs(occ.cpu().numpy(), ref['cano_pts_ov'][:, 0]) < 1e-4
    assert maxabs(occ.cpu().numpy(), o['occ'].cpu().numpy()) < 2e-5
    # mesh parity on the evaluated field: vertex count within 1e-3, Chamfer within 1e-3 m
    from oracle import mesh_oracle as mo
    vol_g = o['occ'].reshape(64, 64, 64)
    v, f, n = eng.extract_mesh(vol_g, fr['cano_bounds'], 0.0)
    rv, rf, rn = _cached('config1_mesh', lambda: mo.recon_mesh(ref['cano_pts_ov'][:, 0].reshape(64, 64, 64), (64, 64, 64), fr['cano_bounds'], 0.0))
    assert abs(v.shape[0] - rv.shape[0]) <= 1e-3 * rv.shape[0] + 1
    assert mo.chamfer(v.cpu().numpy(), rv) < 1e-3


@pytest.mark.parametrize('impl', IMPLS)
def test_dense_grid_entry_equals_point_list(eng, scene, impl):
    """avc_eval_occupancy_grid / avc_eval_recon_grid: coordinates from the point index (SURVEY 8b `pts_or_grid_desc`) must give the
    bits of make_grid + the point-list entry -- whole grids, x-slabs, non-cubic and odd extents, outputs into a caller buffer"""
    _impl_ok(eng, impl)
    fr = scene['frame']
    eng.load_avatar(scene['avatar_sd']); eng.load_recon(scene['recon_sd'])
    eng.set_pose_feature_map(scene['pose_map']); eng.set_image_feature_map(scene['image_map'])
    for res, x0, xc in (((24, 20, 16), 0, None), ((37, 21, 19), 0, None), ((37, 21, 19), 11, 9), ((2, 3, 1), 0, None), ((64, 64, 32), 40, 24)):
        pts = eng.make_grid(fr['cano_bounds'], res, x0, xc)
        a = eng.eval_occupancy(pts, fr['cano_smpl_center'], want_texture=True, impl=impl)
        buf = torch.full((pts.shape[0] + 5,), -7.0, device=eng.device)
        b = eng.eval_occupancy_grid(fr['cano_bounds'], res, fr['cano_smpl_center'], x0, xc, want_texture=True, impl=impl, out_occ=buf[:pts.shape[0]])
        for k in ('occ', 'off', 'rgb', 'alpha'):
            assert torch.equal(a[k], b[k]), (res, k)
        assert float(buf[pts.shape[0]:].min()) == -7.0                               # nothing written past the slab
        assert torch.equal(eng.eval_recon(pts, fr['cano_smpl_center'], impl=impl), eng.eval_recon_grid(fr['cano_bounds'], res, fr['cano_smpl_center'], x0, xc, impl=impl))
    with pytest.raises(Exception):
        eng.eval_occupancy_grid(fr['cano_bounds'], (8, 8, 8), fr['cano_smpl_center'], 5, 4, impl=impl)   # slab beyond the grid


@pytest.mark.parametrize('impl', IMPLS)
def test_config2_dense_256_properties(eng, impl):
    """BASELINE config 2 at full size: 256^3 dense. The oracle cannot finish 16.7 M points in seconds, so:
    (a) a seeded 1 Mi-point subsample (">= 1 M-point seeded subsample") is checked against the oracle, (b) the full-grid result must equal the same points
    evaluated as an independent ragged list (tile-position independence), (c) evaluation is deterministic."""
    _impl_ok(eng, impl)
    from oracle import field_oracle as fo
    s = tpose_scene(256)
    eng.load_avatar(s['avatar_sd']); eng.set_pose_feature_map(s['pose_map'])
    fr = s['frame']
    pts = eng.make_grid(fr['cano_bounds'], (256, 256, 256))
    o = eng.eval_occupancy(pts, fr['cano_smpl_center'], want_texture=True, impl=impl)
    o2 = eng.eval_occupancy(pts, fr['cano_smpl_center'], want_offsets=False, impl=impl)
    assert torch.equal(o['occ'], o2['occ'])
    og = eng.eval_occupancy_grid(fr['cano_bounds'], (256, 256, 256), fr['cano_smpl_center'], want_texture=True, impl=impl)   # the bench's entry point
    assert all(torch.equal(o[k], og[k]) for k in ('occ', 'off', 'rgb', 'alpha'))
    del og
    rs = np.random.RandomState(9)
    sel = np.sort(rs.choice(256 ** 3, 1 << 20, replace=False))
    sel_t = torch.from_numpy(sel).to(eng.device)
    sub = eng.eval_occupancy(pts[sel_t], fr['cano_smpl_center'], impl=impl)
    assert maxabs(sub['occ'].cpu().numpy(), o['occ'][sel_t].cpu().numpy()) < 2e-6
    ref = _cached('config2', lambda: fo.occupancy_query(s['avatar_sd'], pts[sel_t].cpu().numpy(), s['pose_map'], fr['cano_smpl_center'], with_texture=True))
    assert maxabs(o['occ'][sel_t].cpu().numpy(), ref['cano_pts_ov'][:, 0]) < 1e-4
    assert maxabs(o['rgb'][sel_t].cpu().numpy(), ref['rgb']) < 1e-5
    assert maxabs(o['alpha'][sel_t].cpu().numpy(), ref['alpha'][:, 0]) < 1e-4 * max(1.0, float(np.abs(ref['alpha']).max()))


@pytest.mark.parametrize('res', [(96, 96, 48), (256, 256, 256)])
def test_config3_full_frame_masked(eng, res):
    """BASELINE config 3: masked query -> scatter (+-1 fill) -> recon_mesh(iso 0) -> LBS, and recon decoder -> scatter ->
    recon_mesh(iso 0.5) -> LBS, vs the oracle pipeline -- on a reduced grid (seconds) and at the STATED size, 256^3 (the oracle
    evaluates every valid point, ~3 M of them: about a minute of host time): volume <= 1e-4, vertex count and Chamfer <= 1e-3."""
    from oracle import field_oracle as fo
    from oracle import mesh_oracle as mo
    from avatarcap_b200 import pipeline
    big = int(np.prod(res)) > 2_000_000
    s = tpose_scene(256 if big else 128)
    s['frame'] = synth.make_frame(s['body'], synth.random_pose(3, 0.3))
    fr = s['frame']
    eng.load_avatar(s['avatar_sd']); eng.load_recon(s['recon_sd'])
    grid = eng.make_grid(fr['cano_bounds'], res)
    flag = pipeline.valid_points_flag(eng, grid, fr['cano_smpl_v'])
    if not big:
        ref_flag = fo.valid_points_flag(grid.cpu().numpy(), fr['cano_smpl_v'])
        assert np.array_equal(flag.cpu().numpy(), ref_flag)
    else:
        # the brute-force oracle flag costs 1e11 distance evaluations at 256^3: exact check on a seeded 300 k subsample, and the whole
        # grid against a float64 k-d tree outside a +-1e-5 m band around the radius (inside the band float32 rounding decides)
        from scipy.spatial import cKDTree
        g_np = grid.cpu().numpy(); ref_flag = flag.cpu().numpy()
        sub = np.sort(np.random.RandomState(5).choice(len(g_np), 300_000, replace=False))
        assert np.array_equal(ref_flag[sub], fo.valid_points_flag(g_np[sub], fr['cano_smpl_v']))
        d, _ = cKDTree(fr['cano_smpl_v'].astype(np.float64)).query(g_np.astype(np.float64), workers=-1)
        sure = np.abs(d - 0.1) > 1e-5
        assert np.array_equal(ref_flag[sure], (d < 0.1)[sure])
        del d, sure
    inside = synth.body_inside(grid.cpu().numpy()[~ref_flag], synth.cano_pose())
    fill = (2.0 * inside.astype(np.float32) - 1.0)                                   # avatarcap_dataset.py:123-124
    pts = grid[flag]
    frame_dev = {k: torch.from_numpy(fr[k]).to(eng.device) for k in ('cano_smpl_v', 'smpl_skinning_weights', 'cano2live_jnt_mats')}
    frame_dev.update(cano_bounds=fr['cano_bounds'], cano_smpl_center=fr['cano_smpl_center'])
    for kind in ('avatar', 'recon'):
        if kind == 'avatar':
            out = pipeline.avatar_frame(eng, frame_dev, s['pose_map'], res, flag, pts, torch.from_numpy(fill), iso=0.0)
            rvals = fo.occupancy_query(s['avatar_sd'], pts.cpu().numpy(), s['pose_map'], fr['cano_smpl_center'])['cano_pts_ov'][:, 0]
            iso = 0.0
        else:
            out = pipeline.recon_frame(eng, frame_dev, s['image_map'], res, flag, pts, torch.from_numpy(fill), iso=0.5)
            rvals = fo.recon_infer(s['recon_sd'], pts.cpu().numpy(), s['image_map'], fr['cano_smpl_center'])
            iso = 0.5
        rvol = fo.scatter_fill(ref_flag, rvals, fill, res)
        assert maxabs(out['volume'].cpu().numpy(), rvol) < 1e-4
        rv, rf, rn = mo.recon_mesh(rvol, res, fr['cano_bounds'], iso)
        nv = out['verts'].shape[0]
        print('%s frame %s: %d valid points (%.1f%%), verts %d (oracle %d), faces %d (oracle %d)' % (kind, 'x'.join(map(str, res)), int(ref_flag.sum()), 100 * ref_flag.mean(), nv, rv.shape[0], out['faces'].shape[0], rf.shape[0]))
        assert abs(nv - rv.shape[0]) <= 1e-3 * rv.shape[0] + 1
        assert mo.chamfer(out['verts'].cpu().numpy(), rv) < 1e-3
        if not big:
            lbs = fo.calculate_lbs(rv, fr['cano_smpl_v'], fr['smpl_skinning_weights'])
            live, _ = fo.skinning(rv, lbs, fr['cano2live_jnt_mats'])
            assert mo.chamfer(out['live_verts'].cpu().numpy(), live) < 1e-3
        else:           # oracle LBS (brute-force KNN) on a seeded 50 k subset of its vertices: each must have a skinned GPU vertex within 1e-3 m
            from scipy.spatial import cKDTree
            pick = np.sort(np.random.RandomState(6).choice(len(rv), 50_000, replace=False))
            lbs = fo.calculate_lbs(rv[pick], fr['cano_smpl_v'], fr['smpl_skinning_weights'])
            live, _ = fo.skinning(rv[pick], lbs, fr['cano2live_jnt_mats'])
            dd, _ = cKDTree(out['live_verts'].cpu().numpy()).query(live, workers=-1)
            assert float(dd.mean()) < 1e-3 and float(dd.max()) < 2e-2
        del out, rvol, rv, rf, rn


@pytest.mark.parametrize('pinned', [False, True])
def test_host_buffer_entry_points(eng, scene, pinned):
    """avc_eval_occupancy_host / avc_eval_recon_host (the end-to-end entry bench.py times): host in, host out, chunked
    3-stream pipeline; must equal the device-pointer entry bit for bit, for pageable and for page-locked buffers."""
    g = load_golden('avatar_golden.npz')
    n = (1 << 21) + 12345                                   # more than one pipeline chunk, ragged tail
    reps = n // len(g['pts']) + 1
    pts = np.tile(g['pts'], (reps, 1))[:n].copy()
    pts += np.random.RandomState(1).normal(0, 0.01, pts.shape).astype(np.float32)
    eng.set_pose_feature_map(scene['pose_map']); eng.set_image_feature_map(scene['image_map'])

    def buf(shape):
        t = torch.empty(shape, dtype=torch.float32)
        return (t.pin_memory() if pinned else t).numpy()
    p_h = buf((n, 3)); p_h[:] = pts
    occ, off, rgb, al, ov = buf(n), buf((n, 3)), buf((n, 3)), buf(n), buf(n)
    eng.eval_occupancy_host(p_h, g['center'], occ, off, rgb, al)
    eng.eval_recon_host(p_h, g['center'], ov)
    d = eng.eval_occupancy(pts, g['center'], want_texture=True)
    assert np.array_equal(occ, d['occ'].cpu().numpy()) and np.array_equal(off, d['off'].cpu().numpy())
    assert np.array_equal(rgb, d['rgb'].cpu().numpy()) and np.array_equal(al, d['alpha'].cpu().numpy())
    assert np.array_equal(ov, eng.eval_recon(pts, g['center']).cpu().numpy())


def test_nerf_vertex_colours_vs_golden(eng, scene):
    """'next' row 2: NerfRenderer.render + raw2outputs through the library, driven like main.py:464-478."""
    from avatarcap_b200 import api
    g = load_golden('nerf_golden.npz'); fr = scene['frame']; dev = eng.device
    wvol = torch.from_numpy(synth.blend_weight_volume(fr)).to(dev)
    batch = {k: torch.from_numpy(fr[k])[None].to(dev) for k in ('cano_smpl_center', 'cano_bounds', 'cano2live_jnt_mats', 'live_smpl_v')}
    R = len(g['verts'])
    rend = api.NerfRenderer.for_engine(eng, torch.from_numpy(scene['pose_map'])[None].to(dev), torch.from_numpy(fr['smpl_skinning_weights']).to(dev),
                                       torch.from_numpy(fr['cano_smpl_v']).to(dev), wvol)
    items = dict(batch)
    v = torch.from_numpy(g['verts']).to(dev); n = torch.from_numpy(g['normals']).to(dev)
    items['ray_o'] = (v + n)[None]; items['ray_d'] = -n[None]
    items['depth'] = torch.ones((1, R), device=dev); items['near'] = items['depth'] - 0.05; items['far'] = items['depth'] + 0.05
    out = rend.render(items, pts_space='cano', near_dist=0.02, far_dist=0.05)
    assert set(out) == {'rgb_map', 'acc_map', 'depth_map', 'raw', 'occ', 'nonrigid_offset'}
    assert np.array_equal(items['near'][0].cpu().numpy(), g['near_after']) and np.array_equal(items['far'][0].cpu().numpy(), g['far_after'])
    assert maxabs(out['raw'][0].cpu().numpy(), g['raw']) < 1e-4
    assert maxabs(out['rgb_map'][0].cpu().numpy(), g['rgb_map']) < 1e-4
    assert maxabs(out['acc_map'][0].cpu().numpy(), g['acc_map']) < 1e-4 and maxabs(out['depth_map'][0].cpu().numpy(), g['depth_map']) < 1e-4
    col = api.vertex_colors(rend, batch, v, n)
    assert maxabs(col.cpu().numpy(), g['rgb_map'][:, [2, 1, 0]]) < 1e-4
    # colour transfer to another vertex set (main.py:480-484)
    tgt = v[:100] + 0.001
    assert torch.equal(api.transfer_colors(eng, tgt, v, col), col[eng.knn(tgt, v, 1)[1][:, 0]])
    # compositing alone, bit-level vs the oracle's raw2outputs on the same raw
    from oracle import field_oracle as fo
    pts, z, dists = eng.ray_samples(items['ray_o'][0], items['ray_d'][0], items['near'][0], items['far'][0], 64)
    rgb_o, acc_o, dep_o = fo.raw2outputs(out['raw'][0].cpu().reshape(R, 64, 4), z.cpu())
    rgb_k, acc_k, dep_k = eng.composite(out['raw'][0], z)
    assert maxabs(rgb_k.cpu().numpy(), rgb_o.numpy()) < 2e-6 and maxabs(acc_k.cpu().numpy(), acc_o.numpy()) < 2e-6


def test_near_flag_and_grid_knn_far_queries(eng, scene):
    """uniform-grid KNN: exact also for queries far from every vertex (fallback scan) and for the bounded near-flag search"""
    from oracle import field_oracle as fo
    fr = scene['frame']
    rs = np.random.RandomState(12)
    bmin, bmax = fr['cano_bounds']
    q = (rs.uniform(0, 1, (20000, 3)) * (bmax - bmin) * 1.6 + bmin - 0.3 * (bmax - bmin)).astype(np.float32)     # many far outside the body
    rd, ri = fo.knn_points(torch.from_numpy(q), torch.from_numpy(fr['cano_smpl_v']), 4)
    import os
    for rmax in (None, '1', '3', '12'):                  # shells walked before the brute-force fallback: the result must not depend on it
        if rmax is not None:
            os.environ['AVC_KNN_RMAX'] = rmax
        try:
            d2, idx = eng.knn(q, fr['cano_smpl_v'], 4)
        finally:
            os.environ.pop('AVC_KNN_RMAX', None)
        assert np.array_equal(idx.cpu().numpy(), ri.numpy()) and np.array_equal(d2.cpu().numpy(), rd.numpy())
    for radius in (0.1, 0.08, 0.03):
        flag = eng.near_flag(q, fr['cano_smpl_v'], radius).cpu().numpy()
        assert np.array_equal(flag, (rd[:, 0] < radius ** 2).numpy())            # torch semantics: float32(radius ** 2)


from helpers import uv_sphere as _uv_sphere  # noqa: E402


def test_inside_volume_fill(eng):
    """'next' row 3: trimesh.contains on the dense grid, by crossing parity; vs the CPU ray-parity oracle and the analytic shape"""
    from oracle import mesh_oracle as mo
    from oracle import field_oracle as fo
    from avatarcap_b200 import pipeline
    bounds = np.array([[-0.9, -1.0, -0.35], [0.95, 0.9, 0.3]], np.float32)
    v1, f1 = _uv_sphere([0.1, -0.2, 0.0], 0.27); v2, f2 = _uv_sphere([-0.45, 0.5, 0.05], 0.2, 16, 20)
    verts = np.concatenate([v1, v2], 0); faces = np.concatenate([f1, f2 + len(v1)], 0)
    res = (48, 56, 24)
    inside = eng.inside_volume(verts, faces, bounds, res).cpu().numpy()
    pts = fo.generate_volume_points(bounds, res)
    ref = mo.contains_points(verts, faces, pts).reshape(res)
    assert np.array_equal(inside, ref)
    d1 = np.linalg.norm(pts - np.array([0.1, -0.2, 0.0]), axis=1).reshape(res); d2 = np.linalg.norm(pts - np.array([-0.45, 0.5, 0.05]), axis=1).reshape(res)
    sure_in = (d1 < 0.25) | (d2 < 0.18); sure_out = (d1 > 0.28) & (d2 > 0.21)
    assert inside[sure_in].all() and not inside[sure_out].any() and inside.sum() > 100
    flag = torch.from_numpy((d1 < 0.31).reshape(-1)).to(eng.device)
    fill = pipeline.invalid_points_fill(eng, verts, faces, bounds, res, flag).cpu().numpy()
    assert np.array_equal(fill, 2.0 * ref.reshape(-1)[~flag.cpu().numpy()].astype(np.float32) - 1.0)
