"""GPU (-m gpu): off-screen rasteriser, normal canonicalisation and fusion stage through the C ABI,
against the numpy oracle (same arithmetic contract -> coverage and depth resolution must be bit-identical) and against the goldens
made by the reference's own render_cano_mesh / canonicalize_normal_map / merge_normal_images (tests/golden/gen_raster_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, tpose_scene
from avatarcap_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from avatarcap_b200.engine import Engine
    e = Engine()
    yield e
    e.close()


@pytest.fixture(scope='module')
def g():
    return load_golden('raster_golden.npz')


def _cmp(dev_img, ref, what):
    d = dev_img.cpu().numpy()
    assert d.shape == ref.shape, (d.shape, ref.shape)
    cov_d, cov_r = d[..., 3] > 0, ref[..., 3] > 0
    assert np.array_equal(cov_d, cov_r), '%s: %d pixels differ in coverage' % (what, int((cov_d != cov_r).sum()))
    err = float(np.abs(d - ref).max())
    print('%s: coverage %.1f%%, max-abs %.3g, bit-identical %s' % (what, 100 * cov_r.mean(), err, np.array_equal(d, ref)))
    assert err < 1e-6


@pytest.mark.parametrize('size', [512, 96])
def test_rasterize_vs_oracle_cano_views(eng, g, size):
    from oracle import raster_oracle as ro
    fm, bm = ro.cano_view_matrices(g['center'])
    _cmp(eng.rasterize(g['v'], g['f'], g['n'], fm, size, size), ro.rasterize(g['v'], g['f'], g['n'], fm, size, size), 'front %d' % size)
    back = eng.rasterize(g['v'], g['f'], g['n'], bm, size, size, flip_x=True)
    _cmp(back, np.ascontiguousarray(ro.rasterize(g['v'], g['f'], g['n'], bm, size, size)[:, ::-1]), 'back %d' % size)
    rgb = eng.rasterize(g['v'], g['f'], g['n'], fm, size, size, channels=3)
    assert torch.equal(rgb, eng.rasterize(g['v'], g['f'], g['n'], fm, size, size)[..., :3])


def test_rasterize_soup_perspective_big_triangles_and_edge_cases(eng):
    """triangle soup + position shader + perspective; large triangles take the cooperative pass; culling on/off; non-square target"""
    from oracle import raster_oracle as ro
    from avatarcap_b200._lib import AvcError
    rs = np.random.RandomState(0)
    v = rs.uniform(-1, 1, (60, 3)).astype(np.float32); v[:, 2] += 3
    proj = ro.gl_perspective_projection_matrix(300, 300, 128, 128, 256, 200, gl_space=False)
    for cull in (True, False):
        d = eng.rasterize(v, None, None, proj, 256, 200, bg=(0.25, 0.5, 0.75), cull=cull)
        _cmp(d, ro.rasterize(v, None, None, proj, 256, 200, bg=(0.25, 0.5, 0.75), cull=cull), 'soup cull=%s' % cull)
    # behind the eye / outside the depth range / out-of-range indices: dropped, never read; empty mesh = background
    v2 = np.array([[0, 0, 3], [1, 0, 3], [0, 1, -1], [0, 0, 500], [1, 0, 500], [0, 1, 500]], np.float32)
    assert not eng.rasterize(v2, None, None, proj, 64, 64, cull=False).any()
    bad = np.array([[0, 1, 7], [-1, 0, 1]], np.int32)
    assert not eng.rasterize(v2[:3] * np.float32([1, 1, 1]), bad, None, proj, 64, 64, cull=False).any()
    e = eng.rasterize(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), None, proj, 8, 8, bg=(1, 0, 0))
    assert float(e[..., 0].min()) == 1.0 and not e[..., 3].any()
    # equal depth: the first triangle wins (GL_LESS); shared edges are drawn once (top-left rule)
    ortho = np.identity(4, np.float32)
    two = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]] * 2, np.float32)
    col = np.array([[1, 0, 0]] * 3 + [[0, 1, 0]] * 3, np.float32)
    h = eng.rasterize(two, None, col, ortho, 16, 16)
    assert float(h[..., 0].max()) == 1.0 and float(h[..., 1].max()) == 0.0
    quad = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0.5, 0.5, 0], [-0.5, 0.5, 0]], np.float32)
    cnt = sum(eng.rasterize(quad, np.array([t], np.int32), quad, ortho, 8, 8)[..., 3] for t in ([0, 1, 2], [0, 2, 3]))
    assert float(cnt.max()) == 1.0 and float(cnt.sum()) == 16.0
    with pytest.raises(AvcError):
        eng.rasterize(v, None, None, proj, 64, 64, channels=2)
    with pytest.raises(ValueError):
        eng.rasterize(v, None, v[:10], proj, 64, 64)


def test_renderer_mirror_and_render_cano_mesh_vs_reference_golden(eng, g):
    from avatarcap_b200 import render
    img = int(g['img'])
    r = render.Renderer(img, img, shader_name='vertex_attribute', window_name='Normal', engine=eng)
    front, back = render.render_cano_mesh(r, g['v'], g['n'], g['f'], g['center'])
    assert front.shape == (img, img, 3) and front.dtype == np.float32
    assert np.array_equal(np.packbits(np.linalg.norm(front, axis=-1) > 0), g['front_mask'])
    assert np.array_equal(np.packbits(np.linalg.norm(back, axis=-1) > 0), g['back_mask'])
    assert np.abs(front - g['front'].astype(np.float32)).max() < 1e-3 and np.abs(back - g['back'].astype(np.float32)).max() < 1e-3
    # the reference's own calling convention: triangle soup through set_model / set_mvp_mat / render (visualize_util.py:12-48)
    fm, _ = render.cano_view_matrices(g['center'])
    r.set_model(g['v'][g['f'].reshape(-1)], g['n'][g['f'].reshape(-1)])
    r.set_mvp_mat(fm); r.set_mv_mat(np.identity(4))
    soup = r.render()
    assert soup.shape == (img, img, 4) and np.array_equal(soup[..., :3], front) and np.array_equal(soup[..., 3] > 0, np.linalg.norm(front, axis=-1) > 0)
    with pytest.raises(ValueError):
        render.Renderer(8, 8, shader_name='bogus', engine=eng)


def test_canonicalize_normal_map_vs_reference_golden(eng, g):
    from avatarcap_b200 import render
    img = int(g['img']); fx, fy, cx, cy = [float(x) for x in g['cam']]
    nm = g['normal_map'].astype(np.float32)
    front, back, vn = render.canonicalize_normal_map_device(eng, g['v'], g['live_v'], g['f'], nm, g['vert_mats'], g['mv'], fx, fy, cx, cy, g['center'], img)
    vn = vn.cpu().numpy()
    bad = np.abs(vn - g['vn']).max(-1) > 1e-5        # a vertex projecting onto a pixel boundary may round to the neighbouring pixel
    print('canonicalised normals: %d valid, %d of %d vertices differ' % (int((np.linalg.norm(vn, axis=-1) > 0).sum()), int(bad.sum()), len(bad)))
    assert bad.mean() < 5e-3
    fi, bi = g['fi'].astype(np.float32), g['bi'].astype(np.float32)
    assert (np.abs(front.cpu().numpy() - fi).max(-1) > 2e-3).mean() < 1e-3 and (np.abs(back.cpu().numpy() - bi).max(-1) > 2e-3).mean() < 1e-3
    # numpy-in / numpy-out mirror with the reference's argument list
    pr = render.Renderer(img, img, shader_name='position', engine=eng); ar = render.Renderer(img, img, shader_name='vertex_attribute', engine=eng)
    f2, b2 = render.canonicalize_normal_map(pr, ar, g['v'], g['live_v'], g['f'], nm, torch.from_numpy(g['vert_mats']), g['mv'], fx, fy, cx, cy, g['center'])
    assert np.array_equal(f2, front.cpu().numpy()) and np.array_equal(b2, back.cpu().numpy())


def test_fusion_on_device_vs_reference_golden(eng, g):
    from avatarcap_b200 import render
    img = int(g['img'])
    front, _ = render.render_cano_mesh_device(eng, g['v'], g['n'], g['f'], g['center'], img)
    fi, _ = render.render_cano_mesh_device(eng, g['v'], g['vn'], g['f'], g['center'], img)
    cover = render.merge_normal_images_cover(front.clone(), fi)
    assert np.abs(cover.cpu().numpy()[::4, ::4] - g['cover_sub']).max() < 2e-5


def test_full_frame_with_fusion_stage(eng):
    """avatar field -> mesh -> avatar normal maps + canonicalised image normals -> HGFilter encoder -> recon field -> mesh -> LBS,
    nothing leaves the device; the normal maps must equal the oracle rasteriser's on the same (device-extracted) mesh."""
    from oracle import raster_oracle as ro
    from avatarcap_b200 import pipeline, encoders, render
    s = tpose_scene(128)
    s['frame'] = synth.make_frame(s['body'], synth.random_pose(3, 0.3))
    fr = s['frame']; res = (64, 64, 32); img = 256
    eng.load_avatar(s['avatar_sd']); eng.load_recon(s['recon_sd'])
    frame_dev = {k: torch.from_numpy(fr[k]).to(eng.device) for k in ('cano_smpl_v', 'smpl_skinning_weights', 'cano2live_jnt_mats')}
    frame_dev.update(cano_bounds=fr['cano_bounds'], cano_smpl_center=fr['cano_smpl_center'])
    # a smooth closed body instead of the random-weight field, so that the maps look like a person
    grid = eng.make_grid(fr['cano_bounds'], res)
    vol = torch.from_numpy(synth.body_sdf(grid.cpu().numpy(), synth.cano_pose()).reshape(res)).to(eng.device)
    v, f, n = eng.extract_mesh(vol, fr['cano_bounds'], 0.0)
    avatar = {'verts': v, 'faces': f, 'normals': n}
    live_c = 0.5 * (fr['live_smpl_v'].max(0) + fr['live_smpl_v'].min(0))
    mv = np.identity(4, np.float32); mv[:3, :3] = np.diag([1., -1., -1.]).astype(np.float32); mv[:3, 3] = -(mv[:3, :3] @ live_c) + np.float32([0, 0, 2.6])
    cam = dict(fx=280.0, fy=280.0, cx=img / 2, cy=img / 2)
    normal_map = torch.zeros((img, img, 3), device=eng.device); normal_map[..., 2] = -1.0       # "every pixel sees a camera-facing normal"
    out = pipeline.fused_normal_maps(eng, avatar, frame_dev, normal_map, cam, mv, img=img)
    assert tuple(out['front_normal'].shape) == (1, 3, img, img) and tuple(out['back_normal'].shape) == (1, 3, img, img)
    rf, rb = ro.render_cano_mesh(v.cpu().numpy(), n.cpu().numpy(), f.cpu().numpy(), fr['cano_smpl_center'], img)
    assert np.array_equal(out['front_avatar_normal'].cpu().numpy(), rf)
    assert np.array_equal(out['back_normal'][0].permute(1, 2, 0).cpu().numpy(), rb)
    fin = out['front_image_normal']
    assert 0.02 < float((fin.norm(dim=-1) > 0).float().mean()) < 0.5
    # visible vertices got the image normal rotated back: mv^-1 * (0, 0, +1) = world -z ... i.e. canonical normals of unit length
    ln = fin[fin.norm(dim=-1) > 0].norm(dim=-1)
    assert 0.9 < float(ln.median()) < 1.1 and float(ln.max()) < 1.2     # unit normals (pixels between a visible and a hidden vertex are shorter)
    ie = encoders.ImageFeatureEncoder(synth.hgfilter_state_dict(), device=eng.device, use_graph=False, benchmark=False)   # no cuDNN autotune: one call only
    fmap = ie(torch.cat([out['front_normal'], out['back_normal']], 1))                        # arch_recon.py:51-52
    assert tuple(fmap.shape) == (1, 32, img // 2, img // 2)
    rec = pipeline.recon_frame(eng, frame_dev, fmap, res, iso=0.5)
    assert rec['volume'].shape == res and torch.isfinite(rec['volume']).all()
    assert rec['live_verts'].shape == rec['verts'].shape


def test_c_client_of_the_abi(tmp_path):
    """tests/abi_smoke.c: a plain C99 program (no torch, no Python) drives grid, marching cubes, the rasteriser and the error paths"""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    if shutil.which('gcc') is None or not os.path.exists(os.path.join(cuda, 'include', 'cuda_runtime_api.h')):
        pytest.skip('no gcc / CUDA headers on this box')
    exe = str(tmp_path / 'abi_smoke')
    libdir = os.path.join(root, 'avatarcap_b200')
    cmd = ['gcc', '-std=c99', '-O1', os.path.join(root, 'tests', 'abi_smoke.c'), '-I', os.path.join(root, 'include'), '-I', os.path.join(cuda, 'include'),
           '-L', libdir, '-lavatarcap_b200', '-L', os.path.join(cuda, 'lib64'), '-lcudart', '-lm', '-Wl,-rpath,' + libdir, '-Wl,-rpath,' + os.path.join(cuda, 'lib64'),
           '-o', exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout.strip(), r.stderr.strip())
    assert r.returncode == 0 and 'abi_smoke ok' in r.stdout, r.stdout + r.stderr
