"""CPU: the rasterisation / normal-fusion stage.
  * oracle (numpy) against the goldens produced by the reference's own render_cano_mesh / canonicalize_normal_map /
    merge_normal_images / save_mesh_as_ply (tests/golden/gen_raster_golden.py);
  * the product's rasteriser arithmetic (csrc/raster_core.h, the header the CUDA kernels are built from) compiled for the
    host and compared bit for bit with the oracle -- host logic only, the CUDA path itself is covered by the -m gpu tests;
  * the host-side mirrors that need no GPU (PLY writer, fusion optimiser, view matrices)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import raster_oracle as ro  # noqa: E402
from helpers import load_golden  # noqa: E402


@pytest.fixture(scope='module')
def g():
    return load_golden('raster_golden.npz')


@pytest.fixture(scope='module')
def host_lib(tmp_path_factory):
    if shutil.which('g++') is None:
        pytest.skip('no g++')
    out = str(tmp_path_factory.mktemp('rh') / 'raster_host.so')
    src = os.path.join(ROOT, 'avatarcap_b200', 'csrc', 'raster_host.cpp')
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-w', '-shared', '-fPIC', src, '-o', out], check=True)
    return C.CDLL(out)


def host_rasterize(lib, v, f, a, mvp, W, H, cull=True, flip_x=False, ch=4, bg=None):
    v = np.ascontiguousarray(v, np.float32)
    f = None if f is None else np.ascontiguousarray(f, np.int32)
    a = None if a is None else np.ascontiguousarray(a, np.float32)
    mvp = np.ascontiguousarray(mvp, np.float32)
    bgv = None if bg is None else np.ascontiguousarray(bg, np.float32)
    out = np.empty((H, W, ch), np.float32)
    P = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)   # noqa: E731
    nf = len(f) if f is not None else len(v) // 3
    rc = lib.rc_host_rasterize(P(v), C.c_long(len(v)), P(f), C.c_long(nf), P(a), P(mvp), W, H, P(bgv), int(cull), int(flip_x), ch, P(out))
    assert rc == 0
    return out


def test_view_matrices_match_oracle_and_mirror(g):
    from avatarcap_b200 import render
    fo_, bo_ = ro.cano_view_matrices(g['center'])
    fm, bm = render.cano_view_matrices(g['center'])
    assert np.array_equal(fo_, fm) and np.array_equal(bo_, bm)
    for kw in ({}, {'gl_space': True}):
        assert np.array_equal(ro.gl_perspective_projection_matrix(550., 540., 250., 260., 512, 480, **kw),
                              render.gl_perspective_projection_matrix(550., 540., 250., 260., 512, 480, **kw))
    assert np.array_equal(ro.gl_orthographic_projection_matrix(), render.gl_orthographic_projection_matrix())


def test_oracle_render_cano_mesh_vs_reference_golden(g):
    """coverage exact (the golden was rendered by the reference's render_cano_mesh around this rasteriser), values to fp16 storage"""
    f, b = ro.render_cano_mesh(g['v'], g['n'], g['f'], g['center'], int(g['img']))
    assert np.array_equal(np.packbits(np.linalg.norm(f, axis=-1) > 0), g['front_mask'])
    assert np.array_equal(np.packbits(np.linalg.norm(b, axis=-1) > 0), g['back_mask'])
    assert np.abs(f - g['front'].astype(np.float32)).max() < 1e-3 and np.abs(b - g['back'].astype(np.float32)).max() < 1e-3
    # closed surface seen from both sides: the mirrored back silhouette equals the front one up to boundary pixels
    fm = np.linalg.norm(f, axis=-1) > 0; bm = np.linalg.norm(b, axis=-1) > 0
    assert (fm != bm).mean() < 2e-3


def test_oracle_canonicalize_vs_reference_golden(g):
    img = int(g['img']); fx, fy, cx, cy = [float(x) for x in g['cam']]
    proj = ro.gl_perspective_projection_matrix(fx, fy, cx, cy, img, img, gl_space=False)
    pos = ro.rasterize(g['live_v'], g['f'], None, np.dot(proj, g['mv']), img, img)
    vn = ro.canonicalize_vertex_normals(g['live_v'], g['normal_map'].astype(np.float32), pos, g['vert_mats'], g['mv'], fx, fy, cx, cy)
    assert np.abs(vn - g['vn']).max() < 1e-6
    valid = np.linalg.norm(vn, axis=-1) > 0
    assert 0.3 < valid.mean() < 0.8                       # roughly the camera-facing half
    fi, bi = ro.render_cano_mesh(g['v'], vn, g['f'], g['center'], img)
    assert np.abs(fi - g['fi'].astype(np.float32)).max() < 1e-3 and np.abs(bi - g['bi'].astype(np.float32)).max() < 1e-3


@pytest.mark.parametrize('size', [512, 96])
def test_host_build_of_raster_core_equals_oracle(g, host_lib, size):
    fm, bm = ro.cano_view_matrices(g['center'])
    h = host_rasterize(host_lib, g['v'], g['f'], g['n'], fm, size, size)
    o = ro.rasterize(g['v'], g['f'], g['n'], fm, size, size)
    assert np.array_equal(h, o)
    h = host_rasterize(host_lib, g['v'], g['f'], g['n'], bm, size, size, flip_x=True, ch=3)
    assert np.array_equal(h, ro.rasterize(g['v'], g['f'], g['n'], bm, size, size)[:, ::-1, :3])


def test_host_build_soup_perspective_cull_and_edge_cases(host_lib):
    rs = np.random.RandomState(0)
    v = rs.uniform(-1, 1, (60, 3)).astype(np.float32); v[:, 2] += 3
    proj = ro.gl_perspective_projection_matrix(300, 300, 128, 128, 256, 200, gl_space=False)
    for cull in (True, False):
        h = host_rasterize(host_lib, v, None, None, proj, 256, 200, cull=cull, bg=(0.25, 0.5, 0.75))
        o = ro.rasterize(v, None, None, proj, 256, 200, bg=(0.25, 0.5, 0.75), cull=cull)
        assert np.array_equal(h, o) and (h[..., 3] > 0).mean() > 0.1
    # a vertex behind the eye drops its triangle; depth outside [0,1] drops the fragment; empty mesh = background
    v2 = np.array([[0, 0, 3], [1, 0, 3], [0, 1, -1], [0, 0, 500], [1, 0, 500], [0, 1, 500]], np.float32)
    h = host_rasterize(host_lib, v2, None, None, proj, 64, 64, cull=False)
    assert np.array_equal(h, ro.rasterize(v2, None, None, proj, 64, 64, cull=False)) and not h.any()
    h = host_rasterize(host_lib, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), None, proj, 8, 8, bg=(1, 0, 0))
    assert np.array_equal(h[..., 0], np.ones((8, 8), np.float32)) and not h[..., 3].any()
    # two triangles sharing an edge through pixel centres: every pixel of the quad is drawn exactly once (top-left rule)
    quad = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0.5, 0.5, 0], [-0.5, 0.5, 0]], np.float32)
    ortho = np.identity(4, np.float32)
    cnt = np.zeros((8, 8))
    for tri in ([0, 1, 2], [0, 2, 3]):
        cnt += host_rasterize(host_lib, quad, np.array([tri], np.int32), quad, ortho, 8, 8)[..., 3]
    assert cnt[2:6, 2:6].min() == 1 and cnt.max() == 1 and cnt.sum() == 16
    # equal depth: the triangle drawn first wins (GL_LESS)
    two = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]] * 2, np.float32)
    col = np.array([[1, 0, 0]] * 3 + [[0, 1, 0]] * 3, np.float32)
    h = host_rasterize(host_lib, two, None, col, ortho, 16, 16)
    assert h[..., 0].max() == 1 and h[..., 1].max() == 0


def test_nearest_border_rounding(host_lib):
    """grid_sample(nearest, border, align_corners=True): half-way cases round to even, outside clamps"""
    import torch
    import torch.nn.functional as F
    size = 9
    img = torch.arange(size, dtype=torch.float32).reshape(1, 1, 1, size)
    gs = np.concatenate([np.linspace(-1.3, 1.3, 41), (np.arange(size - 1) + 0.5) / (size - 1) * 2 - 1]).astype(np.float32)
    grid = torch.stack([torch.from_numpy(gs), torch.zeros(len(gs))], -1).reshape(1, 1, -1, 2)
    ref = F.grid_sample(img, grid, 'nearest', 'border', True).reshape(-1).numpy().astype(int)
    host_lib.rc_host_nearest_border.argtypes = [C.c_float, C.c_int]
    got = np.array([host_lib.rc_host_nearest_border(float(x), size) for x in gs])
    assert np.array_equal(got, ref)
    assert np.array_equal(ro.nearest_border_sample(np.arange(size, dtype=np.float32).reshape(1, size, 1), gs, np.zeros_like(gs))[:, 0].astype(int), ref)


def test_ply_writer_bytes_equal_reference(tmp_path):
    from avatarcap_b200 import mesh_io
    p = load_golden('ply_golden.npz')
    path = str(tmp_path / 'm.ply')
    for name, (n, c) in {'plain': (None, None), 'n': (p['n'], None), 'c': (None, p['c']), 'nc': (p['n'], p['c'])}.items():
        mesh_io.save_mesh_as_ply(path, p['v'], p['f'], n, c)
        assert np.array_equal(np.frombuffer(open(path, 'rb').read(), np.uint8), p['bytes_' + name]), name
    mesh_io.save_mesh_as_ply(path, p['v'], None, p['n'], (p['c'] * 255).astype(np.float32))
    assert np.array_equal(np.frombuffer(open(path, 'rb').read(), np.uint8), p['bytes_nofaces_c255'])


def test_fusion_mirror_vs_reference_golden(g):
    """merge_normal_images_cover against the reference's own output (the Adam-based merge_normal_images stays the reference's)"""
    from avatarcap_b200 import render
    img = int(g['img'])
    front, _ = ro.render_cano_mesh(g['v'], g['n'], g['f'], g['center'], img)
    fi, _ = ro.render_cano_mesh(g['v'], g['vn'], g['f'], g['center'], img)
    cover = render.merge_normal_images_cover(front.copy(), fi)
    assert np.abs(cover[::4, ::4] - g['cover_sub']).max() < 2e-5   # the golden's per-vertex normals came from torch (reference), these from numpy
    assert np.array_equal(cover, ro.merge_normal_images_cover(front, fi))


def math_pi():
    import math
    return math.pi
