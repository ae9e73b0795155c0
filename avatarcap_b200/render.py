"""Host-side mirror of the reference's off-screen render and normal-fusion helpers -- the stage between the two field
evaluations of a frame. The rasterisation and the per-vertex canonicalisation run in the CUDA
library (csrc/raster.cu) on meshes that are already in HBM; the reference round-trips through the host and OpenGL.

    reference symbol                                                   mirror
    -----------------------------------------------------------------  -----------------------------------------------
    Renderer(img_w, img_h, mvp, shader_name, bg_color, window_name)    Renderer            utils/renderer.py:326-451
    gl_orthographic_projection_matrix / gl_perspective_projection_..   same names          utils/renderer.py:296-323
    render_cano_mesh(renderer, vertices, normals, faces, center)       render_cano_mesh    utils/visualize_util.py:11-52
    canonicalize_normal_map(pos_renderer, attri_renderer, ...)         canonicalize_normal_map   normal_fusion.py:12-66
    merge_normal_images_cover(src, tar)                                merge_normal_images_cover normal_fusion.py:158-167

`*_device` variants keep everything on the GPU (torch tensors in, torch tensors out) for pipeline.py.
The phong shaders (renderer.py:54-292) only draw the JPEG previews main.py writes next to the meshes: out of scope.
normal_fusion.merge_normal_images (normal_fusion.py:91-155, the 100-step Adam registration on a 64x64 rotation grid) is an
optimiser, not a kernel of the replaced path, and SURVEY.md section 2 marks it out of scope: `patch.install()` leaves the
reference's own function in place and pipeline.fused_normal_maps takes it as a callable.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .engine import Engine, default_engine


# ------------------------------------------------------------------------------------------------------------ matrices
def gl_perspective_projection_matrix(fx, fy, cx, cy, img_w, img_h, far=100.0, near=0.1, gl_space=False) -> np.ndarray:
    """utils/renderer.py:296-313 (model in the usual camera space unless gl_space)."""
    proj_mat = np.zeros((4, 4), dtype=np.float32)
    proj_mat[0, 0] = 2 * fx / img_w
    proj_mat[0, 2] = (2 * cx - img_w) / img_w
    proj_mat[1, 1] = -2 * fy / img_h
    proj_mat[1, 2] = (img_h - 2 * cy) / img_h
    proj_mat[2, 2] = (far + near) / (far - near)
    proj_mat[2, 3] = 2 * near * far / (near - far)
    proj_mat[3, 2] = 1.
    if gl_space:
        real2gl = np.identity(4, dtype=np.float32)
        real2gl[1, 1] = -1
        real2gl[2, 2] = -1
        proj_mat = np.dot(proj_mat, real2gl)
    return proj_mat


def gl_orthographic_projection_matrix(far=-100.0, near=-0.1) -> np.ndarray:
    """utils/renderer.py:316-323 (model in the OpenGL camera space)."""
    proj_mat = np.zeros((4, 4), dtype=np.float32)
    proj_mat[0, 0] = 1.
    proj_mat[1, 1] = 1.
    proj_mat[2, 2] = 2 / (far - near)
    proj_mat[2, 3] = -(far + near) / (far - near)
    proj_mat[3, 3] = 1.
    return proj_mat


def _rot_y_pi() -> np.ndarray:
    """cv.Rodrigues([0, pi, 0]) stored into a float32 matrix (visualize_util.py:29-30)."""
    th = math.pi
    return np.array([[math.cos(th), 0., math.sin(th)], [0., 1., 0.], [-math.sin(th), 0., math.cos(th)]], np.float64).astype(np.float32)


def cano_view_matrices(mesh_center) -> Tuple[np.ndarray, np.ndarray]:
    """front / back MVP of render_cano_mesh (visualize_util.py:15-37): orthographic, camera 10 m in front of / behind the centre."""
    c = np.asarray(mesh_center, np.float32).reshape(3)
    model_RT = np.identity(4, dtype=np.float32)
    model_RT[:3, 3] = -c
    model_RT[2, 3] -= 10
    proj_mat = gl_orthographic_projection_matrix()
    front_mvp = np.dot(proj_mat, model_RT)
    trans_cen = np.identity(4, np.float32); trans_cen[:3, 3] = -c
    rot_y = np.identity(4, np.float32); rot_y[:3, :3] = _rot_y_pi()
    trans_z = np.identity(4, np.float32); trans_z[2, 3] = -10
    back_mvp = np.dot(proj_mat, np.dot(trans_z, np.dot(rot_y, trans_cen)))
    return front_mvp, back_mvp


# ------------------------------------------------------------------------------------------------------------ Renderer
class Renderer:
    """utils/renderer.py:326 without a window or a GL context: same constructor arguments and methods. set_model takes what the
    reference passes -- a triangle soup (3 consecutive vertices per triangle) with per-vertex attributes -- or, through
    set_indexed_model, the indexed mesh straight from marching cubes."""

    def __init__(self, img_w: int, img_h: int, mvp: Optional[np.ndarray] = None, shader_name: str = 'vertex_attribute', bg_color=(0, 0, 0),
                 window_name: str = '', engine: Optional[Engine] = None):
        if shader_name in ('phong_geometry', 'phong_color'):
            raise NotImplementedError('the phong preview shaders (utils/renderer.py:54-292) are outside the replaced path')
        if shader_name not in ('vertex_attribute', 'position'):
            raise ValueError('Invalid shader name!')                     # renderer.py:351
        self.img_w, self.img_h, self.shader_name, self.bg_color = int(img_w), int(img_h), shader_name, tuple(float(b) for b in bg_color)
        self.engine = engine if engine is not None else default_engine()
        self.mvp = np.identity(4, np.float32) if mvp is None or np.ndim(mvp) != 2 else np.asarray(mvp, np.float32)
        self.mv = np.identity(4, np.float32)
        self._v = self._a = self._f = None
        self.vnum = 0

    def set_mvp_mat(self, mvp) -> None:
        self.mvp = np.asarray(mvp, np.float32).reshape(4, 4)

    def set_mv_mat(self, mv) -> None:
        self.mv = np.asarray(mv, np.float32).reshape(4, 4)               # only the phong shaders read it

    def set_model(self, vertices, vertex_attributes=None, vertex_attributes_2=None) -> None:
        """attribute order as in the reference (1. normal, 2. colour): the vertex_attribute shader shows attribute 1 (:13)."""
        self._v = self.engine._f32(vertices, 3)
        self._a = None if vertex_attributes is None else self.engine._f32(vertex_attributes, 3)
        self._f = None
        self.vnum = int(self._v.shape[0])

    def set_indexed_model(self, vertices, faces, vertex_attributes=None) -> None:
        self._v = self.engine._f32(vertices, 3)
        self._a = None if vertex_attributes is None else self.engine._f32(vertex_attributes, 3)
        self._f = faces
        self.vnum = int(self._v.shape[0])

    def render_device(self, flip_x: bool = False, channels: int = 4) -> torch.Tensor:
        if self._v is None:
            raise RuntimeError('Renderer.render() before set_model()')
        attrs = None if self.shader_name == 'position' else self._a
        if self.shader_name == 'vertex_attribute' and attrs is None:
            raise ValueError('the vertex_attribute shader needs vertex attributes')
        return self.engine.rasterize(self._v, self._f, attrs, self.mvp, self.img_w, self.img_h, self.bg_color, cull=True, flip_x=flip_x,
                                     channels=channels)

    def render(self) -> np.ndarray:
        """(img_h, img_w, 4) float32, row 0 = top (renderer.py:444-451)."""
        return self.render_device().cpu().numpy()


# ------------------------------------------------------------------------------------------------------------ render_cano_mesh
def render_cano_mesh_device(engine: Engine, vertices, normals, faces, mesh_center, img: int = 512) -> Tuple[torch.Tensor, torch.Tensor]:
    """visualize_util.py:11-52 on an indexed device mesh: front / back (img,img,3) normal maps; the back view mirrored (:51)."""
    front_mvp, back_mvp = cano_view_matrices(mesh_center)
    front = engine.rasterize(vertices, faces, normals, front_mvp, img, img, channels=3)
    back = engine.rasterize(vertices, faces, normals, back_mvp, img, img, flip_x=True, channels=3)
    return front, back


def render_cano_mesh(renderer: Renderer, vertices, normals, faces, mesh_center=np.zeros(3), colors=None):
    """Same call as the reference's (numpy in, numpy out). `colors` is accepted and, as in the reference, not shown: the
    vertex_attribute shader outputs attribute 1, the normals (renderer.py:13-19)."""
    eng = renderer.engine
    f, b = render_cano_mesh_device(eng, np.asarray(vertices, np.float32), np.asarray(normals, np.float32), np.asarray(faces), mesh_center,
                                   renderer.img_w) if renderer.img_w == renderer.img_h else _render_cano_rect(renderer, vertices, normals, faces, mesh_center)
    return f.cpu().numpy(), b.cpu().numpy()


def _render_cano_rect(renderer: Renderer, vertices, normals, faces, mesh_center):
    front_mvp, back_mvp = cano_view_matrices(mesh_center)
    e = renderer.engine
    return (e.rasterize(vertices, faces, normals, front_mvp, renderer.img_w, renderer.img_h, channels=3),
            e.rasterize(vertices, faces, normals, back_mvp, renderer.img_w, renderer.img_h, flip_x=True, channels=3))


# ------------------------------------------------------------------------------------------------------------ canonicalize_normal_map
def canonicalize_normal_map_device(engine: Engine, cano_vertices, live_vertices, faces, normal_map, vert_mats, mv, fx, fy, cx, cy,
                                   cano_smpl_center, cano_img: int = 512):
    """normal_fusion.py:12-66 without leaving the device: position map of the live mesh through the pinhole camera, per-vertex
    visibility + nearest-sampled image normal rotated back to the canonical space, then the canonical front / back normal maps.
    -> (front (S,S,3), back (S,S,3), per-vertex canonical normals (V,3))."""
    nm = engine._f32(normal_map, 3)
    img_h, img_w = int(nm.shape[0]), int(nm.shape[1])
    proj_mat = gl_perspective_projection_matrix(fx, fy, cx, cy, img_w, img_h, gl_space=False)
    mvp = np.dot(proj_mat, np.asarray(mv, np.float32))
    position_map = engine.rasterize(live_vertices, faces, None, mvp, img_w, img_h, channels=4)
    proj_n = engine.canonicalize_normals(live_vertices, vert_mats, mv, fx, fy, cx, cy, position_map, nm)
    front, back = render_cano_mesh_device(engine, cano_vertices, proj_n, faces, cano_smpl_center, cano_img)
    return front, back, proj_n


def canonicalize_normal_map(pos_renderer: Renderer, attri_renderer: Renderer, cano_vertices, live_vertices, faces, normal_map, vert_mats, mv,
                            fx, fy, cx, cy, cano_smpl_center):
    """Same call as normal_fusion.canonicalize_normal_map (numpy / torch in, two numpy images out)."""
    eng = attri_renderer.engine
    vm = vert_mats if isinstance(vert_mats, torch.Tensor) else torch.as_tensor(np.asarray(vert_mats))
    f, b, _ = canonicalize_normal_map_device(eng, np.asarray(cano_vertices, np.float32), np.asarray(live_vertices, np.float32), np.asarray(faces),
                                             np.asarray(normal_map, np.float32), vm, mv, fx, fy, cx, cy, cano_smpl_center, attri_renderer.img_w)
    return f.cpu().numpy(), b.cpu().numpy()


# ------------------------------------------------------------------------------------------------------------ fusion
def merge_normal_images_cover(src_img, tar_img):
    """normal_fusion.py:158-167: cover the avatar normal with the image-observed one where that is valid (in place, like the
    reference; works on numpy arrays and on torch tensors)."""
    if isinstance(src_img, torch.Tensor):
        valid = torch.linalg.norm(tar_img, dim=-1) > 1e-6
        src_img[valid] = tar_img[valid]
        return src_img
    valid_mask = np.linalg.norm(tar_img, axis=-1) > 1e-6
    src_img[valid_mask] = tar_img[valid_mask]
    return src_img
