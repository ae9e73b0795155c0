"""Host-side mirror of the reference's off-screen render and normal-fusion helpers -- the stage between the two field
evaluations of a frame. The rasterisation and the per-vertex canonicalisation run in the CUDA
library (csrc/raster.cu) on meshes that are already in HBM; the reference round-trips through the host and OpenGL.

    reference symbol                                                   mirror
    -----------------------------------------------------------------  -----------------------------------------------
    Renderer(img_w, img_h, mvp, shader_name, bg_color, window_name)    Renderer            utils/renderer.py:326-451
    gl_orthographic_projection_matrix / gl_perspective_projection_..   same names          utils/renderer.py:296-323
    render_cano_mesh(renderer, vertices, normals, faces, center)       render_cano_mesh    utils/visualize_util.py:11-52
    canonicalize_normal_map(pos_renderer, attri_renderer, ...)         canonicalize_normal_map   normal_fusion.py:12-66
    merge_normal_images(src, tar, iter_num, neck_xy)                   merge_normal_images       normal_fusion.py:91-155
    merge_normal_images_cover(src, tar)                                merge_normal_images_cover normal_fusion.py:158-167

`*_device` variants keep everything on the GPU (torch tensors in, torch tensors out) for pipeline.py.
The phong shaders (renderer.py:54-292) only draw the JPEG previews main.py writes next to the meshes: out of scope.
merge_normal_images is an Adam loop over a 64x64 rotation grid: it stays a PyTorch autograd program like the reference's
(no kernel of ours on that path), run on the engine's device.
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
def axis_angle_to_matrix(axis_angle: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.axis_angle_to_matrix (pytorch3d==0.6.0, requirements.txt:6; not vendored): axis-angle -> unit
    quaternion (sin(x/2)/x by its Taylor series below 1e-6 rad) -> rotation matrix."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half_angles = angles * 0.5
    small = angles.abs() < 1e-6
    sin_half_over_angle = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half_angles) / torch.where(small, torch.ones_like(angles), angles))
    q = torch.cat([torch.cos(half_angles), axis_angle * sin_half_over_angle], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def _neighbor_images(img: torch.Tensor, win_size: int = 3):
    """get_neighbor_images (normal_fusion.py:69-82): the 8 one-texel shifts by nearest sampling with zero padding."""
    H, W, _ = img.shape
    half = win_size // 2
    out = []
    for i in range(-half, half + 1):
        for j in range(-half, half + 1):
            if i == 0 and j == 0:
                continue
            theta = torch.tensor([[1, 0, j / (H / 2)], [0, 1, i / (W / 2)]], dtype=torch.float32, device=img.device)
            grid = F.affine_grid(theta.unsqueeze(0), torch.Size((1, 1, H, W)), align_corners=True)
            a = F.grid_sample(input=img.permute((2, 0, 1)).unsqueeze(0), grid=grid, mode='nearest', align_corners=True)
            out.append(a.squeeze(0).permute((1, 2, 0)))
    return out


def _resize_img(src: torch.Tensor, tar_shape) -> torch.Tensor:
    """resize_img (normal_fusion.py:85-90): bilinear, border, align_corners."""
    theta = torch.tensor([[1, 0, 0], [0, 1, 0]], dtype=torch.float32, device=src.device)
    grid = F.affine_grid(theta.unsqueeze(0), torch.Size((1, 1, tar_shape[0], tar_shape[1])), align_corners=True)
    return F.grid_sample(src.permute((2, 0, 1)).unsqueeze(0), grid, 'bilinear', 'border', True).squeeze(0).permute((1, 2, 0))


def merge_normal_images(src_img, tar_img, iter_num: int, neck_xy, device: Optional[torch.device] = None) -> np.ndarray:
    """normal_fusion.merge_normal_images (:91-155): rotation-grid registration of the avatar normals (src) to the image-observed
    normals (tar), then a distance-transform blend; the face rectangle keeps the avatar normals. Autograd + Adam as in the
    reference (this is an optimiser, not a kernel of the replaced path); cv2 erode / distanceTransform on the host, as there."""
    import cv2 as cv
    dev = torch.device(device) if device is not None else (torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu'))
    with torch.enable_grad():
        as_t = lambda x: (x.detach() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(dev, torch.float32)   # noqa: E731
        src_img = as_t(src_img)
        tar_img = as_t(tar_img)
        src_mask = torch.linalg.norm(src_img, dim=-1) > 0.
        tar_mask = torch.linalg.norm(tar_img, dim=-1) > 0.
        kernel = cv.getStructuringElement(cv.MORPH_RECT, (3, 3))
        tar_mask = cv.erode(tar_mask.cpu().numpy().astype(np.uint8), kernel, iterations=3)
        dt_tar_mask = torch.from_numpy(cv.distanceTransform(tar_mask, cv.DIST_L1, 3)).to(dev)
        tar_mask = torch.from_numpy(tar_mask > 0).to(dev)
        valid_mask = torch.logical_and(src_mask, tar_mask)
        src_img = src_img.clone().requires_grad_()
        init_src_img = src_img.detach().clone()
        rot_aa_img = torch.zeros((64, 64, 3), dtype=torch.float32, device=dev, requires_grad=True)
        optm_rot = torch.optim.Adam([rot_aa_img], lr=1e-2)
        optm_normal = torch.optim.Adam([src_img], lr=1e-1)
        smooth_lambda = 1.
        for iter_idx in range(iter_num):
            rot_mat_img = axis_angle_to_matrix(_resize_img(rot_aa_img, (512, 512)))
            data_loss = torch.square(torch.einsum('ijab,ijb->ija', rot_mat_img, src_img) - tar_img)[valid_mask].mean()
            smooth_loss = 0.
            for nb in _neighbor_images(rot_aa_img):
                smooth_loss = smooth_loss + torch.square(nb - rot_aa_img).mean()
            total_loss = data_loss + smooth_lambda * smooth_loss
            opt = optm_rot if iter_idx < iter_num / 2 else optm_normal
            opt.zero_grad()
            total_loss.backward()
            opt.step()
        with torch.no_grad():
            dt = dt_tar_mask[..., None] / 5.
            w0 = torch.ones_like(dt)
            w0[dt > 1.] = 0.
            out = (src_img * dt + init_src_img * w0) / (dt + w0)
            r = [neck_xy[1] - 90, neck_xy[0] - 35, neck_xy[1], neck_xy[0] + 35]
            out[r[0]: r[2], r[1]: r[3]] = init_src_img[r[0]: r[2], r[1]: r[3]]
        return out.detach().cpu().numpy()


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
