"""ORACLE (test infrastructure only) -- CPU restatement of the reference's off-screen GL passes that sit between the two
field evaluations of a frame:

  utils/renderer.py:326-451      Renderer ('vertex_attribute' / 'position' shaders, :9-51): depth-tested, back-face-culled
                                 triangle rasterisation into an RGBA32F frame buffer, read back flipped (row 0 = top, :448)
  utils/renderer.py:300-323      gl_perspective_projection_matrix / gl_orthographic_projection_matrix
  utils/visualize_util.py:11-52  render_cano_mesh (front / back orthographic normal maps of the canonical mesh)
  normal_fusion/normal_fusion.py:12-66, 158-167   canonicalize_normal_map, merge_normal_images_cover

Only tests/ (and tests/golden/gen_raster_golden.py) may import this.

PARITY UNPINNED at one boundary: the reference rasterises with OpenGL (glDrawArrays into an FBO); there is no GL context,
driver or golden image anywhere in the container, and the reference has no tests. OpenGL's rules are restated as published:
fragments are generated for pixel CENTRES inside the triangle, shared edges are drawn exactly once (top-left rule), window
depth is linear in window space and quantised to the 24-bit depth attachment (:392), GL_LESS keeps the first of two equal
depths, GL_CULL_FACE drops clockwise (back-facing) triangles, varyings are perspective-correct. A real GPU snaps window
coordinates to a sub-pixel grid first, so single boundary pixels can differ from GL. Everything around the rasteriser
(matrices, flips, channel handling, visibility test, canonicalisation) is pinned by running the reference's own functions
with this rasteriser plugged in as `Renderer` (tests/golden/raster_golden.npz).

Arithmetic contract (shared with avatarcap_b200/csrc/raster.cu so that coverage is bit-identical):
  * clip = mvp (row-major) * (x,y,z,1) in float32, summed left to right without FMA; ndc = clip.xyz / clip.w (float32);
    window x = (ndc.x + 1) * (W/2), y = (ndc.y + 1) * (H/2) (GL: y up), depth = (ndc.z + 1) * 0.5, all float32;
  * edge functions, barycentrics and interpolation in float64 on those float32 values, in the order written below.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def gl_orthographic_projection_matrix(far: float = -100.0, near: float = -0.1) -> np.ndarray:
    """utils/renderer.py:316-323 (the model is already in the GL camera space)."""
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 1.; m[1, 1] = 1.
    m[2, 2] = 2 / (far - near)
    m[2, 3] = -(far + near) / (far - near)
    m[3, 3] = 1.
    return m


def gl_perspective_projection_matrix(fx, fy, cx, cy, img_w, img_h, far: float = 100.0, near: float = 0.1, gl_space: bool = False) -> np.ndarray:
    """utils/renderer.py:293-313."""
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 2 * fx / img_w
    m[0, 2] = (2 * cx - img_w) / img_w
    m[1, 1] = -2 * fy / img_h
    m[1, 2] = (img_h - 2 * cy) / img_h
    m[2, 2] = (far + near) / (far - near)
    m[2, 3] = 2 * near * far / (near - far)
    m[3, 2] = 1.
    if gl_space:
        real2gl = np.identity(4, np.float32); real2gl[1, 1] = -1; real2gl[2, 2] = -1
        m = np.dot(m, real2gl)
    return m


def transform(verts: np.ndarray, mvp: np.ndarray, W: int, H: int) -> np.ndarray:
    """(V,3) float32 -> (V,4) float32 [window x, window y, depth, 1/w] (see the arithmetic contract)."""
    v = np.ascontiguousarray(verts, np.float32)
    m = np.ascontiguousarray(mvp, np.float32)
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    clip = [((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3] for r in range(4)]
    with np.errstate(divide='ignore', invalid='ignore'):
        out = np.empty((v.shape[0], 4), np.float32)
        out[:, 0] = (clip[0] / clip[3] + np.float32(1)) * np.float32(0.5 * W)
        out[:, 1] = (clip[1] / clip[3] + np.float32(1)) * np.float32(0.5 * H)
        out[:, 2] = (clip[2] / clip[3] + np.float32(1)) * np.float32(0.5)
        out[:, 3] = np.float32(1) / clip[3]
    out[:, 3][~(clip[3] > 0)] = np.float32(0)      # marks "behind the eye": triangles touching such a vertex are dropped
    return out


def _edge(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def _top_left(ax, ay, bx, by) -> bool:
    dx, dy = bx - ax, by - ay
    return dy < 0 or (dy == 0 and dx < 0)


def rasterize(verts: np.ndarray, faces: Optional[np.ndarray], attrs: Optional[np.ndarray], mvp: np.ndarray, W: int, H: int,
              bg=(0., 0., 0.), cull: bool = True) -> np.ndarray:
    """-> (H,W,4) float32, row 0 = top (utils/renderer.py:444-448). attrs None = the 'position' shader (:32-51): the
    attribute is the object-space vertex position. faces None = glDrawArrays over a triangle soup (:442)."""
    verts = np.ascontiguousarray(verts, np.float32)
    if faces is None:
        faces = np.arange(verts.shape[0] // 3 * 3, dtype=np.int64).reshape(-1, 3)
    A = verts if attrs is None else np.ascontiguousarray(attrs, np.float32)
    tv = transform(verts, mvp, W, H).astype(np.float64)
    zbuf = np.full((H, W), EMPTY, np.uint64)
    setups = {}
    for t, (i0, i1, i2) in enumerate(np.asarray(faces, np.int64)):
        s = _setup(tv, int(i0), int(i1), int(i2), cull)
        if s is None:
            continue
        (x0, y0, z0), (x1, y1, z1), (x2, y2, z2), area2 = s[0], s[1], s[2], s[3]
        px0 = max(int(math.ceil(min(x0, x1, x2) - 0.5)), 0); px1 = min(int(math.floor(max(x0, x1, x2) - 0.5)), W - 1)
        py0 = max(int(math.ceil(min(y0, y1, y2) - 0.5)), 0); py1 = min(int(math.floor(max(y0, y1, y2) - 0.5)), H - 1)
        if px1 < px0 or py1 < py0:
            continue
        cx = (np.arange(px0, px1 + 1, dtype=np.float64) + 0.5)[None, :]
        cy = (np.arange(py0, py1 + 1, dtype=np.float64) + 0.5)[:, None]
        w0 = _edge(x1, y1, x2, y2, cx, cy); w1 = _edge(x2, y2, x0, y0, cx, cy); w2 = _edge(x0, y0, x1, y1, cx, cy)
        ins = ((w0 > 0) | ((w0 == 0) & _top_left(x1, y1, x2, y2))) & ((w1 > 0) | ((w1 == 0) & _top_left(x2, y2, x0, y0))) & \
              ((w2 > 0) | ((w2 == 0) & _top_left(x0, y0, x1, y1)))
        z = ((w0 * z0 + w1 * z1) + w2 * z2) / area2
        ins &= (z >= 0) & (z <= 1)
        if not ins.any():
            continue
        z24 = np.floor(np.where(ins, z, 0.) * 16777215.0 + 0.5).astype(np.uint64)
        key = (z24 << np.uint64(32)) | np.uint64(t)
        sub = zbuf[py0:py1 + 1, px0:px1 + 1]
        upd = ins & (key < sub)
        sub[upd] = key[upd]
        setups[t] = s
    out = np.zeros((H, W, 4), np.float32)
    out[..., 0] = bg[0]; out[..., 1] = bg[1]; out[..., 2] = bg[2]
    ys, xs = np.nonzero(zbuf != EMPTY)
    for py, px in zip(ys, xs):
        t = int(zbuf[py, px] & np.uint64(0xFFFFFFFF))
        (x0, y0, _), (x1, y1, _), (x2, y2, _), area2, (j0, j1, j2) = setups[t]
        cx, cy = px + 0.5, py + 0.5
        w0 = _edge(x1, y1, x2, y2, cx, cy); w1 = _edge(x2, y2, x0, y0, cx, cy); w2 = _edge(x0, y0, x1, y1, cx, cy)
        q0 = (w0 / area2) * tv[j0, 3]; q1 = (w1 / area2) * tv[j1, 3]; q2 = (w2 / area2) * tv[j2, 3]
        den = (q0 + q1) + q2
        a0, a1, a2 = A[j0].astype(np.float64), A[j1].astype(np.float64), A[j2].astype(np.float64)
        out[H - 1 - py, px, :3] = (((q0 * a0 + q1 * a1) + q2 * a2) / den).astype(np.float32)
        out[H - 1 - py, px, 3] = 1.0
    return out


def _setup(tv, i0, i1, i2, cull):
    """Triangle set-up: None if dropped; else ((x,y,z) x3 in CCW order, 2*area, vertex ids in that order)."""
    if tv[i0, 3] <= 0 or tv[i1, 3] <= 0 or tv[i2, 3] <= 0:
        return None
    x0, y0 = tv[i0, 0], tv[i0, 1]; x1, y1 = tv[i1, 0], tv[i1, 1]; x2, y2 = tv[i2, 0], tv[i2, 1]
    if not (np.isfinite([x0, y0, x1, y1, x2, y2]).all()):
        return None
    area2 = _edge(x0, y0, x1, y1, x2, y2)
    if area2 == 0 or (cull and area2 < 0):
        return None
    if area2 < 0:                         # clockwise with culling off: swap to counter-clockwise
        i1, i2 = i2, i1
        x1, y1, x2, y2 = x2, y2, x1, y1
        area2 = -area2
    return (x0, y0, tv[i0, 2]), (x1, y1, tv[i1, 2]), (x2, y2, tv[i2, 2]), area2, (i0, i1, i2)


class OracleRenderer:
    """Duck-typed stand-in for utils/renderer.py `Renderer` (set_model / set_mvp_mat / set_mv_mat / render), so that the
    reference's own render_cano_mesh / canonicalize_normal_map can run on the CPU without a GL context."""

    def __init__(self, img_w: int, img_h: int, shader_name: str = 'vertex_attribute', bg_color=(0, 0, 0)):
        if shader_name not in ('vertex_attribute', 'position'):
            raise ValueError('Invalid shader name!')
        self.img_w, self.img_h, self.shader_name, self.bg_color = img_w, img_h, shader_name, bg_color
        self.mvp = np.identity(4, np.float32); self.v = None; self.a = None

    def set_mvp_mat(self, mvp): self.mvp = np.asarray(mvp, np.float32)
    def set_mv_mat(self, mv): pass

    def set_model(self, vertices, vertex_attributes=None, vertex_attributes_2=None):
        self.v = np.asarray(vertices, np.float32)
        self.a = None if (vertex_attributes is None or self.shader_name == 'position') else np.asarray(vertex_attributes, np.float32)

    def render(self):
        return rasterize(self.v, None, self.a, self.mvp, self.img_w, self.img_h, self.bg_color, cull=True)


# ----------------------------------------------------------------------------------------------------------------
def rodrigues_y_pi() -> np.ndarray:
    """cv.Rodrigues([0, pi, 0]) in float32 as visualize_util.py:30 stores it (float64 result assigned into a float32 matrix)."""
    th = math.pi
    R = np.array([[math.cos(th), 0., math.sin(th)], [0., 1., 0.], [-math.sin(th), 0., math.cos(th)]], np.float64)
    return R.astype(np.float32)


def cano_view_matrices(mesh_center) -> Tuple[np.ndarray, np.ndarray]:
    """front / back MVP of render_cano_mesh (visualize_util.py:15-37)."""
    c = np.asarray(mesh_center, np.float32)
    model = np.identity(4, np.float32); model[:3, 3] = -c; model[2, 3] -= 10
    proj = gl_orthographic_projection_matrix()
    front = np.dot(proj, model)
    trans_cen = np.identity(4, np.float32); trans_cen[:3, 3] = -c
    rot_y = np.identity(4, np.float32); rot_y[:3, :3] = rodrigues_y_pi()
    trans_z = np.identity(4, np.float32); trans_z[2, 3] = -10
    back = np.dot(proj, np.dot(trans_z, np.dot(rot_y, trans_cen)))
    return front, back


def render_cano_mesh(vertices, normals, faces, mesh_center=np.zeros(3), img=512, colors=None):
    """visualize_util.py:11-52 with the rasteriser above: -> front (H,W,3), back (H,W,3) mirrored left-right (:51).
    (With `colors` the reference still shows attribute 1 = normals, :43-44 + shader :13-19.)"""
    front_mvp, back_mvp = cano_view_matrices(mesh_center)
    f = rasterize(vertices, faces, normals, front_mvp, img, img)[..., :3]
    b = rasterize(vertices, faces, normals, back_mvp, img, img)[..., :3][:, ::-1]
    return np.ascontiguousarray(f), np.ascontiguousarray(b)


def nearest_border_sample(img_hwc: np.ndarray, gx: np.ndarray, gy: np.ndarray) -> np.ndarray:
    """F.grid_sample(mode='nearest', padding_mode='border', align_corners=True) at normalised (gx, gy): float32 like ATen
    (unnormalise ((g+1)/2)*(size-1), clip, round half to even)."""
    H, W = img_hwc.shape[:2]
    ix = ((gx.astype(np.float32) + np.float32(1)) / np.float32(2)) * np.float32(W - 1)
    iy = ((gy.astype(np.float32) + np.float32(1)) / np.float32(2)) * np.float32(H - 1)
    ix = np.clip(ix, np.float32(0), np.float32(W - 1)); iy = np.clip(iy, np.float32(0), np.float32(H - 1))
    return img_hwc[np.rint(iy).astype(np.int64), np.rint(ix).astype(np.int64)]


def canonicalize_vertex_normals(live_vertices, normal_map, position_map, vert_mats, mv, fx, fy, cx, cy):
    """normal_fusion.py:27-62: per-vertex image normal, checked for visibility against the rendered position map and
    rotated back to the canonical space. -> (V,3) float32 (zeros where not valid)."""
    v = np.asarray(live_vertices, np.float32); mv = np.asarray(mv, np.float32)
    H, W = normal_map.shape[:2]
    cam = (v @ mv[:3, :3].T + mv[:3, 3][None]).astype(np.float32)
    coord_x = cam[:, 0] / cam[:, 2] * np.float32(fx) + np.float32(cx)
    coord_y = cam[:, 1] / cam[:, 2] * np.float32(fy) + np.float32(cy)
    coord_x = np.float32(2.) * (coord_x / np.float32(W)) - np.float32(1.)
    coord_y = np.float32(2.) * (coord_y / np.float32(H)) - np.float32(1.)
    proj_v = nearest_border_sample(np.asarray(position_map, np.float32), coord_x, coord_y)[:, :3]
    vis = np.linalg.norm(v - proj_v, axis=-1) < 0.05
    proj_n = nearest_border_sample(np.asarray(normal_map, np.float32), coord_x, coord_y)[:, :3].copy()
    valid = vis & (np.linalg.norm(proj_n, axis=-1) > 1e-6)
    proj_n[:, 1:] *= -1
    proj_n = proj_n @ np.linalg.inv(mv.astype(np.float64))[:3, :3].T
    inv_vm = np.linalg.inv(np.asarray(vert_mats, np.float64))[:, :3, :3]
    proj_n = np.einsum('vij,vj->vi', inv_vm, proj_n)
    proj_n[~valid] = 0.
    return proj_n.astype(np.float32)


def merge_normal_images_cover(src_img, tar_img):
    """normal_fusion.py:158-167."""
    src = np.array(src_img, np.float32, copy=True)
    m = np.linalg.norm(tar_img, axis=-1) > 1e-6
    src[m] = np.asarray(tar_img, np.float32)[m]
    return src
