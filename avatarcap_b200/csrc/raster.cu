// Off-screen triangle rasteriser and per-vertex normal canonicalisation: the stage between the two field evaluations of a frame
// (SURVEY.md section 8f row 4). The reference copies the mesh to the host, expands it to a triangle soup, uploads it to OpenGL,
// draws into an FBO and reads the image back (utils/renderer.py:403-451, utils/visualize_util.py:11-52) -- twice per frame for the
// avatar normal maps and three more times inside canonicalize_normal_map (normal_fusion/normal_fusion.py:12-66). Here the mesh
// never leaves HBM: marching-cubes output -> these kernels -> (H,W,C) image that the HGFilter encoder reads.
//
//   Renderer.render, 'vertex_attribute' / 'position' shaders   utils/renderer.py:9-51, 428-451
//   canonicalize_normal_map (per-vertex part)                  normal_fusion/normal_fusion.py:27-62
//
// Bound: HBM / L2 (3 index loads + 3 x 16 B transformed vertices per triangle, one 64-bit atomicMin per covered pixel; the
// canonical meshes have ~3 M triangles of 1-4 pixels each). Arithmetic rules: raster_core.h.
#include "common.cuh"
#include "raster_core.h"

namespace {

constexpr int RS_NT = 256;
constexpr int RS_BIG = 64;            // triangles whose pixel box is larger go to the cooperative pass (one CTA per triangle)

__global__ void __launch_bounds__(RS_NT) raster_transform_kernel(const float* __restrict__ verts, int64_t n, RcMat M, int W, int H,
                                                                 RcVtx* __restrict__ tv) {
  const int64_t i = (int64_t)blockIdx.x * RS_NT + threadIdx.x;
  if (i >= n) return;
  tv[i] = rc_transform(M, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], W, H);
}

__device__ __forceinline__ RcTri load_tri(const RcVtx* __restrict__ tv, const int32_t* __restrict__ faces, int64_t t, int64_t n_verts, bool cull) {
  int i0, i1, i2;
  if (faces) { i0 = faces[3 * t]; i1 = faces[3 * t + 1]; i2 = faces[3 * t + 2]; } else { i0 = (int)(3 * t); i1 = i0 + 1; i2 = i0 + 2; }
  RcTri tri; tri.ok = false;
  if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= n_verts || i1 >= n_verts || i2 >= n_verts) return tri;   // out-of-range index: dropped, never read
  return rc_setup(tv[i0], tv[i1], tv[i2], i0, i1, i2, cull);
}

__global__ void __launch_bounds__(RS_NT) raster_tri_kernel(const RcVtx* __restrict__ tv, const int32_t* __restrict__ faces, int64_t n_faces,
                                                           int64_t n_verts, int cull, int W, int H, unsigned long long* __restrict__ zbuf,
                                                           int* __restrict__ big_list, int* __restrict__ big_count) {
  const int64_t t = (int64_t)blockIdx.x * RS_NT + threadIdx.x;
  if (t >= n_faces) return;
  const RcTri tri = load_tri(tv, faces, t, n_verts, cull != 0);
  if (!tri.ok) return;
  int px0, px1, py0, py1;
  if (!rc_bbox(tri, W, H, px0, px1, py0, py1)) return;
  if ((int64_t)(px1 - px0 + 1) * (py1 - py0 + 1) > RS_BIG) { big_list[atomicAdd(big_count, 1)] = (int)t; return; }
  for (int py = py0; py <= py1; ++py)
    for (int px = px0; px <= px1; ++px) {
      uint32_t z24;
      if (rc_cover(tri, px, py, z24)) atomicMin(&zbuf[(size_t)py * W + px], ((unsigned long long)z24 << 32) | (unsigned long long)(uint32_t)t);
    }
}

// large triangles (test scenes, close-ups): one CTA per triangle, threads stride over its pixel box
__global__ void __launch_bounds__(RS_NT) raster_big_kernel(const RcVtx* __restrict__ tv, const int32_t* __restrict__ faces, int64_t n_verts, int cull,
                                                           int W, int H, unsigned long long* __restrict__ zbuf, const int* __restrict__ big_list,
                                                           const int* __restrict__ big_count) {
  const int nb = *big_count;
  for (int b = blockIdx.x; b < nb; b += gridDim.x) {
    const int64_t t = big_list[b];
    const RcTri tri = load_tri(tv, faces, t, n_verts, cull != 0);
    int px0, px1, py0, py1;
    if (!tri.ok || !rc_bbox(tri, W, H, px0, px1, py0, py1)) continue;
    const int bw = px1 - px0 + 1;
    const int64_t np = (int64_t)bw * (py1 - py0 + 1);
    for (int64_t p = threadIdx.x; p < np; p += RS_NT) {
      const int px = px0 + (int)(p % bw), py = py0 + (int)(p / bw);
      uint32_t z24;
      if (rc_cover(tri, px, py, z24)) atomicMin(&zbuf[(size_t)py * W + px], ((unsigned long long)z24 << 32) | (unsigned long long)(uint32_t)t);
    }
  }
}

// one thread per pixel: winner's attributes, perspective-correct; image row 0 = top (renderer.py:448), optional left-right mirror
__global__ void __launch_bounds__(RS_NT) raster_resolve_kernel(const RcVtx* __restrict__ tv, const int32_t* __restrict__ faces, int64_t n_verts,
                                                               const float* __restrict__ attrs, int cull, int W, int H,
                                                               const unsigned long long* __restrict__ zbuf, float bg0, float bg1, float bg2,
                                                               int flip_x, int channels, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * RS_NT + threadIdx.x;
  if (p >= (int64_t)W * H) return;
  const int px = (int)(p % W), py = (int)(p / W);
  const unsigned long long key = zbuf[p];
  float r[4] = {bg0, bg1, bg2, 0.f};                       // glClearColor(bg, 0)  renderer.py:434
  if (key != RC_EMPTY) {
    const int64_t t = (int64_t)(key & 0xFFFFFFFFull);
    const RcTri tri = load_tri(tv, faces, t, n_verts, cull != 0);
    rc_shade(tri, tv[tri.i0].iw, tv[tri.i1].iw, tv[tri.i2].iw, attrs + 3 * (size_t)tri.i0, attrs + 3 * (size_t)tri.i1, attrs + 3 * (size_t)tri.i2, px,
             py, r);
    r[3] = 1.f;                                            // vec4(attributes, 1)  renderer.py:17
  }
  const int ox = flip_x ? W - 1 - px : px, oy = H - 1 - py;
  float* o = out + ((size_t)oy * W + ox) * channels;
  for (int c = 0; c < channels; ++c) o[c] = r[c];
}

// normal_fusion.py:27-62 for one vertex
struct CanonArgs {
  float mv[16];          // world -> camera (row-major)
  float imv[9];          // inv(mv)[:3,:3]
  float fx, fy, cx, cy;
  int H, W, pos_ch;
};
__global__ void __launch_bounds__(RS_NT) canonicalize_normals_kernel(const float* __restrict__ live_v, const float* __restrict__ vert_mats, int64_t n,
                                                                     CanonArgs a, const float* __restrict__ position_map,
                                                                     const float* __restrict__ normal_map, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * RS_NT + threadIdx.x;
  if (i >= n) return;
  const float vx = live_v[3 * i], vy = live_v[3 * i + 1], vz = live_v[3 * i + 2];
  float cam[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) cam[r] = a.mv[4 * r] * vx + a.mv[4 * r + 1] * vy + a.mv[4 * r + 2] * vz + a.mv[4 * r + 3];   // :29
  float gx = RC_FADD(RC_FMUL(RC_FDIV(cam[0], cam[2]), a.fx), a.cx), gy = RC_FADD(RC_FMUL(RC_FDIV(cam[1], cam[2]), a.fy), a.cy);   // :30-31
  gx = RC_FADD(RC_FMUL(2.f, RC_FDIV(gx, (float)a.W)), -1.f); gy = RC_FADD(RC_FMUL(2.f, RC_FDIV(gy, (float)a.H)), -1.f);           // :32-33
  const int ix = rc_nearest_border(gx, a.W), iy = rc_nearest_border(gy, a.H);                                                      // :36,49
  const float* pv = position_map + ((size_t)iy * a.W + ix) * a.pos_ch;
  const float dx = vx - pv[0], dy = vy - pv[1], dz = vz - pv[2];
  const bool vis = sqrtf(dx * dx + dy * dy + dz * dz) < 0.05f;                                                                     // :37
  const float* pn = normal_map + ((size_t)iy * a.W + ix) * 3;
  float nx = pn[0], ny = pn[1], nz = pn[2];
  const bool valid = vis && sqrtf(nx * nx + ny * ny + nz * nz) > 1e-6f;                                                            // :50
  ny = -ny; nz = -nz;                                                                                                              // :60
  const float wx = a.imv[0] * nx + a.imv[1] * ny + a.imv[2] * nz, wy = a.imv[3] * nx + a.imv[4] * ny + a.imv[5] * nz,
              wz = a.imv[6] * nx + a.imv[7] * ny + a.imv[8] * nz;                                                                  // :61
  // inv(vert_mats)[:3,:3] == inverse of the 3x3 block (affine matrices), by the adjugate                                          // :62
  const float* M = vert_mats + 16 * i;
  const float m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[4], m11 = M[5], m12 = M[6], m20 = M[8], m21 = M[9], m22 = M[10];
  const float c00 = m11 * m22 - m12 * m21, c01 = m02 * m21 - m01 * m22, c02 = m01 * m12 - m02 * m11;
  const float c10 = m12 * m20 - m10 * m22, c11 = m00 * m22 - m02 * m20, c12 = m02 * m10 - m00 * m12;
  const float c20 = m10 * m21 - m11 * m20, c21 = m01 * m20 - m00 * m21, c22 = m00 * m11 - m01 * m10;
  const float det = m00 * c00 + m01 * c10 + m02 * c20;
  const float id = 1.f / det;
  float ox = (c00 * wx + c01 * wy + c02 * wz) * id, oy = (c10 * wx + c11 * wy + c12 * wz) * id, oz = (c20 * wx + c21 * wy + c22 * wz) * id;
  if (!valid) { ox = 0.f; oy = 0.f; oz = 0.f; }                                                                                    // :63
  out[3 * i] = ox; out[3 * i + 1] = oy; out[3 * i + 2] = oz;
}

}  // namespace

extern "C" int avc_rasterize(avc_ctx* ctx, const float* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces, const float* attrs,
                             const float mvp[16], int width, int height, const float bg[3], int flags, int channels, float* out_image,
                             void* stream) {
  if (!ctx || !mvp || !out_image || (!verts && n_verts > 0) || (!verts && !faces && n_faces > 0)) return avc_fail(ctx, AVC_EINVAL, "avc_rasterize: NULL argument");
  if (width <= 0 || height <= 0 || width > 16384 || height > 16384) return avc_fail(ctx, AVC_EINVAL, "avc_rasterize: bad image size %dx%d", width, height);
  if (channels != 3 && channels != 4) return avc_fail(ctx, AVC_EINVAL, "avc_rasterize: channels must be 3 or 4");
  if (n_verts < 0 || n_faces < 0 || n_verts > 0x7fffffffLL || n_faces > 0x7fffffffLL) return avc_fail(ctx, AVC_EINVAL, "avc_rasterize: mesh too large for 32-bit indices");
  if (!faces && n_faces * 3 > n_verts) return avc_fail(ctx, AVC_EINVAL, "avc_rasterize: triangle soup needs 3 vertices per face");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)width * height;
  const size_t off_z = ((size_t)n_verts * sizeof(RcVtx) + 255) & ~(size_t)255;
  const size_t off_big = off_z + npix * sizeof(unsigned long long);
  const size_t need = off_big + ((size_t)n_faces + 4) * sizeof(int);
  int rc = avc_ensure_scratch(ctx, need);
  if (rc) return rc;
  char* base = (char*)ctx->d_scratch;
  RcVtx* tv = (RcVtx*)base;
  unsigned long long* zbuf = (unsigned long long*)(base + off_z);
  int* big_count = (int*)(base + off_big); int* big_list = big_count + 4;
  AVC_CUDA(ctx, cudaMemsetAsync(zbuf, 0xFF, npix * sizeof(unsigned long long), st));
  AVC_CUDA(ctx, cudaMemsetAsync(big_count, 0, 4 * sizeof(int), st));
  RcMat M; memcpy(M.m, mvp, sizeof(M.m));
  const int cull = (flags & AVC_RASTER_CULL_BACK) ? 1 : 0;
  if (n_verts > 0 && n_faces > 0) {
    raster_transform_kernel<<<(unsigned)((n_verts + RS_NT - 1) / RS_NT), RS_NT, 0, st>>>(verts, n_verts, M, width, height, tv);
    AVC_LAUNCH_CHECK(ctx, "raster_transform_kernel");
    raster_tri_kernel<<<(unsigned)((n_faces + RS_NT - 1) / RS_NT), RS_NT, 0, st>>>(tv, faces, n_faces, n_verts, cull, width, height, zbuf, big_list, big_count);
    AVC_LAUNCH_CHECK(ctx, "raster_tri_kernel");
    raster_big_kernel<<<ctx->sm_count * 4, RS_NT, 0, st>>>(tv, faces, n_verts, cull, width, height, zbuf, big_list, big_count);
    AVC_LAUNCH_CHECK(ctx, "raster_big_kernel");
  }
  const float b0 = bg ? bg[0] : 0.f, b1 = bg ? bg[1] : 0.f, b2 = bg ? bg[2] : 0.f;
  raster_resolve_kernel<<<(unsigned)((npix + RS_NT - 1) / RS_NT), RS_NT, 0, st>>>(tv, faces, n_verts, attrs ? attrs : verts, cull, width, height, zbuf, b0,
                                                                                 b1, b2, (flags & AVC_RASTER_FLIP_X) ? 1 : 0, channels, out_image);
  AVC_LAUNCH_CHECK(ctx, "raster_resolve_kernel");
  return AVC_OK;
}

extern "C" int avc_canonicalize_normals(avc_ctx* ctx, const float* live_verts, const float* vert_mats, int64_t n, const float mv[16], float fx,
                                        float fy, float cx, float cy, const float* position_map, int pos_channels, const float* normal_map,
                                        int height, int width, float* out_normals, void* stream) {
  if (!ctx || !live_verts || !vert_mats || !mv || !position_map || !normal_map || !out_normals)
    return avc_fail(ctx, AVC_EINVAL, "avc_canonicalize_normals: NULL argument");
  if (height <= 0 || width <= 0 || (pos_channels != 3 && pos_channels != 4)) return avc_fail(ctx, AVC_EINVAL, "avc_canonicalize_normals: bad image description");
  if (n == 0) return AVC_OK;
  CanonArgs a;
  memcpy(a.mv, mv, sizeof(a.mv));
  {  // inv(mv)[:3,:3]: mv is a rigid/affine world->camera matrix, so this is the inverse of its 3x3 block (double, adjugate)
    const double m00 = mv[0], m01 = mv[1], m02 = mv[2], m10 = mv[4], m11 = mv[5], m12 = mv[6], m20 = mv[8], m21 = mv[9], m22 = mv[10];
    const double c00 = m11 * m22 - m12 * m21, c01 = m02 * m21 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c10 = m12 * m20 - m10 * m22, c11 = m00 * m22 - m02 * m20, c12 = m02 * m10 - m00 * m12;
    const double c20 = m10 * m21 - m11 * m20, c21 = m01 * m20 - m00 * m21, c22 = m00 * m11 - m01 * m10;
    const double det = m00 * c00 + m01 * c10 + m02 * c20;
    if (det == 0.0) return avc_fail(ctx, AVC_EVALUE, "avc_canonicalize_normals: singular model-view matrix");
    const double c[9] = {c00, c01, c02, c10, c11, c12, c20, c21, c22};
    for (int k = 0; k < 9; ++k) a.imv[k] = (float)(c[k] / det);
  }
  a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.H = height; a.W = width; a.pos_ch = pos_channels;
  canonicalize_normals_kernel<<<(unsigned)((n + RS_NT - 1) / RS_NT), RS_NT, 0, (cudaStream_t)stream>>>(live_verts, vert_mats, n, a, position_map,
                                                                                                       normal_map, out_normals);
  AVC_LAUNCH_CHECK(ctx, "canonicalize_normals_kernel");
  return AVC_OK;
}
