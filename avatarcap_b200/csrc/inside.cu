// Inside/outside classification of the dense canonical grid against a closed triangle mesh (SURVEY.md section 8f,
// "next" row 3: the +-1 fill of the voxels the networks do not evaluate).
//
// Reference call site replaced: dataset/avatarcap_dataset.py:120-124
//     cano_smpl_trimesh = trimesh.Trimesh(verts, faces, use_embree=True); invalid_pts_ov = 2*contains(invalid_pts) - 1
// (trimesh + embree ray casting on the CPU, once per dataset). Here: every triangle toggles one bit per grid column (x_i, y_j)
// whose vertical line it crosses, at the first grid index above the crossing; a suffix-XOR along z then gives the crossing
// parity above every grid point == inside. O(triangles x covered columns + R^3), no acceleration structure needed because
// the query set is the regular grid itself. The 2D containment test uses the top-left fill rule in double precision, so a
// column through a shared edge or vertex is counted exactly once.
#include "common.cuh"

namespace {

// torch.linspace(0,1,steps) float32 (same helper as make_grid_kernel / nerf.cu)
__device__ __forceinline__ float lin01i(int q, int steps) {
  if (steps <= 1) return 0.f;
  const float step = __fdiv_rn(1.f, (float)(steps - 1));
  return q < steps / 2 ? __fmul_rn(step, (float)q) : __fsub_rn(1.f, __fmul_rn(step, (float)(steps - 1 - q)));
}
__device__ __forceinline__ float grid_coord(int q, int steps, float b, float len) { return __fadd_rn(__fmul_rn(lin01i(q, steps), len), b); }

struct InsideArgs { float bx, by, bz, lx, ly, lz; int rx, ry, rz; };

// edge function with the top-left rule: > 0 strictly inside; == 0 counts only for "top" or "left" edges
__device__ __forceinline__ bool edge_in(double ax, double ay, double bx, double by, double px, double py, double sgn) {
  const double e = sgn * ((bx - ax) * (py - ay) - (by - ay) * (px - ax));
  if (e != 0.0) return e > 0.0;
  const double dx = sgn * (bx - ax), dy = sgn * (by - ay);
  return (dy == 0.0 && dx < 0.0) || (dy < 0.0);         // top edge, or left edge (counter-clockwise orientation after sgn)
}

__global__ void tri_toggle_kernel(const float* __restrict__ verts, const int32_t* __restrict__ faces, int nf, InsideArgs a,
                                  unsigned int* __restrict__ bits) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  double p[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { const int v = faces[3 * f + c]; p[c][0] = verts[3 * v]; p[c][1] = verts[3 * v + 1]; p[c][2] = verts[3 * v + 2]; }
  const double area = (p[1][0] - p[0][0]) * (p[2][1] - p[0][1]) - (p[1][1] - p[0][1]) * (p[2][0] - p[0][0]);
  if (area == 0.0) return;                                // triangle parallel to the ray: no crossing
  const double sgn = area > 0.0 ? 1.0 : -1.0;
  const double xmin = fmin(p[0][0], fmin(p[1][0], p[2][0])), xmax = fmax(p[0][0], fmax(p[1][0], p[2][0]));
  const double ymin = fmin(p[0][1], fmin(p[1][1], p[2][1])), ymax = fmax(p[0][1], fmax(p[1][1], p[2][1]));
  // conservative index range of the columns inside the bounding box (grid spacing len/(R-1))
  const int i0 = max(0, (int)floor((xmin - a.bx) / a.lx * (a.rx - 1)) - 1), i1 = min(a.rx - 1, (int)ceil((xmax - a.bx) / a.lx * (a.rx - 1)) + 1);
  const int j0 = max(0, (int)floor((ymin - a.by) / a.ly * (a.ry - 1)) - 1), j1 = min(a.ry - 1, (int)ceil((ymax - a.by) / a.ly * (a.ry - 1)) + 1);
  const int stride = a.rz + 1;
  for (int i = i0; i <= i1; ++i) {
    const double px = grid_coord(i, a.rx, a.bx, a.lx);
    if (px < xmin || px > xmax) continue;
    for (int j = j0; j <= j1; ++j) {
      const double py = grid_coord(j, a.ry, a.by, a.ly);
      if (py < ymin || py > ymax) continue;
      if (!edge_in(p[0][0], p[0][1], p[1][0], p[1][1], px, py, sgn) || !edge_in(p[1][0], p[1][1], p[2][0], p[2][1], px, py, sgn) ||
          !edge_in(p[2][0], p[2][1], p[0][0], p[0][1], px, py, sgn)) continue;
      // z of the crossing (barycentric interpolation)
      const double w0 = ((p[1][0] - px) * (p[2][1] - py) - (p[1][1] - py) * (p[2][0] - px)) / area;
      const double w1 = ((p[2][0] - px) * (p[0][1] - py) - (p[2][1] - py) * (p[0][0] - px)) / area;
      const double zc = w0 * p[0][2] + w1 * p[1][2] + (1.0 - w0 - w1) * p[2][2];
      // k0 = number of grid points of this column strictly below the crossing
      int k0 = (int)floor((zc - a.bz) / a.lz * (a.rz - 1)) + 2;
      k0 = min(max(k0, 0), a.rz);
      while (k0 > 0 && (double)grid_coord(k0 - 1, a.rz, a.bz, a.lz) >= zc) --k0;
      while (k0 < a.rz && (double)grid_coord(k0, a.rz, a.bz, a.lz) < zc) ++k0;
      const long long bit = ((long long)i * a.ry + j) * stride + k0;
      atomicXor(&bits[bit >> 5], 1u << (bit & 31));
    }
  }
}

// inside[k] = parity of the crossings above grid point k = XOR of the toggles at indices k0 > k
__global__ void column_parity_kernel(const unsigned int* __restrict__ bits, InsideArgs a, uint8_t* __restrict__ out) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= (long long)a.rx * a.ry) return;
  const int stride = a.rz + 1;
  unsigned par = 0;
  for (int k = a.rz - 1; k >= 0; --k) {
    const long long bit = col * stride + (k + 1);
    par ^= (bits[bit >> 5] >> (bit & 31)) & 1u;
    out[col * a.rz + k] = (uint8_t)par;
  }
}

}  // namespace

extern "C" int avc_inside_volume(avc_ctx* ctx, const float* verts, int nv, const int32_t* faces, int nf, const float bounds[6], const int res[3],
                                 uint8_t* out_inside, void* stream) {
  if (!ctx || !verts || !faces || !bounds || !res || !out_inside) return avc_fail(ctx, AVC_EINVAL, "avc_inside_volume: NULL argument");
  if (nv < 3 || nf < 1 || res[0] < 2 || res[1] < 2 || res[2] < 2) return avc_fail(ctx, AVC_EINVAL, "avc_inside_volume: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  InsideArgs a;
  a.bx = bounds[0]; a.by = bounds[1]; a.bz = bounds[2]; a.lx = bounds[3] - bounds[0]; a.ly = bounds[4] - bounds[1]; a.lz = bounds[5] - bounds[2];
  a.rx = res[0]; a.ry = res[1]; a.rz = res[2];
  const long long nbits = (long long)res[0] * res[1] * (res[2] + 1);
  const size_t words = (size_t)((nbits + 31) / 32);
  int rc = avc_ensure_scratch(ctx, words * 4 + 64);
  if (rc) return rc;
  unsigned int* bits = (unsigned int*)ctx->d_scratch;
  AVC_CUDA(ctx, cudaMemsetAsync(bits, 0, words * 4, st));
  tri_toggle_kernel<<<(nf + 127) / 128, 128, 0, st>>>(verts, faces, nf, a, bits);
  AVC_LAUNCH_CHECK(ctx, "tri_toggle_kernel");
  const long long cols = (long long)res[0] * res[1];
  column_parity_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, st>>>(bits, a, out_inside);
  AVC_LAUNCH_CHECK(ctx, "column_parity_kernel");
  return AVC_OK;
}
