// KNN against the SMPL vertex set, LBS weights and skinning.
//
// Reference call sites replaced:
//   pytorch3d.ops.knn_points / knn_gather   utils/smpl_util.py:33,37 ; network/arch_avatar.py:190,197,208 ; dataset/avatarcap_dataset.py:114
//   SmplUtil.calculate_lbs                  utils/smpl_util.py:24-39
//   SmplUtil.skinning / skinning_normal     utils/smpl_util.py:58-81
//   GeoTexAvatar.forward posed->cano warp   network/arch_avatar.py:189-205 ; CanoBlendWeightVolume.forward :152-165
//
// The reference set (6 890 SMPL vertices) is staged through shared memory in tiles; every query thread keeps its
// K<=4 best candidates in registers (brute force, like pytorch3d's kernel), so the (B,N,K,24) gather tensor and the
// (B,N,4,4) per-point matrices of the reference are never materialised unless the caller asks for them.
#include "common.cuh"

namespace {

constexpr int KNN_NT = 256;
constexpr int KNN_TILE = 2048;   // reference points per shared-memory tile (24 KB)

struct Top4 { float d[4]; int i[4]; };

__device__ __forceinline__ void top_init(Top4& t) {
#pragma unroll
  for (int k = 0; k < 4; ++k) { t.d[k] = 3.4e38f; t.i[k] = 0; }
}
template <int K>
__device__ __forceinline__ void top_insert(Top4& t, float d, int idx) {
  if (d < t.d[K - 1]) {           // strict: the earlier index wins ties
    t.d[K - 1] = d; t.i[K - 1] = idx;
#pragma unroll
    for (int k = K - 1; k > 0; --k)
      if (t.d[k] < t.d[k - 1]) { const float td = t.d[k]; t.d[k] = t.d[k - 1]; t.d[k - 1] = td; const int ti = t.i[k]; t.i[k] = t.i[k - 1]; t.i[k - 1] = ti; }
  }
}

// brute-force scan of the whole reference set for one query per thread; all threads of the CTA take part in the staging
template <int K>
__device__ __forceinline__ void knn_scan(const float* __restrict__ ref, int m, float qx, float qy, float qz, bool active, Top4& best,
                                         float* s_ref) {
  top_init(best);
  for (int base = 0; base < m; base += KNN_TILE) {
    const int cnt = min(KNN_TILE, m - base);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * 3; t += blockDim.x) s_ref[t] = __ldg(ref + (size_t)base * 3 + t);
    __syncthreads();
    if (active) {
      for (int r = 0; r < cnt; ++r) {
        const float dx = __fsub_rn(qx, s_ref[3 * r]), dy = __fsub_rn(qy, s_ref[3 * r + 1]), dz = __fsub_rn(qz, s_ref[3 * r + 2]);
        // squared L2, summed x,y,z in order without FMA contraction (bit-matches the CPU oracle)
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        top_insert<K>(best, d, base + r);
      }
    }
  }
}

template <int K>
__global__ void __launch_bounds__(KNN_NT) knn_kernel(const float* __restrict__ q, int64_t n, const float* __restrict__ ref, int m,
                                                     float* __restrict__ out_d2, int64_t* __restrict__ out_idx) {
  __shared__ float s_ref[KNN_TILE * 3];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = q[g * 3]; qy = q[g * 3 + 1]; qz = q[g * 3 + 2]; }
  Top4 best; knn_scan<K>(ref, m, qx, qy, qz, active, best, s_ref);
  if (active) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (out_d2) out_d2[g * K + k] = best.d[k];
      if (out_idx) out_idx[g * K + k] = best.i[k];
    }
  }
}

// Gaussian KNN-4 blend of the SMPL skinning weights   smpl_util.py:33-38
__device__ __forceinline__ void lbs_from_knn(const Top4& best, const float* __restrict__ skin_w, float lbs[24]) {
  float w[4], s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { w[k] = expf(-best.d[k] / (2.f * 0.05f * 0.05f)); s += w[k]; }   // exp(-d2 / (2 r^2)), r = 0.05
  s += 1e-16f;
#pragma unroll
  for (int j = 0; j < 24; ++j) lbs[j] = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float wk = w[k] / s;
    const float4* row = reinterpret_cast<const float4*>(skin_w + (size_t)best.i[k] * 24);
#pragma unroll
    for (int q4 = 0; q4 < 6; ++q4) {
      const float4 t = __ldg(row + q4);
      lbs[4 * q4] += t.x * wk; lbs[4 * q4 + 1] += t.y * wk; lbs[4 * q4 + 2] += t.z * wk; lbs[4 * q4 + 3] += t.w * wk;
    }
  }
}

// M = sum_j lbs_j * J_j (row-major 4x4, rows 0..ROWS-1)   smpl_util.py:67
template <int ROWS>
__device__ __forceinline__ void blend_mats(const float lbs[24], const float* s_mats /*24*16 in smem*/, float M[ROWS * 4]) {
#pragma unroll
  for (int e = 0; e < ROWS * 4; ++e) M[e] = 0.f;
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    const float w = lbs[j];
#pragma unroll
    for (int e = 0; e < ROWS * 4; ++e) M[e] = fmaf(w, s_mats[j * 16 + e], M[e]);
  }
}

__global__ void __launch_bounds__(KNN_NT) lbs_weights_kernel(const float* __restrict__ pts, int64_t n, const float* __restrict__ cano_v, int m,
                                                             const float* __restrict__ skin_w, float* __restrict__ out_lbs) {
  __shared__ float s_ref[KNN_TILE * 3];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = pts[g * 3]; qy = pts[g * 3 + 1]; qz = pts[g * 3 + 2]; }
  Top4 best; knn_scan<4>(cano_v, m, qx, qy, qz, active, best, s_ref);
  if (!active) return;
  float lbs[24]; lbs_from_knn(best, skin_w, lbs);
#pragma unroll
  for (int j = 0; j < 24; ++j) out_lbs[g * 24 + j] = lbs[j];
}

__global__ void __launch_bounds__(KNN_NT) skin_kernel(const float* __restrict__ pts, const float* __restrict__ lbs_g, const float* __restrict__ jm,
                                                      int64_t n, float* __restrict__ out_pts, float* __restrict__ out_mats, int normal_mode) {
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = jm[t];
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  if (g >= n) return;
  float lbs[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) lbs[j] = lbs_g[g * 24 + j];
  float M[16]; blend_mats<4>(lbs, s_m, M);
  const float x = pts[g * 3], y = pts[g * 3 + 1], z = pts[g * 3 + 2];
  const float t = normal_mode ? 0.f : 1.f;       // skinning_normal: rotation block only (smpl_util.py:80)
  out_pts[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + t * M[3];
  out_pts[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + t * M[7];
  out_pts[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + t * M[11];
  if (out_mats)
#pragma unroll
    for (int e = 0; e < 16; ++e) out_mats[g * 16 + e] = M[e];
}

__global__ void __launch_bounds__(KNN_NT) skin_mesh_kernel(const float* __restrict__ verts, const float* __restrict__ normals, int64_t n,
                                                           const float* __restrict__ cano_v, int m, const float* __restrict__ skin_w,
                                                           const float* __restrict__ jm, float* __restrict__ out_v, float* __restrict__ out_n) {
  __shared__ float s_ref[KNN_TILE * 3];
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = jm[t];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float x = 0, y = 0, z = 0;
  if (active) { x = verts[g * 3]; y = verts[g * 3 + 1]; z = verts[g * 3 + 2]; }
  Top4 best; knn_scan<4>(cano_v, m, x, y, z, active, best, s_ref);
  if (!active) return;
  float lbs[24]; lbs_from_knn(best, skin_w, lbs);
  float M[12]; blend_mats<3>(lbs, s_m, M);
  out_v[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + M[3];
  out_v[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
  out_v[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
  if (normals && out_n) {
    const float nx = normals[g * 3], ny = normals[g * 3 + 1], nz = normals[g * 3 + 2];
    out_n[g * 3 + 0] = M[0] * nx + M[1] * ny + M[2] * nz;
    out_n[g * 3 + 1] = M[4] * nx + M[5] * ny + M[6] * nz;
    out_n[g * 3 + 2] = M[8] * nx + M[9] * ny + M[10] * nz;
  }
}

struct P2C {
  float bmin[3], inv_unused[3], len[3];
  int vd[3];
};

__global__ void __launch_bounds__(KNN_NT) posed_to_cano_kernel(const float* __restrict__ wpts, int64_t n, const float* __restrict__ live_v, int m,
                                                               const float* __restrict__ skin_w, const float* __restrict__ l2c, P2C p,
                                                               const float* __restrict__ wvol, float* __restrict__ out_cano,
                                                               uint8_t* __restrict__ out_near) {
  __shared__ float s_ref[KNN_TILE * 3];
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = l2c[t];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float x = 0, y = 0, z = 0;
  if (active) { x = wpts[g * 3]; y = wpts[g * 3 + 1]; z = wpts[g * 3 + 2]; }
  Top4 best; knn_scan<1>(live_v, m, x, y, z, active, best, s_ref);     // arch_avatar.py:190
  if (!active) return;
  if (out_near) out_near[g] = best.d[0] < 0.08f * 0.08f ? 1 : 0;         // :191
  float lbs[24];
#pragma unroll
  for (int j = 0; j < 24; ++j) lbs[j] = __ldg(skin_w + (size_t)best.i[0] * 24 + j);   // :197-198
  float M[12]; blend_mats<3>(lbs, s_m, M);
  float c[3];                                                             // :200
  c[0] = M[0] * x + M[1] * y + M[2] * z + M[3]; c[1] = M[4] * x + M[5] * y + M[6] * z + M[7]; c[2] = M[8] * x + M[9] * y + M[10] * z + M[11];
  // normalise to [0,1] by the canonical bounds (:201-203) then grid = 2p-1 and unnormalise with align_corners=True, border (:154-160)
  float idx[3]; int i0[3]; float f[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float p01 = (c[a] - p.bmin[a]) / p.len[a];
    const float gg = 2.f * p01 - 1.f;
    float s = ((gg + 1.f) / 2.f) * (float)(p.vd[a] - 1);
    s = fminf((float)(p.vd[a] - 1), fmaxf(s, 0.f));
    idx[a] = s; i0[a] = (int)floorf(s); f[a] = s - (float)i0[a];
  }
#pragma unroll
  for (int j = 0; j < 24; ++j) lbs[j] = 0.f;
#pragma unroll
  for (int dx = 0; dx < 2; ++dx)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dz = 0; dz < 2; ++dz) {
        const int xi = i0[0] + dx, yi = i0[1] + dy, zi = i0[2] + dz;
        if (xi > p.vd[0] - 1 || yi > p.vd[1] - 1 || zi > p.vd[2] - 1) continue;   // zero-weight taps
        const float w = (dx ? f[0] : 1.f - f[0]) * (dy ? f[1] : 1.f - f[1]) * (dz ? f[2] : 1.f - f[2]);
        const float4* row = reinterpret_cast<const float4*>(wvol + (((size_t)xi * p.vd[1] + yi) * p.vd[2] + zi) * 24);
#pragma unroll
        for (int q4 = 0; q4 < 6; ++q4) {
          const float4 t = __ldg(row + q4);
          lbs[4 * q4] += t.x * w; lbs[4 * q4 + 1] += t.y * w; lbs[4 * q4 + 2] += t.z * w; lbs[4 * q4 + 3] += t.w * w;
        }
      }
  blend_mats<3>(lbs, s_m, M);                                             // :205
  out_cano[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + M[3];
  out_cano[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
  out_cano[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
}

inline int nblocks(int64_t n) { return (int)((n + KNN_NT - 1) / KNN_NT); }

}  // namespace

extern "C" int avc_knn(avc_ctx* ctx, const float* query, int64_t n, const float* ref, int m, int K, float* out_d2, int64_t* out_idx, void* stream) {
  if (!ctx || !query || !ref) return avc_fail(ctx, AVC_EINVAL, "avc_knn: NULL argument");
  if (K < 1 || K > 4 || m < K) return avc_fail(ctx, AVC_EINVAL, "avc_knn: K must be 1..4 and m >= K");
  if (n == 0) return AVC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (K) {
    case 1: knn_kernel<1><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx); break;
    case 2: knn_kernel<2><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx); break;
    case 3: knn_kernel<3><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx); break;
    default: knn_kernel<4><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx); break;
  }
  AVC_LAUNCH_CHECK(ctx, "knn_kernel");
  return AVC_OK;
}

extern "C" int avc_lbs_weights(avc_ctx* ctx, const float* pts, int64_t n, const float* cano_verts, int m, const float* skin_weights,
                               float* out_lbs, void* stream) {
  if (!ctx || !pts || !cano_verts || !skin_weights || !out_lbs) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: NULL argument");
  if (m < 4) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: need at least 4 reference vertices");
  if (n == 0) return AVC_OK;
  lbs_weights_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(pts, n, cano_verts, m, skin_weights, out_lbs);
  AVC_LAUNCH_CHECK(ctx, "lbs_weights_kernel");
  return AVC_OK;
}

extern "C" int avc_skin_points(avc_ctx* ctx, const float* pts, const float* lbs, const float* jnt_mats, int64_t n, float* out_pts,
                               float* out_mats, void* stream) {
  if (!ctx || !pts || !lbs || !jnt_mats || !out_pts) return avc_fail(ctx, AVC_EINVAL, "avc_skin_points: NULL argument");
  if (n == 0) return AVC_OK;
  skin_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(pts, lbs, jnt_mats, n, out_pts, out_mats, 0);
  AVC_LAUNCH_CHECK(ctx, "skin_kernel");
  return AVC_OK;
}

extern "C" int avc_skin_normals(avc_ctx* ctx, const float* normals, const float* lbs, const float* jnt_mats, int64_t n, float* out_normals,
                                void* stream) {
  if (!ctx || !normals || !lbs || !jnt_mats || !out_normals) return avc_fail(ctx, AVC_EINVAL, "avc_skin_normals: NULL argument");
  if (n == 0) return AVC_OK;
  skin_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(normals, lbs, jnt_mats, n, out_normals, nullptr, 1);
  AVC_LAUNCH_CHECK(ctx, "skin_kernel(normals)");
  return AVC_OK;
}

extern "C" int avc_skin_mesh(avc_ctx* ctx, const float* verts, const float* normals, int64_t n, const float* cano_verts, int m,
                             const float* skin_weights, const float* jnt_mats, float* out_verts, float* out_normals, void* stream) {
  if (!ctx || !verts || !cano_verts || !skin_weights || !jnt_mats || !out_verts) return avc_fail(ctx, AVC_EINVAL, "avc_skin_mesh: NULL argument");
  if (m < 4) return avc_fail(ctx, AVC_EINVAL, "avc_skin_mesh: need at least 4 reference vertices");
  if (n == 0) return AVC_OK;
  skin_mesh_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(verts, normals, n, cano_verts, m, skin_weights, jnt_mats, out_verts, out_normals);
  AVC_LAUNCH_CHECK(ctx, "skin_mesh_kernel");
  return AVC_OK;
}

extern "C" int avc_posed_to_cano(avc_ctx* ctx, const float* wpts, int64_t n, const float* live_verts, int m, const float* skin_weights,
                                 const float* live2cano_mats, const float bounds[6], const float* weight_volume, const int vdims[3],
                                 float* out_cano, uint8_t* out_near, void* stream) {
  if (!ctx || !wpts || !live_verts || !skin_weights || !live2cano_mats || !bounds || !weight_volume || !vdims || !out_cano)
    return avc_fail(ctx, AVC_EINVAL, "avc_posed_to_cano: NULL argument");
  if (m < 1 || vdims[0] < 1 || vdims[1] < 1 || vdims[2] < 1) return avc_fail(ctx, AVC_EINVAL, "avc_posed_to_cano: bad sizes");
  if (n == 0) return AVC_OK;
  P2C p;
  for (int a = 0; a < 3; ++a) { p.bmin[a] = bounds[a]; p.len[a] = bounds[3 + a] - bounds[a]; p.vd[a] = vdims[a]; p.inv_unused[a] = 0.f; }
  posed_to_cano_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(wpts, n, live_verts, m, skin_weights, live2cano_mats, p, weight_volume,
                                                                       out_cano, out_near);
  AVC_LAUNCH_CHECK(ctx, "posed_to_cano_kernel");
  return AVC_OK;
}
