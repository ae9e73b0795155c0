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
  for (int k = 0; k < 4; ++k) { t.d[k] = 3.4e38f; t.i[k] = 0x7fffffff; }
}
template <int K>
__device__ __forceinline__ void top_insert(Top4& t, float d, int idx) {
  // ordered by (distance, index): the lower index wins ties whatever the visiting order (brute force and grid agree bit for bit)
  if (d < t.d[K - 1] || (d == t.d[K - 1] && idx < t.i[K - 1])) {
    t.d[K - 1] = d; t.i[K - 1] = idx;
#pragma unroll
    for (int k = K - 1; k > 0; --k)
      if (t.d[k] < t.d[k - 1] || (t.d[k] == t.d[k - 1] && t.i[k] < t.i[k - 1])) {
        const float td = t.d[k]; t.d[k] = t.d[k - 1]; t.d[k - 1] = td; const int ti = t.i[k]; t.i[k] = t.i[k - 1]; t.i[k - 1] = ti;
      }
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

// ---------------------------------------------------------------------------------------------------------------
// Uniform grid over the reference vertices: exact KNN in a few cells instead of a 6 890-vertex scan per query.
// After the shells of Chebyshev radius 0..r around the query's (clamped) cell have been visited, every unvisited vertex is
// at least r*h away (projection onto the grid box is non-expansive), so the search stops as soon as the K-th best squared
// distance is <= (r*h)^2; queries that are far from every vertex fall back to a scan of the whole set.
constexpr int GRID_MAX_CELLS = 1 << 18;
constexpr int GRID_RMAX = 8;    // default number of shells before the brute-force fallback (AVC_KNN_RMAX overrides, 1..12)
struct GridDesc { float ox, oy, oz, h, inv_h; int dx, dy, dz, m, cells, rmax; };

__device__ __forceinline__ int grid_cell(const GridDesc& G, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = min(max((int)floorf((x - G.ox) * G.inv_h), 0), G.dx - 1);
  cy = min(max((int)floorf((y - G.oy) * G.inv_h), 0), G.dy - 1);
  cz = min(max((int)floorf((z - G.oz) * G.inv_h), 0), G.dz - 1);
  return (cx * G.dy + cy) * G.dz + cz;
}

__global__ void __launch_bounds__(1024) grid_bounds_kernel(const float* __restrict__ ref, int m, float h_min, int rmax, GridDesc* __restrict__ G, int* __restrict__ cnt) {
  __shared__ float s_mn[3][32], s_mx[3][32];
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = threadIdx.x; i < m; i += blockDim.x)
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float v = ref[3 * i + c]; mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v); }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    for (int o = 16; o > 0; o >>= 1) { mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o)); mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o)); }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) { s_mn[c][threadIdx.x >> 5] = mn[c]; s_mx[c][threadIdx.x >> 5] = mx[c]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) for (int w = 0; w < 32; ++w) { mn[c] = fminf(mn[c], s_mn[c][w]); mx[c] = fmaxf(mx[c], s_mx[c][w]); }
    const float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    GridDesc g;
    g.h = fmaxf(h_min, ext / 60.f); g.inv_h = 1.f / g.h;
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
    g.dx = (int)floorf((mx[0] - mn[0]) * g.inv_h) + 1; g.dy = (int)floorf((mx[1] - mn[1]) * g.inv_h) + 1; g.dz = (int)floorf((mx[2] - mn[2]) * g.inv_h) + 1;
    g.rmax = rmax; g.m = m; g.cells = g.dx * g.dy * g.dz;          // <= 61^3 < GRID_MAX_CELLS
    *G = g;
  }
  for (int i = threadIdx.x; i < GRID_MAX_CELLS; i += blockDim.x) cnt[i] = 0;
}
__global__ void grid_count_kernel(const float* __restrict__ ref, int m, const GridDesc* __restrict__ Gp, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const GridDesc G = *Gp; int cx, cy, cz;
  atomicAdd(&cnt[grid_cell(G, ref[3 * i], ref[3 * i + 1], ref[3 * i + 2], cx, cy, cz)], 1);
}
__global__ void __launch_bounds__(1024) grid_scan_kernel(const GridDesc* __restrict__ Gp, int* __restrict__ cnt, int* __restrict__ start) {
  __shared__ int wsum[32]; __shared__ int carry;
  const int cells = Gp->cells;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < cells; base += 1024) {
    const int c = base + threadIdx.x;
    const int x = c < cells ? cnt[c] : 0;
    int inc = x;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const int sw = wsum[w]; if (w < wid) pre += sw; tot += sw; }
    if (c < cells) { start[c] = carry + pre + inc - x; cnt[c] = 0; }      // cnt becomes the fill cursor
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[cells] = carry;
}
__global__ void grid_fill_kernel(const float* __restrict__ ref, int m, const GridDesc* __restrict__ Gp, const int* __restrict__ start,
                                 int* __restrict__ cursor, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const GridDesc G = *Gp; int cx, cy, cz;
  const float x = ref[3 * i], y = ref[3 * i + 1], z = ref[3 * i + 2];
  const int c = grid_cell(G, x, y, z, cx, cy, cz);
  sorted[start[c] + atomicAdd(&cursor[c], 1)] = make_float4(x, y, z, __int_as_float(i));
}

struct GridView { const GridDesc* G; const int* start; const float4* sorted; };

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, const float4& r) {
  const float dx = __fsub_rn(qx, r.x), dy = __fsub_rn(qy, r.y), dz = __fsub_rn(qz, r.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// all vertices of the cells (i, j, k0..k1): cells that differ only in k are adjacent in the sorted array, so a whole z-run is ONE
// [start, end) range -- two index loads per run instead of two per cell
template <int K>
__device__ __forceinline__ void knn_visit_run(const GridView& V, const GridDesc& G, int i, int j, int k0, int k1, float qx, float qy, float qz, Top4& best) {
  const int c = (i * G.dy + j) * G.dz;
  const int e = __ldg(V.start + c + k1 + 1);
  for (int p = __ldg(V.start + c + k0); p < e; ++p) { const float4 v = __ldg(V.sorted + p); top_insert<K>(best, dist2_rn(qx, qy, qz, v), __float_as_int(v.w)); }
}

template <int K>
__device__ __forceinline__ bool knn_grid(const GridView& V, float qx, float qy, float qz, Top4& best) {
  const GridDesc G = *V.G;
  top_init(best);
  int cx, cy, cz; grid_cell(G, qx, qy, qz, cx, cy, cz);
  // shells 0 and 1 together (shell 0 alone can never satisfy the stop test): the 3x3x3 cube as 9 z-runs
  for (int i = max(cx - 1, 0); i <= min(cx + 1, G.dx - 1); ++i)
    for (int j = max(cy - 1, 0); j <= min(cy + 1, G.dy - 1); ++j)
      knn_visit_run<K>(V, G, i, j, max(cz - 1, 0), min(cz + 1, G.dz - 1), qx, qy, qz, best);
  float rh = G.h * 0.999f;                         // 0.1 % slack for the float rounding of the cell assignment
  bool done = best.d[K - 1] <= rh * rh;
  for (int r = 2; r <= G.rmax && !done; ++r) {
    for (int i = max(cx - r, 0); i <= min(cx + r, G.dx - 1); ++i)
      for (int j = max(cy - r, 0); j <= min(cy + r, G.dy - 1); ++j) {
        if ((abs(i - cx) == r) || (abs(j - cy) == r)) {            // a row of the shell's side faces: the whole z-run
          knn_visit_run<K>(V, G, i, j, max(cz - r, 0), min(cz + r, G.dz - 1), qx, qy, qz, best);
        } else {                                                   // interior row: only the two end cells belong to the shell
          if (cz - r >= 0) knn_visit_run<K>(V, G, i, j, cz - r, cz - r, qx, qy, qz, best);
          if (cz + r <= G.dz - 1) knn_visit_run<K>(V, G, i, j, cz + r, cz + r, qx, qy, qz, best);
        }
      }
    rh = (float)r * G.h * 0.999f;
    done = best.d[K - 1] <= rh * rh;
  }
  return done;          // false: far from every vertex -> the caller falls back to the block-cooperative brute-force scan
}

// min squared distance < r2 ?  (dataset/avatarcap_dataset.py:114-116 valid flag) -- bounded search, exact
__global__ void __launch_bounds__(KNN_NT) near_flag_kernel(const float* __restrict__ q, int64_t n, GridView V, float r2, uint8_t* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  if (g >= n) return;
  const GridDesc G = *V.G;
  const float qx = q[g * 3], qy = q[g * 3 + 1], qz = q[g * 3 + 2];
  // vertices within sqrt(r2) of q lie in cells within ceil(sqrt(r2)/h) of the cell of q's projection onto the grid box;
  // a query farther than that from the box itself cannot have any
  const float px = fminf(fmaxf(qx, G.ox), G.ox + G.dx * G.h), py = fminf(fmaxf(qy, G.oy), G.oy + G.dy * G.h), pz = fminf(fmaxf(qz, G.oz), G.oz + G.dz * G.h);
  const float ob = (qx - px) * (qx - px) + (qy - py) * (qy - py) + (qz - pz) * (qz - pz);
  bool hit = false;
  if (ob < r2 * 1.0001f) {
    const int R = (int)floorf(sqrtf(r2) * G.inv_h) + 1;   // >= ceil, with a full cell of slack when radius is a multiple of h
    int cx, cy, cz; grid_cell(G, qx, qy, qz, cx, cy, cz);
    const int k0 = max(cz - R, 0), k1 = min(cz + R, G.dz - 1);          // a z-run of cells is one contiguous range of the sorted array
    for (int i = max(cx - R, 0); i <= min(cx + R, G.dx - 1) && !hit; ++i)
      for (int j = max(cy - R, 0); j <= min(cy + R, G.dy - 1) && !hit; ++j) {
        const int c = (i * G.dy + j) * G.dz;
        const int e = __ldg(V.start + c + k1 + 1);
        for (int p = __ldg(V.start + c + k0); p < e; ++p) if (dist2_rn(qx, qy, qz, __ldg(V.sorted + p)) < r2) { hit = true; break; }
      }
  }
  out[g] = hit ? 1 : 0;
}

template <int K>
__global__ void __launch_bounds__(KNN_NT) knn_kernel(const float* __restrict__ q, int64_t n, const float* __restrict__ ref, int m,
                                                     float* __restrict__ out_d2, int64_t* __restrict__ out_idx, GridView V) {
  __shared__ float s_ref[KNN_TILE * 3];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = q[g * 3]; qy = q[g * 3 + 1]; qz = q[g * 3 + 2]; }
  Top4 best;
  bool need = active;
  if (V.G) { top_init(best); need = active && !knn_grid<K>(V, qx, qy, qz, best); }
  if (__syncthreads_or(need)) { Top4 b2; knn_scan<K>(ref, m, qx, qy, qz, need, b2, s_ref); if (need) best = b2; }
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
                                                             const float* __restrict__ skin_w, float* __restrict__ out_lbs, GridView V) {
  __shared__ float s_ref[KNN_TILE * 3];
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // block size is a launch-time knob (knn_block)
  const bool active = g < n;
  float qx = 0, qy = 0, qz = 0;
  if (active) { qx = pts[g * 3]; qy = pts[g * 3 + 1]; qz = pts[g * 3 + 2]; }
  Top4 best;
  bool need = active;
  if (V.G) { top_init(best); need = active && !knn_grid<4>(V, qx, qy, qz, best); }
  if (__syncthreads_or(need)) { Top4 b2; knn_scan<4>(cano_v, m, qx, qy, qz, need, b2, s_ref); if (need) best = b2; }
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
                                                           const float* __restrict__ jm, float* __restrict__ out_v, float* __restrict__ out_n, GridView V) {
  __shared__ float s_ref[KNN_TILE * 3];
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = jm[t];
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // block size is a launch-time knob (knn_block)
  const bool active = g < n;
  float x = 0, y = 0, z = 0;
  if (active) { x = verts[g * 3]; y = verts[g * 3 + 1]; z = verts[g * 3 + 2]; }
  Top4 best;
  bool need = active;
  if (V.G) { top_init(best); need = active && !knn_grid<4>(V, x, y, z, best); }
  if (__syncthreads_or(need)) { Top4 b2; knn_scan<4>(cano_v, m, x, y, z, need, b2, s_ref); if (need) best = b2; }
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
                                                               uint8_t* __restrict__ out_near, GridView V) {
  __shared__ float s_ref[KNN_TILE * 3];
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = l2c[t];
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  const bool active = g < n;
  float x = 0, y = 0, z = 0;
  if (active) { x = wpts[g * 3]; y = wpts[g * 3 + 1]; z = wpts[g * 3 + 2]; }
  Top4 best;                                                           // arch_avatar.py:190
  bool need = active;
  if (V.G) { top_init(best); need = active && !knn_grid<1>(V, x, y, z, best); }
  if (__syncthreads_or(need)) { Top4 b2; knn_scan<1>(live_v, m, x, y, z, need, b2, s_ref); if (need) best = b2; }
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
// KNN-4 kernels end their grid search at a block-wide vote (cooperative fallback): smaller blocks wait less for their slowest thread
inline int knn_block() { const char* e = getenv("AVC_KNN_BLOCK"); const int v = e ? atoi(e) : 0; return (v == 64 || v == 128 || v == 256) ? v : 256; }
inline int nblocks_b(int64_t n, int b) { return (int)((n + b - 1) / b); }

// (re)build the uniform grid over `ref` on the stream (4 tiny kernels); small sets keep the shared-memory brute force
int build_grid(avc_ctx* ctx, const float* ref, int m, cudaStream_t st, GridView* gv) {
  gv->G = nullptr; gv->start = nullptr; gv->sorted = nullptr;
  if (m < 512) return AVC_OK;
  const size_t need = 256 + (size_t)GRID_MAX_CELLS * 4 + ((size_t)GRID_MAX_CELLS + 4) * 4 + (size_t)m * sizeof(float4) + 64;
  if (need > ctx->grid_cap) {
    if (ctx->d_grid) cudaFree(ctx->d_grid);
    ctx->d_grid = nullptr; ctx->grid_cap = 0;
    AVC_CUDA(ctx, cudaMalloc(&ctx->d_grid, need));
    ctx->grid_cap = need;
  }
  char* base = (char*)ctx->d_grid;
  GridDesc* G = (GridDesc*)base; int* cnt = (int*)(base + 256); int* start = cnt + GRID_MAX_CELLS;
  float4* sorted = (float4*)(base + 256 + (size_t)GRID_MAX_CELLS * 4 + ((size_t)GRID_MAX_CELLS + 4) * 4);
  const float h_min = [] { const char* e = getenv("AVC_KNN_CELL"); const float v = e ? (float)atof(e) : 0.f; return v >= 0.01f && v <= 1.f ? v : 0.04f; }();   // tuning knob (metres), read per call
  const int rmax = [] { const char* e = getenv("AVC_KNN_RMAX"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 12 ? v : GRID_RMAX; }();
  grid_bounds_kernel<<<1, 1024, 0, st>>>(ref, m, h_min, rmax, G, cnt);
  AVC_LAUNCH_CHECK(ctx, "grid_bounds_kernel");
  grid_count_kernel<<<(m + 255) / 256, 256, 0, st>>>(ref, m, G, cnt);
  AVC_LAUNCH_CHECK(ctx, "grid_count_kernel");
  grid_scan_kernel<<<1, 1024, 0, st>>>(G, cnt, start);
  AVC_LAUNCH_CHECK(ctx, "grid_scan_kernel");
  grid_fill_kernel<<<(m + 255) / 256, 256, 0, st>>>(ref, m, G, start, cnt, sorted);
  AVC_LAUNCH_CHECK(ctx, "grid_fill_kernel");
  gv->G = G; gv->start = start; gv->sorted = sorted;
  return AVC_OK;
}

}  // namespace

extern "C" int avc_knn(avc_ctx* ctx, const float* query, int64_t n, const float* ref, int m, int K, float* out_d2, int64_t* out_idx, void* stream) {
  if (!ctx || !query || !ref) return avc_fail(ctx, AVC_EINVAL, "avc_knn: NULL argument");
  if (K < 1 || K > 4 || m < K) return avc_fail(ctx, AVC_EINVAL, "avc_knn: K must be 1..4 and m >= K");
  if (n == 0) return AVC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GridView gv; int rc = build_grid(ctx, ref, m, st, &gv);
  if (rc) return rc;
  switch (K) {
    case 1: knn_kernel<1><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx, gv); break;
    case 2: knn_kernel<2><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx, gv); break;
    case 3: knn_kernel<3><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx, gv); break;
    default: knn_kernel<4><<<nblocks(n), KNN_NT, 0, st>>>(query, n, ref, m, out_d2, out_idx, gv); break;
  }
  AVC_LAUNCH_CHECK(ctx, "knn_kernel");
  return AVC_OK;
}

extern "C" int avc_lbs_weights(avc_ctx* ctx, const float* pts, int64_t n, const float* cano_verts, int m, const float* skin_weights,
                               float* out_lbs, void* stream) {
  if (!ctx || !pts || !cano_verts || !skin_weights || !out_lbs) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: NULL argument");
  if (m < 4) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: need at least 4 reference vertices");
  if (n == 0) return AVC_OK;
  GridView gv; int rc = build_grid(ctx, cano_verts, m, (cudaStream_t)stream, &gv);
  if (rc) return rc;
  const int kb = knn_block();
  lbs_weights_kernel<<<nblocks_b(n, kb), kb, 0, (cudaStream_t)stream>>>(pts, n, cano_verts, m, skin_weights, out_lbs, gv);
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
  GridView gv; int rc = build_grid(ctx, cano_verts, m, (cudaStream_t)stream, &gv);
  if (rc) return rc;
  const int kb = knn_block();
  skin_mesh_kernel<<<nblocks_b(n, kb), kb, 0, (cudaStream_t)stream>>>(verts, normals, n, cano_verts, m, skin_weights, jnt_mats, out_verts, out_normals, gv);
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
  GridView gv; int rc = build_grid(ctx, live_verts, m, (cudaStream_t)stream, &gv);
  if (rc) return rc;
  posed_to_cano_kernel<<<nblocks(n), KNN_NT, 0, (cudaStream_t)stream>>>(wpts, n, live_verts, m, skin_weights, live2cano_mats, p, weight_volume,
                                                                       out_cano, out_near, gv);
  AVC_LAUNCH_CHECK(ctx, "posed_to_cano_kernel");
  return AVC_OK;
}

extern "C" int avc_near_flag(avc_ctx* ctx, const float* query, int64_t n, const float* ref, int m, double radius, uint8_t* out_flag, void* stream) {
  if (!ctx || !query || !ref || !out_flag) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: NULL argument");
  if (m < 1 || !(radius > 0.0)) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: bad sizes");
  if (n == 0) return AVC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GridView gv; int rc = build_grid(ctx, ref, m, st, &gv);
  if (rc) return rc;
  if (!gv.G) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: needs at least 512 reference vertices (use avc_knn for small sets)");
  near_flag_kernel<<<nblocks(n), KNN_NT, 0, st>>>(query, n, gv, (float)(radius * radius), out_flag);   // float32(0.1 ** 2) like torch (dist < 0.1 ** 2)
  AVC_LAUNCH_CHECK(ctx, "near_flag_kernel");
  return AVC_OK;
}
