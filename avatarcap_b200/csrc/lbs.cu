// KNN against the SMPL vertex set, LBS weights and skinning.
//
// Reference call sites replaced:
//   pytorch3d.ops.knn_points / knn_gather   utils/smpl_util.py:33,37 ; network/arch_avatar.py:190,197,208 ; dataset/avatarcap_dataset.py:114
//   SmplUtil.calculate_lbs                  utils/smpl_util.py:24-39
//   SmplUtil.skinning / skinning_normal     utils/smpl_util.py:58-81
//   GeoTexAvatar.forward posed->cano warp   network/arch_avatar.py:189-205 ; CanoBlendWeightVolume.forward :152-165
//
// Every query thread keeps its K<=4 best candidates in registers and finds them through a two-level uniform grid over the
// reference set (6 890 SMPL vertices); the consumers of the neighbours (Gaussian weight blend, matrix blend, skinning, the
// posed->canonical warp) are fused behind the search as epilogues, so the (B,N,K,24) gather tensor and the (B,N,4,4) per-point
// matrices of the reference are never materialised unless the caller asks for them.
#include "common.cuh"

namespace {

constexpr int KNN_NT = 256;
#ifndef KNN_MIN_BLOCKS_K4
#define KNN_MIN_BLOCKS_K4 3
#endif

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

// per-thread brute force over a SMALL reference set (m < 512: no grid is built); every lane reads the same address -> broadcast loads
template <int K>
__device__ __forceinline__ void knn_brute_thread(const float* __restrict__ ref, int m, float qx, float qy, float qz, Top4& best) {
  top_init(best);
  for (int r = 0; r < m; ++r) {
    const float dx = __fsub_rn(qx, __ldg(ref + 3 * r)), dy = __fsub_rn(qy, __ldg(ref + 3 * r + 1)), dz = __fsub_rn(qz, __ldg(ref + 3 * r + 2));
    // squared L2, summed x,y,z in order without FMA contraction (bit-matches the CPU oracle)
    top_insert<K>(best, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), r);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Two-level uniform grid over the reference vertices: exact KNN in a few cells instead of a 6 890-vertex scan per query.
// After the shells of Chebyshev radius 0..r around the query's (clamped) cell have been visited, every unvisited vertex is
// at least r*h away (projection onto the grid box is non-expansive), so the search stops as soon as the K-th best squared
// distance is <= (r*h)^2. Inside a shell, a z-run of cells (one contiguous range of the cell-sorted vertex array) is skipped when
// its column is farther than the current K-th best, and clipped in z to the cells that can still hold a closer vertex -- the
// cube of a shell shrinks to the ball that matters. Queries that the fine level (4 cm cells) cannot finish within its shell
// limit restart on the coarse level (cells 4x as wide); the few that are far from everything even there go to a compacted
// list and a second launch (one WARP per query, brute force) -- no block-wide vote, no 24 KB staging tile in the main kernel.
constexpr int GRID_MAX_CELLS = 1 << 18;
constexpr int GRID1_MAX_CELLS = 1 << 13;       // coarse level: <= 17^3 cells
constexpr int GRID_RMAX = 4;    // shells walked per level before falling through to the next one (AVC_KNN_RMAX overrides, 1..12)
struct GridDesc { float ox, oy, oz, h, inv_h; int dx, dy, dz, m, cells, rmax; };

__device__ __forceinline__ int grid_cell(const GridDesc& G, float x, float y, float z, int& cx, int& cy, int& cz) {
  cx = min(max((int)floorf((x - G.ox) * G.inv_h), 0), G.dx - 1);
  cy = min(max((int)floorf((y - G.oy) * G.inv_h), 0), G.dy - 1);
  cz = min(max((int)floorf((z - G.oz) * G.inv_h), 0), G.dz - 1);
  return (cx * G.dy + cy) * G.dz + cz;
}

// level l: descriptor G[l], counters / cell starts at cnt + l * GRID_MAX_CELLS (the coarse level uses a prefix of its slot)
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
    for (int l = 0; l < 2; ++l) {
      GridDesc g;
      g.h = fmaxf(h_min, ext / 60.f) * (l ? 4.f : 1.f); g.inv_h = 1.f / g.h;
      g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
      g.dx = (int)floorf((mx[0] - mn[0]) * g.inv_h) + 1; g.dy = (int)floorf((mx[1] - mn[1]) * g.inv_h) + 1; g.dz = (int)floorf((mx[2] - mn[2]) * g.inv_h) + 1;
      g.rmax = rmax; g.m = m; g.cells = g.dx * g.dy * g.dz;          // <= 61^3 < GRID_MAX_CELLS, <= 16^3 < GRID1_MAX_CELLS
      G[l] = g;
    }
  }
}
__global__ void grid_count_kernel(const float* __restrict__ ref, int m, const GridDesc* __restrict__ Gp, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float x = ref[3 * i], y = ref[3 * i + 1], z = ref[3 * i + 2];
  int cx, cy, cz;
  atomicAdd(&cnt[grid_cell(Gp[0], x, y, z, cx, cy, cz)], 1);
  atomicAdd(&cnt[GRID_MAX_CELLS + grid_cell(Gp[1], x, y, z, cx, cy, cz)], 1);
}
// one block per level; a thread owns 16 consecutive cells per round (a 48 x 48 x 15 grid is 3 rounds instead of 34)
__global__ void __launch_bounds__(1024) grid_scan_kernel(const GridDesc* __restrict__ Gp, int* __restrict__ cnt_all, int* __restrict__ start_all) {
  __shared__ int wsum[32]; __shared__ int carry;
  constexpr int CPT = 16;
  const int level = blockIdx.x;
  int* cnt = cnt_all + level * GRID_MAX_CELLS; int* start = start_all + level * (GRID_MAX_CELLS + 4);
  const int cells = Gp[level].cells;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < cells; base += 1024 * CPT) {
    const int c0 = base + threadIdx.x * CPT;
    int x[CPT]; int sum = 0;
#pragma unroll
    for (int e = 0; e < CPT; ++e) { x[e] = c0 + e < cells ? cnt[c0 + e] : 0; sum += x[e]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < 32; ++w) { const int sw = wsum[w]; if (w < wid) pre += sw; tot += sw; }
    int run = carry + pre + inc - sum;
#pragma unroll
    for (int e = 0; e < CPT; ++e) { if (c0 + e < cells) { start[c0 + e] = run; cnt[c0 + e] = 0; } run += x[e]; }      // cnt becomes the fill cursor
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[cells] = carry;
}
__global__ void grid_fill_kernel(const float* __restrict__ ref, int m, const GridDesc* __restrict__ Gp, const int* __restrict__ start_all,
                                 int* __restrict__ cursor_all, float4* __restrict__ sorted_all) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float x = ref[3 * i], y = ref[3 * i + 1], z = ref[3 * i + 2];
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    int cx, cy, cz;
    const int c = grid_cell(Gp[l], x, y, z, cx, cy, cz);
    const int* start = start_all + l * (GRID_MAX_CELLS + 4);
    sorted_all[(size_t)l * m + start[c] + atomicAdd(&cursor_all[l * GRID_MAX_CELLS + c], 1)] = make_float4(x, y, z, __int_as_float(i));
  }
}

struct GridLevel { const GridDesc* G; const int* start; const float4* sorted; };
struct GridView { const GridDesc* G; GridLevel L[2]; int* far_count; int* far_list; int far_cap; };   // G == NULL: no grid (small sets)

__device__ __forceinline__ float dist2_rn(float qx, float qy, float qz, const float4& r) {
  const float dx = __fsub_rn(qx, r.x), dy = __fsub_rn(qy, r.y), dz = __fsub_rn(qz, r.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// distance from q to the slab [lo, lo + h) of one axis, made SMALLER by a slack of h/512: the cell of a vertex is computed in float
// and may be off by a rounding at a cell boundary -- pruning must never discard a cell that could hold a closer (or equal) vertex
__device__ __forceinline__ float axis_gap(float q, float lo, float h) { return fmaxf(fmaxf(lo - q, q - (lo + h)) - h * (1.f / 512.f), 0.f); }

// all vertices of the cells (i, j, k0..k1): cells that differ only in k are adjacent in the sorted array, so a whole z-run is ONE
// [start, end) range -- two index loads per run instead of two per cell. The run is skipped / clipped against the K-th best.
template <int K>
__device__ __forceinline__ void knn_visit_run(const GridLevel& V, const GridDesc& G, int i, int j, int k0, int k1, float qx, float qy, float qz, Top4& best) {
  const float bound = best.d[K - 1];
  if (bound < 3.0e38f) {
    const float gx = axis_gap(qx, G.ox + (float)i * G.h, G.h), gy = axis_gap(qy, G.oy + (float)j * G.h, G.h);
    const float dxy2 = gx * gx + gy * gy;
    if (dxy2 > bound) return;                                    // the whole column is farther than the K-th best
    const float rz = sqrtf(bound - dxy2) + G.h * (1.f / 256.f);
    k0 = max(k0, (int)floorf((qz - rz - G.oz) * G.inv_h));
    k1 = min(k1, (int)floorf((qz + rz - G.oz) * G.inv_h));
    if (k0 > k1) return;
  }
  const int c = (i * G.dy + j) * G.dz;
  const int e = __ldg(V.start + c + k1 + 1);
  for (int p = __ldg(V.start + c + k0); p < e; ++p) { const float4 v = __ldg(V.sorted + p); top_insert<K>(best, dist2_rn(qx, qy, qz, v), __float_as_int(v.w)); }
}

template <int K>
__device__ __forceinline__ bool knn_level(const GridLevel& V, float qx, float qy, float qz, Top4& best) {
  const GridDesc G = *V.G;
  top_init(best);
  int cx, cy, cz; grid_cell(G, qx, qy, qz, cx, cy, cz);
  // shells 0 and 1 together (shell 0 alone can never satisfy the stop test). The 9 z-runs' [start, end) pairs are fetched first, as
  // 18 independent loads (the search is bound by the latency of dependent index loads), then the centre column is scanned before the
  // other 8 so that it tightens the bound they are pruned against.
  {
    const int k0 = max(cz - 1, 0), k1 = min(cz + 1, G.dz - 1);
    int rs[9], re[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int i = cx + t / 3 - 1, j = cy + t % 3 - 1;
      const bool ok = i >= 0 && i < G.dx && j >= 0 && j < G.dy;
      const int c = ok ? (i * G.dy + j) * G.dz : 0;
      rs[t] = ok ? __ldg(V.start + c + k0) : 0; re[t] = ok ? __ldg(V.start + c + k1 + 1) : 0;
    }
#pragma unroll
    for (int u = 0; u < 9; ++u) {
      const int t = u == 0 ? 4 : (u <= 4 ? u - 1 : u);                 // centre (t = 4) first
      if (u > 0 && best.d[K - 1] < 3.0e38f) {
        const float gx = axis_gap(qx, G.ox + (float)(cx + t / 3 - 1) * G.h, G.h), gy = axis_gap(qy, G.oy + (float)(cy + t % 3 - 1) * G.h, G.h);
        if (gx * gx + gy * gy > best.d[K - 1]) continue;
      }
      for (int p = rs[t]; p < re[t]; ++p) { const float4 v = __ldg(V.sorted + p); top_insert<K>(best, dist2_rn(qx, qy, qz, v), __float_as_int(v.w)); }
    }
  }
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
  return done;
}

// fine level, then the coarse one; false: far from every vertex on both -> the caller queues the query for the brute-force launch
template <int K>
__device__ __forceinline__ bool knn_grid(const GridView& V, float qx, float qy, float qz, Top4& best) {
  if (knn_level<K>(V.L[0], qx, qy, qz, best)) return true;
  return knn_level<K>(V.L[1], qx, qy, qz, best);
}

// min squared distance < r2 ?  (dataset/avatarcap_dataset.py:114-116 valid flag) -- bounded search, exact, nearest columns first
__global__ void __launch_bounds__(KNN_NT) near_flag_kernel(const float* __restrict__ q, int64_t n, GridView V, float r2, uint8_t* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  if (g >= n) return;
  const float qx = q[g * 3], qy = q[g * 3 + 1], qz = q[g * 3 + 2];
  const float rad = sqrtf(r2);
  bool hit = false, maybe = true;
  {
    // coarse reject: no vertex in any coarse cell the ball can touch -> nothing within the radius (most of the grid points that are
    // not near the body end here after a handful of index loads)
    const GridDesc G = *V.L[1].G;
    const float px = fminf(fmaxf(qx, G.ox), G.ox + G.dx * G.h), py = fminf(fmaxf(qy, G.oy), G.oy + G.dy * G.h), pz = fminf(fmaxf(qz, G.oz), G.oz + G.dz * G.h);
    const float ob = (qx - px) * (qx - px) + (qy - py) * (qy - py) + (qz - pz) * (qz - pz);
    if (ob >= r2 * 1.0001f) maybe = false;
    else {
      const int R = (int)floorf(rad * G.inv_h) + 1;
      int cx, cy, cz; grid_cell(G, qx, qy, qz, cx, cy, cz);
      const int k0 = max(cz - R, 0), k1 = min(cz + R, G.dz - 1);
      int any = 0;
      for (int i = max(cx - R, 0); i <= min(cx + R, G.dx - 1); ++i)
        for (int j = max(cy - R, 0); j <= min(cy + R, G.dy - 1); ++j) {
          const int c = (i * G.dy + j) * G.dz;
          any |= __ldg(V.L[1].start + c + k1 + 1) - __ldg(V.L[1].start + c + k0);
        }
      maybe = any != 0;
    }
  }
  if (maybe) {
    const GridDesc G = *V.L[0].G;
    const int R = (int)floorf(rad * G.inv_h) + 1;   // >= ceil, with a full cell of slack when radius is a multiple of h
    int cx, cy, cz; grid_cell(G, qx, qy, qz, cx, cy, cz);
    // one z-run of cells (i, j, .): skipped when the column is farther than the radius, clipped in z to the cells the ball reaches
    auto visit = [&](int i, int j) {
      const float gx = axis_gap(qx, G.ox + (float)i * G.h, G.h), gy = axis_gap(qy, G.oy + (float)j * G.h, G.h);
      const float dxy2 = gx * gx + gy * gy;
      if (dxy2 >= r2) return;
      const float rz = sqrtf(r2 - dxy2) + G.h * (1.f / 256.f);
      const int k0 = max(max(cz - R, 0), (int)floorf((qz - rz - G.oz) * G.inv_h)), k1 = min(min(cz + R, G.dz - 1), (int)floorf((qz + rz - G.oz) * G.inv_h));
      if (k0 > k1) return;
      const int c = (i * G.dy + j) * G.dz;
      const int e = __ldg(V.L[0].start + c + k1 + 1);
      for (int p = __ldg(V.L[0].start + c + k0); p < e; ++p) if (dist2_rn(qx, qy, qz, __ldg(V.L[0].sorted + p)) < r2) { hit = true; break; }
    };
    // rings of columns around the query's column, nearest first: a near query hits in ring 0 or 1
    for (int ring = 0; ring <= R && !hit; ++ring)
      for (int i = max(cx - ring, 0); i <= min(cx + ring, G.dx - 1) && !hit; ++i) {
        if (abs(i - cx) == ring) {                                 // a full row of the ring
          for (int j = max(cy - ring, 0); j <= min(cy + ring, G.dy - 1) && !hit; ++j) visit(i, j);
        } else {                                                   // interior row: its two end columns
          if (cy - ring >= 0) visit(i, cy - ring);
          if (!hit && cy + ring <= G.dy - 1) visit(i, cy + ring);
        }
      }
  }
  out[g] = hit ? 1 : 0;
}

// M (rows 0..ROWS-1 of the blended 4x4, row-major) = sum_j lbs_j * J_j with lbs = sum_k w_k * skin_w[i_k]   smpl_util.py:33-38, :67
// Same order of operations as computing the 24 blend weights first and the matrix second, but four joints at a time: only
// 4 + ROWS*4 live accumulators instead of 24 + ROWS*4 (the round-1 kernels spilled 100-170 bytes here). out_lbs != NULL also
// stores the 24 weights (float4 stores).
template <int ROWS, int NW>
__device__ __forceinline__ void blend_rows(const float* const* rows, const float* w, const float* __restrict__ s_mats /*24*16 smem*/,
                                           float* M /*ROWS*4*/, float* __restrict__ out_lbs) {
#pragma unroll
  for (int e = 0; e < ROWS * 4; ++e) M[e] = 0.f;
#pragma unroll
  for (int q4 = 0; q4 < 6; ++q4) {
    float l[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(rows[k]) + q4);
      l[0] += t.x * w[k]; l[1] += t.y * w[k]; l[2] += t.z * w[k]; l[3] += t.w * w[k];
    }
    if (out_lbs) *reinterpret_cast<float4*>(out_lbs + 4 * q4) = make_float4(l[0], l[1], l[2], l[3]);
    if (ROWS > 0) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int e = 0; e < ROWS * 4; ++e) M[e] = fmaf(l[jj], s_mats[(4 * q4 + jj) * 16 + e], M[e]);
    }
  }
}

// Gaussian KNN-4 weights   smpl_util.py:33-36: exp(-d2 / (2 r^2)), r = 0.05, normalised with + 1e-16
__device__ __forceinline__ void gauss_weights(const Top4& best, float w[4]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { w[k] = expf(-best.d[k] / (2.f * 0.05f * 0.05f)); s += w[k]; }
  s += 1e-16f;
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = w[k] / s;
}

// ---- epilogues: what happens to a query's neighbours. K = neighbours needed, MATS = stage 24 joint matrices in shared memory
struct EpiKnn {            // pytorch3d.ops.knn_points: squared distances ascending + int64 indices
  float* out_d2; int64_t* out_idx; int k;
  static constexpr bool MATS = false;
  static constexpr int MIN_BLOCKS = 2;       // 128 registers: the K = 4 insertion network spills badly below that
  __device__ __forceinline__ const float* mats() const { return nullptr; }
  __device__ __forceinline__ void operator()(int64_t g, float, float, float, const Top4& best, const float*) const {
    for (int i = 0; i < k; ++i) {
      if (out_d2) out_d2[g * k + i] = best.d[i];
      if (out_idx) out_idx[g * k + i] = best.i[i];
    }
  }
};
struct EpiLbs {            // SmplUtil.calculate_lbs   smpl_util.py:24-39
  const float* skin_w; float* out_lbs;
  static constexpr bool MATS = false;
  static constexpr int MIN_BLOCKS = KNN_MIN_BLOCKS_K4;       // 3 = 80 registers, no spills: the search is latency bound, occupancy pays
  __device__ __forceinline__ const float* mats() const { return nullptr; }
  __device__ __forceinline__ void operator()(int64_t g, float, float, float, const Top4& best, const float*) const {
    float w[4]; gauss_weights(best, w);
    const float* rows[4] = {skin_w + (size_t)best.i[0] * 24, skin_w + (size_t)best.i[1] * 24, skin_w + (size_t)best.i[2] * 24, skin_w + (size_t)best.i[3] * 24};
    float M[1];
    blend_rows<0, 4>(rows, w, nullptr, M, out_lbs + g * 24);
  }
};
struct EpiSkinMesh {       // calculate_lbs + skinning (+ skinning_normal)   main.py:385-389, smpl_util.py:58-81
  const float* skin_w; const float* jm; const float* normals; float* out_v; float* out_n;
  static constexpr bool MATS = true;
  static constexpr int MIN_BLOCKS = KNN_MIN_BLOCKS_K4;
  __device__ __forceinline__ const float* mats() const { return jm; }
  __device__ __forceinline__ void operator()(int64_t g, float x, float y, float z, const Top4& best, const float* s_m) const {
    float w[4]; gauss_weights(best, w);
    const float* rows[4] = {skin_w + (size_t)best.i[0] * 24, skin_w + (size_t)best.i[1] * 24, skin_w + (size_t)best.i[2] * 24, skin_w + (size_t)best.i[3] * 24};
    float M[12];
    blend_rows<3, 4>(rows, w, s_m, M, nullptr);
    out_v[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + M[3];
    out_v[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    out_v[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
    if (normals && out_n) {                                     // rotation block only, no renormalisation (smpl_util.py:76-81)
      const float nx = normals[g * 3], ny = normals[g * 3 + 1], nz = normals[g * 3 + 2];
      out_n[g * 3 + 0] = M[0] * nx + M[1] * ny + M[2] * nz;
      out_n[g * 3 + 1] = M[4] * nx + M[5] * ny + M[6] * nz;
      out_n[g * 3 + 2] = M[8] * nx + M[9] * ny + M[10] * nz;
    }
  }
};
struct P2C {
  float bmin[3], len[3];
  int vd[3];
};
struct EpiPosedToCano {    // GeoTexAvatar.forward posed branch   arch_avatar.py:189-205, CanoBlendWeightVolume.forward :152-165
  const float* skin_w; const float* l2c; P2C p; const float* wvol; float* out_cano; uint8_t* out_near;
  static constexpr bool MATS = true;
  static constexpr int MIN_BLOCKS = 2;
  __device__ __forceinline__ const float* mats() const { return l2c; }
  __device__ __forceinline__ void operator()(int64_t g, float x, float y, float z, const Top4& best, const float* s_m) const {
    if (out_near) out_near[g] = best.d[0] < 0.08f * 0.08f ? 1 : 0;         // :191
    float M[12];
    {
      const float one[1] = {1.f};
      const float* rows[1] = {skin_w + (size_t)best.i[0] * 24};                // :197-198
      blend_rows<3, 1>(rows, one, s_m, M, nullptr);
    }
    float c[3];                                                             // :200
    c[0] = M[0] * x + M[1] * y + M[2] * z + M[3]; c[1] = M[4] * x + M[5] * y + M[6] * z + M[7]; c[2] = M[8] * x + M[9] * y + M[10] * z + M[11];
    // normalise to [0,1] by the canonical bounds (:201-203) then grid = 2p-1 and unnormalise with align_corners=True, border (:154-160)
    int i0[3]; float f[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p01 = (c[a] - p.bmin[a]) / p.len[a];
      const float gg = 2.f * p01 - 1.f;
      float s = ((gg + 1.f) / 2.f) * (float)(p.vd[a] - 1);
      s = fminf((float)(p.vd[a] - 1), fmaxf(s, 0.f));
      i0[a] = (int)floorf(s); f[a] = s - (float)i0[a];
    }
    const float* rows[8]; float w[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int dx = t >> 2, dy = (t >> 1) & 1, dz = t & 1;
      int xi = i0[0] + dx, yi = i0[1] + dy, zi = i0[2] + dz;
      const bool ok = xi <= p.vd[0] - 1 && yi <= p.vd[1] - 1 && zi <= p.vd[2] - 1;   // zero-weight taps
      xi = min(xi, p.vd[0] - 1); yi = min(yi, p.vd[1] - 1); zi = min(zi, p.vd[2] - 1);
      w[t] = ok ? (dx ? f[0] : 1.f - f[0]) * (dy ? f[1] : 1.f - f[1]) * (dz ? f[2] : 1.f - f[2]) : 0.f;
      rows[t] = wvol + (((size_t)xi * p.vd[1] + yi) * p.vd[2] + zi) * 24;
    }
    blend_rows<3, 8>(rows, w, s_m, M, nullptr);                             // :205
    out_cano[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + M[3];
    out_cano[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    out_cano[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
  }
};

// main launch: one query per thread through the grid; queries the grid cannot finish are appended to the far list
template <int K, class Epi>
__global__ void __launch_bounds__(KNN_NT, Epi::MIN_BLOCKS) knn_main_kernel(const float* __restrict__ q, int64_t n, const float* __restrict__ ref, int m, GridView V, Epi epi) {
  __shared__ float s_m[Epi::MATS ? 24 * 16 : 1];
  if (Epi::MATS) {
    for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = epi.mats()[t];
    __syncthreads();
  }
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  if (g >= n) return;
  const float x = q[g * 3], y = q[g * 3 + 1], z = q[g * 3 + 2];
  Top4 best;
  if (!V.G) knn_brute_thread<K>(ref, m, x, y, z, best);
  else if (!knn_grid<K>(V, x, y, z, best)) {
    const int slot = atomicAdd(V.far_count, 1);
    if (slot < V.far_cap) V.far_list[slot] = (int)g;          // far_cap == n: cannot overflow
    return;
  }
  epi(g, x, y, z, best, s_m);
}

// second launch: one WARP per far query, brute force over the whole reference set. Lane l scans references l, l+32, ... keeping its
// own top-K; K rounds of a lexicographic (distance, index) warp arg-min pop the global top-K in order -- the same total order as
// the per-thread insertion, so the result is bit-identical to the grid search and to the brute-force oracle.
template <int K, class Epi>
__global__ void __launch_bounds__(KNN_NT) knn_far_kernel(const float* __restrict__ q, const float* __restrict__ ref, int m, GridView V, Epi epi) {
  __shared__ float s_m[Epi::MATS ? 24 * 16 : 1];
  if (Epi::MATS) {
    for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = epi.mats()[t];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int n_far = min(*V.far_count, V.far_cap);
  const int warps = (gridDim.x * KNN_NT) >> 5;
  for (int fq = (blockIdx.x * KNN_NT + threadIdx.x) >> 5; fq < n_far; fq += warps) {
    const int64_t g = V.far_list[fq];
    const float x = q[g * 3], y = q[g * 3 + 1], z = q[g * 3 + 2];
    Top4 mine; top_init(mine);
    for (int r = lane; r < m; r += 32) {
      const float dx = __fsub_rn(x, __ldg(ref + 3 * r)), dy = __fsub_rn(y, __ldg(ref + 3 * r + 1)), dz = __fsub_rn(z, __ldg(ref + 3 * r + 2));
      top_insert<K>(mine, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), r);
    }
    Top4 best; top_init(best);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float d = mine.d[0]; int i = mine.i[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, o); const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (od < d || (od == d && oi < i)) { d = od; i = oi; }
      }
      best.d[k] = d; best.i[k] = i;
      if (mine.i[0] == i && mine.d[0] == d) {                    // the winning lane pops its head
#pragma unroll
        for (int t = 0; t < K - 1; ++t) { mine.d[t] = mine.d[t + 1]; mine.i[t] = mine.i[t + 1]; }
        mine.d[K - 1] = 3.4e38f; mine.i[K - 1] = 0x7fffffff;
      }
    }
    if (lane == 0) epi(g, x, y, z, best, s_m);
    __syncwarp();
  }
}

// skinning with given weights   smpl_util.py:58-81
__global__ void __launch_bounds__(KNN_NT) skin_kernel(const float* __restrict__ pts, const float* __restrict__ lbs_g, const float* __restrict__ jm,
                                                      int64_t n, float* __restrict__ out_pts, float* __restrict__ out_mats, int normal_mode) {
  __shared__ float s_m[24 * 16];
  for (int t = threadIdx.x; t < 24 * 16; t += blockDim.x) s_m[t] = jm[t];
  __syncthreads();
  const int64_t g = (int64_t)blockIdx.x * KNN_NT + threadIdx.x;
  if (g >= n) return;
  const float one[1] = {1.f};
  const float* rows[1] = {lbs_g + g * 24};
  float M[16]; blend_rows<4, 1>(rows, one, s_m, M, nullptr);
  const float x = pts[g * 3], y = pts[g * 3 + 1], z = pts[g * 3 + 2];
  const float t = normal_mode ? 0.f : 1.f;       // skinning_normal: rotation block only (smpl_util.py:80)
  out_pts[g * 3 + 0] = M[0] * x + M[1] * y + M[2] * z + t * M[3];
  out_pts[g * 3 + 1] = M[4] * x + M[5] * y + M[6] * z + t * M[7];
  out_pts[g * 3 + 2] = M[8] * x + M[9] * y + M[10] * z + t * M[11];
  if (out_mats)
#pragma unroll
    for (int e = 0; e < 4; ++e) *reinterpret_cast<float4*>(out_mats + g * 16 + 4 * e) = make_float4(M[4 * e], M[4 * e + 1], M[4 * e + 2], M[4 * e + 3]);
}

inline int nblocks(int64_t n) { return (int)((n + KNN_NT - 1) / KNN_NT); }

// (re)build the two-level grid over `ref` on the stream (4 small kernels) and reserve the far list for n queries; small sets
// (m < 512) keep the per-thread brute force
int build_grid(avc_ctx* ctx, const float* ref, int m, int64_t n_query, cudaStream_t st, GridView* gv) {
  gv->G = nullptr; gv->far_count = nullptr; gv->far_list = nullptr; gv->far_cap = 0;
  for (int l = 0; l < 2; ++l) { gv->L[l].G = nullptr; gv->L[l].start = nullptr; gv->L[l].sorted = nullptr; }
  if (m < 512) return AVC_OK;
  if (n_query > 0x7fffffffLL) return avc_fail(ctx, AVC_EINVAL, "too many query points for one call (%lld)", (long long)n_query);
  const size_t off_cnt = 256, off_start = off_cnt + (size_t)2 * GRID_MAX_CELLS * 4, off_sorted = off_start + (size_t)2 * (GRID_MAX_CELLS + 4) * 4;
  const size_t need = off_sorted + (size_t)2 * m * sizeof(float4) + 64;
  if (need > ctx->grid_cap) {
    if (ctx->d_grid) cudaFree(ctx->d_grid);
    ctx->d_grid = nullptr; ctx->grid_cap = 0;
    AVC_CUDA(ctx, cudaMalloc(&ctx->d_grid, need));
    ctx->grid_cap = need;
  }
  char* base = (char*)ctx->d_grid;
  GridDesc* G = (GridDesc*)base; int* cnt = (int*)(base + off_cnt); int* start = (int*)(base + off_start);
  float4* sorted = (float4*)(base + off_sorted);
  if (n_query > 0) {                     // far list: a counter + one int per query, in the context scratch
    int rc = avc_ensure_scratch(ctx, 64 + (size_t)n_query * sizeof(int));
    if (rc) return rc;
    gv->far_count = (int*)ctx->d_scratch; gv->far_list = (int*)((char*)ctx->d_scratch + 64); gv->far_cap = (int)n_query;
    AVC_CUDA(ctx, cudaMemsetAsync(gv->far_count, 0, sizeof(int), st));
  }
  const float h_min = [] { const char* e = getenv("AVC_KNN_CELL"); const float v = e ? (float)atof(e) : 0.f; return v >= 0.01f && v <= 1.f ? v : 0.04f; }();   // tuning knob (metres), read per call
  const int rmax = [] { const char* e = getenv("AVC_KNN_RMAX"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 12 ? v : GRID_RMAX; }();
  AVC_CUDA(ctx, cudaMemsetAsync(cnt, 0, (size_t)2 * GRID_MAX_CELLS * sizeof(int), st));      // the copy engine zeroes the counters (was a 24 us single-block loop)
  grid_bounds_kernel<<<1, 1024, 0, st>>>(ref, m, h_min, rmax, G, cnt);
  AVC_LAUNCH_CHECK(ctx, "grid_bounds_kernel");
  grid_count_kernel<<<(m + 255) / 256, 256, 0, st>>>(ref, m, G, cnt);
  AVC_LAUNCH_CHECK(ctx, "grid_count_kernel");
  grid_scan_kernel<<<2, 1024, 0, st>>>(G, cnt, start);
  AVC_LAUNCH_CHECK(ctx, "grid_scan_kernel");
  grid_fill_kernel<<<(m + 255) / 256, 256, 0, st>>>(ref, m, G, start, cnt, sorted);
  AVC_LAUNCH_CHECK(ctx, "grid_fill_kernel");
  gv->G = G;
  for (int l = 0; l < 2; ++l) { gv->L[l].G = G + l; gv->L[l].start = start + l * (GRID_MAX_CELLS + 4); gv->L[l].sorted = sorted + (size_t)l * m; }
  return AVC_OK;
}

// main launch over all queries + (with a grid) the far-query launch on a fixed persistent grid
template <int K, class Epi>
int run_knn(avc_ctx* ctx, const float* q, int64_t n, const float* ref, int m, const Epi& epi, cudaStream_t st, const char* name) {
  GridView gv; int rc = build_grid(ctx, ref, m, n, st, &gv);
  if (rc) return rc;
  knn_main_kernel<K, Epi><<<nblocks(n), KNN_NT, 0, st>>>(q, n, ref, m, gv, epi);
  AVC_LAUNCH_CHECK(ctx, name);
  if (gv.G) {
    int blocks = ctx->sm_count * 4; const int64_t want = (n * 32 + KNN_NT - 1) / KNN_NT;
    if (want < blocks) blocks = (int)want;
    knn_far_kernel<K, Epi><<<blocks, KNN_NT, 0, st>>>(q, ref, m, gv, epi);
    AVC_LAUNCH_CHECK(ctx, "knn_far_kernel");
  }
  return AVC_OK;
}

}  // namespace

extern "C" int avc_knn(avc_ctx* ctx, const float* query, int64_t n, const float* ref, int m, int K, float* out_d2, int64_t* out_idx, void* stream) {
  if (!ctx || !query || !ref) return avc_fail(ctx, AVC_EINVAL, "avc_knn: NULL argument");
  if (K < 1 || K > 4 || m < K) return avc_fail(ctx, AVC_EINVAL, "avc_knn: K must be 1..4 and m >= K");
  if (n == 0) return AVC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  EpiKnn epi; epi.out_d2 = out_d2; epi.out_idx = out_idx; epi.k = K;
  switch (K) {
    case 1: return run_knn<1>(ctx, query, n, ref, m, epi, st, "knn_main_kernel<1>");
    case 2: return run_knn<2>(ctx, query, n, ref, m, epi, st, "knn_main_kernel<2>");
    case 3: return run_knn<3>(ctx, query, n, ref, m, epi, st, "knn_main_kernel<3>");
    default: return run_knn<4>(ctx, query, n, ref, m, epi, st, "knn_main_kernel<4>");
  }
}

extern "C" int avc_lbs_weights(avc_ctx* ctx, const float* pts, int64_t n, const float* cano_verts, int m, const float* skin_weights,
                               float* out_lbs, void* stream) {
  if (!ctx || !pts || !cano_verts || !skin_weights || !out_lbs) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: NULL argument");
  if (m < 4) return avc_fail(ctx, AVC_EINVAL, "avc_lbs_weights: need at least 4 reference vertices");
  if (n == 0) return AVC_OK;
  EpiLbs epi; epi.skin_w = skin_weights; epi.out_lbs = out_lbs;
  return run_knn<4>(ctx, pts, n, cano_verts, m, epi, (cudaStream_t)stream, "knn_main_kernel<lbs>");
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
  EpiSkinMesh epi; epi.skin_w = skin_weights; epi.jm = jnt_mats; epi.normals = normals; epi.out_v = out_verts; epi.out_n = out_normals;
  return run_knn<4>(ctx, verts, n, cano_verts, m, epi, (cudaStream_t)stream, "knn_main_kernel<skin_mesh>");
}

extern "C" int avc_posed_to_cano(avc_ctx* ctx, const float* wpts, int64_t n, const float* live_verts, int m, const float* skin_weights,
                                 const float* live2cano_mats, const float bounds[6], const float* weight_volume, const int vdims[3],
                                 float* out_cano, uint8_t* out_near, void* stream) {
  if (!ctx || !wpts || !live_verts || !skin_weights || !live2cano_mats || !bounds || !weight_volume || !vdims || !out_cano)
    return avc_fail(ctx, AVC_EINVAL, "avc_posed_to_cano: NULL argument");
  if (m < 1 || vdims[0] < 1 || vdims[1] < 1 || vdims[2] < 1) return avc_fail(ctx, AVC_EINVAL, "avc_posed_to_cano: bad sizes");
  if (n == 0) return AVC_OK;
  EpiPosedToCano epi;
  for (int a = 0; a < 3; ++a) { epi.p.bmin[a] = bounds[a]; epi.p.len[a] = bounds[3 + a] - bounds[a]; epi.p.vd[a] = vdims[a]; }
  epi.skin_w = skin_weights; epi.l2c = live2cano_mats; epi.wvol = weight_volume; epi.out_cano = out_cano; epi.out_near = out_near;
  return run_knn<1>(ctx, wpts, n, live_verts, m, epi, (cudaStream_t)stream, "knn_main_kernel<posed_to_cano>");
}

extern "C" int avc_near_flag(avc_ctx* ctx, const float* query, int64_t n, const float* ref, int m, double radius, uint8_t* out_flag, void* stream) {
  if (!ctx || !query || !ref || !out_flag) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: NULL argument");
  if (m < 1 || !(radius > 0.0)) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: bad sizes");
  if (n == 0) return AVC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  GridView gv; int rc = build_grid(ctx, ref, m, 0, st, &gv);
  if (rc) return rc;
  if (!gv.G) return avc_fail(ctx, AVC_EINVAL, "avc_near_flag: needs at least 512 reference vertices (use avc_knn for small sets)");
  near_flag_kernel<<<nblocks(n), KNN_NT, 0, st>>>(query, n, gv, (float)(radius * radius), out_flag);   // float32(0.1 ** 2) like torch (dist < 0.1 ** 2)
  AVC_LAUNCH_CHECK(ctx, "near_flag_kernel");
  return AVC_OK;
}
