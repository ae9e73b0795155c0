// Per-frame image encoder (HGFilter, network/HGFilters.py:124-219; ReconNetwork.get_feat_maps, network/arch_recon.py:41-43) on the
// 5th-gen tensor cores: every 3x3 / 1x1 convolution is an implicit GEMM (M = 128 pixels, N = output channels, K = taps x input
// channels) issued as tcgen05.mma with fp32 accumulators in TMEM, operands moved by TMA TENSOR loads:
//   * activations live in HBM as (H, W, C) fp16 hi / lo planes (x = hi + lo to ~2^-22, the same split-precision scheme as the field
//     kernel: single-pass fp16 / tf32 are 2.7e-3 off on the 32x256x256 feature map, the 3-pass product is as exact as f32,
//     profiles/r1_hgfilter_split_precision.txt). A 3-D tensor map (C, W, H) with a (64 ch, bw px, bh rows) box, bw*bh = 128, loads the
//     A tile of one filter tap as a SHIFTED box: tap (dy, dx) = box origin (x0 + dx, y0 + dy); rows / columns outside the image are
//     zero-filled by the TMA unit -- that IS the convolution's zero padding, no im2col buffer, no halo logic. 128-byte swizzle;
//   * weights are (C_out, taps, C_in) fp16 hi / lo, box (64, 1, C_out);
//   * 3 MMAs per product (hi*hi + hi*lo + lo*hi), K = 16 per instruction, 4 k-steps per 128-byte swizzle atom;
//   * persistent CTAs, warp-specialised: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2-5 epilogue. Two TMEM
//     accumulator buffers: the epilogue of tile t (TMEM -> registers -> f32 NHWC slice, optional bias / accumulate) overlaps the
//     main loop of tile t+1.
// Around the convolutions: GroupNorm statistics (two-stage, fixed-order double reduction: deterministic), normalise + ReLU + hi/lo
// split, residual add, 2x2 average pool, bicubic x2 up-sampling (+ skip add), and the 7x7 stride-2 stem on the CUDA cores.
// The network structure itself is a PROGRAM of these ops built on the host side (avatarcap_b200/encoders.py) from the reference's
// state_dict and interpreted here; the whole program is captured once into a CUDA graph.
#include "common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <math.h>

#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(void* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) {      // a protocol bug must abort the launch instead of hanging the GPU
      printf("avatarcap_b200: conv mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, a, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, void* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ss1(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit1(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, 128-byte swizzle: rows of 128 B (64 fp16), 8-row groups 1024 B apart (SBO), layout type 2 = SWIZZLE_128B,
// descriptor version 1 (bit 46). The k-th 16-element step inside the swizzle atom advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D fp32 (bit 4), A/B fp16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 128)
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
      "%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ implicit-GEMM convolution
constexpr int CV_THREADS = 192;      // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int CV_MAX_STAGES = 4;
constexpr int CV_A_BYTES = 128 * 128;   // one A plane of a stage: 128 pixels x 64 channels fp16

struct ConvArgs {
  int W, bw, bh, tiles_x, n_tiles;     // tile = bh rows x bw pixels (bw * bh == 128)
  int kslabs, taps, N, stages;         // 64-channel slabs of the input, filter taps (1 or 9), output channels, ring depth
  float* out; int ldc, c_off;          // f32 (P, ldc) destination, channel offset of the slice this convolution writes
  int accumulate;                      // out += conv (the residual path of ConvBlock, HGFilters.py:69-73) instead of out = conv
  const float* bias;                   // per output channel or NULL
  float scale;                         // 2^-s: the packer scales the weights by 2^s so that their fp16 lo parts stay normal
};

struct __align__(8) ConvBars {
  unsigned long long full[CV_MAX_STAGES], empty[CV_MAX_STAGES], acc_full[2], acc_empty[2];
  unsigned int tmem_base; unsigned int pad;
};

__global__ void __launch_bounds__(CV_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                                                               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                                                               const ConvArgs a) {
  extern __shared__ unsigned char cv_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cv_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = a.N * 128;                                  // one B plane of a stage
  const int stage_bytes = 2 * CV_A_BYTES + 2 * b_bytes;
  ConvBars& S = *reinterpret_cast<ConvBars*>(ring + (size_t)a.stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.acc_full[b], 1); mbar_init(&S.acc_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const int k_iters = a.taps * a.kslabs;

  if (warp == 0) {
    // ===================================================== TMA producer
    int s = 0; uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const int x0 = (tile % a.tiles_x) * a.bw, y0 = (tile / a.tiles_x) * a.bh;
      for (int tap = 0; tap < a.taps; ++tap) {
        const int dy = a.taps == 9 ? tap / 3 - 1 : 0, dx = a.taps == 9 ? tap % 3 - 1 : 0;
        for (int ks = 0; ks < a.kslabs; ++ks) {
          mbar_wait(&S.empty[s], ph ^ 1);
          if (elect_one()) {
            unsigned char* st = ring + (size_t)s * stage_bytes;
            mbar_expect_tx(&S.full[s], (uint32_t)stage_bytes);
            tma_load_3d(st, &tm_a_hi, ks * 64, x0 + dx, y0 + dy, &S.full[s]);                   // shifted box: the tap; OOB = zero padding
            tma_load_3d(st + CV_A_BYTES, &tm_a_lo, ks * 64, x0 + dx, y0 + dy, &S.full[s]);
            tma_load_3d(st + 2 * CV_A_BYTES, &tm_b_hi, ks * 64, tap, 0, &S.full[s]);
            tma_load_3d(st + 2 * CV_A_BYTES + b_bytes, &tm_b_lo, ks * 64, tap, 0, &S.full[s]);
          }
          __syncwarp();
          if (++s == a.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint32_t idesc = make_idesc(a.N);
    int s = 0; uint32_t ph = 0; int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&S.acc_empty[buf], (uint32_t)((it >> 1) & 1) ^ 1u);       // the epilogue has drained this accumulator
      tc_fence_after();
      // Two accumulators per tile when they fit (N <= 128): the hi*hi products in one, the two 2^-11-sized cross terms in the other,
      // summed in fp32 by the epilogue. The tensor core adds into the fp32 accumulator with truncation, a bias that grows with the
      // number of accumulation steps (3 x 144 for a 3x3 convolution over 256 channels); keeping the small terms apart takes two
      // thirds of the steps -- and their rounding -- off the main sum.
      const bool dual = a.N <= 128;
      const uint32_t d_addr = tmem + (uint32_t)(buf * 256), d_cross = dual ? d_addr + 128u : d_addr;
      uint32_t acc = 0, acc_x = dual ? 0u : 1u;
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&S.full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(ring + (size_t)s * stage_bytes);
          const uint64_t ah = sw128_desc(st), al = sw128_desc(st + CV_A_BYTES);
          const uint64_t bh = sw128_desc(st + 2 * CV_A_BYTES), bl = sw128_desc(st + 2 * CV_A_BYTES + b_bytes);
#pragma unroll
          for (int u = 0; u < 4; ++u) {                                  // 4 x K16 inside the 128-byte atom: +32 B = +2 in the address field
            mma_ss1(d_addr, ah + 2 * u, bh + 2 * u, idesc, acc); acc = 1u;
            mma_ss1(d_cross, ah + 2 * u, bl + 2 * u, idesc, acc_x); acc_x = 1u;
            mma_ss1(d_cross, al + 2 * u, bh + 2 * u, idesc, 1u);
          }
          tc_commit1(&S.empty[s]);                                       // the stage is free once these MMAs have read it
          if (k == k_iters - 1) tc_commit1(&S.acc_full[buf]);            // ... and the tile's accumulator is complete
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5 <-> TMEM lane quadrants 2, 3, 0, 1)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;                                      // row of the tile == TMEM lane
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int x = (tile % a.tiles_x) * a.bw + (m % a.bw), y = (tile / a.tiles_x) * a.bh + m / a.bw;
      float* dst = a.out + ((size_t)y * a.W + x) * a.ldc + a.c_off;
      mbar_wait(&S.acc_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t t_row = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256);
      for (int c0 = 0; c0 < a.N; c0 += 32) {
        float v[32];
        tmem_ld32(t_row + (uint32_t)c0, v);
        if (a.N <= 128) {                                                // + the cross-term accumulator (see the MMA issuer)
          float x[32];
          tmem_ld32(t_row + 128u + (uint32_t)c0, x);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += x[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= a.scale;
        if (a.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(a.bias + c0 + i);
        }
        float4* d4 = reinterpret_cast<float4*>(dst + c0);
        if (a.accumulate) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float4 o = d4[i]; v[4 * i] += o.x; v[4 * i + 1] += o.y; v[4 * i + 2] += o.z; v[4 * i + 3] += o.w; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ element-wise stages (f32 NHWC)
// GroupNorm(32, C) statistics of the (P, C) slice [c_off, c_off + C) of a (P, ld) tensor (HGFilters.py:45-49, eps 1e-5).
// Stage 1: each block reduces a chunk of pixels to per-group (sum, sum of squares) in double, fixed order; the last block to finish
// (atomic ticket) folds the per-block partials in block order -> bit-reproducible. C in {32, 64, 128, 256}: a thread's channel is fixed.
constexpr int GN_PX = 128;
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, int P, int C, int ld, int c_off, double* __restrict__ partial,
                                                       unsigned int* __restrict__ ticket, float* __restrict__ stats /*32 x {mean, rstd}*/) {
  __shared__ float s_sum[1024], s_sq[1024];           // [thread][j]: the 4 channels a thread owns
  __shared__ bool s_last;
  const int tid = threadIdx.x, C4 = C / 4, c4 = tid % C4, rows = 256 / C4;    // a pixel's C channels are C/4 float4 loads; `rows` pixels per sweep
  const int p0 = blockIdx.x * GN_PX, p1 = min(P, p0 + GN_PX);
  float sum[4] = {0.f, 0.f, 0.f, 0.f}, sq[4] = {0.f, 0.f, 0.f, 0.f};
  const float* base = x + c_off + 4 * c4;
#pragma unroll 4
  for (int p = p0 + tid / C4; p < p1; p += rows) {       // independent 16-byte loads, four in flight per thread
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * ld));
    sum[0] += v.x; sum[1] += v.y; sum[2] += v.z; sum[3] += v.w;
    sq[0] = fmaf(v.x, v.x, sq[0]); sq[1] = fmaf(v.y, v.y, sq[1]); sq[2] = fmaf(v.z, v.z, sq[2]); sq[3] = fmaf(v.w, v.w, sq[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { s_sum[4 * tid + j] = sum[j]; s_sq[4 * tid + j] = sq[j]; }
  __syncthreads();
  if (tid < 32) {
    // group g = channels [g*cg, (g+1)*cg); entry (t, j) holds channel 4*(t % C4) + j: fold the 32 entries of the group in a fixed order
    const int cg = C / 32;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < rows; ++r)
      for (int k = 0; k < cg; ++k) { const int ch = tid * cg + k; const int e = 4 * (r * C4 + ch / 4) + (ch & 3); a += (double)s_sum[e]; b += (double)s_sq[e]; }
    partial[((size_t)blockIdx.x * 32 + tid) * 2] = a; partial[((size_t)blockIdx.x * 32 + tid) * 2 + 1] = b;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    // the last block folds the per-block partials: 8 threads per group take interleaved blocks, then a fixed-order fold of the 8
    __shared__ double s_a[256], s_b[256];
    const int g = tid & 31, part = tid >> 5;
    double a = 0.0, b = 0.0;
    for (unsigned int blk = part; blk < gridDim.x; blk += 8) { a += partial[((size_t)blk * 32 + g) * 2]; b += partial[((size_t)blk * 32 + g) * 2 + 1]; }
    s_a[tid] = a; s_b[tid] = b;
    __syncthreads();
    if (tid < 32) {
      a = 0.0; b = 0.0;
      for (int q = 0; q < 8; ++q) { a += s_a[q * 32 + tid]; b += s_b[q * 32 + tid]; }
      const double cnt = (double)P * (double)(C / 32);
      const double mean = a / cnt, var = fmax(b / cnt - mean * mean, 0.0);
      stats[2 * tid] = (float)mean; stats[2 * tid + 1] = (float)(1.0 / sqrt(var + 1e-5));
    }
  }
  if (tid == 0) *ticket = 0;
}

// y = [relu]( (x - mean_g) * rstd_g * gamma_c + beta_c ) (stats == NULL: y = x), written as fp16 hi / lo planes (P, cpad) and / or f32
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, int P, int C, int ld, int c_off, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                       __half* __restrict__ hi, __half* __restrict__ lo, int cpad, float* __restrict__ y32, int ld32) {
  const int64_t n4 = (int64_t)P * (C / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / (C / 4)), c = (int)(i % (C / 4)) * 4;
    const float4 v4 = *reinterpret_cast<const float4*>(x + (size_t)p * ld + c_off + c);
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
    if (stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gg = (c + j) / (C / 32);
        v[j] = (v[j] - stats[2 * gg]) * stats[2 * gg + 1] * __ldg(gamma + c + j) + __ldg(beta + c + j);
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (y32) *reinterpret_cast<float4*>(y32 + (size_t)p * ld32 + c) = make_float4(v[0], v[1], v[2], v[3]);
    if (hi) {
      __half h[4], l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { h[j] = __float2half_rn(v[j]); l[j] = __float2half_rn(v[j] - __half2float(h[j])); }
      *reinterpret_cast<uint2*>(hi + (size_t)p * cpad + c) = *reinterpret_cast<const uint2*>(h);
      *reinterpret_cast<uint2*>(lo + (size_t)p * cpad + c) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

// dst (P, C) += src (P, C): the identity residual of ConvBlock (HGFilters.py:73 with downsample None)
__global__ void __launch_bounds__(256) add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i]; const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}

// F.avg_pool2d(x, 2, stride=2) (HGFilters.py:105): (H, W, C) -> (H/2, W/2, C)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int64_t n = (int64_t)Ho * Wo * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4; const int64_t t = i / C4; const int xo = (int)(t % Wo), yo = (int)(t / Wo);
    const float* p = x + ((size_t)(2 * yo) * W + 2 * xo) * C + c;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + C);
    const float4 d = *reinterpret_cast<const float4*>(p + (size_t)W * C), e = *reinterpret_cast<const float4*>(p + (size_t)W * C + C);
    *reinterpret_cast<float4*>(y + ((size_t)yo * Wo + xo) * C + c) =
        make_float4((a.x + b.x + d.x + e.x) * 0.25f, (a.y + b.y + d.y + e.y) * 0.25f, (a.z + b.z + d.z + e.z) * 0.25f, (a.w + b.w + d.w + e.w) * 0.25f);
  }
}

// out = up1 + F.interpolate(low, scale_factor=2, mode='bicubic', align_corners=True)   (HGFilters.py:115-117); ATen's cubic
// convolution coefficients (A = -0.75), source index = dst * (in - 1) / (out - 1), taps clamped to the border
__device__ __forceinline__ void cubic_coeffs(float t, float w[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x3 = (1.f - t) + 1.f, x2 = 1.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
__global__ void __launch_bounds__(256) bicubic_up2_add_kernel(const float* __restrict__ up1, const float* __restrict__ low, float* __restrict__ out,
                                                              int h, int w, int C) {
  const int H = 2 * h, W = 2 * w, C4 = C / 4;
  const float sy = (float)(h - 1) / (float)(H - 1), sx = (float)(w - 1) / (float)(W - 1);
  const int64_t n = (int64_t)H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4; const int64_t t = i / C4; const int ox = (int)(t % W), oy = (int)(t / W);
    const float ry = sy * (float)oy, rx = sx * (float)ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float wy[4], wx[4];
    cubic_coeffs(ry - (float)iy, wy); cubic_coeffs(rx - (float)ix, wx);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), h - 1);
      float row[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), w - 1);
        const float4 v = *reinterpret_cast<const float4*>(low + ((size_t)yy * w + xx) * C + c);
        row[0] += wx[b] * v.x; row[1] += wx[b] * v.y; row[2] += wx[b] * v.z; row[3] += wx[b] * v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] += wy[a] * row[j];
    }
    const float4 u = *reinterpret_cast<const float4*>(up1 + ((size_t)oy * W + ox) * C + c);
    *reinterpret_cast<float4*>(out + ((size_t)oy * W + ox) * C + c) = make_float4(u.x + acc[0], u.y + acc[1], u.z + acc[2], u.w + acc[3]);
  }
}

// nn.Upsample(mode='bilinear', scale_factor=2, align_corners=False) of [relu](x) written as fp16 hi / lo planes -- the input stage of
// UpConv2DBlock(up_mode='upsample') (network/unets.py:41-44, 47-49): src = (dst + 0.5) / 2 - 0.5 clamped at 0, second tap clamped to the
// last row / column (ATen upsample_bilinear2d). x: (h, w) pixels, channels [c_off, c_off + C) of rows of length ld; planes: (2h * 2w, cpad).
__global__ void __launch_bounds__(256) bilinear_up2_split_kernel(const float* __restrict__ x, int h, int w, int C, int ld, int c_off, int relu,
                                                                 __half* __restrict__ hi, __half* __restrict__ lo, int cpad) {
  const int H = 2 * h, W = 2 * w, C4 = C / 4;
  const int64_t n = (int64_t)H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4; const int64_t t = i / C4; const int ox = (int)(t % W), oy = (int)(t / W);
    const float sy = fmaxf(((float)oy + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf(((float)ox + 0.5f) * 0.5f - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx, y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* base = x + c_off + c;
    float4 a = *reinterpret_cast<const float4*>(base + ((size_t)y0 * w + x0) * ld), b = *reinterpret_cast<const float4*>(base + ((size_t)y0 * w + x1) * ld);
    float4 d = *reinterpret_cast<const float4*>(base + ((size_t)y1 * w + x0) * ld), e = *reinterpret_cast<const float4*>(base + ((size_t)y1 * w + x1) * ld);
    if (relu) {
      a = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); b = make_float4(fmaxf(b.x, 0.f), fmaxf(b.y, 0.f), fmaxf(b.z, 0.f), fmaxf(b.w, 0.f));
      d = make_float4(fmaxf(d.x, 0.f), fmaxf(d.y, 0.f), fmaxf(d.z, 0.f), fmaxf(d.w, 0.f)); e = make_float4(fmaxf(e.x, 0.f), fmaxf(e.y, 0.f), fmaxf(e.z, 0.f), fmaxf(e.w, 0.f));
    }
    // ATen: hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11)
    const float v[4] = {hy * (hx * a.x + lx * b.x) + ly * (hx * d.x + lx * e.x), hy * (hx * a.y + lx * b.y) + ly * (hx * d.y + lx * e.y),
                        hy * (hx * a.z + lx * b.z) + ly * (hx * d.z + lx * e.z), hy * (hx * a.w + lx * b.w) + ly * (hx * d.w + lx * e.w)};
    __half hh[4], ll[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { hh[j] = __float2half_rn(v[j]); ll[j] = __float2half_rn(v[j] - __half2float(hh[j])); }
    const size_t o = ((size_t)oy * W + ox) * cpad + c;
    *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(hh);
    *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(ll);
  }
}

// dst[:, c_off : c_off + C] = src (P, C): the skip half of torch.cat([conv, skip], 1) (unets.py:55-56) lands in its channel slice
__global__ void __launch_bounds__(256) copy_slice_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t P, int C, int ld, int c_off) {
  const int C4 = C / 4;
  const int64_t n = P * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C4; const int c = (int)(i % C4) * 4;
    *reinterpret_cast<float4*>(dst + p * ld + c_off + c) = *reinterpret_cast<const float4*>(src + p * C + c);
  }
}

// ------------------------------------------------------------------------------------------------ 4x4 stride-2 (transposed) convolutions
// The UNet's encoder (Conv2DBlock: 4x4, stride 2, padding 1, unets.py:8-29, 201-207) and shared decoder (UpConv2DBlock with
// ConvTranspose2d 4x4, stride 2, padding 1, unets.py:32-58, 211-215) work on maps of 128^2 down to 2^2 pixels with up to 512 input channels:
// 2.5 GFLOP of weight streaming over a handful of pixels. One split-K gather-GEMM on the CUDA cores, fp32 like the reference:
//   out[m, n] = sum_k A[m, k] * Wk[cls][k, n],   k = tap * Ci + ci
// stride-2 convolution: m = output pixel (qy, qx), 16 taps (ky, kx), input pixel (2 qy + ky - 1, 2 qx + kx - 1), one class;
// transposed: four output-parity classes (ry, rx), m = (qy, qx) of the INPUT lattice, output pixel (2 qy + ry, 2 qx + rx), 4 taps (ty, tx)
//   reading input pixel (qy + ry - ty, qx + rx - tx) with kernel element (ky, kx) = (1 - ry + 2 ty, 1 - rx + 2 tx) (packed by the host).
// Block = 64 x 64 outputs over one K slice (grid.y = slices), thread = 4 x 4; partial sums go to a workspace that conv4_reduce_kernel
// folds in slice order (deterministic), adding the bias (eval BatchNorm folded) and the LeakyReLU(0.2) the reference applies in place.
struct Conv4Args {
  const float* src; const float* w; float* part;
  long long sp, sc;                  // source strides in floats: pixel, channel ((H,W,C) buffer: ld, 1; the caller's (C,H,W) input: 1, H*W)
  int c_off, Hin, Win, Ci, Co, K, kc, wq, Mcls, Mtot, Wout, transposed, in_relu;
};
constexpr int C4_BM = 64, C4_BN = 64, C4_BK = 16;

__global__ void __launch_bounds__(256) conv4_gemm_kernel(const Conv4Args a) {
  __shared__ __align__(16) float As[C4_BK][C4_BM + 4];
  __shared__ __align__(16) float Bs[C4_BK][C4_BN];
  const int t = threadIdx.x;
  const int n0 = blockIdx.x * C4_BN;
  const int mtiles = (a.Mcls + C4_BM - 1) / C4_BM;
  const int cls = blockIdx.z / mtiles, m0 = (blockIdx.z % mtiles) * C4_BM;
  const int ry = cls >> 1, rx = cls & 1;
  const int k_begin = blockIdx.y * a.kc, k_end = min(a.K, k_begin + a.kc);
  // loader roles: A element (row (t >> 4) + 16 i, k = t & 15); B element (k = (t >> 6) + 4 i, column t & 63)
  const int ak = t & 15, ar = t >> 4;
  int qy[4], qx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ar + 16 * i;
    qy[i] = m < a.Mcls ? m / a.wq : -(1 << 20);       // rows past the end gather nothing
    qx[i] = m < a.Mcls ? m % a.wq : -(1 << 20);
  }
  const float* wbase = a.w + (size_t)cls * a.K * a.Co;
  const int bn = t & 63, bk = t >> 6;
  const int ty4 = (t >> 4) * 4, tx4 = (t & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += C4_BK) {
    {
      const int k = k0 + ak;
      const bool kin = k < k_end;
      const int tap = kin ? k / a.Ci : 0, ci = k - tap * a.Ci;
      int dy, dx;
      if (a.transposed) { dy = ry - (tap >> 1); dx = rx - (tap & 1); }
      else { dy = (tap >> 2) - 1; dx = (tap & 3) - 1; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int iy = a.transposed ? qy[i] + dy : 2 * qy[i] + dy, ix = a.transposed ? qx[i] + dx : 2 * qx[i] + dx;
        float v = 0.f;
        if (kin && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) v = __ldg(a.src + ((size_t)iy * a.Win + ix) * a.sp + a.c_off + (size_t)ci * a.sc);
        As[ak][ar + 16 * i] = a.in_relu ? fmaxf(v, 0.f) : v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = k0 + bk + 4 * i;
        Bs[bk + 4 * i][bn] = (kk < k_end && n0 + bn < a.Co) ? __ldg(wbase + (size_t)kk * a.Co + n0 + bn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < C4_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx4]);
      const float am[4] = {av.x, av.y, av.z, av.w}, bm[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bm[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (n0 + tx4 >= a.Co) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty4 + i;
    if (m >= a.Mcls) continue;
    const int my = m / a.wq, mx = m % a.wq;
    const int opix = a.transposed ? (2 * my + ry) * a.Wout + 2 * mx + rx : m;
    *reinterpret_cast<float4*>(a.part + ((size_t)blockIdx.y * a.Mtot + opix) * a.Co + n0 + tx4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// dst[m, c_off + n] = act(sum_s part[s][m][n] + bias[n]); act = LeakyReLU(0.2) when `leaky`. A group of L lanes (a power of two <= 32) owns
// one float4 of the output: lane j adds slices j, j + L, ... in order, the group folds by xor shuffles -- a fixed order, so deterministic.
__global__ void __launch_bounds__(256) conv4_reduce_kernel(const float4* __restrict__ part, int S, int64_t mn4, int Co, const float* __restrict__ bias, int leaky,
                                                           float* __restrict__ dst, int ld, int c_off, int L) {
  const int lane = threadIdx.x & (L - 1), gpb = 256 / L;
  for (int64_t base = (int64_t)blockIdx.x * gpb; base < mn4; base += (int64_t)gridDim.x * gpb) {      // uniform per block: every lane reaches the shuffles
    const int64_t i = base + threadIdx.x / L;
    const bool ok = i < mn4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok)
      for (int s = lane; s < S; s += L) { const float4 p = part[(int64_t)s * mn4 + i]; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
    for (int o = L >> 1; o > 0; o >>= 1) {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
      v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
    }
    if (!ok || lane) continue;
    const int64_t e = i * 4; const int n = (int)(e % Co); const int64_t m = e / Co;
    if (bias) { v.x += bias[n]; v.y += bias[n + 1]; v.z += bias[n + 2]; v.w += bias[n + 3]; }
    if (leaky) { v.x = v.x > 0.f ? v.x : 0.2f * v.x; v.y = v.y > 0.f ? v.y : 0.2f * v.y; v.z = v.z > 0.f ? v.z : 0.2f * v.z; v.w = v.w > 0.f ? v.w : 0.2f * v.w; }
    *reinterpret_cast<float4*>(dst + m * ld + c_off + n) = v;
  }
}

// HGFilter.conv1: 7x7, stride 2, padding 3, 6 -> 64 channels, with bias (HGFilters.py:136, 180), fp32 on the CUDA cores (1.2 GMAC).
// in: (6, Hin, Win) f32 (the reference's NCHW input), out: (Hin/2, Win/2, 64) f32. Block = 16x16 output pixels, thread = 1 pixel,
// weights staged per 8-output-channel group.
__global__ void __launch_bounds__(256) stem7x7_kernel(const float* __restrict__ in, const float* __restrict__ w /*(64,6,7,7)*/, const float* __restrict__ bias,
                                                      float* __restrict__ out, int Hin, int Win) {
  __shared__ float s_in[6][37][38];          // 16*2 + 5 = 37 input rows / columns per block
  __shared__ float4 s_w[6 * 49][2];         // [tap][8 output channels]: two 16-byte broadcast loads per tap
  const int Ho = Hin / 2, Wo = Win / 2;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox0 = blockIdx.x * 16, oy0 = blockIdx.y * 16;
  for (int i = threadIdx.x; i < 6 * 37 * 37; i += 256) {
    const int ci = i / (37 * 37), r = (i / 37) % 37, cidx = i % 37;
    const int yy = 2 * oy0 - 3 + r, xx = 2 * ox0 - 3 + cidx;
    s_in[ci][r][cidx] = (yy >= 0 && yy < Hin && xx >= 0 && xx < Win) ? in[((size_t)ci * Hin + yy) * Win + xx] : 0.f;
  }
  const int ox = ox0 + tx, oy = oy0 + ty;
  for (int cg = 0; cg < 8; ++cg) {
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 294; i += 256) reinterpret_cast<float*>(s_w)[(i % 294) * 8 + i / 294] = w[(size_t)(cg * 8 + i / 294) * 294 + i % 294];
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ci = 0; ci < 6; ++ci)
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float v = s_in[ci][2 * ty + ky][2 * tx + kx];
          const int wi = ci * 49 + ky * 7 + kx;
          const float4 w0 = s_w[wi][0], w1 = s_w[wi][1];
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
    if (ox < Wo && oy < Ho) {
      float* d = out + ((size_t)oy * Wo + ox) * 64 + cg * 8;
#pragma unroll
      for (int j = 0; j < 8; j += 4)
        *reinterpret_cast<float4*>(d + j) = make_float4(acc[j] + bias[cg * 8 + j], acc[j + 1] + bias[cg * 8 + j + 1], acc[j + 2] + bias[cg * 8 + j + 2],
                                                        acc[j + 3] + bias[cg * 8 + j + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}
// fp16 (d2, d1, d0) tensor, d0 contiguous; box (b0, b1, b2); 128-byte swizzle; out-of-bounds elements read as zero
int make_tmap(avc_ctx* ctx, CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return avc_fail(ctx, AVC_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstr[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return avc_fail(ctx, AVC_ECUDA, "cuTensorMapEncodeTiled failed (%d) for dims (%llu,%llu,%llu) box (%u,%u,%u)", (int)r,
                                         (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1, b2);
  return AVC_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ the op program
// One op = 16 int32: [kind, a0..a14]. Buffers are numbered; f32 buffers and fp16 plane pairs live in one arena each.
enum { ENC_OP_STEM = 1, ENC_OP_GN = 2, ENC_OP_CONV = 3, ENC_OP_ADD = 4, ENC_OP_POOL = 5, ENC_OP_UPADD = 6, ENC_OP_INPUT = 7, ENC_OP_UPSPLIT = 8, ENC_OP_COPY = 9,
       ENC_OP_CONV4 = 10 };

struct EncConv { CUtensorMap a_hi, a_lo, b_hi, b_lo; ConvArgs args; size_t smem; int grid; };
struct EncConv4 { Conv4Args args; dim3 grid; int S; };

struct avc_encoder {
  avc_ctx* ctx = nullptr;
  std::vector<int32_t> ops;                 // n_ops x 16
  std::vector<float*> f32_bufs;             // views into d_f32
  std::vector<__half*> plane_hi, plane_lo;  // views into d_planes
  std::vector<int> plane_cpad;
  float* d_f32 = nullptr; __half* d_planes = nullptr; unsigned char* d_weights = nullptr; float* d_params = nullptr;
  double* d_partial = nullptr; unsigned int* d_ticket = nullptr; float* d_stats = nullptr;
  std::vector<EncConv> convs;               // one per ENC_OP_CONV, in op order
  std::vector<EncConv4> conv4s;             // one per ENC_OP_CONV4, in op order
  float* d_ws = nullptr;                    // split-K partial sums of the 4x4 convolutions
  int in_c = 0, in_h = 0, in_w = 0, out_buf = 0, out_c = 0, out_h = 0, out_w = 0;
  cudaGraphExec_t graph = nullptr; const float* graph_in = nullptr; float* graph_out = nullptr;
  int64_t launches_per_run = 0;
};

static void enc_free(avc_encoder* e) {
  if (!e) return;
  if (e->graph) cudaGraphExecDestroy(e->graph);
  if (e->d_f32) cudaFree(e->d_f32);
  if (e->d_planes) cudaFree(e->d_planes);
  if (e->d_weights) cudaFree(e->d_weights);
  if (e->d_params) cudaFree(e->d_params);
  if (e->d_partial) cudaFree(e->d_partial);
  if (e->d_ticket) cudaFree(e->d_ticket);
  if (e->d_stats) cudaFree(e->d_stats);
  if (e->d_ws) cudaFree(e->d_ws);
  delete e;
}

extern "C" void avc_encoder_destroy(avc_encoder* e) {
  if (e && e->ctx) cudaSetDevice(e->ctx->device);
  enc_free(e);
}

// program layout (int32 words), see avatarcap_b200/encoders.py build_hgfilter_program():
//   header: [magic 'AVCE', n_f32_bufs, n_planes, n_ops, in_c, in_h, in_w, out_buf, out_c, out_h, out_w, 0...] (16 words)
//   f32 buffer sizes in floats (n_f32_bufs words), plane descriptors (n_planes x 2 words: pixels, cpad), ops (n_ops x 16 words)
// weights: fp16 blob (hi / lo planes of every convolution, (C_out, taps, C_in_pad) each); params: f32 blob (biases, GroupNorm gamma / beta,
// the stem's weights)
extern "C" int avc_encoder_create(avc_ctx* ctx, const int32_t* program, int64_t n_words, const void* weights_f16, size_t weight_bytes,
                                  const float* params_f32, int64_t n_params, avc_encoder** out) {
  if (!ctx || !program || !weights_f16 || !params_f32 || !out) return avc_fail(ctx, AVC_EINVAL, "avc_encoder_create: NULL argument");
  *out = nullptr;
  if (!avc_tc_available(ctx)) return avc_fail(ctx, AVC_ESTATE, "the tensor-core encoder needs an sm_100 device");
  if (n_words < 16 || program[0] != 0x45435641) return avc_fail(ctx, AVC_EFORMAT, "encoder program: bad header");
  const int nb = program[1], np = program[2], nops = program[3];
  if (nb < 1 || np < 1 || nops < 1 || (int64_t)16 + nb + 2 * np + 16 * (int64_t)nops != n_words) return avc_fail(ctx, AVC_EFORMAT, "encoder program: bad sizes");
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  avc_encoder* e = new avc_encoder();
  e->ctx = ctx;
  e->in_c = program[4]; e->in_h = program[5]; e->in_w = program[6]; e->out_buf = program[7]; e->out_c = program[8]; e->out_h = program[9]; e->out_w = program[10];
  const int32_t* sizes = program + 16; const int32_t* pl = sizes + nb; const int32_t* ops = pl + 2 * np;
  e->ops.assign(ops, ops + 16 * (size_t)nops);
  auto fail = [&](int rc) { enc_free(e); return rc; };
#define ENC_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return fail(avc_check_cuda(ctx, _e, #call)); } while (0)
  size_t f32_total = 0;
  for (int i = 0; i < nb; ++i) { if (sizes[i] <= 0) return fail(avc_fail(ctx, AVC_EFORMAT, "encoder program: bad buffer size")); f32_total += ((size_t)sizes[i] + 63) & ~(size_t)63; }
  ENC_CUDA(cudaMalloc(&e->d_f32, f32_total * sizeof(float)));
  { size_t off = 0; for (int i = 0; i < nb; ++i) { e->f32_bufs.push_back(e->d_f32 + off); off += ((size_t)sizes[i] + 63) & ~(size_t)63; } }
  size_t pl_total = 0;
  for (int i = 0; i < np; ++i) pl_total += 2 * (((size_t)pl[2 * i] * pl[2 * i + 1] + 511) & ~(size_t)511);
  ENC_CUDA(cudaMalloc(&e->d_planes, pl_total * sizeof(__half)));
  ENC_CUDA(cudaMemset(e->d_planes, 0, pl_total * sizeof(__half)));          // channel padding (C < 64) stays zero for ever
  { size_t off = 0; for (int i = 0; i < np; ++i) { const size_t n = ((size_t)pl[2 * i] * pl[2 * i + 1] + 511) & ~(size_t)511;
      e->plane_hi.push_back(e->d_planes + off); e->plane_lo.push_back(e->d_planes + off + n); e->plane_cpad.push_back(pl[2 * i + 1]); off += 2 * n; } }
  ENC_CUDA(cudaMalloc(&e->d_weights, weight_bytes));
  ENC_CUDA(cudaMemcpy(e->d_weights, weights_f16, weight_bytes, cudaMemcpyHostToDevice));
  ENC_CUDA(cudaMalloc(&e->d_params, (size_t)n_params * sizeof(float)));
  ENC_CUDA(cudaMemcpy(e->d_params, params_f32, (size_t)n_params * sizeof(float), cudaMemcpyHostToDevice));
  ENC_CUDA(cudaMalloc(&e->d_partial, (size_t)1024 * 32 * 2 * sizeof(double)));
  ENC_CUDA(cudaMalloc(&e->d_ticket, 64)); ENC_CUDA(cudaMemset(e->d_ticket, 0, 64));
  ENC_CUDA(cudaMalloc(&e->d_stats, 64 * sizeof(float)));
  // tensor maps + launch geometry of every convolution; buffer / plane indices of every op
  size_t ws_floats = 0;
  for (int o = 0; o < nops; ++o) {
    const int32_t* op = ops + 16 * o;
    auto bad = [&](const char* what) { return fail(avc_fail(ctx, AVC_EFORMAT, "encoder program: op %d (%d): %s", o, op[0], what)); };
    auto buf_ok = [&](int b) { return b >= 0 && b < nb; };
    switch (op[0]) {
      case ENC_OP_STEM: if (!buf_ok(op[3])) return bad("buffer index"); break;
      case ENC_OP_GN: if (!buf_ok(op[1]) || op[9] >= np || (op[10] >= 0 && !buf_ok(op[10])) || (op[3] & 31) || op[3] > 256) return bad("buffer / plane index or channel count"); break;
      case ENC_OP_ADD: if (!buf_ok(op[1]) || !buf_ok(op[2]) || op[3] > sizes[op[1]] || op[3] > sizes[op[2]]) return bad("buffer index / size"); break;
      case ENC_OP_POOL: if (!buf_ok(op[1]) || !buf_ok(op[2])) return bad("buffer index"); break;
      case ENC_OP_UPADD: if (!buf_ok(op[1]) || !buf_ok(op[2]) || !buf_ok(op[3])) return bad("buffer index"); break;
      case ENC_OP_INPUT: if (!buf_ok(op[1]) || op[2] < 0 || op[3] < 0 || op[3] > sizes[op[1]]) return bad("buffer index / size"); break;
      case ENC_OP_UPSPLIT: if (!buf_ok(op[1]) || op[8] < 0 || op[8] >= np || (op[4] & 3) || (int64_t)4 * op[2] * op[3] != pl[2 * op[8]]) return bad("buffer / plane"); break;
      case ENC_OP_COPY: if (!buf_ok(op[1]) || !buf_ok(op[2]) || (op[4] & 3) || (int64_t)op[3] * op[5] > sizes[op[2]]) return bad("buffer index / size"); break;
      case ENC_OP_CONV: break;
      case ENC_OP_CONV4: {
        // [10, src_buf (-1: the caller's (C,H,W) input), dst_buf, Hin, Win, Ci, ld_src, c_off_src, Co, ld_dst, c_off_dst, w_off (params), bias_off (-1),
        //  flags: 1 = transposed (ConvTranspose2d 4x4 s2 p1: out 2Hin x 2Win; else Conv2d 4x4 s2 p1: out Hin/2 x Win/2), 2 = ReLU on the input, 4 = LeakyReLU(0.2) out]
        const int src = op[1], Hin = op[3], Win = op[4], Ci = op[5], lds = op[6], cos = op[7], Co = op[8], ldd = op[9], cod = op[10], flags = op[13];
        const bool tr = flags & 1;
        if (src < -1 || src >= nb || !buf_ok(op[2]) || src == op[2]) return bad("buffer index");
        if (Hin < 1 || Win < 1 || Ci < 1 || Co < 4 || (Co & 3) || (!tr && ((Hin | Win) & 1)) || (flags & ~7)) return bad("shape / flags");
        const int Ho = tr ? 2 * Hin : Hin / 2, Wo = tr ? 2 * Win : Win / 2;
        if (src >= 0 ? (cos < 0 || cos + Ci > lds || (int64_t)Hin * Win * lds > sizes[src]) : (Ci != e->in_c || Hin != e->in_h || Win != e->in_w)) return bad("input slice");
        if ((ldd & 3) || (cod & 3) || cod < 0 || cod + Co > ldd || (int64_t)Ho * Wo * ldd > sizes[op[2]]) return bad("output slice");
        const int64_t K = (int64_t)(tr ? 4 : 16) * Ci, nw = (tr ? 4 : 1) * K * Co;
        if ((K & 15) || op[11] < 0 || op[11] + nw > n_params || (op[12] >= 0 && op[12] + Co > n_params)) return bad("weight / bias offset");
        EncConv4 c;
        Conv4Args& a = c.args;
        a.src = src >= 0 ? e->f32_bufs[src] : nullptr; a.w = e->d_params + op[11]; a.part = nullptr;
        a.sp = src >= 0 ? lds : 1; a.sc = src >= 0 ? 1 : (long long)Hin * Win; a.c_off = src >= 0 ? cos : 0;
        a.Hin = Hin; a.Win = Win; a.Ci = Ci; a.Co = Co; a.K = (int)K; a.wq = tr ? Win : Wo; a.Mcls = tr ? Hin * Win : Ho * Wo; a.Mtot = Ho * Wo; a.Wout = Wo;
        a.transposed = tr; a.in_relu = (flags >> 1) & 1;
        const int tiles = ((Co + C4_BN - 1) / C4_BN) * ((a.Mcls + C4_BM - 1) / C4_BM) * (tr ? 4 : 1);
        // K slices: one wave of 4 resident blocks per SM (64 registers x 256 threads), at least 64 k per slice
        int S = (4 * ctx->sm_count) / tiles;
        if (S > (int)(K / 64)) S = (int)(K / 64);
        if (S < 1) S = 1;
        a.kc = (int)(((K + S - 1) / S + C4_BK - 1) / C4_BK) * C4_BK;
        S = (int)((K + a.kc - 1) / a.kc);
        c.S = S; c.grid = dim3((Co + C4_BN - 1) / C4_BN, S, ((a.Mcls + C4_BM - 1) / C4_BM) * (tr ? 4 : 1));
        const size_t need = (size_t)S * a.Mtot * Co;
        if (need > ws_floats) ws_floats = need;
        e->conv4s.push_back(c);
        break;
      }
      default: return bad("unknown op");
    }
    if (op[0] != ENC_OP_CONV) continue;
    // [3, plane, w_off_bytes, out_buf, H, W, cin_pad, N, taps, c_off, ldc, accumulate, bias_off(-1), weight scale exponent s]
    const int plane = op[1], H = op[4], W = op[5], cin = op[6], N = op[7], taps = op[8];
    if (plane < 0 || plane >= np || op[3] < 0 || op[3] >= nb) return bad("buffer index");
    if (cin % 64 || cin != e->plane_cpad[plane] || N % 32 || N < 32 || N > 256 || (taps != 1 && taps != 9)) return bad("channels / taps");
    int bw = W >= 128 ? 128 : W, bh = 128 / bw;
    if (W % bw || H % bh || (bw & (bw - 1))) return bad("image extent");
    const size_t wplane = (size_t)N * taps * cin * sizeof(__half);
    if ((size_t)op[2] + 2 * wplane > weight_bytes || (op[2] & 127)) return bad("weight offset");
    EncConv c;
    int rc = make_tmap(ctx, &c.a_hi, e->plane_hi[plane], (uint64_t)cin, (uint64_t)W, (uint64_t)H, 64, (uint32_t)bw, (uint32_t)bh);
    if (!rc) rc = make_tmap(ctx, &c.a_lo, e->plane_lo[plane], (uint64_t)cin, (uint64_t)W, (uint64_t)H, 64, (uint32_t)bw, (uint32_t)bh);
    if (!rc) rc = make_tmap(ctx, &c.b_hi, e->d_weights + op[2], (uint64_t)cin, (uint64_t)taps, (uint64_t)N, 64, 1, (uint32_t)N);
    if (!rc) rc = make_tmap(ctx, &c.b_lo, e->d_weights + op[2] + wplane, (uint64_t)cin, (uint64_t)taps, (uint64_t)N, 64, 1, (uint32_t)N);
    if (rc) return fail(rc);
    ConvArgs& a = c.args;
    a.W = W; a.bw = bw; a.bh = bh; a.tiles_x = W / bw; a.n_tiles = (W / bw) * (H / bh);
    a.kslabs = cin / 64; a.taps = taps; a.N = N;
    const int stage_bytes = 2 * CV_A_BYTES + 2 * N * 128;
    int stages = (200 * 1024) / stage_bytes; if (stages > CV_MAX_STAGES) stages = CV_MAX_STAGES; if (stages < 2) return bad("stage does not fit");
    a.stages = stages;
    a.out = e->f32_bufs[op[3]]; a.ldc = op[10]; a.c_off = op[9]; a.accumulate = op[11];
    a.bias = op[12] >= 0 ? e->d_params + op[12] : nullptr;
    if (op[13] < -40 || op[13] > 40) return bad("weight scale exponent");
    a.scale = ldexpf(1.f, -op[13]);
    if ((a.ldc & 3) || (a.c_off & 3) || a.c_off + N > a.ldc || (int64_t)H * W * a.ldc > sizes[op[3]]) return bad("output slice");
    c.smem = (size_t)stages * stage_bytes + sizeof(ConvBars) + 1024;
    c.grid = a.n_tiles < ctx->sm_count ? a.n_tiles : ctx->sm_count;
    e->convs.push_back(c);
  }
  if (ws_floats) {
    ENC_CUDA(cudaMalloc(&e->d_ws, ws_floats * sizeof(float)));
    for (auto& c : e->conv4s) c.args.part = e->d_ws;
  }
  cudaError_t ce = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
  if (ce != cudaSuccess) return fail(avc_check_cuda(ctx, ce, "cudaFuncSetAttribute(conv_tc_kernel)"));
#undef ENC_CUDA
  *out = e;
  return AVC_OK;
}

// enqueue the whole program on `st` (plain launches; avc_encoder_run wraps this in a CUDA graph)
static int enc_enqueue(avc_encoder* e, const float* in, float* outp, cudaStream_t st, int64_t* n_launch) {
  avc_ctx* ctx = e->ctx;
  const int nops = (int)(e->ops.size() / 16);
  int conv_i = 0, conv4_i = 0;
  int64_t nl = 0;
  auto blocks_for = [&](int64_t n) { int64_t b = (n + 255) / 256; const int64_t cap = (int64_t)ctx->sm_count * 8; return (int)(b < cap ? (b > 0 ? b : 1) : cap); };
  for (int o = 0; o < nops; ++o) {
    const int32_t* op = e->ops.data() + 16 * o;
    switch (op[0]) {
      case ENC_OP_STEM: {       // [1, w_off(params), bias_off, out_buf, Hin, Win]
        dim3 grid((op[5] / 2 + 15) / 16, (op[4] / 2 + 15) / 16);
        stem7x7_kernel<<<grid, 256, 0, st>>>(in, e->d_params + op[1], e->d_params + op[2], e->f32_bufs[op[3]], op[4], op[5]);
        ++nl; break;
      }
      case ENC_OP_GN: {         // [2, src_buf, P, C, ld, c_off, gamma_off(-1: no norm), beta_off, relu, plane(-1), dst32_buf(-1), ld32]
        const float* src = e->f32_bufs[op[1]];
        const int P = op[2], C = op[3], ld = op[4], c_off = op[5];
        const bool norm = op[6] >= 0;
        if (norm) {
          const int nblk = (P + GN_PX - 1) / GN_PX;
          if (nblk > 1024) return avc_fail(ctx, AVC_EFORMAT, "encoder program: GroupNorm over too many pixels");
          gn_stats_kernel<<<nblk, 256, 0, st>>>(src, P, C, ld, c_off, e->d_partial, e->d_ticket, e->d_stats);
          ++nl;
        }
        const int plane = op[9];
        gn_apply_kernel<<<blocks_for((int64_t)P * C / 4), 256, 0, st>>>(src, P, C, ld, c_off, norm ? e->d_stats : nullptr, norm ? e->d_params + op[6] : nullptr,
                                                                       norm ? e->d_params + op[7] : nullptr, op[8], plane >= 0 ? e->plane_hi[plane] : nullptr,
                                                                       plane >= 0 ? e->plane_lo[plane] : nullptr, plane >= 0 ? e->plane_cpad[plane] : 0,
                                                                       op[10] >= 0 ? e->f32_bufs[op[10]] : nullptr, op[11]);
        ++nl; break;
      }
      case ENC_OP_CONV: {
        EncConv& c = e->convs[conv_i++];
        ConvArgs a = c.args;
        if (op[3] == e->out_buf && outp) a.out = outp;            // the program's last convolution writes straight into the caller's buffer
        conv_tc_kernel<<<c.grid, CV_THREADS, c.smem, st>>>(c.a_hi, c.a_lo, c.b_hi, c.b_lo, a);
        ++nl; break;
      }
      case ENC_OP_ADD: {        // [4, dst_buf, src_buf, n_floats]
        add_kernel<<<blocks_for(op[3] / 4), 256, 0, st>>>(e->f32_bufs[op[1]], e->f32_bufs[op[2]], op[3] / 4);
        ++nl; break;
      }
      case ENC_OP_POOL: {       // [5, src_buf, dst_buf, H, W, C]
        avgpool2_kernel<<<blocks_for((int64_t)op[3] / 2 * (op[4] / 2) * op[5] / 4), 256, 0, st>>>(e->f32_bufs[op[1]], e->f32_bufs[op[2]], op[3], op[4], op[5]);
        ++nl; break;
      }
      case ENC_OP_UPADD: {      // [6, up1_buf, low_buf, dst_buf, h, w, C]  (low is h x w, up1 / dst are 2h x 2w)
        bicubic_up2_add_kernel<<<blocks_for((int64_t)4 * op[4] * op[5] * op[6] / 4), 256, 0, st>>>(e->f32_bufs[op[1]], e->f32_bufs[op[2]], e->f32_bufs[op[3]], op[4], op[5], op[6]);
        ++nl; break;
      }
      case ENC_OP_INPUT: {      // [7, dst_buf, offset (floats) into the caller's input, n_floats]: programs with several input tensors
        if (cudaMemcpyAsync(e->f32_bufs[op[1]], in + op[2], (size_t)op[3] * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
          return avc_check_cuda(ctx, cudaGetLastError(), "encoder input copy");
        break;
      }
      case ENC_OP_UPSPLIT: {    // [8, src_buf, h, w, C, ld, c_off, relu, plane]: bilinear x2 (align_corners=False) of [relu](src) -> fp16 planes
        const int plane = op[8];
        bilinear_up2_split_kernel<<<blocks_for((int64_t)4 * op[2] * op[3] * op[4] / 4), 256, 0, st>>>(e->f32_bufs[op[1]], op[2], op[3], op[4], op[5], op[6], op[7],
                                                                                                   e->plane_hi[plane], e->plane_lo[plane], e->plane_cpad[plane]);
        ++nl; break;
      }
      case ENC_OP_COPY: {       // [9, src_buf, dst_buf, P, C, ld_dst, c_off_dst]
        copy_slice_kernel<<<blocks_for((int64_t)op[3] * op[4] / 4), 256, 0, st>>>(e->f32_bufs[op[1]], e->f32_bufs[op[2]], (int64_t)op[3], op[4], op[5], op[6]);
        ++nl; break;
      }
      case ENC_OP_CONV4: {      // see avc_encoder_create
        EncConv4& c = e->conv4s[conv4_i++];
        Conv4Args a = c.args;
        if (op[1] < 0) a.src = in;
        conv4_gemm_kernel<<<c.grid, 256, 0, st>>>(a);
        const int64_t mn4 = (int64_t)a.Mtot * a.Co / 4;
        int L = 1; while (L < c.S && L < 32) L <<= 1;
        conv4_reduce_kernel<<<blocks_for(mn4 * L), 256, 0, st>>>(reinterpret_cast<const float4*>(a.part), c.S, mn4, a.Co, op[12] >= 0 ? e->d_params + op[12] : nullptr,
                                                                (op[13] >> 2) & 1, e->f32_bufs[op[2]], op[9], op[10], L);
        nl += 2; break;
      }
      default: return avc_fail(ctx, AVC_EFORMAT, "encoder program: unknown op %d", op[0]);
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return avc_check_cuda(ctx, le, "encoder launch");
  }
  *n_launch = nl;
  return AVC_OK;
}

// in: [dev] (in_c, in_h, in_w) f32 (the reference's NCHW input, batch 1); out: [dev] (out_h, out_w, out_c) f32 -- the (H, W, C) layout the
// gather kernels read (avc_set_feature_map_hwc). The program is captured into a CUDA graph on first use (and re-captured when the
// pointers change); use_graph = 0 launches the kernels one by one.
extern "C" int avc_encoder_run(avc_encoder* e, const float* in, float* out, int use_graph, void* stream) {
  if (!e || !in || !out) return e ? avc_fail(e->ctx, AVC_EINVAL, "avc_encoder_run: NULL argument") : AVC_EINVAL;
  avc_ctx* ctx = e->ctx;
  cudaStream_t st = (cudaStream_t)stream;
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t nl = 0;
  if (!use_graph) {
    int rc = enc_enqueue(e, in, out, st, &nl);
    if (rc) return rc;
    ctx->launches += nl;
    return AVC_OK;
  }
  if (!e->graph || e->graph_in != in || e->graph_out != out) {
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    cudaStream_t cs;
    AVC_CUDA(ctx, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    int rc = AVC_OK;
    if (ce == cudaSuccess) {
      rc = enc_enqueue(e, in, out, cs, &nl);
      ce = cudaStreamEndCapture(cs, &g);
    }
    if (rc == AVC_OK && ce == cudaSuccess) ce = cudaGraphInstantiate(&e->graph, g, 0);
    if (g) cudaGraphDestroy(g);
    cudaStreamDestroy(cs);
    if (rc) return rc;
    if (ce != cudaSuccess) { e->graph = nullptr; return avc_check_cuda(ctx, ce, "encoder graph capture"); }
    e->graph_in = in; e->graph_out = out; e->launches_per_run = nl;
  }
  AVC_CUDA(ctx, cudaGraphLaunch(e->graph, st));
  ctx->launches += e->launches_per_run;
  return AVC_OK;
}

// debugging / tests: copy f32 buffer `buf` of the program (contents after the last run) to dst [dev]
extern "C" int avc_encoder_read_buffer(avc_encoder* e, int buf, float* dst, int64_t n_floats, void* stream) {
  if (!e || !dst) return e ? avc_fail(e->ctx, AVC_EINVAL, "avc_encoder_read_buffer: NULL argument") : AVC_EINVAL;
  if (buf < 0 || buf >= (int)e->f32_bufs.size() || n_floats < 0) return avc_fail(e->ctx, AVC_EINVAL, "avc_encoder_read_buffer: bad buffer %d", buf);
  AVC_CUDA(e->ctx, cudaMemcpyAsync(dst, e->f32_bufs[buf], (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return AVC_OK;
}

extern "C" int avc_encoder_shape(const avc_encoder* e, int in_chw[3], int out_hwc[3]) {
  if (!e) return AVC_EINVAL;
  if (in_chw) { in_chw[0] = e->in_c; in_chw[1] = e->in_h; in_chw[2] = e->in_w; }
  if (out_hwc) { out_hwc[0] = e->out_h; out_hwc[1] = e->out_w; out_hwc[2] = e->out_c; }
  return AVC_OK;
}
