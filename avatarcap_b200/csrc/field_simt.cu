// fp32 CUDA-core implementation of the fused per-point networks (AVC_IMPL_SIMT).
//
// One CTA evaluates a tile of TP points through the WHOLE network with activations resident in shared memory
// (the reference round-trips every 256-channel activation through HBM: network/mlp.py:56-72, 101-112).
// It is the numerically closest path to the reference's fp32 math and the on-device cross-check for the
// tcgen05 kernel (field_tc2.cu).
//
// Reference call sites restated here:
//   WarpingField.query      network/arch_avatar.py:113-140  (bilinear gather :133, OffsetDecoder mlp.py:101-112, out :138)
//   DoubleTNet.forward      network/arch_avatar.py:65-83    (PE utils/net_util.py:28-37, MLP.forward mlp.py:56-72)
//   OccupancyNet.query      network/arch_avatar.py:356-381
//   ReconNetwork.infer      network/arch_recon.py:55-76
#include "common.cuh"

namespace {

constexpr int TP = 64;          // points per tile
constexpr int NT = 256;         // threads per CTA
constexpr int BUF_A = 512 * TP; // floats
constexpr int BUF_B = 256 * TP;
constexpr int BUF_S = 80 * TP;
constexpr size_t SMEM_BYTES = (size_t)(BUF_A + BUF_B + BUF_S) * sizeof(float);

__device__ __forceinline__ float act_fn(float v, int act) {
  switch (act) {
    case AVC_ACT_RELU: return fmaxf(v, 0.f);
    case AVC_ACT_LRELU: return v > 0.f ? v : v * 0.02f;                     // nn.LeakyReLU(0.02) mlp.py:11
    case AVC_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));           // nn.Softplus(beta=1, threshold=20)
    case AVC_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// dst[c][p] = act(scale[c] * sum_k W^T[k][c] * src[k][p] + bias[c]);  src = concat(src0[k0], src1[k1]) along k.
// Thread (tx = tid & 15, ty = tid >> 4) owns points 4*tx..4*tx+3 and channels nb + 16*ty .. +15.
__device__ void dense_layer(const float* __restrict__ f32, const AvcLayerDesc& L, const float* src0, const float* src1,
                            float* dst) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n_out = L.n;
  const float* __restrict__ Wt = f32 + L.wt_off;
  const float* __restrict__ sc = f32 + L.sb_off;
  const float* __restrict__ bi = sc + n_out;
  for (int nb = 0; nb < n_out; nb += 256) {
    const int c0 = nb + ty * 16;
    if (c0 < n_out) {
      float acc[16][4];
#pragma unroll
      for (int j = 0; j < 16; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
      const float* wrow = Wt + c0;
      for (int seg = 0; seg < 2; ++seg) {
        const float* src = seg == 0 ? src0 : src1;
        const int kk = seg == 0 ? L.k0 : L.k1;
#pragma unroll 2
        for (int k = 0; k < kk; ++k) {
          const float4 a = *reinterpret_cast<const float4*>(src + k * TP + 4 * tx);
          float w[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(wrow) + q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            acc[j][0] = fmaf(w[j], a.x, acc[j][0]);
            acc[j][1] = fmaf(w[j], a.y, acc[j][1]);
            acc[j][2] = fmaf(w[j], a.z, acc[j][2]);
            acc[j][3] = fmaf(w[j], a.w, acc[j][3]);
          }
          wrow += n_out;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = __ldg(sc + c0 + j), b = __ldg(bi + c0 + j);
        float4 o;
        o.x = act_fn(fmaf(acc[j][0], s, b), L.act);
        o.y = act_fn(fmaf(acc[j][1], s, b), L.act);
        o.z = act_fn(fmaf(acc[j][2], s, b), L.act);
        o.w = act_fn(fmaf(acc[j][3], s, b), L.act);
        *reinterpret_cast<float4*>(dst + (c0 + j) * TP + 4 * tx) = o;
      }
    }
  }
  __syncthreads();
}

// Small head: n <= 4 outputs, W stored [n][k] row-major. Result (pre-activation, scale*acc+bias) for point pt, output o.
__device__ __forceinline__ float head_dot(const float* __restrict__ f32, const AvcLayerDesc& L, const float* src, int pt, int o) {
  const float* __restrict__ w = f32 + L.wt_off + o * L.k0;
  float acc = 0.f;
  for (int k = 0; k < L.k0; ++k) acc = fmaf(__ldg(w + k), src[k * TP + pt], acc);
  return fmaf(acc, __ldg(f32 + L.sb_off + o), __ldg(f32 + L.sb_off + L.n + o));
}

// Bilinear tap setup, F.grid_sample(mode='bilinear', padding_mode='border', align_corners=True)
// (ATen GridSampler.h: grid_sampler_compute_source_index + clip_coordinates), arch_avatar.py:133 / arch_recon.py:68.
struct Taps { int i00, i01, i10, i11; float w00, w01, w10, w11; };
__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float x1 = x0 + 1.f, y1 = y0 + 1.f;
  Taps t;
  t.w00 = (x1 - ix) * (y1 - iy);   // nw
  t.w01 = (ix - x0) * (y1 - iy);   // ne
  t.w10 = (x1 - ix) * (iy - y0);   // sw
  t.w11 = (ix - x0) * (iy - y0);   // se
  const int xi0 = (int)x0, yi0 = (int)y0;
  int xi1 = xi0 + 1, yi1 = yi0 + 1;
  if (xi1 > W - 1) { xi1 = W - 1; t.w01 = 0.f; t.w11 = 0.f; }   // out-of-range taps contribute 0
  if (yi1 > H - 1) { yi1 = H - 1; t.w10 = 0.f; t.w11 = 0.f; }
  t.i00 = yi0 * W + xi0; t.i01 = yi0 * W + xi1; t.i10 = yi1 * W + xi0; t.i11 = yi1 * W + xi1;
  return t;
}

// gather C channels (multiple of 16) of an (H,W,C) map into dst rows [row0 .. row0+C)
__device__ void gather_features(const float* __restrict__ hwc, int C, int H, int W, const float* sp /*smem xyz of tile*/,
                                float cx, float cy, float* dst, int row0) {
  const int tid = threadIdx.x, pt = tid & (TP - 1), part = tid >> 6;   // 4 parts
  const float px = sp[pt] - cx, py = sp[TP + pt] - cy;
  const Taps t = make_taps(px, -py, H, W);
  const int cper = C / 4;
  for (int c = part * cper; c < (part + 1) * cper; c += 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i00 * C + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i01 * C + c));
    const float4 d = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i10 * C + c));
    const float4 e = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i11 * C + c));
    // accumulation order nw, ne, sw, se as ATen does
    dst[(row0 + c + 0) * TP + pt] = ((a.x * t.w00 + b.x * t.w01) + d.x * t.w10) + e.x * t.w11;
    dst[(row0 + c + 1) * TP + pt] = ((a.y * t.w00 + b.y * t.w01) + d.y * t.w10) + e.y * t.w11;
    dst[(row0 + c + 2) * TP + pt] = ((a.z * t.w00 + b.z * t.w01) + d.z * t.w10) + e.z * t.w11;
    dst[(row0 + c + 3) * TP + pt] = ((a.w * t.w00 + b.w * t.w01) + d.w * t.w10) + e.w * t.w11;
  }
}

struct AvatarArgs {
  const float* pts; int64_t n;
  float cx, cy, cz;
  const float* map; int mC, mH, mW;
  float* out_occ; float* out_off; float* out_rgb; float* out_alpha;
  int if_type, mode;
};

__global__ void __launch_bounds__(NT, 1)
avatar_simt_kernel(const float* __restrict__ f32, const AvcBlobHeader* __restrict__ hdr_g, AvatarArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem; float* bufB = smem + BUF_A; float* bufS = smem + BUF_A + BUF_B;
  __shared__ AvcLayerDesc L[20];
  __shared__ float s_p[3 * TP];   // tile points (x row, y row, z row)
  __shared__ float s_q[3 * TP];   // p + offset
  const int tid = threadIdx.x;
  for (int i = tid; i < 20 * (int)(sizeof(AvcLayerDesc) / 4); i += NT)
    reinterpret_cast<int32_t*>(L)[i] = reinterpret_cast<const int32_t*>(hdr_g->layers)[i];
  const int64_t n_tiles = (a.n + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t base = tile * TP;
    if (tid < 3 * TP) {
      const int pt = tid / 3, d = tid % 3;
      const int64_t g = base + pt;
      s_p[d * TP + pt] = g < a.n ? a.pts[g * 3 + d] : 0.f;
    }
    __syncthreads();
    const int pt = tid & (TP - 1), part = tid >> 6;
    if (a.mode != AVC_MODE_TEMPLATE_ONLY) {
      // ---- WarpingField.query: h0 = [p(3), bilinear(pose_feat_map)(64)]  arch_avatar.py:121-136 (PE freq 0 = identity)
      if (tid < 3 * TP) bufS[(tid / TP) * TP + (tid % TP)] = s_p[tid];
      gather_features(a.map, a.mC, a.mH, a.mW, s_p, a.cx, a.cy, bufS, 3);
      __syncthreads();
      dense_layer(f32, L[0], bufS, nullptr, bufA);      // conv1+bn1+softplus   mlp.py:102
      dense_layer(f32, L[1], bufA, nullptr, bufB);      // :103
      dense_layer(f32, L[2], bufB, nullptr, bufA);      // :104
      dense_layer(f32, L[3], bufA, nullptr, bufB);      // :105  x4 in B
      dense_layer(f32, L[4], bufS, bufB, bufA);         // :106  cat([x, x4])
      dense_layer(f32, L[5], bufA, nullptr, bufB);      // :109
      dense_layer(f32, L[6], bufB, nullptr, bufA);      // :110  x7 in A
      if (part < 3) {                                   // out_layer_coord_affine  arch_avatar.py:138
        const float off = head_dot(f32, L[7], bufA, pt, part);
        s_q[part * TP + pt] = s_p[part * TP + pt] + off;   // cano_pts_chunk + offset_chunk  :372
        const int64_t g = base + pt;
        if (a.out_off && g < a.n) a.out_off[g * 3 + part] = off;
      }
      __syncthreads();
    } else {
      if (tid < 3 * TP) s_q[tid] = s_p[tid];
      __syncthreads();
    }
    if (a.mode == AVC_MODE_WARP_ONLY) continue;
    // ---- DoubleTNet.forward: positional encoding  net_util.py:28-37: [x, sin(2^k x), cos(2^k x)]_k
    {
      if (tid < 3 * TP) bufS[tid] = s_q[tid];
      for (int k = part; k < 10; k += 4) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float v = s_q[d * TP + pt] * f;
          bufS[(3 + 6 * k + d) * TP + pt] = sinf(v);
          bufS[(3 + 6 * k + 3 + d) * TP + pt] = cosf(v);
        }
      }
      __syncthreads();
    }
    dense_layer(f32, L[8], bufS, nullptr, bufA);        // shared fc0
    dense_layer(f32, L[9], bufA, nullptr, bufB);
    dense_layer(f32, L[10], bufB, nullptr, bufA);
    dense_layer(f32, L[11], bufA, nullptr, bufB);       // s4 in B
    dense_layer(f32, L[12], bufB, bufS, bufA);          // fc4: cat([x, tmpx])  mlp.py:60-61
    dense_layer(f32, L[13], bufA, nullptr, bufB);
    dense_layer(f32, L[14], bufB, nullptr, bufA);       // fc6 linear -> shared feature in A
    dense_layer(f32, L[15], bufA, nullptr, bufB);       // geo fc0 256->128 leaky relu
    if (part < 2) {                                     // geo fc1 128->2
      const float v = head_dot(f32, L[16], bufB, pt, part);
      const int64_t g = base + pt;
      if (g < a.n) {
        if (part == 0) {
          if (a.out_occ) a.out_occ[g] = a.if_type == AVC_IF_OCCUPANCY ? 1.f / (1.f + expf(-v)) : v;   // arch_avatar.py:77-80
        } else if (a.out_alpha) a.out_alpha[g] = fmaxf(v, 0.f);                                    // :76
      }
    }
    if (a.out_rgb) {
      __syncthreads();
      dense_layer(f32, L[17], bufA, nullptr, bufB);     // clr fc0 256->256 relu
      dense_layer(f32, L[18], bufB, nullptr, bufA);     // clr fc1 256->128 relu
      if (part < 3) {
        const float v = head_dot(f32, L[19], bufA, pt, part);
        const int64_t g = base + pt;
        if (g < a.n) a.out_rgb[g * 3 + part] = 1.f / (1.f + expf(-v));    // sigmoid  :75
      }
    }
  }
}

struct ReconArgs {
  const float* pts; int64_t n;
  float cx, cy, cz;
  const float* map; int mC, mH, mW;
  float* out_ov;
};

__global__ void __launch_bounds__(NT, 1)
recon_simt_kernel(const float* __restrict__ f32, const AvcBlobHeader* __restrict__ hdr_g, ReconArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem; float* bufB = smem + BUF_A; float* bufS = smem + BUF_A + BUF_B;
  __shared__ AvcLayerDesc L[4];
  __shared__ float s_p[3 * TP];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * (int)(sizeof(AvcLayerDesc) / 4); i += NT)
    reinterpret_cast<int32_t*>(L)[i] = reinterpret_cast<const int32_t*>(hdr_g->layers)[i];
  const int64_t n_tiles = (a.n + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t base = tile * TP;
    if (tid < 3 * TP) {
      const int pt = tid / 3, d = tid % 3;
      const int64_t g = base + pt;
      s_p[d * TP + pt] = g < a.n ? a.pts[g * 3 + d] : 0.f;
    }
    __syncthreads();
    // h0 = [bilinear(img_feat_map)(32), z - cz]   arch_recon.py:62-70
    gather_features(a.map, a.mC, a.mH, a.mW, s_p, a.cx, a.cy, bufS, 0);
    if (tid < TP) bufS[32 * TP + tid] = s_p[2 * TP + tid] - a.cz;
    __syncthreads();
    dense_layer(f32, L[0], bufS, nullptr, bufA);        // 33 -> 512
    dense_layer(f32, L[1], bufA, bufS, bufB);           // cat([y1, h0]) 545 -> 256   mlp.py:60-61
    dense_layer(f32, L[2], bufB, bufS, bufA);           // cat([y2, h0]) 289 -> 128
    if (tid < TP) {
      const float v = head_dot(f32, L[3], bufA, tid, 0);
      const int64_t g = base + tid;
      if (g < a.n) a.out_ov[g] = 1.f / (1.f + expf(-v));   // last_op sigmoid  mlp.py:49-50,64-65
    }
  }
}

}  // namespace

int avc_simt_eval_avatar(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_occ, float* out_off,
                         float* out_rgb, float* out_alpha, int if_type, int mode, cudaStream_t st) {
  if (n == 0) return AVC_OK;
  AvatarArgs a;
  a.pts = pts; a.n = n; a.cx = center[0]; a.cy = center[1]; a.cz = center[2];
  a.map = ctx->maps[AVC_MAP_POSE].d_hwc; a.mC = ctx->maps[AVC_MAP_POSE].C; a.mH = ctx->maps[AVC_MAP_POSE].H;
  a.mW = ctx->maps[AVC_MAP_POSE].W;
  a.out_occ = out_occ; a.out_off = out_off; a.out_rgb = out_rgb; a.out_alpha = out_alpha; a.if_type = if_type; a.mode = mode;
  static bool attr_set = false;
  if (!attr_set) {
    AVC_CUDA(ctx, cudaFuncSetAttribute(avatar_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = (n + TP - 1) / TP;
  const int grid = (int)(tiles < (int64_t)ctx->sm_count ? tiles : ctx->sm_count);
  avatar_simt_kernel<<<grid, NT, SMEM_BYTES, st>>>(ctx->avatar.d_f32, reinterpret_cast<const AvcBlobHeader*>(ctx->avatar.d_blob), a);
  AVC_LAUNCH_CHECK(ctx, "avatar_simt_kernel");
  return AVC_OK;
}

int avc_simt_eval_recon(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_ov, cudaStream_t st) {
  if (n == 0) return AVC_OK;
  ReconArgs a;
  a.pts = pts; a.n = n; a.cx = center[0]; a.cy = center[1]; a.cz = center[2];
  a.map = ctx->maps[AVC_MAP_IMAGE].d_hwc; a.mC = ctx->maps[AVC_MAP_IMAGE].C; a.mH = ctx->maps[AVC_MAP_IMAGE].H;
  a.mW = ctx->maps[AVC_MAP_IMAGE].W;
  a.out_ov = out_ov;
  static bool attr_set = false;
  if (!attr_set) {
    AVC_CUDA(ctx, cudaFuncSetAttribute(recon_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int64_t tiles = (n + TP - 1) / TP;
  const int grid = (int)(tiles < (int64_t)ctx->sm_count ? tiles : ctx->sm_count);
  recon_simt_kernel<<<grid, NT, SMEM_BYTES, st>>>(ctx->recon.d_f32, reinterpret_cast<const AvcBlobHeader*>(ctx->recon.d_blob), a);
  AVC_LAUNCH_CHECK(ctx, "recon_simt_kernel");
  return AVC_OK;
}
