// C-ABI entry points: context, weights, feature maps, field-evaluation dispatch, host-buffer (end-to-end) variants.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"
#include <cstdlib>
#include <cstdio>

static std::string g_create_error;

int avc_fail(avc_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

int avc_check_cuda(avc_ctx* ctx, cudaError_t e, const char* what) {
  return avc_fail(ctx, AVC_ECUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

int avc_ensure_scratch(avc_ctx* ctx, size_t bytes) {
  ctx->mc_last.valid = false;      // whoever asks for the scratch is about to overwrite it
  if (bytes <= ctx->scratch_cap) return AVC_OK;
  if (ctx->d_scratch) { cudaFree(ctx->d_scratch); ctx->d_scratch = nullptr; ctx->scratch_cap = 0; }
  const size_t cap = bytes + (bytes >> 3) + 4096;
  AVC_CUDA(ctx, cudaMalloc(&ctx->d_scratch, cap));
  ctx->scratch_cap = cap;
  return AVC_OK;
}

extern "C" int avc_abi_version(void) { return AVC_ABI_VERSION; }

extern "C" int avc_ctx_create(int device, avc_ctx** out) {
  if (!out) return avc_fail(nullptr, AVC_EINVAL, "avc_ctx_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return avc_fail(nullptr, AVC_ECUDA, "avc_ctx_create: no CUDA device available (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= count) return avc_fail(nullptr, AVC_EINVAL, "avc_ctx_create: device %d out of range (0..%d)", device, count - 1);
  avc_ctx* ctx = new avc_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete ctx;
    return avc_fail(nullptr, AVC_ECUDA, "avc_ctx_create: %s", cudaGetErrorString(e));
  }
  ctx->sm_count = prop.multiProcessorCount; ctx->cc_major = prop.major; ctx->cc_minor = prop.minor;
  if ((e = cudaMallocHost(&ctx->h_counts, 8 * sizeof(int64_t))) != cudaSuccess) {
    delete ctx;
    return avc_fail(nullptr, AVC_ECUDA, "avc_ctx_create: cudaMallocHost: %s", cudaGetErrorString(e));
  }
  cudaStreamCreateWithFlags(&ctx->s_copy_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_copy_out, cudaStreamNonBlocking);
  *out = ctx;
  return AVC_OK;
}

static void free_weights(AvcWeights& w) {
  if (w.d_blob) cudaFree(w.d_blob);
  w = AvcWeights();
}

extern "C" void avc_ctx_destroy(avc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int k = 0; k < 2; ++k) for (int s = 0; s < AVC_WEIGHT_SLOTS; ++s) free_weights(ctx->slots[k][s]);
  for (int i = 0; i < 2; ++i) { if (ctx->maps[i].d_hwc) cudaFree(ctx->maps[i].d_hwc); if (ctx->maps[i].ready) cudaEventDestroy(ctx->maps[i].ready); }
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_scratch2) cudaFree(ctx->d_scratch2);
  if (ctx->d_grid) cudaFree(ctx->d_grid);
  if (ctx->d_gridpts) cudaFree(ctx->d_gridpts);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  if (ctx->s_copy_in) cudaStreamDestroy(ctx->s_copy_in);
  if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
  if (ctx->s_copy_out) cudaStreamDestroy(ctx->s_copy_out);
  delete ctx;
}

extern "C" const char* avc_last_error(const avc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int avc_has_tensor_core_path(const avc_ctx* ctx) { return ctx ? avc_tc_available(ctx) : 0; }
extern "C" int64_t avc_launch_count(const avc_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void avc_reset_launch_count(avc_ctx* ctx) { if (ctx) ctx->launches = 0; }
extern "C" int avc_debug_set_trace(avc_ctx* ctx, void* dev_buf, int flags) { if (!ctx) return AVC_EINVAL; ctx->d_trace = dev_buf; ctx->dbg_flags = flags; return AVC_OK; }

// -----------------------------------------------------------------------------------------------------------------
static int load_weights(avc_ctx* ctx, AvcWeights& w, const void* blob, size_t nbytes, uint32_t kind, uint32_t n_layers) {
  if (!ctx || !blob) return avc_fail(ctx, AVC_EINVAL, "load weights: NULL argument");
  if (nbytes < sizeof(AvcBlobHeader)) return avc_fail(ctx, AVC_EFORMAT, "weight blob too small (%zu bytes)", nbytes);
  AvcBlobHeader h; memcpy(&h, blob, sizeof(h));
  if (h.magic != AVC_MAGIC) return avc_fail(ctx, AVC_EFORMAT, "weight blob: bad magic 0x%08x", h.magic);
  if (h.version != AVC_BLOB_VERSION) return avc_fail(ctx, AVC_EFORMAT, "weight blob: version %u, library expects %u", h.version, AVC_BLOB_VERSION);
  if (h.kind != kind || h.n_layers != n_layers) return avc_fail(ctx, AVC_EFORMAT, "weight blob: kind %u with %u layers, expected kind %u with %u", h.kind, h.n_layers, kind, n_layers);
  if (h.f32_off + h.f32_bytes > nbytes || h.f16_off + h.f16_bytes > nbytes || (h.f32_off & 15) || (h.f16_off & 127))
    return avc_fail(ctx, AVC_EFORMAT, "weight blob: section out of range or misaligned");
  for (uint32_t l = 0; l < n_layers; ++l) {
    const AvcLayerDesc& L = h.layers[l];
    const int64_t ktot = (int64_t)L.k0 + L.k1;
    if (L.k0 <= 0 || L.k1 < 0 || L.n <= 0 || L.wt_off < 0 || L.sb_off < 0 ||
        ((int64_t)L.wt_off + ktot * L.n) * 4 > (int64_t)h.f32_bytes || ((int64_t)L.sb_off + 2 * L.n) * 4 > (int64_t)h.f32_bytes)
      return avc_fail(ctx, AVC_EFORMAT, "weight blob: layer %u descriptor out of range", l);
  }
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  free_weights(w);
  AVC_CUDA(ctx, cudaMalloc(&w.d_blob, nbytes));
  AVC_CUDA(ctx, cudaMemcpy(w.d_blob, blob, nbytes, cudaMemcpyHostToDevice));
  w.hdr = h; w.d_f32 = reinterpret_cast<const float*>(w.d_blob + h.f32_off); w.d_f16 = w.d_blob + h.f16_off; w.loaded = true;
  return AVC_OK;
}

extern "C" int avc_select_weights(avc_ctx* ctx, int kind, int slot) {
  if (!ctx) return AVC_EINVAL;
  if ((kind != AVC_KIND_AVATAR && kind != AVC_KIND_RECON) || slot < 0 || slot >= AVC_WEIGHT_SLOTS)
    return avc_fail(ctx, AVC_EINVAL, "avc_select_weights: bad kind %d / slot %d", kind, slot);
  if (!ctx->slots[kind][slot].loaded) return avc_fail(ctx, AVC_ESTATE, "avc_select_weights: slot %d holds no weights", slot);
  (kind == AVC_KIND_AVATAR ? ctx->avatar : ctx->recon) = ctx->slots[kind][slot];      // pointer swap: the slot keeps owning the blob
  ctx->slot_sel[kind] = slot;
  return AVC_OK;
}

extern "C" int avc_load_weights_slot(avc_ctx* ctx, int kind, int slot, const void* blob, size_t nbytes) {
  if (!ctx) return AVC_EINVAL;
  if ((kind != AVC_KIND_AVATAR && kind != AVC_KIND_RECON) || slot < 0 || slot >= AVC_WEIGHT_SLOTS)
    return avc_fail(ctx, AVC_EINVAL, "avc_load_weights_slot: bad kind %d / slot %d", kind, slot);
  // kernels of earlier calls may still read the blob this slot is about to free
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->slots[kind][slot].loaded) AVC_CUDA(ctx, cudaDeviceSynchronize());
  int rc = load_weights(ctx, ctx->slots[kind][slot], blob, nbytes, (uint32_t)kind, kind == AVC_KIND_AVATAR ? 20 : 4);
  if (rc) { if (ctx->slot_sel[kind] == slot) (kind == AVC_KIND_AVATAR ? ctx->avatar : ctx->recon) = AvcWeights(); return rc; }
  return avc_select_weights(ctx, kind, slot);
}

extern "C" int avc_load_avatar_weights(avc_ctx* ctx, const void* blob, size_t nbytes) {
  return avc_load_weights_slot(ctx, AVC_KIND_AVATAR, ctx ? ctx->slot_sel[AVC_KIND_AVATAR] : 0, blob, nbytes);
}
extern "C" int avc_load_recon_weights(avc_ctx* ctx, const void* blob, size_t nbytes) {
  return avc_load_weights_slot(ctx, AVC_KIND_RECON, ctx ? ctx->slot_sel[AVC_KIND_RECON] : 0, blob, nbytes);
}

// (C,H,W) -> (H,W,C): one bilinear tap becomes one contiguous C*4-byte read
__global__ void chw_to_hwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && p < HW) ? in[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[(size_t)p * C + c] = tile[threadIdx.x][r];
  }
}

static int set_feature_map(avc_ctx* ctx, int which, const float* src, int C, int H, int W, bool hwc, cudaStream_t st) {
  if (!ctx || !src) return avc_fail(ctx, AVC_EINVAL, "avc_set_feature_map: NULL argument");
  if (which != AVC_MAP_POSE && which != AVC_MAP_IMAGE) return avc_fail(ctx, AVC_EINVAL, "avc_set_feature_map: bad slot %d", which);
  const int want = which == AVC_MAP_POSE ? 64 : 32;
  if (C != want || H < 1 || W < 1) return avc_fail(ctx, AVC_EINVAL, "avc_set_feature_map: slot %d needs C=%d (got %d), H,W >= 1", which, want, C);
  AvcMap& m = ctx->maps[which];
  const size_t need = (size_t)C * H * W * sizeof(float);
  if (need > m.cap) {
    if (m.d_hwc) cudaFree(m.d_hwc);
    m.d_hwc = nullptr; m.cap = 0;
    AVC_CUDA(ctx, cudaMalloc(&m.d_hwc, need));
    m.cap = need;
  }
  m.C = C; m.H = H; m.W = W;
  if (!m.ready) AVC_CUDA(ctx, cudaEventCreateWithFlags(&m.ready, cudaEventDisableTiming));
  if (hwc) {
    AVC_CUDA(ctx, cudaMemcpyAsync(m.d_hwc, src, need, cudaMemcpyDeviceToDevice, st));
  } else {
    dim3 grid((H * W + 31) / 32, (C + 31) / 32), block(32, 8);
    chw_to_hwc_kernel<<<grid, block, 0, st>>>(src, m.d_hwc, C, H * W);
    AVC_LAUNCH_CHECK(ctx, "chw_to_hwc_kernel");
  }
  // the host-buffer entry points run on internal streams: they order themselves after this copy through the event
  AVC_CUDA(ctx, cudaEventRecord(m.ready, st));
  return AVC_OK;
}

extern "C" int avc_set_feature_map(avc_ctx* ctx, int which, const float* chw, int C, int H, int W, void* stream) {
  return set_feature_map(ctx, which, chw, C, H, W, false, (cudaStream_t)stream);
}
extern "C" int avc_set_feature_map_hwc(avc_ctx* ctx, int which, const float* hwc, int C, int H, int W, void* stream) {
  return set_feature_map(ctx, which, hwc, C, H, W, true, (cudaStream_t)stream);
}

// -----------------------------------------------------------------------------------------------------------------
// AUTO resolves to the paired-CTA tcgen05 kernel (TC2) when the tensor-core path exists, else to the fp32 SIMT kernels; *impl is
// rewritten to the concrete choice.
static int pick_impl(avc_ctx* ctx, int* impl_io, bool* use_tc) {
  int impl = *impl_io;
  if (impl == AVC_IMPL_SIMT) { *use_tc = false; return AVC_OK; }
  if (impl == AVC_IMPL_AUTO) {
    *use_tc = avc_tc_available(ctx) != 0;
    *impl_io = *use_tc ? AVC_IMPL_TC2 : AVC_IMPL_SIMT;
    return AVC_OK;
  }
  if (impl == AVC_IMPL_TC || impl == AVC_IMPL_TC2) {
    if (!avc_tc_available(ctx)) return avc_fail(ctx, AVC_ESTATE, "tensor-core path requested but not available (needs sm_100 and a library built with tcgen05)");
    *use_tc = true; *impl_io = AVC_IMPL_TC2; return AVC_OK;      // AVC_IMPL_TC is kept in the ABI as an alias of the paired kernel
  }
  return avc_fail(ctx, AVC_EINVAL, "bad impl %d", impl);
}

// dense-grid entry on the fp32 SIMT implementation (the cross-check path): the points are materialised once into a context buffer
static int grid_points(avc_ctx* ctx, const AvcGridDesc* g, int64_t n, const float** pts, cudaStream_t st) {
  const size_t need = (size_t)n * 3 * sizeof(float);
  if (need > ctx->gridpts_cap) {
    if (ctx->d_gridpts) { AVC_CUDA(ctx, cudaStreamSynchronize(st)); cudaFree(ctx->d_gridpts); }
    ctx->d_gridpts = nullptr; ctx->gridpts_cap = 0;
    AVC_CUDA(ctx, cudaMalloc(&ctx->d_gridpts, need));
    ctx->gridpts_cap = need;
  }
  const float bounds[6] = {g->bmin[0], g->bmin[1], g->bmin[2], g->bmin[0] + g->len[0], g->bmin[1] + g->len[1], g->bmin[2] + g->len[2]};
  const int64_t plane = (int64_t)g->res[1] * g->res[2];
  int rc = avc_make_grid(ctx, bounds, g->res, g->x_first, (int)(n / plane), (float*)ctx->d_gridpts, st);
  if (rc) return rc;
  *pts = (const float*)ctx->d_gridpts;
  return AVC_OK;
}

static int eval_avatar(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* occ, float* off, float* rgb,
                       float* alpha, int if_type, int impl, int mode, cudaStream_t st) {
  if (!ctx) return AVC_EINVAL;
  if (n < 0 || (n > 0 && !pts && !grid)) return avc_fail(ctx, AVC_EINVAL, "field eval: bad points");
  if (if_type != AVC_IF_SDF && if_type != AVC_IF_OCCUPANCY) return avc_fail(ctx, AVC_EVALUE, "Invalid config.if_type!");   // arch_avatar.py:82
  if (!ctx->avatar.loaded) return avc_fail(ctx, AVC_ESTATE, "avatar weights not loaded");
  if (mode != AVC_MODE_TEMPLATE_ONLY && (!ctx->maps[AVC_MAP_POSE].d_hwc || !center))
    return avc_fail(ctx, AVC_ESTATE, "pose feature map not set (call avc_set_feature_map after WarpingField.precompute_conv)");
  bool use_tc; int rc = pick_impl(ctx, &impl, &use_tc);
  if (rc) return rc;
  const float zero[3] = {0, 0, 0};
  const float* c = center ? center : zero;
  if (use_tc) return avc_tc2_eval_avatar(ctx, pts, grid, n, c, occ, off, rgb, alpha, if_type, mode, st);
  if (grid && n > 0) { rc = grid_points(ctx, grid, n, &pts, st); if (rc) return rc; }
  return avc_simt_eval_avatar(ctx, pts, n, c, occ, off, rgb, alpha, if_type, mode, st);
}

// validates a dense-grid request and fills the descriptor; *n = number of grid points of the slab
static int make_grid_desc(avc_ctx* ctx, const float bounds[6], const int res[3], int x_first, int x_count, AvcGridDesc* g, int64_t* n) {
  if (!bounds || !res) return avc_fail(ctx, AVC_EINVAL, "grid eval: NULL argument");
  if (res[0] < 1 || res[1] < 1 || res[2] < 1 || x_first < 0 || x_count < 0 || x_first + x_count > res[0]) return avc_fail(ctx, AVC_EINVAL, "grid eval: bad grid / slab");
  for (int c = 0; c < 3; ++c) { g->bmin[c] = bounds[c]; g->len[c] = bounds[3 + c] - bounds[c]; g->res[c] = res[c]; }
  g->x_first = x_first;
  *n = (int64_t)x_count * res[1] * res[2];
  return AVC_OK;
}

extern "C" int avc_eval_occupancy(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_occ, float* out_off,
                                  float* out_rgb, float* out_alpha, int if_type, int impl, void* stream) {
  if (ctx && n > 0 && !out_occ) return avc_fail(ctx, AVC_EINVAL, "avc_eval_occupancy: out_occ is NULL");
  return eval_avatar(ctx, pts, nullptr, n, center, out_occ, out_off, out_rgb, out_alpha, if_type, impl, AVC_MODE_QUERY, (cudaStream_t)stream);
}

extern "C" int avc_eval_occupancy_grid(avc_ctx* ctx, const float bounds[6], const int res[3], int x_first, int x_count, const float center[3],
                                       float* out_occ, float* out_off, float* out_rgb, float* out_alpha, int if_type, int impl, void* stream) {
  if (!ctx) return AVC_EINVAL;
  AvcGridDesc g; int64_t n;
  int rc = make_grid_desc(ctx, bounds, res, x_first, x_count, &g, &n);
  if (rc) return rc;
  if (n > 0 && !out_occ) return avc_fail(ctx, AVC_EINVAL, "avc_eval_occupancy_grid: out_occ is NULL");
  return eval_avatar(ctx, nullptr, &g, n, center, out_occ, out_off, out_rgb, out_alpha, if_type, impl, AVC_MODE_QUERY, (cudaStream_t)stream);
}

extern "C" int avc_eval_warp(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_off, int impl, void* stream) {
  if (ctx && n > 0 && !out_off) return avc_fail(ctx, AVC_EINVAL, "avc_eval_warp: out_off is NULL");
  return eval_avatar(ctx, pts, nullptr, n, center, nullptr, out_off, nullptr, nullptr, AVC_IF_SDF, impl, AVC_MODE_WARP_ONLY, (cudaStream_t)stream);
}

extern "C" int avc_eval_template(avc_ctx* ctx, const float* pts, int64_t n, float* out_rgb, float* out_alpha, float* out_occ, int if_type,
                                 int impl, void* stream) {
  return eval_avatar(ctx, pts, nullptr, n, nullptr, out_occ, nullptr, out_rgb, out_alpha, if_type, impl, AVC_MODE_TEMPLATE_ONLY, (cudaStream_t)stream);
}

static int eval_recon(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* out_ov, int impl, cudaStream_t st) {
  if (n < 0 || (n > 0 && ((!pts && !grid) || !out_ov)) || !center) return avc_fail(ctx, AVC_EINVAL, "avc_eval_recon: bad argument");
  if (!ctx->recon.loaded) return avc_fail(ctx, AVC_ESTATE, "recon weights not loaded");
  if (!ctx->maps[AVC_MAP_IMAGE].d_hwc) return avc_fail(ctx, AVC_ESTATE, "image feature map not set");
  bool use_tc; int rc = pick_impl(ctx, &impl, &use_tc);
  if (rc) return rc;
  if (use_tc) return avc_tc2_eval_recon(ctx, pts, grid, n, center, out_ov, st);
  if (grid && n > 0) { rc = grid_points(ctx, grid, n, &pts, st); if (rc) return rc; }
  return avc_simt_eval_recon(ctx, pts, n, center, out_ov, st);
}

extern "C" int avc_eval_recon(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_ov, int impl, void* stream) {
  if (!ctx) return AVC_EINVAL;
  return eval_recon(ctx, pts, nullptr, n, center, out_ov, impl, (cudaStream_t)stream);
}

extern "C" int avc_eval_recon_grid(avc_ctx* ctx, const float bounds[6], const int res[3], int x_first, int x_count, const float center[3],
                                   float* out_ov, int impl, void* stream) {
  if (!ctx) return AVC_EINVAL;
  AvcGridDesc g; int64_t n;
  int rc = make_grid_desc(ctx, bounds, res, x_first, x_count, &g, &n);
  if (rc) return rc;
  return eval_recon(ctx, nullptr, &g, n, center, out_ov, impl, (cudaStream_t)stream);
}

// -----------------------------------------------------------------------------------------------------------------
// Host-buffer variants: chunked, double-buffered  H2D -> kernel -> D2H  on three internal streams.
// Staging layout per slot: pts (chunk,3) | occ (chunk) | off (chunk,3) | rgb (chunk,3) | alpha (chunk)  in pinned host and device memory.
static int ensure_staging(avc_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pinned_cap && bytes <= ctx->stage_cap) return AVC_OK;
  if (ctx->h_pinned) { cudaFreeHost(ctx->h_pinned); ctx->h_pinned = nullptr; ctx->pinned_cap = 0; }
  if (ctx->d_stage) { cudaFree(ctx->d_stage); ctx->d_stage = nullptr; ctx->stage_cap = 0; }
  AVC_CUDA(ctx, cudaMallocHost(&ctx->h_pinned, bytes)); ctx->pinned_cap = bytes;
  AVC_CUDA(ctx, cudaMalloc(&ctx->d_stage, bytes)); ctx->stage_cap = bytes;
  return AVC_OK;
}

// true when `p` is page-locked host memory (cudaMallocHost / cudaHostRegister / torch pin_memory): the DMA engines can use it directly
static bool is_pinned_host(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

static int eval_host(avc_ctx* ctx, bool recon, const float* pts, int64_t n, const float center[3], float* out_a, float* out_off, float* out_rgb,
                     float* out_alpha, int if_type, int impl) {
  if (!ctx) return AVC_EINVAL;
  if (n < 0 || (n > 0 && (!pts || !out_a)) || !center) return avc_fail(ctx, AVC_EINVAL, "host eval: bad argument");
  if (n == 0) return AVC_OK;
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  int64_t chunk = 1 << 21;                             // 2 Mi points per pipeline slot (24 MB in, 8..32 MB out)
  if (const char* e = getenv("AVC_HOST_CHUNK")) {      // tuning knob: points per pipeline slot
    const long long v = atoll(e);
    if (v >= 1024 && v <= (1ll << 26)) chunk = v;
  }
  const size_t slot_floats = (size_t)chunk * 11;
  int rc = ensure_staging(ctx, 2 * slot_floats * sizeof(float));
  if (rc) return rc;
  // buffers that are already page-locked skip the staging copy (DMA straight from / into the caller's memory)
  const bool pin_in = is_pinned_host(pts), pin_a = is_pinned_host(out_a), pin_off = is_pinned_host(out_off), pin_rgb = is_pinned_host(out_rgb),
             pin_al = is_pinned_host(out_alpha);
  cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
  for (int s = 0; s < 2; ++s) { cudaEventCreateWithFlags(&ev_in[s], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_k[s], cudaEventDisableTiming); cudaEventCreateWithFlags(&ev_out[s], cudaEventDisableTiming); }
  const int64_t n_chunks = (n + chunk - 1) / chunk;
  int status = AVC_OK;
  // the feature map was (maybe) written on the caller's stream by avc_set_feature_map: the compute stream must see it complete
  if (cudaEvent_t ready = ctx->maps[recon ? AVC_MAP_IMAGE : AVC_MAP_POSE].ready) cudaStreamWaitEvent(ctx->s_compute, ready, 0);
  // AVC_HOST_TRACE=1: per-chunk timeline (ms from the first H2D) of copy-in, kernel and copy-out, printed to stderr
  const bool trace = getenv("AVC_HOST_TRACE") != nullptr && n_chunks <= 64;
  cudaEvent_t tev[64][5];
  if (trace) for (int64_t c = 0; c < n_chunks; ++c) for (int e = 0; e < 5; ++e) cudaEventCreate(&tev[c][e]);
  // software pipeline: iteration c stages chunk c (memcpy to pinned + H2D), launches its kernel, queues its D2H,
  // and retires chunk c-2's slot (copies its pinned outputs to the caller's buffers) before reusing it.
  for (int64_t c = 0; c < n_chunks + 2 && status == AVC_OK; ++c) {
    if (c >= 2) {            // retire chunk c-2
      const int s = (int)(c & 1); const int64_t b = (c - 2) * chunk; const int64_t m = (n - b < chunk) ? n - b : chunk;
      if (cudaEventSynchronize(ev_out[s]) != cudaSuccess) { status = avc_check_cuda(ctx, cudaGetLastError(), "host eval: D2H"); break; }
      float* hp = (float*)ctx->h_pinned + (size_t)s * slot_floats;
      if (!pin_a) memcpy(out_a + b, hp + (size_t)chunk * 3, (size_t)m * sizeof(float));
      if (out_off && !pin_off) memcpy(out_off + b * 3, hp + (size_t)chunk * 4, (size_t)m * 3 * sizeof(float));
      if (out_rgb && !pin_rgb) memcpy(out_rgb + b * 3, hp + (size_t)chunk * 7, (size_t)m * 3 * sizeof(float));
      if (out_alpha && !pin_al) memcpy(out_alpha + b, hp + (size_t)chunk * 10, (size_t)m * sizeof(float));
    }
    if (c < n_chunks) {
      const int s = (int)(c & 1); const int64_t b = c * chunk; const int64_t m = (n - b < chunk) ? n - b : chunk;
      float* hp = (float*)ctx->h_pinned + (size_t)s * slot_floats;
      float* dp = (float*)ctx->d_stage + (size_t)s * slot_floats;
      if (!pin_in) memcpy(hp, pts + b * 3, (size_t)m * 3 * sizeof(float));
      if (trace) cudaEventRecord(tev[c][0], ctx->s_copy_in);
      cudaMemcpyAsync(dp, pin_in ? pts + b * 3 : hp, (size_t)m * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->s_copy_in);
      cudaEventRecord(ev_in[s], ctx->s_copy_in);
      if (trace) cudaEventRecord(tev[c][1], ctx->s_copy_in);
      cudaStreamWaitEvent(ctx->s_compute, ev_in[s], 0);
      if (trace) cudaEventRecord(tev[c][2], ctx->s_compute);
      float* d_occ = dp + (size_t)chunk * 3; float* d_off = dp + (size_t)chunk * 4;
      float* d_rgb = dp + (size_t)chunk * 7; float* d_alpha = dp + (size_t)chunk * 10;
      status = recon ? avc_eval_recon(ctx, dp, m, center, d_occ, impl, ctx->s_compute)
                     : avc_eval_occupancy(ctx, dp, m, center, d_occ, out_off ? d_off : nullptr, out_rgb ? d_rgb : nullptr,
                                          out_alpha ? d_alpha : nullptr, if_type, impl, ctx->s_compute);
      if (status != AVC_OK) break;
      cudaEventRecord(ev_k[s], ctx->s_compute);
      if (trace) cudaEventRecord(tev[c][3], ctx->s_compute);
      cudaStreamWaitEvent(ctx->s_copy_out, ev_k[s], 0);
      cudaMemcpyAsync(pin_a ? out_a + b : hp + (size_t)chunk * 3, d_occ, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_copy_out);
      if (out_off) cudaMemcpyAsync(pin_off ? out_off + b * 3 : hp + (size_t)chunk * 4, d_off, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_copy_out);
      if (out_rgb) cudaMemcpyAsync(pin_rgb ? out_rgb + b * 3 : hp + (size_t)chunk * 7, d_rgb, (size_t)m * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_copy_out);
      if (out_alpha) cudaMemcpyAsync(pin_al ? out_alpha + b : hp + (size_t)chunk * 10, d_alpha, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, ctx->s_copy_out);
      cudaEventRecord(ev_out[s], ctx->s_copy_out);
      if (trace) cudaEventRecord(tev[c][4], ctx->s_copy_out);
      // The next use of this slot's buffers is chunk c+2; the host waits on ev_out[s] (retire, above) before it enqueues
      // anything of that chunk, which orders it after this D2H. (A stream wait enqueued HERE on s_copy_in would also hold
      // back chunk c+1's H2D -- it serialised the whole pipeline in the first version: see profiles/r1_host_entry_timeline.txt.)
    }
  }
  cudaStreamSynchronize(ctx->s_copy_in); cudaStreamSynchronize(ctx->s_compute); cudaStreamSynchronize(ctx->s_copy_out);
  for (int s = 0; s < 2; ++s) { cudaEventDestroy(ev_in[s]); cudaEventDestroy(ev_k[s]); cudaEventDestroy(ev_out[s]); }
  if (trace) {
    fprintf(stderr, "chunk | h2d start  h2d end | kernel start  kernel end | d2h end   (ms)\n");
    for (int64_t c = 0; c < n_chunks; ++c) {
      float t[5];
      for (int e = 0; e < 5; ++e) { t[e] = 0.f; cudaEventElapsedTime(&t[e], tev[0][0], tev[c][e]); }
      fprintf(stderr, "%5lld | %9.3f %8.3f | %12.3f %11.3f | %7.3f\n", (long long)c, t[0], t[1], t[2], t[3], t[4]);
    }
    for (int64_t c = 0; c < n_chunks; ++c) for (int e = 0; e < 5; ++e) cudaEventDestroy(tev[c][e]);
    cudaGetLastError();
  }
  if (status == AVC_OK) { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) status = avc_check_cuda(ctx, e, "host eval"); }
  return status;
}

extern "C" int avc_eval_occupancy_host(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_occ, float* out_off,
                                       float* out_rgb, float* out_alpha, int if_type, int impl) {
  return eval_host(ctx, false, pts, n, center, out_occ, out_off, out_rgb, out_alpha, if_type, impl);
}
extern "C" int avc_eval_recon_host(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_ov, int impl) {
  return eval_host(ctx, true, pts, n, center, out_ov, nullptr, nullptr, nullptr, AVC_IF_SDF, impl);
}
