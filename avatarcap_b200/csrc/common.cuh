// Shared declarations for the avatarcap_b200 CUDA library (context, weight-blob layout, error helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/avatarcap_b200.h"

#define AVC_MAGIC 0x57435641u /* 'AVCW' */
#define AVC_BLOB_VERSION 3u
#define AVC_MAX_LAYERS 24
#define AVC_WEIGHT_SLOTS 4

// Fixed layer order inside a blob (packer.py writes them in this order).
// avatar: 0..6 warp conv1..7 (BN folded, softplus) | 7 warp out (256->3) | 8..14 shared fc0..6 | 15,16 geo | 17..19 clr
// recon : 0..3 image_decoder fc0..3
enum { AVC_KIND_AVATAR = 0, AVC_KIND_RECON = 1 };
enum { AVC_ACT_NONE = 0, AVC_ACT_RELU = 1, AVC_ACT_LRELU = 2, AVC_ACT_SOFTPLUS = 3, AVC_ACT_SIGMOID = 4 };

struct AvcLayerDesc {
  int32_t k0, k1;      // true K of the first / second (skip) input segment, in the reference's concat order
  int32_t k0p, k1p;    // padded to a multiple of 16 (tensor-core path)
  int32_t n, np;       // output channels, padded to a multiple of 16 (>= 16)
  int32_t act;
  int32_t wt_off;      // f32 section, float index: W^T[(k0+k1)][n] (row k holds n contiguous outputs); heads (n<=4): W[n][k0+k1]
  int32_t sb_off;      // f32 section, float index: scale[n] then bias[n]   (y = acc*scale + bias)
  int32_t tc_w_off;    // f16 section, byte offset of the layer's weight STREAM in tensor-core consumption order (packer.py tc_pieces)
  int32_t tc_sb_off;   // f32 section, float index: scale[np] then bias[np] for the tensor-core path (scale includes 2^-wshift)
  int32_t reserved;
};

struct AvcBlobHeader {
  uint32_t magic, version, kind, n_layers;
  uint64_t f32_off, f32_bytes;    // byte offset / size of the float32 section
  uint64_t f16_off, f16_bytes;    // byte offset / size of the fp16 hi/lo section
  AvcLayerDesc layers[AVC_MAX_LAYERS];
};

struct AvcWeights {
  bool loaded = false;
  AvcBlobHeader hdr;
  unsigned char* d_blob = nullptr;   // whole blob on the device
  const float* d_f32 = nullptr;
  const unsigned char* d_f16 = nullptr;
};

struct AvcMap {
  float* d_hwc = nullptr;   // (H,W,C)
  size_t cap = 0;
  int C = 0, H = 0, W = 0;
  cudaEvent_t ready = nullptr;   // recorded on the caller's stream after the copy / transpose into d_hwc (host entries wait on it)
};

struct avc_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  std::string err;
  // `avatar` / `recon` are the ACTIVE weights (a view: the device blobs are owned by the slots below). avc_select_weights()
  // switches between pre-uploaded blobs with a pointer swap -- the reference's test loop alternates two GeoTexAvatar
  // instances every frame (main.py:307-315).
  AvcWeights avatar, recon;
  AvcWeights slots[2][AVC_WEIGHT_SLOTS];
  int slot_sel[2] = {0, 0};
  AvcMap maps[2];
  int64_t launches = 0;
  // scratch owned by the context (marching cubes scans, host staging)
  void* d_scratch = nullptr; size_t scratch_cap = 0;
  void* d_scratch2 = nullptr; size_t scratch2_cap = 0;   // marching cubes: compact edge list
  void* d_grid = nullptr; size_t grid_cap = 0;           // KNN: uniform grid over the reference vertices
  void* d_gridpts = nullptr; size_t gridpts_cap = 0;     // dense-grid entry on the SIMT implementation: materialised points
  void* h_pinned = nullptr;  size_t pinned_cap = 0;
  void* d_stage = nullptr;   size_t stage_cap = 0;
  cudaStream_t s_copy_in = nullptr, s_compute = nullptr, s_copy_out = nullptr;
  int64_t* h_counts = nullptr;   // pinned, small
  // marching cubes: what the last avc_mc_count scanned (block sums still in d_scratch), for avc_mc_emit_counted
  struct { const void* vol = nullptr; int res[3] = {0, 0, 0}; float iso = 0.f; int lo = 0, hi = 0; int64_t counts[3] = {0, 0, 0}; bool valid = false; } mc_last;
  int dbg_flags = 0;             // avc_debug_set_trace(flags): timing experiments of the tensor-core kernel
  void* d_trace = nullptr;       // optional debug timeline buffer for the tensor-core kernel (avc_debug_set_trace)
};

int avc_fail(avc_ctx* ctx, int code, const char* fmt, ...);
int avc_check_cuda(avc_ctx* ctx, cudaError_t e, const char* what);
int avc_ensure_scratch(avc_ctx* ctx, size_t bytes);

#define AVC_CUDA(ctx, call)                                                        \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return avc_check_cuda((ctx), _e, #call);                \
  } while (0)

#define AVC_LAUNCH_CHECK(ctx, name)                                                \
  do {                                                                             \
    (ctx)->launches++;                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) return avc_check_cuda((ctx), _e, name);                 \
  } while (0)

// ---- kernels implemented in the other translation units -------------------------------------------------
int avc_simt_eval_avatar(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_occ, float* out_off,
                         float* out_rgb, float* out_alpha, int if_type, int mode, cudaStream_t st);
int avc_simt_eval_recon(avc_ctx* ctx, const float* pts, int64_t n, const float center[3], float* out_ov, cudaStream_t st);
int avc_tc_available(const avc_ctx* ctx);
// dense-grid descriptor: point g = grid point of flat index g + x_first*Ry*Rz of generate_volume_points (avatarcap_dataset.py:312-326)
struct AvcGridDesc { float bmin[3], len[3]; int res[3]; int x_first; };
// tcgen05 kernel on CTA pairs (field_tc2.cu); pts == NULL with grid != NULL: coordinates from the index, nothing read per point
int avc_tc2_eval_avatar(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* out_occ, float* out_off,
                        float* out_rgb, float* out_alpha, int if_type, int mode, cudaStream_t st);
int avc_tc2_eval_recon(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* out_ov, cudaStream_t st);

// mode for the avatar evaluation
enum { AVC_MODE_QUERY = 0 /* warp + template */, AVC_MODE_WARP_ONLY = 1, AVC_MODE_TEMPLATE_ONLY = 2 };
