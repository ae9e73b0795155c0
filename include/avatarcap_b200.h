/*
 * avatarcap_b200 -- C ABI of the B200-native implementation of AvatarCap's dense implicit-surface path.
 *
 * The reference (lizhe00/AvatarCap) has no FFI: its seams are Python call sites. Each entry point below
 * names the reference call (file:line, relative to the reference root) whose arithmetic it replaces; the
 * Python mirror that re-binds those call sites lives in avatarcap_b200/patch.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative AVC_E* code otherwise; avc_last_error() gives the text.
 *     No C++ exception crosses this boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Work is enqueued
 *     asynchronously on it; functions that must return a data-dependent count say that they synchronise it.
 *   - pointers marked [dev] are device pointers on the context's device, caller-owned, never retained past
 *     the call (exception: avc_set_feature_map copies into a library-owned buffer). [host] pointers are
 *     ordinary host memory.
 *   - all floating-point data is IEEE float32; points are (n,3) row-major xyz; volumes are (Rx,Ry,Rz)
 *     row-major with z fastest, i.e. flat = (i*Ry + j)*Rz + k  (dataset/avatarcap_dataset.py:317-321, main.py:364).
 *   - one context per device; a context is not re-entrant.
 */
#ifndef AVATARCAP_B200_H
#define AVATARCAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVC_ABI_VERSION   5    /* what avc_abi_version() of a matching library returns */

#define AVC_OK            0
#define AVC_EINVAL       -1   /* bad argument (NULL pointer, bad shape, bad flag)              */
#define AVC_ECUDA        -2   /* a CUDA runtime call or kernel launch failed                   */
#define AVC_ESTATE       -3   /* missing prerequisite (weights / feature map not loaded)       */
#define AVC_ECAPACITY    -4   /* caller-provided output capacity too small (never truncates)   */
#define AVC_EFORMAT      -5   /* weight blob malformed / wrong architecture                    */
#define AVC_EVALUE       -6   /* value error mirrored from the reference (e.g. iso outside the volume's range) */

typedef struct avc_ctx avc_ctx;

/* which implementation of the field evaluation to run */
#define AVC_IMPL_AUTO   0     /* tensor-core kernel when available, else SIMT                  */
#define AVC_IMPL_SIMT   1     /* fp32 CUDA-core kernel (bit-near the reference's fp32 math)    */
#define AVC_IMPL_TC     2     /* alias of AVC_IMPL_TC2 (the single-CTA kernel of ABI <= 4 is gone) */
#define AVC_IMPL_TC2    3     /* tcgen05 kernel on CTA pairs (cta_group::2, M = 256 per MMA), fp16 hi/lo split operands, fp32
                                 accumulate: what AUTO resolves to on sm_100 */

/* implicit-field type, config.py:12-22 */
#define AVC_IF_SDF        0
#define AVC_IF_OCCUPANCY  1

/* feature-map slots */
#define AVC_MAP_POSE   0      /* WarpingField.pose_feat_map, (64,H,W)  arch_avatar.py:109-111   */
#define AVC_MAP_IMAGE  1      /* ReconNetwork HGFilter output, (32,H,W) arch_recon.py:51-52     */

/* ---------------------------------------------------------------------------------------------- */
/* context                                                                                        */
/* ---------------------------------------------------------------------------------------------- */
int  avc_ctx_create(int device, avc_ctx** out);
void avc_ctx_destroy(avc_ctx* ctx);
/* text of the last error on this context (or of the last failed avc_ctx_create when ctx == NULL) */
const char* avc_last_error(const avc_ctx* ctx);
/* ABI version of the library (bumped on any signature change) */
int  avc_abi_version(void);
/* 1 if the tcgen05 kernels were compiled in and the device is sm_100 */
int  avc_has_tensor_core_path(const avc_ctx* ctx);
/* number of kernel launches issued by this context since creation / last reset (for bench.py's gpu_launches) */
int64_t avc_launch_count(const avc_ctx* ctx);
void    avc_reset_launch_count(avc_ctx* ctx);
/* debugging aid: [dev] buffer of 4*24*8 int64 that the tensor-core kernel fills with clock64() stamps of CTA 0's
 * first tiles (op start / issue end / accumulator-ready / epilogue-done); NULL disables it (default). */
int avc_debug_set_trace(avc_ctx* ctx, void* dev_buf /*[dev]|NULL*/, int flags /*0; non-zero = timing experiments, results invalid*/);

/* ---------------------------------------------------------------------------------------------- */
/* weights -- replaces torch.load(...)['network'] + nn.Module parameters (main.py:302-320)        */
/* The blob is produced by avatarcap_b200.packer from the reference's state_dict:
 * BatchNorm(eval) and weight-norm folded into a per-channel scale/bias, skip-concat columns reordered,
 * K padded to the MMA granule, fp16 hi/lo planes in the tcgen05 canonical K-major core-matrix layout.
 * The library copies it ([host] pointer) into device memory it owns.                               */
/* ---------------------------------------------------------------------------------------------- */
int avc_load_avatar_weights(avc_ctx* ctx, const void* blob /*[host]*/, size_t nbytes);
int avc_load_recon_weights(avc_ctx* ctx, const void* blob /*[host]*/, size_t nbytes);
/* The reference's test loop alternates two GeoTexAvatar instances every frame (`network` for the geometry, `network_finetuned`
 * for the colours, main.py:307-315): AVC_WEIGHT_SLOTS blobs per kind stay resident and avc_select_weights switches the active
 * one with a pointer swap. kind: 0 = avatar, 1 = recon. avc_load_*_weights above (re)load the ACTIVE slot. Loading into an
 * occupied slot synchronises the device first (kernels in flight may still read the old blob).                          */
#define AVC_WEIGHT_SLOTS 4
int avc_load_weights_slot(avc_ctx* ctx, int kind, int slot, const void* blob /*[host]*/, size_t nbytes);
int avc_select_weights(avc_ctx* ctx, int kind, int slot);

/* per-frame encoder output (stays in PyTorch; WarpingField.precompute_conv arch_avatar.py:109-111,
 * ReconNetwork.get_feat_maps arch_recon.py:41-43). `chw` is a [dev] (C,H,W) float32 tensor; the library
 * transposes it into an owned (H,W,C) copy so that one bilinear tap is one contiguous read.            */
/* Stream rule: the copy is enqueued on `stream`. Device-pointer entry points called on the SAME stream are ordered after it by
 * the stream; the *_host entry points (internal streams) wait for an event recorded behind the copy. A device-pointer entry
 * point called on a DIFFERENT stream must be ordered by the caller.                                                     */
int avc_set_feature_map(avc_ctx* ctx, int which, const float* chw /*[dev]*/, int C, int H, int W, void* stream);
/* same, for an encoder that already produced the (H,W,C) order (a channels_last torch tensor, avatarcap_b200/encoders.py):
 * plain device copy into the owned buffer, no transpose.                                                          */
int avc_set_feature_map_hwc(avc_ctx* ctx, int which, const float* hwc /*[dev]*/, int C, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* per-frame image encoder on the tensor cores ("next" row 1)                                      */
/* ReconNetwork.get_feat_maps (arch_recon.py:41-43) = HGFilter.forward (network/HGFilters.py:177-219). The network is handed over
 * as a PROGRAM of library ops over numbered buffers (avatarcap_b200/encoders.py build_hgfilter_program: 7x7 stem, GroupNorm + ReLU +
 * fp16 hi/lo split, tcgen05 implicit-GEMM 3x3 / 1x1 convolution fed by TMA tensor loads, residual add, 2x2 average pool, bicubic x2
 * up-sampling) plus a weight blob (fp16 hi / lo planes, (C_out, taps, C_in) each) and a parameter blob (f32). The library owns copies.
 * avc_encoder_run: in [dev] (C,H,W) f32 as the reference feeds it, out [dev] (H/2, W/2, 32) f32 = the (H,W,C) order
 * avc_set_feature_map_hwc takes; use_graph != 0 replays the program from a CUDA graph captured on first use.
 * The same interpreter runs GeoTexAvatar's UNet (WarpingField.precompute_conv, arch_avatar.py:109-111 = UnetNoCond7DS.forward,
 * network/unets.py:201-219; build_unet_program): conv1..7 and upconv1, 2, 3, 3 (4x4 stride-2 / transposed convolutions, eval BatchNorm
 * folded, in-place LeakyReLU) as split-K fp32 gather-GEMMs (CONV4) that write each skip into its slice of the concatenated buffer, then
 * upconvC5/C6/C7 (bilinear x2 + fp16 split = UPSPLIT, tcgen05 3x3 CONV, skip COPY): in [dev] (6,H,W) f32, out [dev] (H,W,64) f32.
 * A program may instead take a flat f32 vector of (H,W,C) tensors that INPUT ops slice into buffers (build_unet_tail_program).       */
typedef struct avc_encoder avc_encoder;
int  avc_encoder_create(avc_ctx* ctx, const int32_t* program /*[host]*/, int64_t n_words, const void* weights_f16 /*[host]*/, size_t weight_bytes,
                        const float* params_f32 /*[host]*/, int64_t n_params, avc_encoder** out);
int  avc_encoder_run(avc_encoder* enc, const float* in /*[dev]*/, float* out /*[dev]*/, int use_graph, void* stream);
int  avc_encoder_shape(const avc_encoder* enc, int in_chw[3], int out_hwc[3]);
/* tests / debugging: copy f32 buffer `buf` of the program (its contents after the last run) to dst */
int  avc_encoder_read_buffer(avc_encoder* enc, int buf, float* dst /*[dev]*/, int64_t n_floats, void* stream);
void avc_encoder_destroy(avc_encoder* enc);

/* ---------------------------------------------------------------------------------------------- */
/* field evaluation                                                                               */
/* ---------------------------------------------------------------------------------------------- */
/* OccupancyNet.query (arch_avatar.py:356-381): off = WarpingField.query(p) (:113-140); (rgb,alpha,occ) =
 * DoubleTNet.forward(p + off) (:65-83). out_occ (n) is required; out_off (n,3), out_rgb (n,3), out_alpha (n)
 * may be NULL (the colour head is skipped when both out_rgb and out_alpha... see below are NULL).
 * out_alpha is relu(geo[1]) as DoubleTNet.forward returns it. if_type selects raw / sigmoid occupancy.     */
int avc_eval_occupancy(avc_ctx* ctx, const float* pts /*[dev] (n,3)*/, int64_t n, const float center[3] /*[host]*/,
                       float* out_occ /*[dev]*/, float* out_off /*[dev]|NULL*/, float* out_rgb /*[dev]|NULL*/,
                       float* out_alpha /*[dev]|NULL*/, int if_type, int impl, void* stream);

/* Dense-grid form of avc_eval_occupancy: the points are those of generate_volume_points (dataset/avatarcap_dataset.py:312-326) for
 * the i-planes [x_first, x_first + x_count) of a (Rx,Ry,Rz) grid over `bounds` -- what main.py:357-360 evaluates when every grid
 * point is valid. The kernel derives the coordinates from the point index (bit-identical to avc_make_grid + avc_eval_occupancy),
 * so nothing is read per point and the 12 B/point list never exists. Outputs are in grid order, (x_count*Ry*Rz) rows.          */
int avc_eval_occupancy_grid(avc_ctx* ctx, const float bounds[6] /*[host]*/, const int res[3], int x_first, int x_count,
                            const float center[3] /*[host]*/, float* out_occ /*[dev]*/, float* out_off /*[dev]|NULL*/,
                            float* out_rgb /*[dev]|NULL*/, float* out_alpha /*[dev]|NULL*/, int if_type, int impl, void* stream);
/* the same for ReconNetwork.infer's decoder (main.py:438-440) */
int avc_eval_recon_grid(avc_ctx* ctx, const float bounds[6] /*[host]*/, const int res[3], int x_first, int x_count,
                        const float center[3] /*[host]*/, float* out_ov /*[dev]*/, int impl, void* stream);

/* WarpingField.query alone (arch_avatar.py:113-140): offsets (n,3) */
int avc_eval_warp(avc_ctx* ctx, const float* pts /*[dev]*/, int64_t n, const float center[3] /*[host]*/,
                  float* out_off /*[dev] (n,3)*/, int impl, void* stream);

/* DoubleTNet.forward alone (arch_avatar.py:65-83): rgb (n,3), alpha (n), occ (n); any output may be NULL */
int avc_eval_template(avc_ctx* ctx, const float* pts /*[dev]*/, int64_t n, float* out_rgb, float* out_alpha,
                      float* out_occ, int if_type, int impl, void* stream);

/* ReconNetwork.infer, per-point part (arch_recon.py:55-76): out_ov (n) = sigmoid(decoder([feat(x,-y), z])) */
int avc_eval_recon(avc_ctx* ctx, const float* pts /*[dev]*/, int64_t n, const float center[3] /*[host]*/,
                   float* out_ov /*[dev]*/, int impl, void* stream);

/* Same as avc_eval_occupancy / avc_eval_recon with HOST buffers: pinned staging, H2D, kernels and D2H are
 * pipelined on internal streams and the call returns when the host outputs are complete. This is the
 * end-to-end entry bench.py times as `e2e`.                                                              */
int avc_eval_occupancy_host(avc_ctx* ctx, const float* pts /*[host]*/, int64_t n, const float center[3],
                            float* out_occ /*[host]*/, float* out_off /*[host]|NULL*/, float* out_rgb /*[host]|NULL*/,
                            float* out_alpha /*[host]|NULL*/, int if_type, int impl);
int avc_eval_recon_host(avc_ctx* ctx, const float* pts /*[host]*/, int64_t n, const float center[3],
                        float* out_ov /*[host]*/, int impl);

/* ---------------------------------------------------------------------------------------------- */
/* grid, mask scatter                                                                             */
/* ---------------------------------------------------------------------------------------------- */
/* generate_volume_points (dataset/avatarcap_dataset.py:312-326): out (Rx*Ry*Rz,3), z fastest.
 * x_first / x_count select a slab of i-planes [x_first, x_first+x_count) (multi-GPU sharding); the
 * coordinates are those of the full grid.                                                         */
int avc_make_grid(avc_ctx* ctx, const float bounds[6] /*[host] min xyz, max xyz*/, const int res[3],
                  int x_first, int x_count, float* out_pts /*[dev]*/, void* stream);

/* main.py:357,362-364 / 438,442-443: vol[flag] = vals (in order); vol[~flag] = fill (in order).
 * flag: n_total bytes (bool); vals: count(flag) floats; fill: n_total - count(flag) floats.            */
int avc_scatter_fill(avc_ctx* ctx, const uint8_t* flag /*[dev]*/, int64_t n_total, const float* vals /*[dev]*/,
                     const float* fill /*[dev]*/, float* out_vol /*[dev]*/, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* mesh extraction -- replaces recon_util.recon_mesh (utils/recon_util.py:51-70), including the
 * skimage marching cubes call (:64), the Sobel normal volume (:9-29) and its trilinear sampling (:32-48).
 * vol: [dev] (Rx,Ry,Rz). Outputs are in the reference's coordinates and conventions: vertices already
 * shifted by bounds[0] + 0.5*voxel (:65), normals negated (:68), faces column-reversed (:69).
 * Vertex order: ascending (owner voxel linear index, axis); face order: ascending (cell, case-table slot).
 * x_halo_lo / x_halo_hi: number of leading / trailing i-planes of `vol` that are halo only (multi-GPU slabs):
 * cells and vertices are emitted for owner voxels with x_halo_lo <= i < Rx - x_halo_hi, but the Sobel
 * stencil reads the halo. x_origin is the global index of plane 0 and gres_x the global Rx (for coordinates;
 * pass 0 and res[0] for a whole volume).
 * Two-phase: avc_mc_count synchronises `stream` and returns the counts; avc_mc_emit fills caller buffers
 * and fails with AVC_ECAPACITY (nothing written) if they are too small.                                   */
int avc_mc_count(avc_ctx* ctx, const float* vol /*[dev]*/, const int res[3], float iso, int x_halo_lo, int x_halo_hi,
                 int64_t* n_verts /*[host]*/, int64_t* n_faces /*[host]*/, void* stream);
int avc_mc_emit(avc_ctx* ctx, const float* vol /*[dev]*/, const int res[3], const float bounds[6] /*[host]*/, float iso,
                int x_halo_lo, int x_halo_hi, int x_origin, int gres_x,
                float* verts /*[dev] (cap_v,3)*/, float* normals /*[dev] (cap_v,3)|NULL*/, int32_t* faces /*[dev] (cap_f,3)*/,
                int64_t cap_v, int64_t cap_f, void* stream);

/* Capacity-bounded, fully ASYNCHRONOUS extraction (no host round trip for the data-dependent sizes): one fused
 * classify + chained-scan pass, one thread per vertex, one thread per triangle, all enqueued on `stream`. counts [dev] receives
 * 4 x int64: {n_verts (owned), n_faces, n_verts including the next slab's first plane (ids >= n_verts, see x_halo_hi), overflow
 * flags: bit 0 = n_verts > cap_v, bit 1 = n_faces > cap_f}. The counts are always exact; when a flag is set the buffers hold the
 * first cap_v vertices / cap_f faces and the caller must re-run with larger buffers (never a silent truncation: the flag is the
 * error). The caller reads `counts` whenever it needs the sizes (after synchronising `stream`).                               */
int avc_mc_extract(avc_ctx* ctx, const float* vol /*[dev]*/, const int res[3], const float bounds[6] /*[host]*/, float iso,
                   int x_halo_lo, int x_halo_hi, int x_origin, int gres_x,
                   float* verts /*[dev] (cap_v,3)*/, float* normals /*[dev] (cap_v,3)|NULL*/, int32_t* faces /*[dev] (cap_f,3)*/,
                   int64_t cap_v, int64_t cap_f, int64_t* counts /*[dev] 4 x int64*/, void* stream);

/* avc_mc_emit for the volume the IMMEDIATELY preceding avc_mc_count call on this context scanned (same pointer, extents, iso, halo):
 * reuses that call's totals instead of counting and synchronising a second time. The caller guarantees the volume was not
 * modified in between; any other library call in between, or any differing argument, silently falls back to avc_mc_emit.       */
int avc_mc_emit_counted(avc_ctx* ctx, const float* vol /*[dev]*/, const int res[3], const float bounds[6] /*[host]*/, float iso,
                        int x_halo_lo, int x_halo_hi, int x_origin, int gres_x,
                        float* verts /*[dev] (cap_v,3)*/, float* normals /*[dev] (cap_v,3)|NULL*/, int32_t* faces /*[dev] (cap_f,3)*/,
                        int64_t cap_v, int64_t cap_f, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* multi-GPU slab exchange over peer memory (one process per GPU)            */
/* The reference evaluates one volume on one GPU (main.py:357-367); sharding it in x-slabs needs ONE exchange step: marching
 * cubes + normals on a slab read 2 planes of the slab below and 3 of the slab above. Each rank owns a padded buffer
 *   [AVC_SHARD_HEADER_BYTES of flags | float data: lo halo | own planes | hi halo]
 * allocated here (cudaMalloc) and exported as a CUDA IPC handle; the neighbours open it and STORE their boundary planes into it
 * over NVLink from a kernel (avc_halo_push) -- no NCCL call, no staging copy, no host synchronisation. Offsets below are in
 * float elements from the start of the data area. `epoch` counts exchanges from 1, identically on all ranks.
 * Call order per volume on every rank, all on one stream:  field evaluation into `own` -> avc_halo_push -> avc_halo_wait ->
 * avc_mc_extract on [lo | own | hi] -> avc_halo_ack.                                                                          */
#define AVC_SHARD_HEADER_BYTES 256
#define AVC_IPC_HANDLE_BYTES 64
int avc_shard_alloc(avc_ctx* ctx, size_t data_bytes, void** out_base /*[dev] buffer incl. header*/, void* out_handle /*[host] 64 B*/);
int avc_shard_open(avc_ctx* ctx, const void* handle /*[host] 64 B, from a peer process*/, void** out_base /*[dev], mapped peer memory*/);
int avc_shard_close(avc_ctx* ctx, void* base /*from avc_shard_open*/);
int avc_shard_free(avc_ctx* ctx, void* base /*from avc_shard_alloc; synchronises the device*/);
/* store my first n_to_lo planes into the lower neighbour's buffer at lo_dst_off and my last n_to_hi planes into the upper
 * neighbour's at hi_dst_off (peer_* = NULL: no such neighbour), after both have acknowledged epoch-1; then publish `epoch` */
int avc_halo_push(avc_ctx* ctx, void* mine, void* peer_lo, void* peer_hi, int64_t plane_elems, int64_t own_off, int nx,
                  int n_to_lo, int64_t lo_dst_off, int n_to_hi, int64_t hi_dst_off, uint64_t epoch, void* stream);
/* stream-ordered wait (a spinning one-thread kernel with a 20 s watchdog) until the neighbours' planes of `epoch` have arrived */
int avc_halo_wait(avc_ctx* ctx, void* mine, int need_lo, int need_hi, uint64_t epoch, void* stream);
/* tell the neighbours that their planes of `epoch` have been consumed (enqueue after the kernels that read the halo) */
int avc_halo_ack(avc_ctx* ctx, void* peer_lo, void* peer_hi, uint64_t epoch, void* stream);
/* seam rule of the slab meshes: local ids < n_own -> + base_own, ids >= n_own (the next slab's first-plane vertices, numbered
 * behind the slab's own by avc_mc_extract) -> - n_own + base_next. In place.                                               */
int avc_renumber_faces(avc_ctx* ctx, int32_t* faces /*[dev] (n_faces,3)*/, int64_t n_faces, int64_t n_own, int64_t base_own,
                       int64_t base_next, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* LBS -- replaces utils/smpl_util.py and the pytorch3d KNN it calls                              */
/* ---------------------------------------------------------------------------------------------- */
/* pytorch3d.ops.knn_points (K<=4) against a small reference set (m <= 16384): squared L2 ascending.
 * out_d2 (n,K) float, out_idx (n,K) int64 (either may be NULL).                                        */
int avc_knn(avc_ctx* ctx, const float* query /*[dev] (n,3)*/, int64_t n, const float* ref /*[dev] (m,3)*/, int m, int K,
            float* out_d2 /*[dev]*/, int64_t* out_idx /*[dev]*/, void* stream);
/* validity flag of the dense grid (dataset/avatarcap_dataset.py:114-116): out_flag[i] = (min_j |q_i - ref_j|^2 < radius^2), exact,
 * bounded search in a uniform grid over the reference vertices (m >= 512). radius = 0.1 in the reference. */
int avc_near_flag(avc_ctx* ctx, const float* query /*[dev] (n,3)*/, int64_t n, const float* ref /*[dev] (m,3)*/, int m, double radius,
                  uint8_t* out_flag /*[dev] (n)*/, void* stream);
/* SmplUtil.calculate_lbs (smpl_util.py:24-39): out (n,24) */
int avc_lbs_weights(avc_ctx* ctx, const float* pts /*[dev]*/, int64_t n, const float* cano_verts /*[dev] (m,3)*/, int m,
                    const float* skin_weights /*[dev] (m,24)*/, float* out_lbs /*[dev] (n,24)*/, void* stream);
/* SmplUtil.skinning (smpl_util.py:58-74): out_pts (n,3), out_mats (n,4,4) or NULL */
int avc_skin_points(avc_ctx* ctx, const float* pts /*[dev]*/, const float* lbs /*[dev] (n,24)*/,
                    const float* jnt_mats /*[dev] (24,4,4)*/, int64_t n, float* out_pts, float* out_mats, void* stream);
/* SmplUtil.skinning_normal (smpl_util.py:76-81): rotation block only */
int avc_skin_normals(avc_ctx* ctx, const float* normals /*[dev]*/, const float* lbs /*[dev]*/,
                     const float* jnt_mats /*[dev]*/, int64_t n, float* out_normals, void* stream);
/* fused calculate_lbs + skinning (+ normals): what main.py:385-389 / 451-453 do per mesh */
int avc_skin_mesh(avc_ctx* ctx, const float* verts /*[dev]*/, const float* normals /*[dev]|NULL*/, int64_t n,
                  const float* cano_verts /*[dev]*/, int m, const float* skin_weights /*[dev]*/,
                  const float* jnt_mats /*[dev]*/, float* out_verts, float* out_normals /*|NULL*/, void* stream);
/* posed -> canonical warp of GeoTexAvatar.forward (arch_avatar.py:189-205): KNN-1 vs live SMPL, gather skin
 * weights, inverse LBS, normalise to bounds, trilinear sample of the (X,Y,Z,24) blend-weight volume
 * (CanoBlendWeightVolume.forward :152-165), inverse LBS again. live2cano = inv(cano2live) is computed by the caller.
 * out_cano (n,3); out_near (n) uint8 = (d2 < 0.08^2).                                                     */
int avc_posed_to_cano(avc_ctx* ctx, const float* wpts /*[dev]*/, int64_t n, const float* live_verts /*[dev] (m,3)*/, int m,
                      const float* skin_weights /*[dev] (m,24)*/, const float* live2cano_mats /*[dev] (24,4,4)*/,
                      const float bounds[6] /*[host]*/, const float* weight_volume /*[dev] (X,Y,Z,24)*/, const int vdims[3],
                      float* out_cano /*[dev]*/, uint8_t* out_near /*[dev]*/, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* dataset-side precompute ("next" row: dataset/avatarcap_dataset.py:110-125)                       */
/* ---------------------------------------------------------------------------------------------- */
/* trimesh.contains on the dense grid (avatarcap_dataset.py:120-124): out_inside (Rx,Ry,Rz) uint8 = 1 where the grid point of
 * generate_volume_points lies inside the closed triangle mesh (verts (nv,3) f32, faces (nf,3) int32), by crossing parity along +z. */
int avc_inside_volume(avc_ctx* ctx, const float* verts /*[dev]*/, int nv, const int32_t* faces /*[dev]*/, int nf, const float bounds[6] /*[host]*/,
                      const int res[3], uint8_t* out_inside /*[dev]*/, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* vertex-colour evaluation driver ("next" row: NerfRenderer.render + raw2outputs, main.py:464-485) */
/* ---------------------------------------------------------------------------------------------- */
/* NerfRenderer.get_wsampling_points / get_density_color (network/arch_avatar.py:244-281), eval mode: out_pts (n_rays*S,3),
 * out_z (n_rays*S) = near*(1-t)+far*t with t = linspace(0,1,S), out_dists (n_rays*S) = z[i+1]-z[i] (last repeated).       */
int avc_ray_samples(avc_ctx* ctx, const float* ray_o /*[dev] (n_rays,3)*/, const float* ray_d /*[dev]*/, const float* near /*[dev]*/,
                    const float* far /*[dev]*/, int64_t n_rays, int n_samples, float* out_pts, float* out_z, float* out_dists, void* stream);
/* GeoTexAvatar.forward post-processing (arch_avatar.py:220-231): alpha = 0 outside cano_bounds or far from SMPL, then
 * 1-exp(-alpha*dist); out_raw (n,4) = [rgb, alpha]. cano_q = warped canonical points (p + offset).                       */
int avc_nerf_raw(avc_ctx* ctx, const float* cano_q /*[dev] (n,3)*/, const uint8_t* near_flag /*[dev] (n)*/, const float* rgb /*[dev] (n,3)*/,
                 const float* alpha_raw /*[dev] (n)*/, const float* dists /*[dev] (n)*/, const float bounds[6] /*[host]*/, int64_t n,
                 float* out_raw /*[dev] (n,4)*/, void* stream);
/* raw2outputs (utils/nerf_util.py:185-212): rgb_map (n_rays,3), acc_map (n_rays), depth_map (n_rays)                     */
int avc_composite(avc_ctx* ctx, const float* raw /*[dev] (n_rays,S,4)*/, const float* z_vals /*[dev] (n_rays,S)*/, int64_t n_rays, int n_samples,
                  int white_bkgd, float* out_rgb, float* out_acc, float* out_depth, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* off-screen rasteriser + normal canonicalisation ("next" row 4: the GL passes between the two field evaluations)   */
/* ---------------------------------------------------------------------------------------------- */
#define AVC_RASTER_CULL_BACK 1   /* glEnable(GL_CULL_FACE): drop clockwise triangles (utils/renderer.py:438) */
#define AVC_RASTER_FLIP_X 2      /* mirror the image left-right (cv.flip(img, 1), utils/visualize_util.py:51) */
/* Renderer.set_model + set_mvp_mat + render (utils/renderer.py:403-451) with the 'vertex_attribute' shader (:9-29), or the
 * 'position' shader (:32-51) when attrs is NULL (attribute = object-space vertex position). Indexed mesh, or a triangle soup
 * (faces NULL: glDrawArrays over 3*n_faces consecutive vertices). mvp row-major (glUniformMatrix4fv(..., GL_TRUE, mvp)).
 * out_image (height, width, channels) float32, channels 3 or 4 (RGBA, alpha 1 where covered, 0 on the background),
 * row 0 = top (data[::-1], :448). Depth test GL_LESS on a 24-bit depth, pixel-centre sampling, top-left fill rule.        */
int avc_rasterize(avc_ctx* ctx, const float* verts /*[dev] (n_verts,3)*/, int64_t n_verts, const int32_t* faces /*[dev] (n_faces,3)|NULL*/,
                  int64_t n_faces, const float* attrs /*[dev] (n_verts,3)|NULL*/, const float mvp[16] /*[host]*/, int width, int height,
                  const float bg[3] /*[host]|NULL*/, int flags, int channels, float* out_image /*[dev]*/, void* stream);
/* per-vertex part of canonicalize_normal_map (normal_fusion/normal_fusion.py:27-62): project the live vertices with the pinhole
 * camera (mv world->camera row-major, fx fy cx cy), nearest-sample the rendered position map (visibility: |v - p| < 0.05) and the
 * image normal map, flip y/z, rotate by inv(mv) and by the inverse of each vertex's skinning matrix; zeros where not valid.
 * position_map (height,width,pos_channels) with pos_channels 3 or 4; normal_map (height,width,3); out_normals (n,3).               */
int avc_canonicalize_normals(avc_ctx* ctx, const float* live_verts /*[dev] (n,3)*/, const float* vert_mats /*[dev] (n,4,4)*/, int64_t n,
                             const float mv[16] /*[host]*/, float fx, float fy, float cx, float cy, const float* position_map /*[dev]*/,
                             int pos_channels, const float* normal_map /*[dev]*/, int height, int width, float* out_normals /*[dev]*/,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVATARCAP_B200_H */
