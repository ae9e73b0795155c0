// Multi-GPU slab exchange over peer memory: the ONE exchange step of the path.
//
// The volume is sharded in x-slabs, one process per GPU. Marching cubes on a slab needs 2 planes of the slab below and 3 planes of
// the slab above (utils/recon_util.py:9-48 normal stencil + the +0.5-voxel quirk :65-66; seam rule in avatarcap_b200/shard.py).
// Round 1 moved them with an all_gather of every rank's 5 planes to every rank followed by a torch.cat of the whole padded slab
// (a 67 MB+ copy). Here every rank owns ONE padded buffer [flags | lo halo | own planes | hi halo] allocated by the library and
// exported with cudaIpcGetMemHandle; the field kernel writes its occupancy straight into the `own` region, and
//
//   halo_push_kernel   stores this rank's boundary planes into the two neighbours' halo regions THROUGH THE MAPPED PEER POINTERS
//                      (st.global over NVLink / NVSwitch), fences at system scope and bumps the neighbours' arrival epochs;
//   halo_wait_kernel   (stream-ordered, before the marching-cubes kernels) spins until both neighbours' epochs have arrived;
//   halo_ack_kernel    (after the marching-cubes kernels) tells the neighbours their pushed planes have been consumed, so that the
//                      next frame's push cannot overwrite planes that are still being read.
//
// No host synchronisation, no NCCL call and no staging copy on the data path. All spins carry a wall-clock watchdog
// (%globaltimer): a peer that never arrives aborts the launch with a sticky error instead of hanging the GPU.
#include "common.cuh"

namespace {

struct __align__(16) ShardFlags {
  unsigned long long from_lo;   // epoch of the planes the LOWER neighbour has stored into my lo-halo region
  unsigned long long from_hi;   // ... the UPPER neighbour into my hi-halo region
  unsigned long long ack_lo;    // epoch up to which the LOWER neighbour has consumed what I pushed to it
  unsigned long long ack_hi;    // ... the UPPER neighbour
  unsigned int done_blocks;     // push kernel: blocks that finished their part of the copy
  unsigned int pad[23];
};
static_assert(sizeof(ShardFlags) <= AVC_SHARD_HEADER_BYTES, "flags must fit the buffer header");

constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void spin_until_ge(const volatile unsigned long long* flag, unsigned long long want, const char* what) {
  if (*flag >= want) return;
  const unsigned long long t0 = gtime_ns();
  while (*flag < want) {
    __nanosleep(200);
    if (gtime_ns() - t0 > SPIN_TIMEOUT_NS) { printf("avatarcap_b200: halo exchange timed out waiting for %s (epoch %llu)\n", what, want); __trap(); }
  }
}

__global__ void __launch_bounds__(256) halo_push_kernel(const float* __restrict__ own, int64_t plane, int nx, int n_to_lo, float* dst_lo, int n_to_hi,
                                                        float* dst_hi, ShardFlags* mine, ShardFlags* peer_lo, ShardFlags* peer_hi,
                                                        unsigned long long epoch) {
  if (threadIdx.x == 0) {
    // the neighbours must have consumed the planes of the previous epoch before they are overwritten
    if (peer_lo) spin_until_ge(&mine->ack_lo, epoch - 1, "the lower neighbour's ack");
    if (peer_hi) spin_until_ge(&mine->ack_hi, epoch - 1, "the upper neighbour's ack");
  }
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (peer_lo) {                                                   // my FIRST n_to_lo planes -> the lower neighbour's hi halo
    const int64_t n = plane * n_to_lo;
    if (((reinterpret_cast<uintptr_t>(own) | reinterpret_cast<uintptr_t>(dst_lo)) & 15) == 0 && (n & 3) == 0) {
      const float4* s = reinterpret_cast<const float4*>(own); float4* d = reinterpret_cast<float4*>(dst_lo);
      for (int64_t i = t0; i < n / 4; i += stride) d[i] = s[i];
    } else {
      for (int64_t i = t0; i < n; i += stride) dst_lo[i] = own[i];
    }
  }
  if (peer_hi) {                                                   // my LAST n_to_hi planes -> the upper neighbour's lo halo
    const int64_t n = plane * n_to_hi;
    const float* src = own + plane * (nx - n_to_hi);
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_hi)) & 15) == 0 && (n & 3) == 0) {
      const float4* s = reinterpret_cast<const float4*>(src); float4* d = reinterpret_cast<float4*>(dst_hi);
      for (int64_t i = t0; i < n / 4; i += stride) d[i] = s[i];
    } else {
      for (int64_t i = t0; i < n; i += stride) dst_hi[i] = src[i];
    }
  }
  __threadfence_system();                                          // my stores are visible system-wide before the arrival epoch is
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&mine->done_blocks, 1u);
    if (prev == gridDim.x - 1) {                                   // last block: every block's planes have been fenced
      mine->done_blocks = 0;
      __threadfence_system();
      if (peer_lo) *reinterpret_cast<volatile unsigned long long*>(&peer_lo->from_hi) = epoch;
      if (peer_hi) *reinterpret_cast<volatile unsigned long long*>(&peer_hi->from_lo) = epoch;
      __threadfence_system();
    }
  }
}

__global__ void halo_wait_kernel(ShardFlags* mine, int need_lo, int need_hi, unsigned long long epoch) {
  if (need_lo) spin_until_ge(&mine->from_lo, epoch, "the lower neighbour's planes");
  if (need_hi) spin_until_ge(&mine->from_hi, epoch, "the upper neighbour's planes");
  __threadfence_system();
}

__global__ void halo_ack_kernel(ShardFlags* peer_lo, ShardFlags* peer_hi, unsigned long long epoch) {
  __threadfence_system();
  if (peer_lo) *reinterpret_cast<volatile unsigned long long*>(&peer_lo->ack_hi) = epoch;     // I am its upper neighbour
  if (peer_hi) *reinterpret_cast<volatile unsigned long long*>(&peer_hi->ack_lo) = epoch;     // I am its lower neighbour
}

// seam rule of the slab meshes (avatarcap_b200/shard.py): local vertex ids >= n_own refer to the NEXT slab's first vertices
__global__ void renumber_faces_kernel(int32_t* __restrict__ faces, int64_t n3, int32_t n_own, int32_t base_own, int32_t base_next) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t f = faces[i];
    faces[i] = f < n_own ? f + base_own : f - n_own + base_next;
  }
}

}  // namespace

extern "C" int avc_shard_alloc(avc_ctx* ctx, size_t data_bytes, void** out_base, void* out_handle) {
  if (!ctx || !out_base || !out_handle) return avc_fail(ctx, AVC_EINVAL, "avc_shard_alloc: NULL argument");
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  void* p = nullptr;
  const size_t bytes = AVC_SHARD_HEADER_BYTES + data_bytes;
  AVC_CUDA(ctx, cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, AVC_SHARD_HEADER_BYTES);
  if (e != cudaSuccess) { cudaFree(p); return avc_check_cuda(ctx, e, "avc_shard_alloc: memset"); }
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == AVC_IPC_HANDLE_BYTES, "IPC handle size");
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); return avc_fail(ctx, AVC_ECUDA, "avc_shard_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
  memcpy(out_handle, &h, sizeof(h));
  AVC_CUDA(ctx, cudaDeviceSynchronize());
  *out_base = p;
  return AVC_OK;
}

extern "C" int avc_shard_open(avc_ctx* ctx, const void* handle, void** out_base) {
  if (!ctx || !handle || !out_base) return avc_fail(ctx, AVC_EINVAL, "avc_shard_open: NULL argument");
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h; memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return avc_fail(ctx, AVC_ECUDA, "avc_shard_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
  *out_base = p;
  return AVC_OK;
}

extern "C" int avc_shard_close(avc_ctx* ctx, void* base) {
  if (!ctx) return AVC_EINVAL;
  if (!base) return AVC_OK;
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  AVC_CUDA(ctx, cudaIpcCloseMemHandle(base));
  return AVC_OK;
}

extern "C" int avc_shard_free(avc_ctx* ctx, void* base) {
  if (!ctx) return AVC_EINVAL;
  if (!base) return AVC_OK;
  AVC_CUDA(ctx, cudaSetDevice(ctx->device));
  AVC_CUDA(ctx, cudaDeviceSynchronize());
  AVC_CUDA(ctx, cudaFree(base));
  return AVC_OK;
}

static inline float* shard_data(void* base) { return reinterpret_cast<float*>(reinterpret_cast<char*>(base) + AVC_SHARD_HEADER_BYTES); }

extern "C" int avc_halo_push(avc_ctx* ctx, void* mine, void* peer_lo, void* peer_hi, int64_t plane_elems, int64_t own_off, int nx, int n_to_lo,
                             int64_t lo_dst_off, int n_to_hi, int64_t hi_dst_off, uint64_t epoch, void* stream) {
  if (!ctx || !mine) return avc_fail(ctx, AVC_EINVAL, "avc_halo_push: NULL argument");
  if (plane_elems <= 0 || nx <= 0 || n_to_lo < 0 || n_to_hi < 0 || n_to_lo > nx || n_to_hi > nx || epoch == 0)
    return avc_fail(ctx, AVC_EINVAL, "avc_halo_push: bad slab description");
  if (!peer_lo && !peer_hi) return AVC_OK;
  const int64_t work = plane_elems * (int64_t)((peer_lo ? n_to_lo : 0) + (peer_hi ? n_to_hi : 0));
  int blocks = (int)((work / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > ctx->sm_count) blocks = ctx->sm_count;     // every block spins on the acks first: all of them must be co-resident
  halo_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(shard_data(mine) + own_off, plane_elems, nx, n_to_lo,
                                                             peer_lo ? shard_data(peer_lo) + lo_dst_off : nullptr, n_to_hi,
                                                             peer_hi ? shard_data(peer_hi) + hi_dst_off : nullptr,
                                                             reinterpret_cast<ShardFlags*>(mine), reinterpret_cast<ShardFlags*>(peer_lo),
                                                             reinterpret_cast<ShardFlags*>(peer_hi), (unsigned long long)epoch);
  AVC_LAUNCH_CHECK(ctx, "halo_push_kernel");
  return AVC_OK;
}

extern "C" int avc_halo_wait(avc_ctx* ctx, void* mine, int need_lo, int need_hi, uint64_t epoch, void* stream) {
  if (!ctx || !mine) return avc_fail(ctx, AVC_EINVAL, "avc_halo_wait: NULL argument");
  if (!need_lo && !need_hi) return AVC_OK;
  halo_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<ShardFlags*>(mine), need_lo, need_hi, (unsigned long long)epoch);
  AVC_LAUNCH_CHECK(ctx, "halo_wait_kernel");
  return AVC_OK;
}

extern "C" int avc_halo_ack(avc_ctx* ctx, void* peer_lo, void* peer_hi, uint64_t epoch, void* stream) {
  if (!ctx) return AVC_EINVAL;
  if (!peer_lo && !peer_hi) return AVC_OK;
  halo_ack_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<ShardFlags*>(peer_lo), reinterpret_cast<ShardFlags*>(peer_hi), (unsigned long long)epoch);
  AVC_LAUNCH_CHECK(ctx, "halo_ack_kernel");
  return AVC_OK;
}

extern "C" int avc_renumber_faces(avc_ctx* ctx, int32_t* faces, int64_t n_faces, int64_t n_own, int64_t base_own, int64_t base_next, void* stream) {
  if (!ctx || (n_faces > 0 && !faces)) return avc_fail(ctx, AVC_EINVAL, "avc_renumber_faces: NULL argument");
  if (n_faces <= 0) return AVC_OK;
  if (base_next > 0x7fffffffLL || base_own > 0x7fffffffLL) return avc_fail(ctx, AVC_EINVAL, "avc_renumber_faces: merged mesh too large for int32 indices");
  const int64_t n3 = n_faces * 3;
  int blocks = (int)((n3 + 255) / 256); if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  renumber_faces_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(faces, n3, (int32_t)n_own, (int32_t)base_own, (int32_t)base_next);
  AVC_LAUNCH_CHECK(ctx, "renumber_faces_kernel");
  return AVC_OK;
}
