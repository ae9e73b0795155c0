// tcgen05 (5th-gen tensor core) implementation of the fused per-point networks for sm_100a -- the product path (AVC_IMPL_TC2, what
// AVC_IMPL_AUTO resolves to on a B200; the single-CTA kernel of round 1 it grew out of is gone, AVC_IMPL_TC is an alias).
//
// Two CTAs of a cluster (= the two SMs of a TPC) evaluate a PAIR of 128-point tiles through the whole network:
//   * every fully-connected layer is a chain of tcgen05.mma.cta_group::2 (M = 256: 128 points per CTA, N <= 256 channels issued as
//     N = 128 halves, K = 16 per instruction) issued by the leader CTA, fp32 accumulators in TMEM; each CTA streams only HALF of every
//     weight slab, so L2->SM weight traffic and MMA instructions per SM halve against the single-CTA kernel;
//   * operands are fp16 hi/lo pairs (x = hi + lo to ~2^-22): each product is 3 MMAs (hi*hi + hi*lo + lo*hi), which keeps the
//     field within the reference's 1e-4 tolerance where a single bf16/fp16/tf32 pass cannot (BASELINE.md precision probe);
//   * activations NEVER leave the SM: the epilogue warps read the fp32 accumulator from TMEM (tcgen05.ld, 16-column pieces, software
//     pipelined), apply scale/bias/activation, split into fp16 hi / negated lo and write it back IN PLACE (tcgen05.st) as the A operand
//     (TS-mode MMA) of the next layer -- 32 fp32 accumulator columns become 16 hi + 16 lo packed columns. The two 256-column halves of
//     TMEM ping-pong between "A of layer l" and "D of layer l";
//   * weights stream from L2 through a 4-stage shared-memory ring filled by the bulk-copy engine (cp.async.bulk; packer.py stores each
//     layer as a stream in exactly the order and canonical K-major core-matrix layout the ring consumes: no tensor map, no swizzle);
//   * the skip-connection inputs (bilinear feature gather / positional encoding) live in shared memory as SS-mode A operands; the next
//     tile's gather is prefetched into a second buffer during the current tile's accumulator waits;
//   * the <= 3-output heads (offsets, geometry, colour, recon) are evaluated on the CUDA cores inside the preceding epilogue.
// Roles (352 threads): warps 0-7 compute/epilogue (warp w owns TMEM lanes 32*(w%4)..+31, i.e. 32 points; warps w and w+4 alternate
//        column chunks), warps 8-9 alternate as MMA issuers on the leader CTA (the peer's warp 8 relays its ring barriers), warp 10
//        streams weights. All roles walk one op program built on the host (build_ops). DESIGN.md 4.1 / 4.1b has the measured history.
//
// Reference call sites restated: see field_simt.cu (same networks, same order of operations per layer).
#include "common.cuh"

#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int TILE = 128;                 // points per tile == UMMA M
constexpr int NT = 352;                   // 8 compute warps + 2 alternating MMA-issuer warps + weight-producer warp
constexpr int N_STAGES = 4;
constexpr int STAGE_KSTEPS = 4;           // k-steps per ring stage (== two 32-column A chunks)
constexpr int STAGE_BYTES = STAGE_KSTEPS * 4096;   // per k-step of one N-half, THIS CTA's 64 of the 128 weight rows: hi 2 KB + lo 2 KB
constexpr int SKIP_KSTEPS = 5;            // up to K=80 of skip input
constexpr int SKIP_BYTES = SKIP_KSTEPS * 8192;   // per k-step: hi slab 4 KB + lo slab 4 KB (128 rows x 16 k x 2 B)
constexpr int MAX_OPS = 24;
constexpr int DOTW_F4_MAX = 520;          // head weights as float4 {w0,w1,w2,0} per input channel + one bias row per head (avatar: 257+129+129)
constexpr int SB_FLOATS_MAX = 8704;       // scale/bias pairs of all layers (avatar: 4272 channels*2)

enum { EPI_HIDDEN = 1, EPI_WARP_OUT = 2, EPI_GEO_OUT = 3, EPI_CLR_OUT = 4, EPI_RECON_OUT = 5 };

struct TcOp {
  int layer;          // index into the blob's layer table
  int n;              // MMA N of this op (<= 256)
  int n_row_off;      // first weight row (R1 is split in two N=256 halves)
  int np;             // padded rows of the layer's slabs
  int ks_smem, ks_smem_w0;   // k-steps fed from the shared-memory skip buffer (issued first) and their first weight k-step
  int ks_tmem, ks_tmem_w0;   // k-steps fed from TMEM and their first weight k-step
  int a_col, d_col;   // TMEM columns of the A operand (packed hi/lo chunks) and of the accumulator
  int accumulate;     // keep the accumulator contents (continuation of a split K loop)
  int wait_epi;       // MMA issue must wait for the compute warps (input staged / head read / PE written)
  int commit_d;       // signal the epilogue when this op's MMAs are complete
  int epi;            // epilogue kind (0 = none)
  int act;
  int sb_off;         // float offset of {scale,bias} pairs in shared memory
  int signal_done;    // compute warps arrive on epi_done after this op's epilogue
  unsigned int w_off; // byte offset of this op's weight stream in the f16 section
  int wait_a;         // the TMEM A chunks are produced by the preceding epilogue (0: already complete, e.g. the colour head re-reads s7)
  int dot_w;          // >= 0: the op's activations feed a tiny (<= 3 output) linear head evaluated on the CUDA cores in this op's
                      // epilogue; float4 index of that head's weights in s_dotw, `epi` then names the OUTPUT stage. -1: plain hidden layer
  int dot_k;          // K of that head (channels of this op)
  int dot_layer;      // the head's layer in the blob (f32 section: W[n][K], scale[n], bias[n])
  int passes;         // 3: hi*hi + hi*lo + lo*hi (the 1e-4 occupancy budget needs it); 1: hi*hi only (the colour head: an 8-bit colour, emulated
                      // error 3e-6 on rgb, tests/diag_pass_pruning.py -> profiles/r2_pass_pruning.txt)
};

struct TcArgs {
  const float* pts; int64_t n;          // pts == NULL: dense-grid mode, point g = grid point of linear index g (+ the slab offset)
  float gb[3], gl[3], gstep[3]; int gr[3]; int gx_first;   // grid: bmin, bmax - bmin, 1/(res-1), resolution, first i-plane   (avatarcap_dataset.py:312-326)
  float cx, cy, cz;
  const float* map; int mC, mH, mW;
  float* out0; float* out_off; float* out_rgb; float* out_alpha;   // out0 = occ (avatar) or ov (recon)
  int if_type, mode, kind;
  const unsigned char* w16; const float* f32; const AvcBlobHeader* hdr;
  int n_ops; TcOp ops[MAX_OPS];   // the op program, built on the host: lives in the constant bank -> uniform registers
  int dbg;            // reserved (debug experiments are compiled out)
  long long* trace;   // optional timeline buffer (debug): [tile<4][op<24][8 events] clock64 stamps of CTA 0
};

struct __align__(16) TcShared {
  unsigned long long full[N_STAGES], empty[N_STAGES];
  unsigned long long peer_full[N_STAGES];   // leader only: the peer CTA's half of the stage has landed
  unsigned long long a_ready[8];
  unsigned long long d_ready[2], epi_done;   // d_ready[h]: N-half h of the current op is complete
  unsigned int tmem_base; int pad[3];
  float4 dot_xch[2][TILE];                  // partial head sums of the two column groups of a row
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    // watchdog: a protocol bug must abort the launch (sticky error reported through the C ABI) instead of hanging the GPU
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 6000000000LL) {
      printf("avatarcap_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, a, parity);
      __trap();
    }
  }
}
// cluster-scope acquire wait: only where the PEER CTA's generic-proxy writes (its skip operand in shared memory) must be observed
__device__ __forceinline__ void mbar_wait_cluster(void* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  long long t0 = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (done) break;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 6000000000LL) { printf("avatarcap_b200: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
// relaxed arrive on the leader's barrier: for hand-offs whose payload lives in TMEM / was moved by the async proxy (the tcgen05 fences and
// the async-proxy completion order it); a release at cluster scope on every chunk costs an L1 invalidation each time
__device__ __forceinline__ void mbar_arrive_leader_relaxed(void* bar, uint32_t rank) {
  if (rank == 0) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the LEADER CTA's copy of a barrier (local arrive on the leader itself, remote arrive from the peer)
__device__ __forceinline__ void mbar_arrive_leader(void* bar, uint32_t rank) {
  if (rank == 0) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0u));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// same with an A-collector hint: 1 = fill (keep A in the collector buffer), 2 = lastuse (take A from the collector buffer)
__device__ __forceinline__ void mma_ts_c(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc, int coll) {
  if (coll == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = 128 B between the two k-halves, SBO = 256 B between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
// kind::f16: D fp32 (bit 4), A/B fp16 (0), both K-major, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)((2 * TILE) >> 4) << 24); }

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
// 16 columns, NO wait: the caller overlaps the load with arithmetic on the previous piece and calls tmem_ld_wait() before use
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
// wait for the outstanding tcgen05.ld's; `r` (the registers of the load being waited for) is threaded through the statement as an
// in/out operand so that no use of those registers can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t r[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t r[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
               "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ math helpers
template <int ACT>
__device__ __forceinline__ float act_tc(float v) {
  if (ACT == AVC_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == AVC_ACT_LRELU) return fmaxf(v, v * 0.02f);          // == v > 0 ? v : 0.02 v   (nn.LeakyReLU(0.02), mlp.py:11)
  if (ACT == AVC_ACT_SOFTPLUS) {
    // softplus(v) = max(v,0) + log1p(exp(-|v|)); exact to ~1e-7 abs, and == v for v > 20 like nn.Softplus(threshold=20)
    const float t = exp2f(-fabsf(v) * 1.4426950408889634f);
    return fmaf(__log2f(1.f + t), 0.6931471805599453f, fmaxf(v, 0.f));
  }
  return v;
}
// sin/cos for |x| < ~1e4 (the PE arguments reach 2^9 * |q| ~ 600): Cody-Waite reduction by pi/2 with three fused steps, then the
// fdlibm single-precision kernels on [-pi/4, pi/4]. Max abs error 9.2e-8 over the PE range (numpy float32 sin: 6.9e-8); ~25
// instructions for the pair, a third of sincosf().
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
  const float k = rintf(x * 0.63661977236758138f);
  float r = fmaf(k, -1.5707964e+00f, x);
  r = fmaf(k, 4.371139e-08f, r);
  r = fmaf(k, 1.7151245e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(2.7557314297e-06f, r2, -1.9841270114e-04f); ps = fmaf(ps, r2, 8.3333337680e-03f); ps = fmaf(ps, r2, -1.6666667163e-01f);
  const float sn = fmaf(r * r2, ps, r);
  float pc = fmaf(-2.7557314297e-07f, r2, 2.4801587642e-05f); pc = fmaf(pc, r2, -1.3888889225e-03f); pc = fmaf(pc, r2, 4.1666667908e-02f);
  const float cs = fmaf(r2 * r2, pc, fmaf(r2, -0.5f, 1.f));
  const int q = (int)k;
  const float a = (q & 1) ? cs : sn, b = (q & 1) ? sn : cs;
  s = (q & 2) ? -a : a;
  c = ((q + 1) & 2) ? -b : b;
}
// split (v0, v1) into packed fp16 hi and lo words (element with the lower k index in the low 16 bits)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// write 8 consecutive k values (one 16-byte core-matrix row) of the skip operand for point row r, k-group g (k = 8g..8g+7)
__device__ __forceinline__ void skip_store8(unsigned char* skip, int r, int g, const float v[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  unsigned char* base = skip + (g >> 1) * 8192 + (r >> 3) * 256 + (g & 1) * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + 4096) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// TMEM-resident A operand: packed fp16 hi, and the NEGATED residual  nlo = hi - v  (one mixed-precision FHADD per value instead of
// convert-back + subtract); the lo*hi pass of the TS-mode MMAs sets the descriptor's negate-A bit, so the product is unchanged.
__device__ __forceinline__ void split2_neg(float v0, float v1, uint32_t& hi, uint32_t& nlo) {
  const __half2 h = __floats2half2_rn(v0, v1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const unsigned short h0 = (unsigned short)(hi & 0xffffu), h1 = (unsigned short)(hi >> 16);
  float r0, r1;
  asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(r0) : "h"(h0), "f"(v0));
  asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(r1) : "h"(h1), "f"(v1));
  const __half2 l = __floats2half2_rn(r0, r1);
  nlo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One 32-column accumulator chunk: TMEM -> scale/bias/activation -> fp16 hi/lo -> TMEM, in place (A operand of the next layer).
// sbc points at {scale,bias} pairs of the chunk's 32 channels.
template <int ACT>
__device__ __forceinline__ void hidden_chunk(uint32_t taddr, const float* __restrict__ sbc) {
  float v[32];
  tmem_ld32(taddr, v);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 s4 = *reinterpret_cast<const float4*>(sbc + 4 * i);   // {scale0, bias0, scale1, bias1}
    v[2 * i] = fmaf(v[2 * i], s4.x, s4.y);
    v[2 * i + 1] = fmaf(v[2 * i + 1], s4.z, s4.w);
  }
  if (ACT == AVC_ACT_SOFTPLUS) {
    // softplus(v) = max(v,0) + log1p(t), t = 2^(-|v| log2e) in (0,1] (ex2.approx.ftz). log1p(t) alternates between the two pipes the
    // epilogue is bound by: even values take ln2 * lg2.approx(1+t) (XU pipe, 3 instructions), odd values a degree-7 near-minimax
    // polynomial t*q(t) (FMA pipe, 7 instructions, max abs error 3.0e-7 -- the same order as lg2.approx). All-XU was XU-bound
    // (2 MUFU/value at 16 lanes/clk/SM), all-polynomial issue-bound.
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = ex2_ftz(-fabsf(v[i]) * 1.4426950408889634f);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float l = lg2_ftz(1.f + t[i]);
      float q = 1.076442841e-02f;
      q = fmaf(q, t[i + 1], -5.514492467e-02f); q = fmaf(q, t[i + 1], 1.346741915e-01f); q = fmaf(q, t[i + 1], -2.258978188e-01f);
      q = fmaf(q, t[i + 1], 3.282421529e-01f); q = fmaf(q, t[i + 1], -4.994717836e-01f); q = fmaf(q, t[i + 1], 9.999811649e-01f);
      v[i] = fmaf(l, 0.6931471805599453f, fmaxf(v[i], 0.f));
      v[i + 1] = fmaf(q, t[i + 1], fmaxf(v[i + 1], 0.f));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = act_tc<ACT>(v[i]);
  }
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) split2_neg(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  tmem_st16(taddr, hi);            // k-steps 2c, 2c+1: hi in columns [0,16) of the chunk
  tmem_st16(taddr + 16u, lo);      //                   -lo in columns [16,32)
  tmem_st_wait();
}

// scale/bias/activation of ONE 16-column piece held in registers (raw accumulator bits in)
template <int ACT>
__device__ __forceinline__ void act_piece(const uint32_t raw[16], const float* __restrict__ sbp, float v[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 s4 = *reinterpret_cast<const float4*>(sbp + 4 * i);   // {scale0, bias0, scale1, bias1}
    v[2 * i] = fmaf(__uint_as_float(raw[2 * i]), s4.x, s4.y);
    v[2 * i + 1] = fmaf(__uint_as_float(raw[2 * i + 1]), s4.z, s4.w);
  }
  if (ACT == AVC_ACT_SOFTPLUS) {           // see hidden_chunk
    float t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = ex2_ftz(-fabsf(v[i]) * 1.4426950408889634f);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float l = lg2_ftz(1.f + t[i]);
      float q = 1.076442841e-02f;
      q = fmaf(q, t[i + 1], -5.514492467e-02f); q = fmaf(q, t[i + 1], 1.346741915e-01f); q = fmaf(q, t[i + 1], -2.258978188e-01f);
      q = fmaf(q, t[i + 1], 3.282421529e-01f); q = fmaf(q, t[i + 1], -4.994717836e-01f); q = fmaf(q, t[i + 1], 9.999811649e-01f);
      v[i] = fmaf(l, 0.6931471805599453f, fmaxf(v[i], 0.f));
      v[i + 1] = fmaf(q, t[i + 1], fmaxf(v[i + 1], 0.f));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = act_tc<ACT>(v[i]);
  }
}
// ... + hi / -lo split (packed words out)
template <int ACT>
__device__ __forceinline__ void hidden_piece(const uint32_t raw[16], const float* __restrict__ sbp, uint32_t hi[8], uint32_t nlo[8]) {
  float v[16];
  act_piece<ACT>(raw, sbp, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) split2_neg(v[2 * i], v[2 * i + 1], hi[i], nlo[i]);
}
// ... + accumulation into a <= 3-output linear head in fp32 (wp: one float4 {w0,w1,w2,0} per channel, broadcast reads)
template <int ACT>
__device__ __forceinline__ void dot_piece(const uint32_t raw[16], const float* __restrict__ sbp, const float4* __restrict__ wp, float acc[3]) {
  float v[16];
  act_piece<ACT>(raw, sbp, v);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 w = wp[i];
    acc[0] = fmaf(v[i], w.x, acc[0]); acc[1] = fmaf(v[i], w.y, acc[1]); acc[2] = fmaf(v[i], w.z, acc[2]);
  }
}
// The last hidden layer of a head (256->3 warp offsets, 128->2 geometry, 128->3 colour, 128->1 recon) followed by its tiny output
// layer: the activations never go back to TMEM, the output layer is 3 FMAs per channel right here. As tensor-core ops these
// N=16 layers cost 3.6-4.2 k cycles of pure dependency latency each (arrive -> 24..48 tiny MMAs -> commit -> wait -> tcgen05.ld).
template <int ACT>
__device__ __forceinline__ void dot_pair(uint32_t t0, uint32_t t1, const float* __restrict__ sb0, const float* __restrict__ sb1,
                                         const float4* __restrict__ w0, const float4* __restrict__ w1, float acc[3]) {
  uint32_t ra[16], rb[16];
  tmem_ld16_issue(t0, ra); tmem_ld_wait(ra);
  tmem_ld16_issue(t0 + 16u, rb);
  dot_piece<ACT>(ra, sb0, w0, acc);
  tmem_ld_wait(rb);
  tmem_ld16_issue(t1, ra);
  dot_piece<ACT>(rb, sb0 + 32, w0 + 16, acc);
  tmem_ld_wait(ra);
  tmem_ld16_issue(t1 + 16u, rb);
  dot_piece<ACT>(ra, sb1, w1, acc);
  tmem_ld_wait(rb);
  dot_piece<ACT>(rb, sb1 + 32, w1 + 16, acc);
}

// Two 32-column chunks (c0, c1) of one accumulator half, software-pipelined in 16-column pieces: the TMEM load of piece p+1 is in
// flight while piece p is computed, and the stores are only waited for once per chunk, right before its a_ready arrive. The
// fixed tcgen05.ld / st / wait latencies (about 600 cycles per chunk when exposed) were half of a ReLU chunk's epilogue time.
// In-place layout of a chunk: piece p (fp32 columns 16p..16p+15) -> hi words in columns 8p..8p+7, -lo words in 16+8p..16+8p+7;
// piece 0's -lo lands on piece 1's fp32 columns, so piece 1 must be in registers (wait::ld) before piece 0 is stored.
template <int ACT>
__device__ __forceinline__ void hidden_pair(uint32_t t0, uint32_t t1, const float* __restrict__ sb0, const float* __restrict__ sb1,
                                            void* bar0, void* bar1, uint32_t rank, int lane) {
  uint32_t ra[16], rb[16], hi[8], nlo[8];
  tmem_ld16_issue(t0, ra); tmem_ld_wait(ra);
  tmem_ld16_issue(t0 + 16u, rb);
  hidden_piece<ACT>(ra, sb0, hi, nlo);
  tmem_ld_wait(rb);                                 // piece (c0,1) is in rb: its columns may now be overwritten
  tmem_st8(t0, hi); tmem_st8(t0 + 16u, nlo);
  tmem_ld16_issue(t1, ra);
  hidden_piece<ACT>(rb, sb0 + 32, hi, nlo);
  tmem_ld_wait(ra);
  tmem_st8(t0 + 8u, hi); tmem_st8(t0 + 24u, nlo);
  tmem_ld16_issue(t1 + 16u, rb);
  hidden_piece<ACT>(ra, sb1, hi, nlo);              // chunk c0's stores drain meanwhile
  tmem_st_wait(); tc_fence_before(); __syncwarp();
  if (lane == 0) mbar_arrive_leader_relaxed(bar0, rank);
  tmem_ld_wait(rb);
  tmem_st8(t1, hi); tmem_st8(t1 + 16u, nlo);
  hidden_piece<ACT>(rb, sb1 + 32, hi, nlo);
  tmem_st8(t1 + 8u, hi); tmem_st8(t1 + 24u, nlo);
  tmem_st_wait(); tc_fence_before(); __syncwarp();
  if (lane == 0) mbar_arrive_leader_relaxed(bar1, rank);
}

// torch.linspace(0, 1, steps)[q] in float32 exactly as ATen computes it (and make_grid_kernel restates it); step = fl(1 / (steps - 1)) comes from
// the host (an IEEE division there, no division subroutine in this kernel: its instruction footprint is on the critical path)
__device__ __forceinline__ float lin_coord(int q, int steps, float step) {
  if (steps <= 1) return 0.f;
  return q < steps / 2 ? __fmul_rn(step, (float)q) : __fsub_rn(1.f, __fmul_rn(step, (float)(steps - 1 - q)));
}

struct Taps { int i00, i01, i10, i11; float w00, w01, w10, w11; };
__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {   // == field_simt.cu (ATen grid_sample, border, align_corners)
  float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float x1 = x0 + 1.f, y1 = y0 + 1.f;
  Taps t;
  t.w00 = (x1 - ix) * (y1 - iy); t.w01 = (ix - x0) * (y1 - iy); t.w10 = (x1 - ix) * (iy - y0); t.w11 = (ix - x0) * (iy - y0);
  const int xi0 = (int)x0, yi0 = (int)y0;
  int xi1 = xi0 + 1, yi1 = yi0 + 1;
  if (xi1 > W - 1) { xi1 = W - 1; t.w01 = 0.f; t.w11 = 0.f; }
  if (yi1 > H - 1) { yi1 = H - 1; t.w10 = 0.f; t.w11 = 0.f; }
  t.i00 = yi0 * W + xi0; t.i01 = yi0 * W + xi1; t.i10 = yi1 * W + xi0; t.i11 = yi1 * W + xi1;
  return t;
}
__device__ __forceinline__ void gather8(const float* __restrict__ hwc, int C, const Taps& t, int c, float v[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i00 * C + c + 4 * h));
    const float4 b = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i01 * C + c + 4 * h));
    const float4 d = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i10 * C + c + 4 * h));
    const float4 e = __ldg(reinterpret_cast<const float4*>(hwc + (size_t)t.i11 * C + c + 4 * h));
    v[4 * h + 0] = ((a.x * t.w00 + b.x * t.w01) + d.x * t.w10) + e.x * t.w11;
    v[4 * h + 1] = ((a.y * t.w00 + b.y * t.w01) + d.y * t.w10) + e.y * t.w11;
    v[4 * h + 2] = ((a.z * t.w00 + b.z * t.w01) + d.z * t.w10) + e.z * t.w11;
    v[4 * h + 3] = ((a.w * t.w00 + b.w * t.w01) + d.w * t.w10) + e.w * t.w11;
  }
}

// point g of the launch: read from the caller's list (GRID = false) or generated from the grid index (GRID = true). Two kernel
// instantiations: the A/B against the round-1 library showed that a few hundred extra instructions in this kernel (a float-division
// subroutine, shuffle-coalesced I/O) cost 7-9 % -- the issuer warps' loop lives off the instruction cache.
template <bool GRID>
__device__ __forceinline__ void fetch_point(const TcArgs& a, int64_t g, float& px, float& py, float& pz);

// One step of the input stage (skip operand of the first layer) for point (px,py,pz) with bilinear taps t:
//   avatar: h0 in tensor-core order [f0..f63, x, y, z, 0...] (packer permutes the 67 columns accordingly)   arch_avatar.py:121-136
//           slices 0..3 = 8 feature channels each of this warp group's 32, slice 4 = the xyz / zero-pad rows (group 1)
//   recon : h0 = [f0..f31, z - cz, 0...]   arch_recon.py:62-70; slices 0..1 = features, slice 2 = z row (group 1)
struct TcArgs;
__device__ __forceinline__ void input_slice(const TcArgs& a, unsigned char* buf, int row, int grp, int sl, const Taps& t, float px, float py, float pz);

// debug timeline: event e of op `oi` in the CTA-local tile number `t` (only CTA 0, first 4 tiles)
__device__ __forceinline__ void trace_ev(long long* trace, int t, int oi, int e) {
  if (trace && blockIdx.x == 0 && t < 4) trace[(t * MAX_OPS + oi) * 8 + e] = clock64();
}

// ------------------------------------------------------------------------------------------------ op program
// TMEM regions: X = columns [0,256), Y = [256,512). See the file header for the ping-pong scheme.
void build_ops(TcArgs& S, const AvcBlobHeader* hdr, int kind, int mode, bool texture) {
  int n = 0, sb = 0;
  // A/B knob: AVC_CLR_PASSES=3 evaluates the colour head with the full split product like everything else
  const int clr_passes = [] { const char* e = getenv("AVC_CLR_PASSES"); return (e && atoi(e) == 3) ? 3 : 1; }();
  int sb_off[AVC_MAX_LAYERS];
  unsigned int stream_pos[AVC_MAX_LAYERS];      // running offset inside each layer's weight stream (pieces in op order)
  for (int l = 0; l < (int)hdr->n_layers; ++l) { sb_off[l] = sb; sb += 2 * hdr->layers[l].np; stream_pos[l] = (unsigned int)hdr->layers[l].tc_w_off; }
  auto add = [&](int layer, int nn, int row_off, int ks_s, int ks_s_w0, int ks_t, int ks_t_w0, int a_col, int d_col, int accum, int wait_epi,
                 int commit, int epi, int signal) {
    TcOp& o = S.ops[n++];
    const AvcLayerDesc& L = hdr->layers[layer];
    o.layer = layer; o.n = nn; o.n_row_off = row_off; o.np = L.np; o.ks_smem = ks_s; o.ks_smem_w0 = ks_s_w0; o.ks_tmem = ks_t; o.ks_tmem_w0 = ks_t_w0;
    o.a_col = a_col; o.d_col = d_col; o.accumulate = accum; o.wait_epi = wait_epi; o.commit_d = commit; o.epi = epi; o.act = L.act;
    o.sb_off = sb_off[layer] + 2 * row_off; o.signal_done = signal; o.wait_a = 1; o.dot_w = -1; o.dot_k = 0; o.dot_layer = -1; o.passes = 3;
    o.w_off = stream_pos[layer]; stream_pos[layer] += (unsigned int)(nn * 64 * (ks_s + ks_t));
  };
  int dot_pos = 0;
  // fold the head `head_layer` (n <= 3 outputs) into the epilogue of the op added last; `epi_out` is the head's output stage
  auto head = [&](int head_layer, int epi_out, int signal) {
    TcOp& o = S.ops[n - 1];
    o.dot_w = dot_pos; o.dot_k = o.n; o.dot_layer = head_layer; o.epi = epi_out; o.signal_done = signal;
    dot_pos += o.n + 1;
  };
  const int X = 0, Y = 256;
  if (kind == AVC_KIND_AVATAR) {
    if (mode != AVC_MODE_TEMPLATE_ONLY) {
      add(0, 256, 0, 5, 0, 0, 0, 0, X, 0, 1, 1, EPI_HIDDEN, 0);        // conv1: h0 (K=80, smem)
      add(1, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);
      add(2, 256, 0, 0, 0, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);
      add(3, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);       // x4 in Y
      add(4, 256, 0, 5, 0, 16, 5, Y, X, 0, 0, 1, EPI_HIDDEN, 0);       // conv5: [h0 | x4]; weight k-steps 0..4 = h0, 5..20 = x4
      add(5, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);
      add(6, 256, 0, 0, 0, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);       // x7 in X
      head(7, EPI_WARP_OUT, 0);                                        // offsets = out_layer(x7) in conv7's epilogue, which then writes the PE
    }
    if (mode != AVC_MODE_WARP_ONLY) {
      add(8, 256, 0, 4, 0, 0, 0, 0, X, 0, 1, 1, EPI_HIDDEN, 0);        // fc0: PE (K=64, smem)
      add(9, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);
      add(10, 256, 0, 0, 0, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);
      add(11, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);      // s4 in Y
      add(12, 256, 0, 4, 16, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);     // fc4: [s4 | PE]; weight k-steps 0..15 = s4, 16..19 = PE
      add(13, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);
      add(14, 256, 0, 0, 0, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);      // shared feature s7 in X (kept for the colour head)
      add(15, 128, 0, 0, 0, 16, 0, X, Y, 0, 0, 1, EPI_HIDDEN, 0);      // geo fc0 -> Y[0,128)
      head(16, EPI_GEO_OUT, texture ? 1 : 0);
      if (texture) {
        add(17, 256, 0, 0, 0, 16, 0, X, Y, 0, 1, 1, EPI_HIDDEN, 0);    // clr fc0 (waits until the geo head has been read out of Y)
        S.ops[n - 1].wait_a = 0;                                       // s7 was completed for geo fc0 already
        S.ops[n - 1].passes = clr_passes;
        add(18, 128, 0, 0, 0, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);    // clr fc1 -> X[0,128)
        S.ops[n - 1].passes = clr_passes;
        head(19, EPI_CLR_OUT, 0);
      }
    }
  } else {
    add(0, 256, 0, 3, 0, 0, 0, 0, X, 0, 1, 1, EPI_HIDDEN, 0);          // fc0 rows 0..255: y1a in X
    add(1, 256, 0, 0, 0, 16, 0, X, Y, 0, 0, 0, 0, 0);                  // fc1 over y1a (no commit: K loop continues)
    add(0, 256, 256, 3, 0, 0, 0, 0, X, 0, 0, 1, EPI_HIDDEN, 0);        // fc0 rows 256..511: y1b in X
    add(1, 256, 0, 3, 32, 16, 16, X, Y, 1, 0, 1, EPI_HIDDEN, 0);       // fc1 over h0 (k-steps 32..34) and y1b (16..31), accumulating
    add(2, 128, 0, 3, 16, 16, 0, Y, X, 0, 0, 1, EPI_HIDDEN, 0);        // fc2: [y2 | h0] -> X[0,128)
    head(3, EPI_RECON_OUT, 0);
  }
  S.n_ops = n;
}

template <bool GRID>
__device__ __forceinline__ void fetch_point(const TcArgs& a, int64_t g, float& px, float& py, float& pz) {
  px = py = pz = 0.f;
  if (g >= a.n) return;
  if (!GRID) { px = a.pts[g * 3]; py = a.pts[g * 3 + 1]; pz = a.pts[g * 3 + 2]; return; }
  // generate_volume_points (avatarcap_dataset.py:312-326): flat = (i*Ry + j)*Rz + k, point = linspace * (bmax - bmin) + bmin
  const unsigned int rz = (unsigned int)a.gr[2], ry = (unsigned int)a.gr[1];
  const unsigned int u = (unsigned int)g;                       // the launcher refuses grid slabs of 2^31 points or more
  const unsigned int t = u / rz; const int k = (int)(u - t * rz);
  const unsigned int ii = t / ry; const int j = (int)(t - ii * ry);
  const int i = (int)ii + a.gx_first;
  px = __fadd_rn(__fmul_rn(lin_coord(i, a.gr[0], a.gstep[0]), a.gl[0]), a.gb[0]);
  py = __fadd_rn(__fmul_rn(lin_coord(j, a.gr[1], a.gstep[1]), a.gl[1]), a.gb[1]);
  pz = __fadd_rn(__fmul_rn(lin_coord(k, a.gr[2], a.gstep[2]), a.gl[2]), a.gb[2]);
}

__device__ __forceinline__ void input_slice(const TcArgs& a, unsigned char* buf, int row, int grp, int sl, const Taps& t, float px, float py, float pz) {
  float v[8];
  if (a.kind == AVC_KIND_AVATAR) {
    if (sl < 4) { gather8(a.map, a.mC, t, grp * 32 + sl * 8, v); skip_store8(buf, row, grp * 4 + sl, v); }
    else if (grp == 1) {
      v[0] = px; v[1] = py; v[2] = pz; v[3] = v[4] = v[5] = v[6] = v[7] = 0.f;
      skip_store8(buf, row, 8, v);
      v[0] = v[1] = v[2] = 0.f;
      skip_store8(buf, row, 9, v);
    }
  } else {
    if (sl < 2) { gather8(a.map, a.mC, t, grp * 16 + sl * 8, v); skip_store8(buf, row, grp * 2 + sl, v); }
    else if (grp == 1) {
      v[0] = pz - a.cz; v[1] = v[2] = v[3] = v[4] = v[5] = v[6] = v[7] = 0.f;
      skip_store8(buf, row, 4, v);
      v[0] = 0.f;
      skip_store8(buf, row, 5, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool GRID>
__global__ void __launch_bounds__(NT, 1) field_tc2_kernel(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) unsigned char dsm[];
  unsigned char* ring = dsm;                                       // N_STAGES * STAGE_BYTES
  unsigned char* skip0 = dsm + N_STAGES * STAGE_BYTES;             // 2 x SKIP_BYTES: tile t uses buffer t & 1, the other one is being
                                                                   // filled with the next tile's input (gather prefetch)
  float* s_sb = reinterpret_cast<float*>(skip0 + 2 * SKIP_BYTES);  // {scale,bias} pairs of every layer
  float4* s_dotw = reinterpret_cast<float4*>(s_sb + SB_FLOATS_MAX);  // head weights {w0,w1,w2,0} per channel (+ a bias row per head)
  TcShared& S = *reinterpret_cast<TcShared*>(reinterpret_cast<unsigned char*>(s_dotw) + DOTW_F4_MAX * sizeof(float4));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool texture = (a.out_rgb != nullptr);
  if (tid == 0) {
    for (int i = 0; i < N_STAGES; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 1); }
    for (int i = 0; i < N_STAGES; ++i) mbar_init(&S.peer_full[i], 1);
    for (int i = 0; i < 8; ++i) mbar_init(&S.a_ready[i], 8);          // 4 quadrant warps of EACH CTA
    mbar_init(&S.d_ready[0], 1); mbar_init(&S.d_ready[1], 1); mbar_init(&S.epi_done, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {   // TMEM: all 512 columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised and both TMEM allocations done before any remote arrive / paired MMA
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const int n_ops = a.n_ops;
  const uint32_t rank = cluster_ctarank();
  {  // stage {scale,bias} interleaved per channel: s_sb[2*c] = scale, s_sb[2*c+1] = bias (tensor-core variants: scale includes 2^-shift)
    int base = 0;
    for (int l = 0; l < (int)a.hdr->n_layers; ++l) {
      const AvcLayerDesc& L = a.hdr->layers[l];
      for (int c = tid; c < L.np; c += NT) {
        s_sb[base + 2 * c] = a.f32[L.tc_sb_off + c];
        s_sb[base + 2 * c + 1] = a.f32[L.tc_sb_off + L.np + c];
      }
      base += 2 * L.np;
    }
    // head weights: the blob keeps W[n][K] for layers with n <= 4 (packer.py); scale folded in, bias row appended
    for (int oi = 0; oi < n_ops; ++oi) {
      const TcOp& o = a.ops[oi];
      if (o.dot_w < 0) continue;
      const AvcLayerDesc& L = a.hdr->layers[o.dot_layer];
      const int K = o.dot_k;
      for (int k = tid; k <= K; k += NT) {
        float w[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < L.n && j < 3; ++j)
          w[j] = k < K ? a.f32[L.wt_off + j * K + k] * a.f32[L.sb_off + j] : a.f32[L.sb_off + L.n + j];
        s_dotw[o.dot_w + k] = make_float4(w[0], w[1], w[2], 0.f);
      }
    }
  }
  __syncthreads();
  const int64_t n_tiles = (a.n + TILE - 1) / TILE;
  const int64_t n_pairs = (n_tiles + 1) / 2;           // a cluster (2 CTAs) evaluates a pair of tiles per iteration
  const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == 10) {
    // ============================================================ weight producer: the layer's weights are stored as a stream in
    // exactly the order and layout the ring consumes them (packer.py), so one bulk copy per stage is all it takes.
    {
      int stage = 0; uint32_t phase = 0;
      for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
        for (int oi = 0; oi < n_ops; ++oi) {
          const TcOp& o = a.ops[oi];
          const int n_halves = o.n == 256 ? 2 : 1;
          const uint32_t rows = (uint32_t)(o.n / n_halves);
          const uint32_t slab_bytes = rows * 32u;               // one hi (or lo) slab of a k-step in the stream
          const uint32_t my_bytes = slab_bytes / 2;             // this CTA's half of its rows (B is split along N across the pair)
          const unsigned char* src = a.w16 + o.w_off;
          for (int h = 0; h < n_halves; ++h) {
            for (int seg = 0; seg < 2; ++seg) {
              const int ks = seg == 0 ? o.ks_smem : o.ks_tmem;
              for (int j = 0; j < ks; j += STAGE_KSTEPS) {
                const int cnt = min(STAGE_KSTEPS, ks - j);
                mbar_wait(&S.empty[stage], phase ^ 1);
                if (elect_one()) {
                  mbar_expect_tx(&S.full[stage], (uint32_t)cnt * 2u * my_bytes);
                  for (int u = 0; u < cnt; ++u) {
                    unsigned char* dst = ring + stage * STAGE_BYTES + u * 2 * my_bytes;
                    const unsigned char* s0 = src + (size_t)u * 2 * slab_bytes + rank * my_bytes;
                    bulk_g2s(dst, s0, my_bytes, &S.full[stage]);                        // hi rows of this CTA
                    bulk_g2s(dst + my_bytes, s0 + slab_bytes, my_bytes, &S.full[stage]);   // lo rows of this CTA
                  }
                }
                __syncwarp();
                src += (size_t)cnt * 2 * slab_bytes;
                if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp >= 8 && rank != 0) {
    // ============================================================ peer CTA: no MMA issue. Warp 8 relays "my half of the stage has landed"
    // to the leader's peer_full barrier; warp 9 idles.
    if (warp == 8) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
        for (int oi = 0; oi < n_ops; ++oi) {
          const TcOp& o = a.ops[oi];
          const int n_halves = o.n == 256 ? 2 : 1;
          for (int h = 0; h < n_halves; ++h)
            for (int seg = 0; seg < 2; ++seg) {
              const int ks = seg == 0 ? o.ks_smem : o.ks_tmem;
              for (int j = 0; j < ks; j += STAGE_KSTEPS) {
                mbar_wait(&S.full[stage], phase);
                if (lane == 0) mbar_arrive_leader_relaxed(&S.peer_full[stage], rank);
                __syncwarp();
                if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
              }
            }
        }
      }
    }
  } else if (warp >= 8) {
    // ============================================================ MMA issuers: warps 8 and 9 walk the same (warp-uniform) program and
    // take the ring stages in turn (stage g belongs to warp 8 + (g & 1)). While one warp issues its 12 MMAs the other is already past
    // its barrier waits and descriptor set-up, so the fixed per-stage latency of a single instruction stream (~600 cycles measured) no
    // longer paces the tensor pipe. Issue ORDER is preserved by a named-barrier hand-off (bar 1: warp 8 may issue, bar 2: warp 9 may).
    {
      const int me = warp - 8;
      int stage = 0; uint32_t phase = 0;
      uint32_t ph_a = 0, ph_epi = 0;   // per-barrier phase bits (tracked by both warps, waited on by the stage owner)
      uint32_t g = 0;                  // global stage counter
      const uint32_t skip_base = smem_u32(skip0), ring_addr = smem_u32(ring);
      const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);   // LBO, SBO, version
      int tl = 0;
      if (me == 1) asm volatile("bar.arrive 1, 64;" ::: "memory");      // warp 8 owns stage 0
      for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
        const uint32_t skip_addr = skip_base + (uint32_t)(tl & 1) * SKIP_BYTES;
        for (int oi = 0; oi < n_ops; ++oi) {
          const TcOp& o = a.ops[oi];
          // A 256-wide layer is issued as two N=128 halves, each over the full K. The epilogue of half 0 (accumulator columns
          // 0..127 -> A chunks 0..3 of the next layer) then overlaps the MMAs of half 1, and the next layer's half 0 can start on
          // chunks 0..3 the moment this layer's half 1 has been issued: the tensor pipe does not wait for the epilogue.
          const int n_halves = o.n == 256 ? 2 : 1;
          const int rows = o.n / n_halves;
          const uint32_t idesc = make_idesc(rows);
          const uint32_t part_bytes = (uint32_t)rows * 16u;      // this CTA's half of the rows of a slab (the peer holds the other half)
          bool need_epi = o.wait_epi != 0;
          if (lane == 0 && (int)(g & 1) == me) trace_ev(a.trace, tl, oi, 0);
          for (int h = 0; h < n_halves; ++h) {
            const uint32_t d_addr = tmem + (uint32_t)(o.d_col + h * 128);
            uint32_t acc = o.accumulate ? 1u : 0u;
            for (int j = 0; j < o.ks_smem; j += STAGE_KSTEPS, ++g) {
              const int cnt = min(STAGE_KSTEPS, o.ks_smem - j);
              if ((int)(g & 1) == me) {
                if (need_epi) { mbar_wait_cluster(&S.epi_done, ph_epi); }
                mbar_wait(&S.full[stage], phase); mbar_wait(&S.peer_full[stage], phase);
                tc_fence_after();
                if (me == 0) asm volatile("bar.sync 1, 64;" ::: "memory"); else asm volatile("bar.sync 2, 64;" ::: "memory");
                if (elect_one()) {
                  for (int u = 0; u < cnt; ++u) {
                    const uint32_t b_addr = ring_addr + stage * STAGE_BYTES + u * 2 * part_bytes;
                    const uint64_t b_hi = desc_hi | (uint64_t)(b_addr >> 4), b_lo = desc_hi | (uint64_t)((b_addr + part_bytes) >> 4);
                    const uint32_t a_addr = skip_addr + (j + u) * 8192;
                    const uint64_t a_hi = desc_hi | (uint64_t)(a_addr >> 4), a_lo = desc_hi | (uint64_t)((a_addr + 4096) >> 4);
                    mma_ss(d_addr, a_hi, b_hi, idesc, acc); acc = 1u;
                    mma_ss(d_addr, a_lo, b_hi, idesc, 1u);
                    mma_ss(d_addr, a_hi, b_lo, idesc, 1u);
                  }
                  tc_commit(&S.empty[stage]);
                }
                __syncwarp();
                if (me == 0) asm volatile("bar.arrive 2, 64;" ::: "memory"); else asm volatile("bar.arrive 1, 64;" ::: "memory");
              }
              if (need_epi) { ph_epi ^= 1; need_epi = false; }
              acc = 1u;
              if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
            }
            for (int c2 = 0; c2 < (o.ks_tmem >> 2); ++c2, ++g) {   // one ring stage = 4 k-steps = two 32-column A chunks
              const bool wa = o.wait_a && h == 0;
              if ((int)(g & 1) == me) {
                // operands first (outside the serialised turn): 8 B descriptors and the A columns of the 4 k-steps
                const uint32_t b0 = ring_addr + stage * STAGE_BYTES;
                const uint32_t a0 = tmem + (uint32_t)o.a_col + (uint32_t)(c2 * 64);
                uint64_t bh[4], bl[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t b_addr = b0 + u * 2 * part_bytes;
                  bh[u] = desc_hi | (uint64_t)(b_addr >> 4); bl[u] = desc_hi | (uint64_t)((b_addr + part_bytes) >> 4);
                }
                if (need_epi) { mbar_wait_cluster(&S.epi_done, ph_epi); }
                if (wa) { mbar_wait(&S.a_ready[2 * c2], (ph_a >> (2 * c2)) & 1u); mbar_wait(&S.a_ready[2 * c2 + 1], (ph_a >> (2 * c2 + 1)) & 1u); }
                mbar_wait(&S.full[stage], phase); mbar_wait(&S.peer_full[stage], phase);
                tc_fence_after();
                if (me == 0) asm volatile("bar.sync 1, 64;" ::: "memory"); else asm volatile("bar.sync 2, 64;" ::: "memory");
                if (elect_one()) {
                  if (o.passes == 1) {                                         // colour head: hi*hi only (one branch per stage, outside the MMA chain)
#pragma unroll
                    for (int u = 0; u < 4; ++u) { mma_ts(d_addr, a0 + (uint32_t)((u >> 1) * 32 + (u & 1) * 8), bh[u], idesc, acc); acc = 1u; }
                  } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                      const uint32_t a_hi = a0 + (uint32_t)((u >> 1) * 32 + (u & 1) * 8), a_lo = a_hi + 16u;
                      // hi*hi and hi*lo back to back with A(hi) held in the collector buffer (one TMEM operand fetch for two MMAs), then lo*hi
                      mma_ts_c(d_addr, a_hi, bh[u], idesc, acc, 1); acc = 1u;
                      mma_ts_c(d_addr, a_hi, bl[u], idesc, 1u, 2);
                      mma_ts(d_addr, a_lo, bh[u], idesc | (1u << 13), 1u);      // negate A: TMEM holds -lo (split2_neg)
                    }
                  }
                  tc_commit(&S.empty[stage]);
                }
                __syncwarp();
                if (me == 0) asm volatile("bar.arrive 2, 64;" ::: "memory"); else asm volatile("bar.arrive 1, 64;" ::: "memory");
              }
              if (wa) ph_a ^= (3u << (2 * c2));
              if (need_epi) { ph_epi ^= 1; need_epi = false; }
              acc = 1u;
              if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
            }
            // the owner of this half's LAST stage signals the epilogue: its commit fires when its own MMAs are done, and the pipe
            // completes MMAs in issue order, so everything before them is done too
            if ((int)((g - 1) & 1) == me) {
              if (o.commit_d) { if (elect_one()) tc_commit(&S.d_ready[h]); __syncwarp(); }
              if (lane == 0) trace_ev(a.trace, tl, oi, 1 + h);
            }
          }
        }
        ++tl;
      }
    }
  } else {
    // ============================================================ compute / epilogue warps (0..7)
    const int quad = warp & 3, grp = warp >> 2;        // TMEM lane quadrant; column-group (0: even chunks, 1: odd chunks)
    const int row = quad * 32 + lane;                  // point within the tile == TMEM lane
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t ph_d0 = 0, ph_d1 = 0;
    int tl = 0;
    float nx_x = 0.f, nx_y = 0.f, nx_z = 0.f;          // the next tile's point (gather prefetch)
    for (int64_t pair = pair0; pair < n_pairs; pair += pair_step) {
      const int64_t tile = pair * 2 + rank;
      const int64_t g = tile * TILE + row;
      const bool valid = g < a.n;
      unsigned char* skip = skip0 + (tl & 1) * SKIP_BYTES;          // this tile's skip operand; the other buffer receives the next tile's
      unsigned char* skip_nx = skip0 + ((tl & 1) ^ 1) * SKIP_BYTES;
      const bool gathers = a.kind == AVC_KIND_RECON || a.mode != AVC_MODE_TEMPLATE_ONLY;     // input = bilinear feature gather
      const int n_slices = a.kind == AVC_KIND_RECON ? 3 : 5;
      float px, py, pz;
      if (tl == 0 || !gathers) {                                      // template-only programs have no prefetch: load every tile's points here
        fetch_point<GRID>(a, g, px, py, pz);
        if (gathers) {
          const Taps t = make_taps(px - a.cx, -(py - a.cy), a.mH, a.mW);
          for (int sl = 0; sl < n_slices; ++sl) input_slice(a, skip, row, grp, sl, t, px, py, pz);
        }
      } else {
        px = nx_x; py = nx_y; pz = nx_z;                            // loaded (and its input staged) during the previous tile
      }
      float qx = px, qy = py, qz = pz;
      bool need_pe = !gathers;                                      // template only: the input stage is the PE of the points themselves
      // gather prefetch of the NEXT tile: one step per op, taken right before the wait for that op's accumulator -- the compute warps
      // idle there (the "tail" between their last A chunk and the first accumulator half of the next layer, ~1.9 k cycles)
      const bool has_next = gathers && pair + pair_step < n_pairs;
      const int pf_first = (a.kind == AVC_KIND_AVATAR && a.mode == AVC_MODE_QUERY) ? 8 : 1;
      int pf = 0;
      Taps nt;
      auto prefetch_step = [&]() {
        if (pf == 0) {
          const int64_t g2 = ((pair + pair_step) * 2 + rank) * TILE + row;
          fetch_point<GRID>(a, g2, nx_x, nx_y, nx_z);
          nt = make_taps(nx_x - a.cx, -(nx_y - a.cy), a.mH, a.mW);
        } else {
          input_slice(a, skip_nx, row, grp, pf - 1, nt, nx_x, nx_y, nx_z);
        }
        ++pf;
      };
      for (int oi = 0; oi <= n_ops; ++oi) {
        if (need_pe) {
          // positional encoding of q into the skip buffer: k = [q(3), {sin(2^f q)(3), cos(2^f q)(3)}_f=0..9, 0]   net_util.py:28-37
          // group 0 writes frequencies 0..4 (+ the identity), group 1 frequencies 5..9 (+ the zero pad)
          unsigned short hi16, lo16;
          auto put = [&](int k, float val) {
            const __half h = __float2half_rn(val); const __half l = __float2half_rn(val - __half2float(h));
            hi16 = *reinterpret_cast<const unsigned short*>(&h); lo16 = *reinterpret_cast<const unsigned short*>(&l);
            unsigned char* base = skip + (k >> 4) * 8192 + (row >> 3) * 256 + ((k >> 3) & 1) * 128 + (row & 7) * 16 + (k & 7) * 2;
            *reinterpret_cast<unsigned short*>(base) = hi16; *reinterpret_cast<unsigned short*>(base + 4096) = lo16;
          };
          const float qq[3] = {qx, qy, qz};
          if (grp == 0) { put(0, qx); put(1, qy); put(2, qz); } else { put(63, 0.f); }
          for (int f = grp * 5; f < grp * 5 + 5; ++f) {
            const float fr = (float)(1 << f);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float s, c; fast_sincos(qq[d] * fr, s, c);
              put(3 + 6 * f + d, s); put(3 + 6 * f + 3 + d, c);
            }
          }
          need_pe = false;
          fence_proxy_async(); __syncwarp();
          if (lane == 0) mbar_arrive_leader(&S.epi_done, rank);
        } else if (oi == 0) {
          fence_proxy_async(); __syncwarp();
          if (lane == 0) mbar_arrive_leader(&S.epi_done, rank);      // input staged
        }
        if (oi == n_ops) break;
        const TcOp& o = a.ops[oi];
        if (!o.commit_d) continue;
        if (has_next && oi >= pf_first && pf <= n_slices) prefetch_step();
        mbar_wait(&S.d_ready[0], ph_d0); ph_d0 ^= 1; tc_fence_after();
        if (tid == 0) trace_ev(a.trace, tl, oi, 3);
        if (o.epi == EPI_HIDDEN) {
          const float* sb = s_sb + o.sb_off;
          const int n_chunks = o.n >> 5;
          for (int c = grp; c < n_chunks; c += 4) {                                                  // one accumulator half per iteration
            if (c == 4 + grp) {                                                                   // second N-half (only 256-wide ops get here)
              if (tid == 0) trace_ev(a.trace, tl, oi, 4);
              mbar_wait(&S.d_ready[1], ph_d1); ph_d1 ^= 1; tc_fence_after();
              if (tid == 0) trace_ev(a.trace, tl, oi, 5);
            }
            const uint32_t t0 = t_lane + (uint32_t)(o.d_col + c * 32), t1 = t0 + 64u;              // chunks c and c+2
            const float* sb0 = sb + 64 * c; const float* sb1 = sb0 + 128;
            switch (o.act) {                                 // one branch per half, none per value
              case AVC_ACT_RELU: hidden_pair<AVC_ACT_RELU>(t0, t1, sb0, sb1, &S.a_ready[c], &S.a_ready[c + 2], rank, lane); break;
              case AVC_ACT_LRELU: hidden_pair<AVC_ACT_LRELU>(t0, t1, sb0, sb1, &S.a_ready[c], &S.a_ready[c + 2], rank, lane); break;
              case AVC_ACT_SOFTPLUS: hidden_pair<AVC_ACT_SOFTPLUS>(t0, t1, sb0, sb1, &S.a_ready[c], &S.a_ready[c + 2], rank, lane); break;
              default: hidden_pair<AVC_ACT_NONE>(t0, t1, sb0, sb1, &S.a_ready[c], &S.a_ready[c + 2], rank, lane); break;
            }
          }
          if (tid == 0) trace_ev(a.trace, tl, oi, 6);
        } else {
          float r[4];
          if (o.dot_w >= 0) {
            // last hidden layer of a head + its <= 3-output linear layer on the CUDA cores (dot_pair)
            float acc[3] = {0.f, 0.f, 0.f};
            const float* sb = s_sb + o.sb_off;
            const int n_chunks = o.n >> 5;
            for (int c = grp; c < n_chunks; c += 4) {
              if (c == 4 + grp) {
                if (tid == 0) trace_ev(a.trace, tl, oi, 4);
                mbar_wait(&S.d_ready[1], ph_d1); ph_d1 ^= 1; tc_fence_after();
                if (tid == 0) trace_ev(a.trace, tl, oi, 5);
              }
              const uint32_t t0 = t_lane + (uint32_t)(o.d_col + c * 32), t1 = t0 + 64u;
              const float* sb0 = sb + 64 * c; const float* sb1 = sb0 + 128;
              const float4* w0 = s_dotw + o.dot_w + c * 32; const float4* w1 = w0 + 64;
              switch (o.act) {
                case AVC_ACT_RELU: dot_pair<AVC_ACT_RELU>(t0, t1, sb0, sb1, w0, w1, acc); break;
                case AVC_ACT_LRELU: dot_pair<AVC_ACT_LRELU>(t0, t1, sb0, sb1, w0, w1, acc); break;
                case AVC_ACT_SOFTPLUS: dot_pair<AVC_ACT_SOFTPLUS>(t0, t1, sb0, sb1, w0, w1, acc); break;
                default: dot_pair<AVC_ACT_NONE>(t0, t1, sb0, sb1, w0, w1, acc); break;
              }
            }
            // the two warps of a lane quadrant hold the sums over alternate 32-column chunks of the same 32 points
            S.dot_xch[grp][row] = make_float4(acc[0], acc[1], acc[2], 0.f);
            asm volatile("bar.sync 3, 256;" ::: "memory");
            const float4 oth = S.dot_xch[grp ^ 1][row];
            const float4 bias = s_dotw[o.dot_w + o.dot_k];
            // fixed summation order (group 0 + group 1) so that both warps of the quadrant get bit-identical results
            const float e0 = grp == 0 ? acc[0] : oth.x, e1 = grp == 0 ? acc[1] : oth.y, e2 = grp == 0 ? acc[2] : oth.z;
            const float f0 = grp == 0 ? oth.x : acc[0], f1 = grp == 0 ? oth.y : acc[1], f2 = grp == 0 ? oth.z : acc[2];
            r[0] = (e0 + f0) + bias.x; r[1] = (e1 + f1) + bias.y; r[2] = (e2 + f2) + bias.z; r[3] = 0.f;
            asm volatile("bar.sync 3, 256;" ::: "memory");          // dot_xch may be rewritten by the next head
            if (tid == 0) trace_ev(a.trace, tl, oi, 6);
          } else {
            float v[4];
            tmem_ld4(t_lane + (uint32_t)o.d_col, v);
            const float* sb = s_sb + o.sb_off;
#pragma unroll
            for (int i = 0; i < 4; ++i) r[i] = fmaf(v[i], sb[2 * i], sb[2 * i + 1]);
          }
          if (o.epi == EPI_WARP_OUT) {
            qx = px + r[0]; qy = py + r[1]; qz = pz + r[2];                       // cano_pts_chunk + offset_chunk  arch_avatar.py:372
            if (grp == 0 && valid && a.out_off) { a.out_off[g * 3] = r[0]; a.out_off[g * 3 + 1] = r[1]; a.out_off[g * 3 + 2] = r[2]; }
            need_pe = (a.mode != AVC_MODE_WARP_ONLY);
          } else if (o.epi == EPI_GEO_OUT) {
            if (grp == 0 && valid) {
              if (a.out0) a.out0[g] = a.if_type == AVC_IF_OCCUPANCY ? 1.f / (1.f + __expf(-r[0])) : r[0];   // arch_avatar.py:77-80
              if (a.out_alpha) a.out_alpha[g] = fmaxf(r[1], 0.f);                                        // :76
            }
          } else if (o.epi == EPI_CLR_OUT) {
            if (grp == 0 && valid && a.out_rgb) {
#pragma unroll
              for (int i = 0; i < 3; ++i) a.out_rgb[g * 3 + i] = 1.f / (1.f + __expf(-r[i]));              // :75
            }
          } else if (o.epi == EPI_RECON_OUT) {
            if (grp == 0 && valid) a.out0[g] = 1.f / (1.f + __expf(-r[0]));                               // mlp.py:49-50
          }
          if (o.signal_done && !need_pe) {
            tc_fence_before(); __syncwarp();
            if (lane == 0) mbar_arrive_leader(&S.epi_done, rank);
          }
        }
      }
      if (has_next) while (pf <= n_slices) prefetch_step();        // short programs have fewer ops than prefetch steps
      ++tl;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA of the pair may exit (or free TMEM) while the other can still signal it or issue paired MMAs
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

constexpr size_t TC_SMEM = (size_t)N_STAGES * STAGE_BYTES + 2 * SKIP_BYTES + SB_FLOATS_MAX * sizeof(float) + DOTW_F4_MAX * sizeof(float4) + sizeof(TcShared) + 64;

int launch_tc2(avc_ctx* ctx, const AvcWeights& w, int kind, const AvcMap* map, const float* pts, const AvcGridDesc* grid, int64_t n,
              const float center[3], float* out0, float* out_off, float* out_rgb, float* out_alpha, int if_type, int mode, cudaStream_t st) {
  if (n == 0) return AVC_OK;
  int sb = 0;
  for (uint32_t l = 0; l < w.hdr.n_layers; ++l) sb += 2 * w.hdr.layers[l].np;
  if (sb > SB_FLOATS_MAX) return avc_fail(ctx, AVC_EFORMAT, "tensor-core path: scale/bias table too large (%d floats)", sb);
  TcArgs a;
  a.pts = pts; a.n = n; a.cx = center[0]; a.cy = center[1]; a.cz = center[2];
  for (int c = 0; c < 3; ++c) {
    a.gb[c] = grid ? grid->bmin[c] : 0.f; a.gl[c] = grid ? grid->len[c] : 0.f; a.gr[c] = grid ? grid->res[c] : 1;
    a.gstep[c] = (grid && grid->res[c] > 1) ? 1.0f / (float)(grid->res[c] - 1) : 0.f;      // IEEE float division == __fdiv_rn of make_grid_kernel
  }
  a.gx_first = grid ? grid->x_first : 0;
  if (grid) a.pts = nullptr;
  if (grid && n >= ((int64_t)1 << 31)) return avc_fail(ctx, AVC_EINVAL, "dense-grid entry: at most 2^31 - 1 points per call (split the slab)");
  a.map = map ? map->d_hwc : nullptr; a.mC = map ? map->C : 0; a.mH = map ? map->H : 1; a.mW = map ? map->W : 1;
  a.out0 = out0; a.out_off = out_off; a.out_rgb = out_rgb; a.out_alpha = out_alpha; a.if_type = if_type; a.mode = mode; a.kind = kind;
  a.w16 = w.d_f16; a.f32 = w.d_f32; a.hdr = reinterpret_cast<const AvcBlobHeader*>(w.d_blob);
  a.trace = reinterpret_cast<long long*>(ctx->d_trace); a.dbg = ctx->dbg_flags;
  build_ops(a, &w.hdr, kind, mode, out_rgb != nullptr);
  AVC_CUDA(ctx, cudaFuncSetAttribute(grid ? field_tc2_kernel<true> : field_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
  const int64_t pairs = ((n + TILE - 1) / TILE + 1) / 2;
  const int max_clusters = ctx->sm_count / 2;
  const int n_ctas = 2 * (int)(pairs < (int64_t)max_clusters ? pairs : max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = TC_SMEM; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (grid) AVC_CUDA(ctx, cudaLaunchKernelEx(&cfg, field_tc2_kernel<true>, a));
  else AVC_CUDA(ctx, cudaLaunchKernelEx(&cfg, field_tc2_kernel<false>, a));
  AVC_LAUNCH_CHECK(ctx, "field_tc2_kernel");
  return AVC_OK;
}

}  // namespace

int avc_tc_available(const avc_ctx* ctx) { return ctx && ctx->cc_major == 10 ? 1 : 0; }

int avc_tc2_eval_avatar(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* out_occ, float* out_off,
                        float* out_rgb, float* out_alpha, int if_type, int mode, cudaStream_t st) {
  // the colour head runs when rgb is requested; alpha comes from the geo head
  return launch_tc2(ctx, ctx->avatar, AVC_KIND_AVATAR, mode == AVC_MODE_TEMPLATE_ONLY ? nullptr : &ctx->maps[AVC_MAP_POSE], pts, grid, n, center,
                    out_occ, out_off, out_rgb, out_alpha, if_type, mode, st);
}

int avc_tc2_eval_recon(avc_ctx* ctx, const float* pts, const AvcGridDesc* grid, int64_t n, const float center[3], float* out_ov, cudaStream_t st) {
  return launch_tc2(ctx, ctx->recon, AVC_KIND_RECON, &ctx->maps[AVC_MAP_IMAGE], pts, grid, n, center, out_ov, nullptr, nullptr, nullptr, AVC_IF_SDF,
                    AVC_MODE_QUERY, st);
}
