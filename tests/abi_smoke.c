/* A plain-C99 client of the C ABI (include/avatarcap_b200.h): no C++, no torch, no Python -- what a host written in any language
 * binds. Compiled with -fsyntax-only by the CPU suite (the header must be valid C) and built + run on the GPU box by
 * tests/test_gpu_raster.py::test_c_client_of_the_abi.
 *   gcc -std=c99 tests/abi_smoke.c -Iinclude -I$CUDA/include -Lavatarcap_b200 -lavatarcap_b200 -L$CUDA/lib64 -lcudart -lm */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "avatarcap_b200.h"

#define CHECK(cond, what) do { if (!(cond)) { fprintf(stderr, "abi_smoke: FAILED %s (line %d): %s\n", what, __LINE__, avc_last_error(ctx)); return 1; } } while (0)

int main(void) {
  avc_ctx* ctx = NULL;
  if (avc_abi_version() != AVC_ABI_VERSION) { fprintf(stderr, "abi_smoke: ABI version %d\n", avc_abi_version()); return 1; }
  if (avc_ctx_create(0, &ctx) != AVC_OK) { fprintf(stderr, "abi_smoke: no context: %s\n", avc_last_error(NULL)); return 2; }

  /* grid of generate_volume_points: last point == bmax */
  const float bounds[6] = {-1.f, -2.f, -3.f, 1.f, 2.f, 3.f};
  const int res[3] = {3, 4, 5};
  float* d_pts = NULL; float h_pts[3 * 4 * 5 * 3];
  CHECK(cudaMalloc((void**)&d_pts, sizeof(h_pts)) == cudaSuccess, "cudaMalloc");
  CHECK(avc_make_grid(ctx, bounds, res, 0, 3, d_pts, NULL) == AVC_OK, "avc_make_grid");
  CHECK(cudaMemcpy(h_pts, d_pts, sizeof(h_pts), cudaMemcpyDeviceToHost) == cudaSuccess, "cudaMemcpy");
  CHECK(h_pts[0] == -1.f && h_pts[1] == -2.f && h_pts[2] == -3.f, "grid first point");
  CHECK(h_pts[177] == 1.f && h_pts[178] == 2.f && h_pts[179] == 3.f, "grid last point");
  CHECK(h_pts[3 + 2] == -1.5f, "z fastest");

  /* marching cubes on an 8^3 ball: two-phase count -> emit, closed surface => F = 2V - 4 */
  enum { R = 8 };
  float h_vol[R * R * R]; float* d_vol = NULL;
  for (int i = 0; i < R; ++i) for (int j = 0; j < R; ++j) for (int k = 0; k < R; ++k)
    h_vol[(i * R + j) * R + k] = 2.6f - sqrtf((i - 3.5f) * (i - 3.5f) + (j - 3.5f) * (j - 3.5f) + (k - 3.5f) * (k - 3.5f));
  const int vres[3] = {R, R, R}; const float vb[6] = {0, 0, 0, 1, 1, 1};
  int64_t nv = 0, nf = 0;
  CHECK(cudaMalloc((void**)&d_vol, sizeof(h_vol)) == cudaSuccess, "cudaMalloc");
  CHECK(cudaMemcpy(d_vol, h_vol, sizeof(h_vol), cudaMemcpyHostToDevice) == cudaSuccess, "cudaMemcpy");
  CHECK(avc_mc_count(ctx, d_vol, vres, 0.f, 0, 0, &nv, &nf, NULL) == AVC_OK, "avc_mc_count");
  CHECK(nv > 0 && nf == 2 * nv - 4, "closed genus-0 surface");
  float *d_v = NULL, *d_n = NULL; int32_t* d_f = NULL;
  CHECK(cudaMalloc((void**)&d_v, (size_t)nv * 12) == cudaSuccess && cudaMalloc((void**)&d_n, (size_t)nv * 12) == cudaSuccess &&
        cudaMalloc((void**)&d_f, (size_t)nf * 12) == cudaSuccess, "cudaMalloc");
  CHECK(avc_mc_emit_counted(ctx, d_vol, vres, vb, 0.f, 0, 0, 0, R, d_v, d_n, d_f, nv, nf, NULL) == AVC_OK, "avc_mc_emit_counted");
  CHECK(avc_mc_emit(ctx, d_vol, vres, vb, 0.f, 0, 0, 0, R, d_v, d_n, d_f, nv - 1, nf, NULL) == AVC_ECAPACITY, "capacity error, no truncation");
  CHECK(strlen(avc_last_error(ctx)) > 0, "error text");

  /* rasterise that mesh with its normals: orthographic front view, alpha 1 inside the silhouette, 0 outside */
  const float mvp[16] = {2, 0, 0, -1,  0, 2, 0, -1,  0, 0, -1, 0,  0, 0, 0, 1};     /* [0,1]^3 -> NDC, looking down -z */
  enum { W = 32, H = 32 };
  float* d_img = NULL; static float h_img[W * H * 4];
  CHECK(cudaMalloc((void**)&d_img, sizeof(h_img)) == cudaSuccess, "cudaMalloc");
  CHECK(avc_rasterize(ctx, d_v, nv, d_f, nf, d_n, mvp, W, H, NULL, AVC_RASTER_CULL_BACK, 4, d_img, NULL) == AVC_OK, "avc_rasterize");
  CHECK(cudaMemcpy(h_img, d_img, sizeof(h_img), cudaMemcpyDeviceToHost) == cudaSuccess, "cudaMemcpy");
  CHECK(h_img[((H / 2) * W + W / 2) * 4 + 3] == 1.f && h_img[3] == 0.f, "silhouette");
  CHECK(fabsf(h_img[((H / 2) * W + W / 2) * 4 + 2]) > 0.9f, "centre normal points along z");
  CHECK(avc_rasterize(ctx, d_v, nv, d_f, nf, d_n, mvp, W, H, NULL, 0, 2, d_img, NULL) == AVC_EINVAL, "bad channel count");

  /* field evaluation without weights is a state error, not a crash */
  const float center[3] = {0, 0, 0};
  CHECK(avc_eval_occupancy(ctx, d_pts, 60, center, d_img, NULL, NULL, NULL, AVC_IF_SDF, AVC_IMPL_AUTO, NULL) == AVC_ESTATE, "no weights loaded");
  CHECK(avc_launch_count(ctx) > 0, "launch counter");

  cudaFree(d_pts); cudaFree(d_vol); cudaFree(d_v); cudaFree(d_n); cudaFree(d_f); cudaFree(d_img);
  avc_ctx_destroy(ctx);
  printf("abi_smoke ok: %lld vertices, %lld faces\n", (long long)nv, (long long)nf);
  return 0;
}
