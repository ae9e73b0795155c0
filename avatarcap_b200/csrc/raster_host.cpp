// HOST build of the rasteriser's arithmetic (raster_core.h) for the CPU test suite: the same per-triangle / per-pixel functions
// the CUDA kernels in raster.cu call, driven by plain loops. NOT part of libavatarcap_b200.so and not a fallback -- it exists
// so that tests/test_raster_host.py can check the coverage, depth and interpolation rules against the numpy oracle in a
// container without a GPU.   g++ -O2 -ffp-contract=off -shared -fPIC raster_host.cpp -o raster_host.so
#include <string.h>

#include <vector>

#include "raster_core.h"

extern "C" int rc_host_rasterize(const float* verts, long n_verts, const int* faces, long n_faces, const float* attrs, const float* mvp, int W, int H,
                                 const float* bg, int cull, int flip_x, int channels, float* out) {
  RcMat M; memcpy(M.m, mvp, sizeof(M.m));
  std::vector<RcVtx> tv(n_verts);
  for (long i = 0; i < n_verts; ++i) tv[i] = rc_transform(M, verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], W, H);
  std::vector<unsigned long long> zbuf((size_t)W * H, RC_EMPTY);
  auto tri_of = [&](long t) {
    int i0, i1, i2;
    if (faces) { i0 = faces[3 * t]; i1 = faces[3 * t + 1]; i2 = faces[3 * t + 2]; } else { i0 = (int)(3 * t); i1 = i0 + 1; i2 = i0 + 2; }
    RcTri tri; tri.ok = false;
    if (i0 < 0 || i1 < 0 || i2 < 0 || i0 >= n_verts || i1 >= n_verts || i2 >= n_verts) return tri;
    return rc_setup(tv[i0], tv[i1], tv[i2], i0, i1, i2, cull != 0);
  };
  for (long t = 0; t < n_faces; ++t) {
    const RcTri tri = tri_of(t);
    int px0, px1, py0, py1;
    if (!tri.ok || !rc_bbox(tri, W, H, px0, px1, py0, py1)) continue;
    for (int py = py0; py <= py1; ++py)
      for (int px = px0; px <= px1; ++px) {
        uint32_t z24;
        if (!rc_cover(tri, px, py, z24)) continue;
        const unsigned long long key = ((unsigned long long)z24 << 32) | (unsigned long long)(uint32_t)t;
        unsigned long long& z = zbuf[(size_t)py * W + px];
        if (key < z) z = key;
      }
  }
  const float* A = attrs ? attrs : verts;
  for (int py = 0; py < H; ++py)
    for (int px = 0; px < W; ++px) {
      float r[4] = {bg ? bg[0] : 0.f, bg ? bg[1] : 0.f, bg ? bg[2] : 0.f, 0.f};
      const unsigned long long key = zbuf[(size_t)py * W + px];
      if (key != RC_EMPTY) {
        const RcTri tri = tri_of((long)(key & 0xFFFFFFFFull));
        rc_shade(tri, tv[tri.i0].iw, tv[tri.i1].iw, tv[tri.i2].iw, A + 3 * (size_t)tri.i0, A + 3 * (size_t)tri.i1, A + 3 * (size_t)tri.i2, px, py, r);
        r[3] = 1.f;
      }
      const int ox = flip_x ? W - 1 - px : px, oy = H - 1 - py;
      for (int c = 0; c < channels; ++c) out[((size_t)oy * W + ox) * channels + c] = r[c];
    }
  return 0;
}

extern "C" int rc_host_nearest_border(float g, int size) { return rc_nearest_border(g, size); }
