// Per-triangle / per-pixel arithmetic of the rasteriser (raster.cu), written once for device and host: the kernels include
// it, and tests/test_raster_host.py compiles the same header with g++ to check the coverage / interpolation rules on the CPU
// against the numpy oracle before any GPU time is spent.
//
// Replaces the reference's off-screen OpenGL passes (utils/renderer.py:326-451, shaders :9-51): depth-tested, back-face
// culled triangles into an RGBA32F target. Rules restated (no GL here -- "parity unpinned", see oracle/raster_oracle.py):
//   * fragments for pixel centres inside the triangle, top-left rule on shared edges, window y up;
//   * window depth linear in window space, quantised to 24 bits (GL_DEPTH24_STENCIL8, renderer.py:392), GL_LESS: of two equal
//     depths the triangle drawn first wins -> 64-bit key (z24 << 32 | triangle index), atomicMin;
//   * perspective-correct varyings (1/w weights).
// Arithmetic contract: float32 vertex transform without FMA, float64 edge functions -- every operation below is written with
// explicit round-to-nearest intrinsics on the device so that device, host build and numpy agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define RC_FMUL(a, b) __fmul_rn((a), (b))
#define RC_FADD(a, b) __fadd_rn((a), (b))
#define RC_FDIV(a, b) __fdiv_rn((a), (b))
#define RC_DMUL(a, b) __dmul_rn((a), (b))
#define RC_DADD(a, b) __dadd_rn((a), (b))
#define RC_DSUB(a, b) __dsub_rn((a), (b))
#define RC_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define RC_FMUL(a, b) ((a) * (b))
#define RC_FADD(a, b) ((a) + (b))
#define RC_FDIV(a, b) ((a) / (b))
#define RC_DMUL(a, b) ((a) * (b))
#define RC_DADD(a, b) ((a) + (b))
#define RC_DSUB(a, b) ((a) - (b))
#define RC_DDIV(a, b) ((a) / (b))
#endif

struct RcVtx { float x, y, z, iw; };            // window x, y (y up), depth in [0,1], 1/w (0: behind the eye)
struct RcMat { float m[16]; };                  // row-major 4x4

#define RC_EMPTY 0xFFFFFFFFFFFFFFFFull

RC_HD RcVtx rc_transform(const RcMat& M, float x, float y, float z, int W, int H) {
  float c[4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
    c[r] = RC_FADD(RC_FADD(RC_FADD(RC_FMUL(M.m[4 * r], x), RC_FMUL(M.m[4 * r + 1], y)), RC_FMUL(M.m[4 * r + 2], z)), M.m[4 * r + 3]);
  RcVtx v;
  v.x = RC_FMUL(RC_FADD(RC_FDIV(c[0], c[3]), 1.f), 0.5f * (float)W);
  v.y = RC_FMUL(RC_FADD(RC_FDIV(c[1], c[3]), 1.f), 0.5f * (float)H);
  v.z = RC_FMUL(RC_FADD(RC_FDIV(c[2], c[3]), 1.f), 0.5f);
  v.iw = c[3] > 0.f ? RC_FDIV(1.f, c[3]) : 0.f;
  return v;
}

struct RcTri {
  double x0, y0, x1, y1, x2, y2, area2;
  double z0, z1, z2;
  int i0, i1, i2;          // vertex ids in counter-clockwise window order
  bool ok;
};

RC_HD double rc_edge(double ax, double ay, double bx, double by, double cx, double cy) {
  return RC_DSUB(RC_DMUL(RC_DSUB(bx, ax), RC_DSUB(cy, ay)), RC_DMUL(RC_DSUB(by, ay), RC_DSUB(cx, ax)));
}
RC_HD bool rc_top_left(double ax, double ay, double bx, double by) {
  const double dx = RC_DSUB(bx, ax), dy = RC_DSUB(by, ay);
  return dy < 0.0 || (dy == 0.0 && dx < 0.0);
}
RC_HD bool rc_finite(float v) { return v == v && fabsf(v) < 3.0e38f; }

RC_HD RcTri rc_setup(const RcVtx& a, const RcVtx& b, const RcVtx& c, int i0, int i1, int i2, bool cull) {
  RcTri t; t.ok = false;
  if (!(a.iw > 0.f) || !(b.iw > 0.f) || !(c.iw > 0.f)) return t;
  if (!rc_finite(a.x) || !rc_finite(a.y) || !rc_finite(b.x) || !rc_finite(b.y) || !rc_finite(c.x) || !rc_finite(c.y)) return t;
  t.x0 = a.x; t.y0 = a.y; t.z0 = a.z; t.i0 = i0;
  double area2 = rc_edge(a.x, a.y, b.x, b.y, c.x, c.y);
  if (area2 == 0.0 || (cull && area2 < 0.0)) return t;
  if (area2 < 0.0) {                       // clockwise with culling off: swap to counter-clockwise
    t.x1 = c.x; t.y1 = c.y; t.z1 = c.z; t.i1 = i2; t.x2 = b.x; t.y2 = b.y; t.z2 = b.z; t.i2 = i1; area2 = -area2;
  } else {
    t.x1 = b.x; t.y1 = b.y; t.z1 = b.z; t.i1 = i1; t.x2 = c.x; t.y2 = c.y; t.z2 = c.z; t.i2 = i2;
  }
  t.area2 = area2; t.ok = true;
  return t;
}

// pixels whose CENTRE can lie inside: ceil(min - 0.5) .. floor(max - 0.5), clamped to the target; false if empty
RC_HD bool rc_bbox(const RcTri& t, int W, int H, int& px0, int& px1, int& py0, int& py1) {
  const double xmin = fmin(t.x0, fmin(t.x1, t.x2)), xmax = fmax(t.x0, fmax(t.x1, t.x2));
  const double ymin = fmin(t.y0, fmin(t.y1, t.y2)), ymax = fmax(t.y0, fmax(t.y1, t.y2));
  // clamp in floating point first: the casts below must not overflow for far-away vertices
  const double lx = fmax(ceil(xmin - 0.5), 0.0), hx = fmin(floor(xmax - 0.5), (double)(W - 1));
  const double ly = fmax(ceil(ymin - 0.5), 0.0), hy = fmin(floor(ymax - 0.5), (double)(H - 1));
  if (hx < lx || hy < ly) return false;
  px0 = (int)lx; px1 = (int)hx; py0 = (int)ly; py1 = (int)hy;
  return true;
}

// coverage + depth of pixel (px,py): true and the 24-bit depth if the centre is covered and inside the depth range
RC_HD bool rc_cover(const RcTri& t, int px, int py, uint32_t& z24) {
  const double cx = (double)px + 0.5, cy = (double)py + 0.5;
  const double w0 = rc_edge(t.x1, t.y1, t.x2, t.y2, cx, cy);
  const double w1 = rc_edge(t.x2, t.y2, t.x0, t.y0, cx, cy);
  const double w2 = rc_edge(t.x0, t.y0, t.x1, t.y1, cx, cy);
  if (!(w0 > 0.0 || (w0 == 0.0 && rc_top_left(t.x1, t.y1, t.x2, t.y2)))) return false;
  if (!(w1 > 0.0 || (w1 == 0.0 && rc_top_left(t.x2, t.y2, t.x0, t.y0)))) return false;
  if (!(w2 > 0.0 || (w2 == 0.0 && rc_top_left(t.x0, t.y0, t.x1, t.y1)))) return false;
  const double z = RC_DDIV(RC_DADD(RC_DADD(RC_DMUL(w0, t.z0), RC_DMUL(w1, t.z1)), RC_DMUL(w2, t.z2)), t.area2);
  if (!(z >= 0.0 && z <= 1.0)) return false;
  z24 = (uint32_t)floor(RC_DADD(RC_DMUL(z, 16777215.0), 0.5));
  return true;
}

// perspective-correct attribute at the centre of (px,py); iw = 1/w of the three vertices, a0..a2 their 3-vectors
RC_HD void rc_shade(const RcTri& t, float iw0, float iw1, float iw2, const float* a0, const float* a1, const float* a2, int px, int py,
                    float out[3]) {
  const double cx = (double)px + 0.5, cy = (double)py + 0.5;
  const double w0 = rc_edge(t.x1, t.y1, t.x2, t.y2, cx, cy);
  const double w1 = rc_edge(t.x2, t.y2, t.x0, t.y0, cx, cy);
  const double w2 = rc_edge(t.x0, t.y0, t.x1, t.y1, cx, cy);
  const double q0 = RC_DMUL(RC_DDIV(w0, t.area2), (double)iw0), q1 = RC_DMUL(RC_DDIV(w1, t.area2), (double)iw1),
               q2 = RC_DMUL(RC_DDIV(w2, t.area2), (double)iw2);
  const double den = RC_DADD(RC_DADD(q0, q1), q2);
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[c] = (float)RC_DDIV(RC_DADD(RC_DADD(RC_DMUL(q0, (double)a0[c]), RC_DMUL(q1, (double)a1[c])), RC_DMUL(q2, (double)a2[c])), den);
}

// F.grid_sample(mode='nearest', padding_mode='border', align_corners=True): pixel index for a normalised coordinate
RC_HD int rc_nearest_border(float g, int size) {
  float i = RC_FMUL(RC_FDIV(RC_FADD(g, 1.f), 2.f), (float)(size - 1));
  i = fminf((float)(size - 1), fmaxf(i, 0.f));
  return (int)nearbyintf(i);               // round half to even, like ATen
}
