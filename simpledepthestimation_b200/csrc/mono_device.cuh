// Device functions shared by the forward and backward MonoDepth2 loss kernels.
#pragma once
#include "mono_params.cuh"

namespace sde {

// Shared-memory plane of a tile + 1-pixel halo: 18 rows x 66 columns.  Column xx is stored at
// index xx + 1 of a 68-float row so that the pixel pair owned by a lane (columns 2*lane+1, 2*lane+2)
// is 8-byte aligned and lands in one register pair with a single LDS.64.
constexpr int kHW = kTileW + 2;       // 66 columns
constexpr int kHH = kTileH + 2;       // 18 rows
constexpr int kPitch = 68;
constexpr int kPlane = kHH * kPitch;  // 1224 floats
constexpr int kPositions = kHH * kHW; // 1188 staged positions

__device__ __forceinline__ int plane_index(int yy, int xx) { return yy * kPitch + xx + 1; }

// Row access of phase 2: centre pair (columns c0+1, c0+2) and outer pair (c0, c0+3).
struct Row4 {
  f2 c, o;
};
__device__ __forceinline__ Row4 ld_row(const float* p) {  // p -> stored index of column c0 (odd index)
  Row4 r;
  r.c = ld2(p + 1);
  r.o = mk2(p[0], p[3]);
  return r;
}

// p0/den and p1/den through one refined reciprocal and a residual correction (<= 1 ulp from the
// IEEE quotient the reference's torch division produces); exotic denominators take the IEEE path.
__device__ __forceinline__ void divide2(float p0, float p1, float den, float& X, float& Y) {
  const float ad = fabsf(den);
  if (ad > 1e-30f && ad < 1e30f) {
    float r = rcp_approx(den);
    r = fmaf(fmaf(-den, r, 1.0f), r, r);
    X = p0 * r;
    Y = p1 * r;
    X = fmaf(fmaf(-X, den, p0), r, X);
    Y = fmaf(fmaf(-Y, den, p1), r, Y);
  } else {
    X = p0 / den;
    Y = p1 / den;
  }
}

// Back-project pixel (gx, gy) with depth d through K^-1, move by (R, t), project with K
// (camera.py:125-163,172-178).  P = K^-1 [x d, y d, d], p = (K R) P + K t, X = p0 / (p2 + 1e-6).
__device__ __forceinline__ void project_full(const Cam& c, const Proj& pj, float gx, float gy, float d, float P[3],
                                             float& den, float& X, float& Y) {
  const float xd = gx * d, yd = gy * d;
  P[0] = c.ki[0] * xd + c.ki[1] * yd + c.ki[2] * d;
  P[1] = c.ki[3] * xd + c.ki[4] * yd + c.ki[5] * d;
  P[2] = c.ki[6] * xd + c.ki[7] * yd + c.ki[8] * d;
  const float p0 = pj.m[0] * P[0] + pj.m[1] * P[1] + pj.m[2] * P[2] + pj.tau[0];
  const float p1 = pj.m[3] * P[0] + pj.m[4] * P[1] + pj.m[5] * P[2] + pj.tau[1];
  const float p2 = pj.m[6] * P[0] + pj.m[7] * P[1] + pj.m[8] * P[2] + pj.tau[2];
  den = p2 + 1e-6f;
  divide2(p0, p1, den, X, Y);
}

// nan_to_num + clamp (camera.py:184-188) and the bilinear cell of grid_sample(align_corners=True).
// fmaxf(NaN, 0) == 0 and +-inf clamp to the borders, which is exactly nan_to_num followed by clamp.
// The cell origin is capped at (w-2, h-2): at ix == w-1 the reference's right tap is out of range
// with weight 0, here the left tap gets weight 0 instead -- the same sample, and all four taps
// stay in range so the loads need no predicate.
struct Cell {
  int off;               // y0 * w + x0
  float ax, ay;          // weights of the right / lower taps
};
__device__ __forceinline__ Cell bilinear_cell(float X, float Y, int w, int h) {
  const float ix = fminf(fmaxf(X, 0.0f), (float)(w - 1));
  const float iy = fminf(fmaxf(Y, 0.0f), (float)(h - 1));
  const int x0 = min((int)ix, w - 2), y0 = min((int)iy, h - 2);   // ix, iy >= 0: truncation == floor
  Cell c;
  c.off = y0 * w + x0;
  c.ax = ix - (float)x0;
  c.ay = iy - (float)y0;
  return c;
}

// one channel: sum of the four taps in ATen's order (nw, ne, sw, se)
__device__ __forceinline__ float tap4(const float* __restrict__ pl, int off, int w, float w00, float w01, float w10,
                                      float w11) {
  const float* r0 = pl + off;
  const float* r1 = r0 + w;
  return __ldg(r0) * w00 + __ldg(r0 + 1) * w01 + __ldg(r1) * w10 + __ldg(r1 + 1) * w11;
}

}  // namespace sde
