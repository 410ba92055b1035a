// Device functions shared by the forward and backward MonoDepth2 loss kernels.
#pragma once
#include "mono_params.cuh"

namespace sde {

constexpr int kHW = kTileW + 2;     // 66 (row pitch == halo'd width: phase-1 stores are linear)
constexpr int kHH = kTileH + 2;     // 18
constexpr int kPlane = kHW * kHH;   // 1188 floats

__device__ __forceinline__ void project_px(const Cam& c, const Proj& pj, float gx, float gy, float d, float& X,
                                           float& Y) {
  const float xd = gx * d, yd = gy * d;
  const float Px = c.ki[0] * xd + c.ki[1] * yd + c.ki[2] * d;
  const float Py = c.ki[3] * xd + c.ki[4] * yd + c.ki[5] * d;
  const float Pz = c.ki[6] * xd + c.ki[7] * yd + c.ki[8] * d;
  const float p0 = pj.m[0] * Px + pj.m[1] * Py + pj.m[2] * Pz + pj.tau[0];
  const float p1 = pj.m[3] * Px + pj.m[4] * Py + pj.m[5] * Pz + pj.tau[1];
  const float p2 = pj.m[6] * Px + pj.m[7] * Py + pj.m[8] * Pz + pj.tau[2];
  const float den = p2 + 1e-6f;
  X = p0 / den;
  Y = p1 / den;
}

// nan_to_num + clamp (camera.py:184-188) then the bilinear taps of grid_sample(align_corners=True).
// fmaxf(NaN, 0) == 0 and +-inf clamp to the borders, which is exactly nan_to_num followed by clamp.
__device__ __forceinline__ void bilinear3(const float* __restrict__ img, int hw, int w, int h, float X, float Y,
                                          float out[3]) {
  const float ix = fminf(fmaxf(X, 0.0f), (float)(w - 1));
  const float iy = fminf(fmaxf(Y, 0.0f), (float)(h - 1));
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
  const float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
  const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);  // clamped taps carry weight 0
  const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
  const int o00 = y0 * w + x0, o01 = y0 * w + x1, o10 = y1 * w + x0, o11 = y1 * w + x1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = img + c * hw;
    out[c] = __ldg(pl + o00) * w00 + __ldg(pl + o01) * w01 + __ldg(pl + o10) * w10 + __ldg(pl + o11) * w11;
  }
}

}  // namespace sde
