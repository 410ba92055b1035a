// Device functions shared by the MotionLearning loss kernels (statistics pre-pass, forward, backward).
//
// One direction A->B of MotionLearningModel.rgbd_consistency_loss (MotionLearning.py:248-291):
//   view_synthesis(cat[frame_B, depth_B], depth_A, K, R, t)  with  t = pose[:3,3] + field(x,y)
//   (camera.py:166-202; the translation is a per-pixel field, MotionLearning.py:143-153),
//   occlusion = (depth_in_B < sampled_depth_B) * valid,
//   depth_error = (depth_in_B - sampled_depth_B)^2,
//   weight = m2_b / (depth_error + m2_b) * valid   with the per-sample second moment m2_b,
//   WeightedSSIM(sampled_rgb, frame_A, weight)  (ssim_loss.py:84-111).
#pragma once
#include "mono_device.cuh"

namespace sde {

struct MotionParams {
  int B, n_dirs, h, w;
  int tiles_x, tiles_y, tiles_per_dir;       // forward: 64x16 tiles
  int btiles_x, btiles_y, btiles_per_dir;    // backward: 62x14 gradient tiles
  int stat_blocks;                           // statistics pre-pass: blocks per (direction, sample)
  float sx, sy;                              // scale_intrinsics factors
  const float* frame_a[SDE_MAX_DIRS];
  const float* frame_b[SDE_MAX_DIRS];
  const float* depth_a[SDE_MAX_DIRS];
  const float* depth_b[SDE_MAX_DIRS];
  const float* pose[SDE_MAX_DIRS];
  const float* field[SDE_MAX_DIRS];
  const float* K;
  float ssim_w, c1, c2;
  int mode;                                  // 0: both SSIM factors, 1: C1 = inf, 2: C2 = inf (ssim_loss.py:97-105)
  float* losses;                             // [n_dirs][4]
  float* stats;                              // [n_dirs][B][4] = m2, sum(occ), mean 1/depth, smoothness
  float* occ[SDE_MAX_DIRS];
  float* weight[SDE_MAX_DIRS];
  float* coords[SDE_MAX_DIRS];
  float* warped[SDE_MAX_DIRS];               // [B,12,h,w]: warped rgb, depth error, valid + 2 occlusion, d rgb/dX, d rgb/dY,
                                             // local smoothness gradient (or null)
  int tma;                                   // loss kernels stage their planes through TMA from `warped`
  // workspace
  unsigned* counters;                        // [0] statistics, [1] forward: images finished, [2] unused
  unsigned* img_counter_f;                   // [n_dirs*B] forward tiles finished per (direction, sample)
  unsigned* img_counter_b;                   // [n_dirs*B] backward tiles finished per (direction, sample)
  double* fin;                               // [n_dirs*B][4] per-image results of the forward pass (l1 sum, ssim sum, smoothness)
  float* stat_partials;                      // [n_dirs*B*stat_blocks][2]
  float* partials;                           // [forward grid][8]
  float* pose_partials;                      // [backward grid][12]
  // backward
  const float* grad_losses;                  // [n_dirs][4]
  float* grad_depth[SDE_MAX_DIRS];
  float* grad_pose[SDE_MAX_DIRS];
  float* grad_field[SDE_MAX_DIRS];
};

constexpr int kMotionSaved = 16;   // planes per sample of a `warped` buffer: 12 kept planes + 4 planes of scratch
constexpr int kMotionPacked = 12;  // first scratch plane: frame B and depth B of the sample interleaved per pixel
                                   // ([h*w] float4 = r, g, b, depth), written by motion_pack_kernel for the gather
constexpr int kStatThreads = 256;
constexpr int kStatPixPerThread = 16;   // 4096 pixels per block: the camera prologue (one thread's global loads + barrier) and the
                                        // ticket epilogue were 20 % of a 1024-pixel block's lifetime
constexpr int kStatPix = kStatThreads * kStatPixPerThread;

// Per-(direction, sample) camera terms: scaled K, K^-1, M = K R, R and the pose translation.
struct MCam {
  Cam cam;
  float k[9];
  float m[9];
  float r[9];
  float t[3];
};

__device__ __forceinline__ void load_mcam(MCam& c, const float* __restrict__ K, const float* __restrict__ pose, int b,
                                          float sx, float sy) {
  load_cam(c.cam, c.k, K, b, sx, sy);
  const float* T = pose + b * 16;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) c.r[i * 3 + j] = T[i * 4 + j];
    c.t[i] = T[i * 4 + 3];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[i * 3 + j] = c.k[i * 3] * c.r[j] + c.k[i * 3 + 1] * c.r[3 + j] + c.k[i * 3 + 2] * c.r[6 + j];
}

// p = (K R) K^-1 [x d, y d, d] + K (t_pose + f)   (camera.py:125-163,172-178 with the translation field)
__device__ __forceinline__ void mproject(const MCam& c, float gx, float gy, float d, float f0, float f1, float f2v,
                                         float P[3], float& den, float& X, float& Y, float& Z) {
  const float xd = gx * d, yd = gy * d;
  P[0] = c.cam.ki[0] * xd + c.cam.ki[1] * yd + c.cam.ki[2] * d;
  P[1] = c.cam.ki[3] * xd + c.cam.ki[4] * yd + c.cam.ki[5] * d;
  P[2] = c.cam.ki[6] * xd + c.cam.ki[7] * yd + c.cam.ki[8] * d;
  const float t0 = c.t[0] + f0, t1 = c.t[1] + f1, t2 = c.t[2] + f2v;
  const float tau0 = c.k[0] * t0 + c.k[1] * t1 + c.k[2] * t2;
  const float tau1 = c.k[3] * t0 + c.k[4] * t1 + c.k[5] * t2;
  const float tau2 = c.k[6] * t0 + c.k[7] * t1 + c.k[8] * t2;
  const float p0 = c.m[0] * P[0] + c.m[1] * P[1] + c.m[2] * P[2] + tau0;
  const float p1 = c.m[3] * P[0] + c.m[4] * P[1] + c.m[5] * P[2] + tau1;
  Z = c.m[6] * P[0] + c.m[7] * P[1] + c.m[8] * P[2] + tau2;
  den = Z + 1e-6f;
  divide2(p0, p1, den, X, Y);
}

// valid_proj_mask of points_to_img (camera.py:153-156); NaN / inf compare false
__device__ __forceinline__ float valid_mask(float X, float Y, float Z, int w, int h) {
  return (X >= 0.0f && X < (float)(w - 1) && Y >= 0.0f && Y < (float)(h - 1) && Z > 0.0f) ? 1.0f : 0.0f;
}

// torch.clamp(Z, min=1e-5) (camera.py:158): NaN stays NaN
__device__ __forceinline__ float clamp_depth(float Z) { return Z < 1e-5f ? 1e-5f : Z; }

__device__ __forceinline__ float bilinear4(const float* __restrict__ q, int w, float w00, float w01, float w10, float w11) {
  const float t00 = __ldg(q), t01 = __ldg(q + 1), t10 = __ldg(q + w), t11 = __ldg(q + w + 1);
  return t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11;   // ATen's order: nw, ne, sw, se
}

// a / b with one refined reciprocal (<= 1 ulp); exotic denominators take the IEEE path
__device__ __forceinline__ float fdiv(float a, float b) {
  const float ab = fabsf(b);
  if (ab > 1e-30f && ab < 1e30f) {
    float r = rcp_approx(b);
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
    const float q = a * r;
    return fmaf(fmaf(-q, b, a), r, q);
  }
  return a / b;
}

// Forward / backward phase 1: everything one staged position of a tile needs.
struct MotionStage {
  const float* __restrict__ depth_a;   // this sample
  const float* __restrict__ depth_b;
  const float* __restrict__ frame_a;
  const float* __restrict__ frame_b;
  const float* __restrict__ field;     // nullptr = rigid
  const float4* __restrict__ packed = nullptr;   // frame_b + depth_b interleaved per pixel (warp mode) or nullptr
  float* planes;
  int oy, ox, h, w, hw;
  float m2;                            // depth_err_2nd_mom of this sample
};

struct MotionSample {
  float S[3], Sd, A[3], d;
  float dSx[3], dSy[3];                // d S_c / d(X, Y), gated as nan_to_num / clamp gate the gradient (want_deriv)
  float Zc, valid, occ, wgt;
  float Xs, Ys;                        // clamped pixel coordinates
};

// Projects pixel `pix` = (gy, gx), gathers rgb + depth of B, loads A; no shared memory involved.
// (the depth of A and the residual translation of the pixel are passed in, so that a caller can have them in flight
// for the next pixel while this one is processed)
__device__ __forceinline__ void motion_sample_pre(const MotionStage& a, const MCam& mc, int gy, int gx, int pix, float dA,
                                                  float f0, float f1, float f2v, bool rgb, MotionSample& o,
                                                  bool load_a = true, bool want_deriv = false);
__device__ __forceinline__ void motion_sample(const MotionStage& a, const MCam& mc, int gy, int gx, int pix, bool rgb,
                                              MotionSample& o, bool load_a = true, bool want_deriv = false) {
  const float dA = __ldg(a.depth_a + pix);
  float f0 = 0.0f, f1 = 0.0f, f2v = 0.0f;
  if (a.field) { f0 = __ldg(a.field + pix); f1 = __ldg(a.field + pix + a.hw); f2v = __ldg(a.field + pix + 2 * a.hw); }
  motion_sample_pre(a, mc, gy, gx, pix, dA, f0, f1, f2v, rgb, o, load_a, want_deriv);
}
__device__ __forceinline__ void motion_sample_pre(const MotionStage& a, const MCam& mc, int gy, int gx, int pix, float dA,
                                                  float f0, float f1, float f2v, bool rgb, MotionSample& o,
                                                  bool load_a, bool want_deriv) {
  o.d = dA;
  float P[3], den, X, Y, Z;
  mproject(mc, (float)gx, (float)gy, o.d, f0, f1, f2v, P, den, X, Y, Z);
  const Cell cell = bilinear_cell(X, Y, a.w, a.h);
  const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
  const float w00 = bx * by, w01 = cell.ax * by, w10 = bx * cell.ay, w11 = cell.ax * cell.ay;
  if (a.packed && rgb) {
    // one 16-byte load per tap brings r, g, b and depth: 4 requests per pixel instead of 16, and every byte of the
    // sectors they touch is used (the planar taps of a 1920-wide noisy-depth gather hit 23 sectors per request)
    const float4* q = a.packed + cell.off;
    const float4 t00 = __ldg(q), t01 = __ldg(q + 1), t10 = __ldg(q + a.w), t11 = __ldg(q + a.w + 1);
    const float gate_x = (X >= 0.0f && X <= (float)(a.w - 1)) ? 1.0f : 0.0f;
    const float gate_y = (Y >= 0.0f && Y <= (float)(a.h - 1)) ? 1.0f : 0.0f;
    o.Sd = t00.w * w00 + t01.w * w01 + t10.w * w10 + t11.w * w11;   // ATen's order: nw, ne, sw, se
    const float a00[3] = {t00.x, t00.y, t00.z}, a01[3] = {t01.x, t01.y, t01.z};
    const float a10[3] = {t10.x, t10.y, t10.z}, a11[3] = {t11.x, t11.y, t11.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o.S[c] = a00[c] * w00 + a01[c] * w01 + a10[c] * w10 + a11[c] * w11;
      if (want_deriv) {
        o.dSx[c] = gate_x * ((a01[c] - a00[c]) * by + (a11[c] - a10[c]) * cell.ay);
        o.dSy[c] = gate_y * ((a10[c] - a00[c]) * bx + (a11[c] - a01[c]) * cell.ax);
      }
      if (load_a) o.A[c] = __ldg(a.frame_a + pix + c * a.hw);
    }
  } else {
  o.Sd = bilinear4(a.depth_b + cell.off, a.w, w00, w01, w10, w11);
  if (rgb) {
    // gradient gates of nan_to_num and clamp (closed interval, camera.py:184-188); false for NaN / +-inf
    const float gate_x = (X >= 0.0f && X <= (float)(a.w - 1)) ? 1.0f : 0.0f;
    const float gate_y = (Y >= 0.0f && Y <= (float)(a.h - 1)) ? 1.0f : 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* q = a.frame_b + cell.off + c * a.hw;
      const float t00 = __ldg(q), t01 = __ldg(q + 1), t10 = __ldg(q + a.w), t11 = __ldg(q + a.w + 1);
      o.S[c] = t00 * w00 + t01 * w01 + t10 * w10 + t11 * w11;   // ATen's order: nw, ne, sw, se
      if (want_deriv) {
        o.dSx[c] = gate_x * ((t01 - t00) * by + (t11 - t10) * cell.ay);
        o.dSy[c] = gate_y * ((t10 - t00) * bx + (t11 - t01) * cell.ax);
      }
      if (load_a) o.A[c] = __ldg(a.frame_a + pix + c * a.hw);
    }
  }
  }
  o.valid = valid_mask(X, Y, Z, a.w, a.h);
  o.Zc = clamp_depth(Z);
  o.occ = (o.Zc < o.Sd ? 1.0f : 0.0f) * o.valid;
  const float e = o.Zc - o.Sd;
  const float derr = e * e;
  o.wgt = fdiv(a.m2, derr + a.m2) * o.valid;
  o.Xs = fminf(fmaxf(X, 0.0f), (float)(a.w - 1));
  o.Ys = fminf(fmaxf(Y, 0.0f), (float)(a.h - 1));
}

__device__ __forceinline__ void decode_motion_tile(int bid, int per_dir, int tx_n, int ty_n, int tw, int th, int& dir,
                                                   int& b, int& x0, int& y0) {
  dir = bid / per_dir;
  int t = bid - dir * per_dir;
  const int tx = t % tx_n;
  t /= tx_n;
  const int ty = t % ty_n;
  b = t / ty_n;
  x0 = tx * tw;
  y0 = ty * th;
}

// Tensor maps of the TMA path (second kernel parameter), box {pitch, 18, 1}
struct alignas(64) MotionTma {
  CUtensorMap frame_a[SDE_MAX_DIRS];   // [B*3, h, w]
  CUtensorMap depth_a[SDE_MAX_DIRS];   // [B, h, w]
  CUtensorMap warped[SDE_MAX_DIRS];    // [B*12, h, w]
};

// proximity weight (MotionLearning.py:279-282) from the planes the warp kernel left: depth error and
// vo = valid + 2 * occlusion
__device__ __forceinline__ float proximity_weight(float derr, float vo, float m2) {
  return fdiv(m2, derr + m2) * (vo != 0.0f ? 1.0f : 0.0f);
}

// Weighted-SSIM terms of one pixel pair from the window sums (ssim_loss.py:84-111).
struct WSsim {
  f2 mx, my, ssim;
  f2 n1, d1, n2, d2;   // factors in use (mode-dependent)
};

}  // namespace sde
