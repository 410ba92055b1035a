// Device functions shared by the forward and backward MonoDepth2 loss kernels.
#pragma once
#include "mono_params.cuh"
#include "tma.cuh"

namespace sde {

// Shared-memory plane of a tile + 1-pixel halo: 18 rows x 66 columns.  Column xx is stored at index
// xx + kColOff of a kPitch-float row; kColOff is odd so that the pixel pair owned by a lane (columns
// 2*lane+1, 2*lane+2) is 8-byte aligned and lands in one register pair with a single LDS.64.
// A translation unit may pick its own (pitch, offset) before including this header: the TMA path needs
// the image column of index 0 to be a multiple of 4 (16-byte aligned box start), which is column
// tile_x0 - 4 with (72, 3) in the forward kernel (tile at columns 1..64) and with (68, 1) in the
// backward kernel (gradient block at columns 3..62).
#ifndef SDE_PITCH
#define SDE_PITCH 68
#endif
#ifndef SDE_COL_OFF
#define SDE_COL_OFF 1
#endif
constexpr int kHW = kTileW + 2;       // 66 columns
constexpr int kHH = kTileH + 2;       // 18 rows
constexpr int kPitch = SDE_PITCH;
constexpr int kColOff = SDE_COL_OFF;
constexpr int kPlane = ((kHH * kPitch * 4 + 127) / 128) * 32;   // floats; every plane starts 128-byte aligned (TMA)
static_assert(kPlane >= kHH * kPitch && (kPlane * 4) % 128 == 0 && (kColOff & 1) == 1 && kHW + kColOff <= kPitch, "plane layout");
constexpr unsigned kPlaneBytesTma = kHH * kPitch * 4;   // bytes one {kPitch, 18, 1} box delivers
constexpr int kPositions = kHH * kHW; // 1188 staged positions

__device__ __forceinline__ int plane_index(int yy, int xx) {
  SDE_CHECK(yy >= 0 && yy < kHH && xx >= -kColOff && xx + kColOff < kPitch);
  return yy * kPitch + xx + kColOff;
}

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
// The variant with `q` also returns the reciprocal itself (refined, <= 1 ulp): the scale of d (X, Y) / d p.
__device__ __forceinline__ void divide2(float p0, float p1, float den, float& X, float& Y, float& q) {
  const float ad = fabsf(den);
  if (ad > 1e-30f && ad < 1e30f) {
    float r = rcp_approx(den);
    r = fmaf(fmaf(-den, r, 1.0f), r, r);
    X = p0 * r;
    Y = p1 * r;
    X = fmaf(fmaf(-X, den, p0), r, X);
    Y = fmaf(fmaf(-Y, den, p1), r, Y);
    q = r;
  } else {
    X = p0 / den;
    Y = p1 / den;
    q = 1.0f / den;
  }
}
__device__ __forceinline__ void divide2(float p0, float p1, float den, float& X, float& Y) {
  float q;
  divide2(p0, p1, den, X, Y, q);
}

// Back-project pixel (gx, gy) with depth d through K^-1, move by (R, t), project with K
// (camera.py:125-163,172-178).  P = K^-1 [x d, y d, d], p = (K R) P + K t, X = p0 / (p2 + 1e-6).
// Two halves, so that a caller that projects one pixel into several sources back-projects once.
__device__ __forceinline__ void backproject(const Cam& c, float gx, float gy, float d, float P[3]) {
  const float xd = gx * d, yd = gy * d;
  P[0] = c.ki[0] * xd + c.ki[1] * yd + c.ki[2] * d;
  P[1] = c.ki[3] * xd + c.ki[4] * yd + c.ki[5] * d;
  P[2] = c.ki[6] * xd + c.ki[7] * yd + c.ki[8] * d;
}
__device__ __forceinline__ void project_point(const Proj& pj, const float P[3], float& den, float& X, float& Y, float& q) {
  const float p0 = pj.m[0] * P[0] + pj.m[1] * P[1] + pj.m[2] * P[2] + pj.tau[0];
  const float p1 = pj.m[3] * P[0] + pj.m[4] * P[1] + pj.m[5] * P[2] + pj.tau[1];
  const float p2 = pj.m[6] * P[0] + pj.m[7] * P[1] + pj.m[8] * P[2] + pj.tau[2];
  den = p2 + 1e-6f;
  divide2(p0, p1, den, X, Y, q);
}
__device__ __forceinline__ void project_point(const Proj& pj, const float P[3], float& den, float& X, float& Y) {
  float q;
  project_point(pj, P, den, X, Y, q);
}
__device__ __forceinline__ void project_full(const Cam& c, const Proj& pj, float gx, float gy, float d, float P[3],
                                             float& den, float& X, float& Y) {
  backproject(c, gx, gy, d, P);
  project_point(pj, P, den, X, Y);
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
  SDE_CHECK(x0 >= 0 && x0 + 1 < w && y0 >= 0 && y0 + 1 < h);
  Cell c;
  c.off = y0 * w + x0;
  c.ax = ix - (float)x0;
  c.ay = iy - (float)y0;
  return c;
}

// Everything phase 1 needs about one (tile, source): where the halo'd tile sits, which planes
// receive what.
struct StageArgs {
  int depth_mode;                    // SDE_DEPTH_IS_*: how stage_target decodes what it loads from `depth`
  float min_disp, disp_range;
  const float* __restrict__ depth;   // [h*w] of this sample
  const float* __restrict__ src;     // [3*h*w] source frame of this sample
  const float* __restrict__ tgt;     // [3*h*w] target frame of this sample
  const uint8_t* __restrict__ amap;  // [h*w] argmin bytes (backward) or nullptr
  float* planes;                     // shared-memory plane array
  uint8_t* arg;                      // shared-memory argmin plane (backward)
  int oy, ox;                        // image coordinates of plane position (0, 0)
  int h, w, hw;
  int plS, plI, plA, plD;            // plane numbers: warped, identity, target, depth
};

// Pins a base pointer in a register pair so that every access is one IMAD.WIDE (index * 4 + base)
// instead of a re-derived 64-bit sum of the sample / plane / pixel offsets.
template <typename T>
__device__ __forceinline__ const T* pinned(const T* p) {
  unsigned long long v = reinterpret_cast<unsigned long long>(p);
  asm volatile("" : "+l"(v));
  return reinterpret_cast<const T*>(v);
}

__device__ __forceinline__ void position_of(int i, int& yy, int& xx) {
  yy = (i * 993) >> 16;  // i / 66 for i < 1188
  xx = i - yy * kHW;
}

template <bool INTERIOR>
__device__ __forceinline__ int pixel_of(const StageArgs& a, int yy, int xx, int& gy, int& gx) {
  gy = a.oy + yy;
  gx = a.ox + xx;
  if (!INTERIOR) {
    gy = reflect_clamp(gy, a.h);
    gx = reflect_clamp(gx, a.w);
  }
  SDE_CHECK(gy >= 0 && gy < a.h && gx >= 0 && gx < a.w);
  return gy * a.w + gx;
}

// 1 / clamp(d, min=1e-6) (smoothness_loss.py:62) with a refined hardware reciprocal (<= 1 ulp);
// the comparison form keeps a NaN depth NaN, as torch.clamp does.
__device__ __forceinline__ float inv_depth(float d) {
  const float c = d < 1e-6f ? 1e-6f : d;
  float r = rcp_approx(c);
  return fmaf(fmaf(-c, r, 1.0f), r, r);
}

// d smoothness / d (1/depth) of one pixel, before the division by the per-image mean
// (smoothness_loss.py:62-80): +-exp(-mean_c |dI|) / N over the four edges of the pixel.
// pd / pa0 point at the pixel in the depth plane / first image plane (planes kPlane apart).
__device__ __forceinline__ float smooth_grad_local(const float* pd, const float* pa0, float ic, int gx, int gy, int w, int h,
                                                   float inx, float iny) {
  float el = 0.0f, er = 0.0f, eu = 0.0f, edn = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pa = pa0 + c * kPlane;
    const float a = pa[0];
    el += fabsf(pa[-1] - a); er += fabsf(a - pa[1]);
    eu += fabsf(pa[-kPitch] - a); edn += fabsf(a - pa[kPitch]);
  }
  auto sgn = [](float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); };
  float G = 0.0f;
  if (gx + 1 < w) G += sgn(ic - inv_depth(pd[1])) * __expf(-er * (1.0f / 3.0f)) * inx;
  if (gx >= 1) G -= sgn(inv_depth(pd[-1]) - ic) * __expf(-el * (1.0f / 3.0f)) * inx;
  if (gy + 1 < h) G += sgn(ic - inv_depth(pd[kPitch])) * __expf(-edn * (1.0f / 3.0f)) * iny;
  if (gy >= 1) G -= sgn(inv_depth(pd[-kPitch]) - ic) * __expf(-eu * (1.0f / 3.0f)) * iny;
  return G;
}

// Edge-aware smoothness of the forward tile kernels (smoothness_loss.py:62-80), evaluated on the edges of this lane's
// pixel pairs (rows r0 .. r0+3 of the tile, columns c0, c0+1; planes with a one-pixel halo).  Adds the sum of inverse
// depths and the un-normalised edge sums of the tile to sinv / smx / smy and, when `gout` (this sample's [h,w] plane)
// is given, stores d smoothness / d (1/depth) before the division by the per-image mean.
__device__ __forceinline__ void tile_smoothness(const float* planes, int plD, int plA, int r0, int c0, int x0, int y0, int h,
                                                int w, int B, float* __restrict__ gout, float& smx, float& smy, float& sinv) {
  const int gx0 = x0 + c0;  // image column of this lane's first pixel
  {
    // Edge-aware smoothness (smoothness_loss.py:62-80) on the edges of this lane's pixel pairs.  An edge between
    // pixels u and v (v right of / below u) with inverse depths iu, iv carries the weight e = exp(-mean_c |dI|);
    // it adds |iu - iv| e to the loss sums and +-sgn(iu - iv) e / N to d loss / d (1/depth) of u and v (before
    // the division by the per-image mean) -- kept in smooth_g for the backward pass.
    // Halo'd rows r0 .. r0+5 (owned rows: r0+1 .. r0+4); columns c0 (left neighbour) .. c0+3 (right neighbour).
    const float* pd = planes + plD * kPlane + plane_index(r0, c0 + 1);
    const float* pa = planes + plA * kPlane + plane_index(r0, c0 + 1);
    // 1 / clamp(d, min=1e-6) through a refined hardware reciprocal (<= 1 ulp; NaN stays NaN), exp through ex2.approx
    // (2 ulp): both far inside the 1e-5 loss tolerance and a fifth of the instructions of the IEEE forms
    auto inv = [](float d) { return inv_depth(d); };
    // sgn(v) * ws for ws >= 0 (sgn(0) = sgn(NaN) = 0): the sign bit of v rides onto ws, one compare selects
    auto sgn_mul = [](float v, float ws) {
      const float sw = __int_as_float(__float_as_int(ws) ^ (__float_as_int(v) & 0x80000000));
      return fabsf(v) > 0.0f ? sw : 0.0f;
    };
    const float inx = 1.0f / ((float)B * (float)h * (float)(w - 1));
    const float iny = 1.0f / ((float)B * (float)(h - 1) * (float)w);
    const bool v0 = gx0 < w, v1 = gx0 + 1 < w, v2 = gx0 + 2 < w, vl = gx0 >= 1 && v0;
    const bool pair_ok = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(gout) & 7) == 0;   // gx0 is even
    // vertical edges between halo'd rows rr and rr + 1 (image rows gyu, gyu + 1): signed gradient terms
    float tv0[kRowsPerWarp + 1], tv1[kRowsPerWarp + 1];
    float ir0[kRowsPerWarp + 2], ir1[kRowsPerWarp + 2];   // inverse depths of this lane's pair in the halo'd rows
    f2 du = ld2(pd);
    float iu0 = inv(lo(du)), iu1 = inv(hi(du));
    ir0[0] = iu0; ir1[0] = iu1;
    f2 au[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) au[c] = ld2(pa + c * kPlane);
#pragma unroll
    for (int rr = 0; rr <= kRowsPerWarp; ++rr) {
      const int gyu = y0 + r0 + rr - 1;
      const f2 dl = ld2(pd + (rr + 1) * kPitch);
      const float il0 = inv(lo(dl)), il1 = inv(hi(dl));
      ir0[rr + 1] = il0; ir1[rr + 1] = il1;
      float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const f2 al = ld2(pa + c * kPlane + (rr + 1) * kPitch);
        e0 += fabsf(lo(au[c]) - lo(al));
        e1 += fabsf(hi(au[c]) - hi(al));
        au[c] = al;
      }
      const bool ve = gyu >= 0 && gyu + 1 < h;
      const float w0 = __expf(-e0 * (1.0f / 3.0f)), w1 = __expf(-e1 * (1.0f / 3.0f));
      tv0[rr] = (ve && v0) ? sgn_mul(iu0 - il0, w0 * iny) : 0.0f;
      tv1[rr] = (ve && v1) ? sgn_mul(iu1 - il1, w1 * iny) : 0.0f;
      if (rr >= 1) {   // the edge below an owned row is counted by this lane
        if (ve && v0) smy += fabsf(iu0 - il0) * w0;
        if (ve && v1) smy += fabsf(iu1 - il1) * w1;
      }
      iu0 = il0; iu1 = il1;
    }
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) {
      const int gy = y0 + r0 + o;
      const float* pdo = pd + (o + 1) * kPitch;
      const float iL = inv(pdo[-1]), i0 = ir0[o + 1], i1 = ir1[o + 1], iR = inv(pdo[2]);
      float eL = 0.0f, eM = 0.0f, eR = 0.0f;  // sum_c |dI|
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pac = pa + c * kPlane + (o + 1) * kPitch;
        const f2 ac = ld2(pac);
        const float aL = pac[-1], a0 = lo(ac), a1 = hi(ac), aR = pac[2];
        eL += fabsf(aL - a0);
        eM += fabsf(a0 - a1);
        eR += fabsf(a1 - aR);
      }
      if (gy < h) {
        const float wL = __expf(-eL * (1.0f / 3.0f)), wM = __expf(-eM * (1.0f / 3.0f)), wR = __expf(-eR * (1.0f / 3.0f));
        if (v0) sinv += i0;
        if (v1) sinv += i1;
        if (v1) smx += fabsf(i0 - i1) * wM;
        if (v2) smx += fabsf(i1 - iR) * wR;
        if (gout) {
          const float tL = vl ? sgn_mul(iL - i0, wL * inx) : 0.0f;
          const float tM = v1 ? sgn_mul(i0 - i1, wM * inx) : 0.0f;
          const float tR = v2 ? sgn_mul(i1 - iR, wR * inx) : 0.0f;
          const float G0 = ((tM - tL) + tv0[o + 1]) - tv0[o];
          const float G1 = ((tR - tM) + tv1[o + 1]) - tv1[o];
          float* po = gout + gy * w + gx0;
          if (v1 && pair_ok) {
            *reinterpret_cast<float2*>(po) = make_float2(G0, G1);
          } else {
            if (v0) po[0] = G0;
            if (v1) po[1] = G1;
          }
        }
      }
    }
  }
}

// positions a thread keeps in flight in the plain (coalesced) staging loops: 5 x 4 loads, two round trips per tile
#ifndef SDE_STAGE_BATCH
#define SDE_STAGE_BATCH 5
#endif
constexpr int kStageBatch = SDE_STAGE_BATCH;

// Phase 0 of both kernels: stage what does not depend on the source -- depth, target (and the
// argmin bytes in the backward kernel) of the halo'd tile.  All loads are independent and issued
// before the first store, so the tile pays one memory round trip here.
template <bool INTERIOR, bool ARG>
__device__ __forceinline__ void stage_target(const StageArgs& a0, int tid, bool reduce_mean) {
  StageArgs a = a0;
  a.depth = pinned(a0.depth);
  a.tgt = pinned(a0.tgt);
  constexpr int kIter = (kPositions + kThreads - 1) / kThreads;  // 10
#pragma unroll 1
  for (int k = 0; k < kIter; k += kStageBatch) {
    float d[kStageBatch], t[kStageBatch][3];
    uint8_t m[kStageBatch];
    int pl[kStageBatch];
    bool ok[kStageBatch];
#pragma unroll
    for (int u = 0; u < kStageBatch; ++u) {
      const int i = tid + (k + u) * kThreads;
      ok[u] = i < kPositions;
      int yy, xx, gy, gx;
      position_of(ok[u] ? i : tid, yy, xx);
      const int pix = pixel_of<INTERIOR>(a, yy, xx, gy, gx);
      pl[u] = plane_index(yy, xx);
      d[u] = decode_depth(__ldg(a.depth + pix), a.depth_mode, a.min_disp, a.disp_range);
#pragma unroll
      for (int c = 0; c < 3; ++c) t[u][c] = __ldg(a.tgt + (pix + c * a.hw));
      if (ARG) {
        const int ty = a.oy + yy, tx = a.ox + xx;
        const bool inside = ty >= 0 && ty < a.h && tx >= 0 && tx < a.w;
        // 255 never matches a candidate: windows centred outside the image do not exist
        m[u] = inside ? (reduce_mean ? (uint8_t)254 : ld_prod(a.amap + pix)) : (uint8_t)255;
      }
    }
#pragma unroll
    for (int u = 0; u < kStageBatch; ++u) {
      if (ok[u]) {
        a.planes[a.plD * kPlane + pl[u]] = d[u];
#pragma unroll
        for (int c = 0; c < 3; ++c) a.planes[(a.plA + c) * kPlane + pl[u]] = t[u][c];
        if (ARG) a.arg[pl[u]] = m[u];
      }
    }
  }
}

// Phase 1 of both kernels: project every staged position into one source (depth comes from shared
// memory), gather its four bilinear taps per channel and store the warped value; NB positions per
// thread are in flight at a time (12 NB independent gather loads).
//   IDENT: also stage the unwarped source (identity / automask candidate)
template <bool INTERIOR, bool IDENT, int NB>
__device__ __forceinline__ void stage_source(const StageArgs& a0, const Cam& cam, const Proj& pj, int tid) {
  StageArgs a = a0;
  a.src = pinned(a0.src);
  const int w = a.w, hw = a.hw;
#pragma unroll 1
  for (int base = tid; base < kPositions; base += NB * kThreads) {
    int pl[NB], off[NB], pix[NB];
    float w00[NB], w01[NB], w10[NB], w11[NB];
    bool ok[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int i = base + u * kThreads;
      ok[u] = i < kPositions;
      int yy, xx, gy, gx;
      position_of(ok[u] ? i : base, yy, xx);
      pix[u] = pixel_of<INTERIOR>(a, yy, xx, gy, gx);
      pl[u] = plane_index(yy, xx);
      const float d = a.planes[a.plD * kPlane + pl[u]];
      float P[3], den, X, Y;
      project_full(cam, pj, (float)gx, (float)gy, d, P, den, X, Y);
      const Cell cell = bilinear_cell(X, Y, w, a.h);
      const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
      off[u] = cell.off;
      w00[u] = bx * by; w01[u] = cell.ax * by; w10[u] = bx * cell.ay; w11[u] = cell.ax * cell.ay;
    }
    float v[NB][3], id[NB][3];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* q = a.src + (off[u] + c * hw);
        const float t00 = __ldg(q), t01 = __ldg(q + 1), t10 = __ldg(q + w), t11 = __ldg(q + w + 1);
        if (IDENT) id[u][c] = __ldg(a.src + (pix[u] + c * hw));
        // ATen's accumulation order: nw, ne, sw, se
        v[u][c] = t00 * w00[u] + t01 * w01[u] + t10 * w10[u] + t11 * w11[u];
      }
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (ok[u]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          a.planes[(a.plS + c) * kPlane + pl[u]] = v[u][c];
          if (IDENT) a.planes[(a.plI + c) * kPlane + pl[u]] = id[u][c];
        }
      }
    }
  }
}

// Phase 1 of the backward kernel when the forward pass kept the warped source: plain coalesced
// loads of the halo'd tile (a reflected halo position takes the warp of the reflected pixel,
// which is what re-projecting it would give).
// (What another kernel of the same step produced -- argmin bytes, warped planes, smoothness gradient -- is read through
// ld_prod, sde_common.cuh: under flow the producer grid may still be running.)
template <bool INTERIOR>
__device__ __forceinline__ void stage_saved(const StageArgs& a0, const float* __restrict__ warped, int tid) {
  StageArgs a = a0;
  const float* wp = pinned(warped);
  constexpr int kIter = (kPositions + kThreads - 1) / kThreads;  // 10
#pragma unroll 1
  for (int k = 0; k < kIter; k += kStageBatch) {
    float v[kStageBatch][3];
    int pl[kStageBatch];
    bool ok[kStageBatch];
#pragma unroll
    for (int u = 0; u < kStageBatch; ++u) {
      const int i = tid + (k + u) * kThreads;
      ok[u] = i < kPositions;
      int yy, xx, gy, gx;
      position_of(ok[u] ? i : tid, yy, xx);
      const int pix = pixel_of<INTERIOR>(a, yy, xx, gy, gx);
      pl[u] = plane_index(yy, xx);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[u][c] = ld_prod(wp + (pix + c * a.hw));
    }
#pragma unroll
    for (int u = 0; u < kStageBatch; ++u)
      if (ok[u]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) a.planes[(a.plS + c) * kPlane + pl[u]] = v[u][c];
      }
  }
}

// A depth plane that arrived through TMA as disparity / logits is decoded in place, every stored
// element once (out-of-image zeros become some finite depth that no valid output reads).
__device__ __forceinline__ void decode_depth_plane(float* plane, int mode, float min_disp, float range, int tid) {
  if (mode == SDE_DEPTH_IS_DEPTH) return;
  for (int i = tid; i < kHH * kPitch; i += kThreads) plane[i] = decode_depth(plane[i], mode, min_disp, range);
}

// After a TMA box load of a tile that touches the image border: out-of-image elements arrived as zeros.
// Positions exactly one pixel outside are nn.ReflectionPad2d(1) positions (ssim_loss.py:32,35) and take the
// mirrored pixel, which always lies inside the same staged tile; positions further out are never read by a
// valid output and stay zero.  n_planes consecutive planes starting at plane0.
__device__ __forceinline__ void reflect_fixup(float* planes, int plane0, int n_planes, int oy, int ox, int h, int w, int tid) {
  // Only the two plane rows / two plane columns that sit one pixel outside the image can hold such positions:
  // walk those lines (2 * 66 + 2 * 18 candidates) instead of the whole tile.  Row lines take the corners (mirrored
  // in both directions); every source position lies inside the image, so no fix-up reads another one's result.
  for (int i = tid; i < 2 * kHW + 2 * kHH; i += kThreads) {
    int yy, xx;
    bool ok;
    if (i < 2 * kHW) {
      const bool top = i < kHW;
      xx = top ? i : i - kHW;
      yy = (top ? -1 : h) - oy;
      const int tx = ox + xx;
      ok = yy >= 0 && yy < kHH && tx >= -1 && tx <= w;
    } else {
      const int k = i - 2 * kHW;
      const bool left = k < kHH;
      yy = left ? k : k - kHH;
      xx = (left ? -1 : w) - ox;
      const int ty = oy + yy;
      ok = xx >= 0 && xx < kHW && ty >= 0 && ty < h;
    }
    if (ok) {
      const int ty = oy + yy, tx = ox + xx;
      const int ry = ty == -1 ? 1 : (ty == h ? h - 2 : ty), rx = tx == -1 ? 1 : (tx == w ? w - 2 : tx);
      const int dst = plane_index(yy, xx), src = plane_index(ry - oy, rx - ox);
      for (int k = 0; k < n_planes; ++k) planes[(plane0 + k) * kPlane + dst] = planes[(plane0 + k) * kPlane + src];
    }
  }
}

// argmin bytes of the halo'd tile (backward): 255 where no window centre exists, 254 = every candidate ('mean').
// All of a thread's loads are issued before the first store (one memory round trip per tile).
__device__ __forceinline__ void stage_arg(uint8_t* arg, const uint8_t* __restrict__ amap, int oy, int ox, int h, int w,
                                          bool reduce_mean, int tid) {
  constexpr int kIter = (kPositions + kThreads - 1) / kThreads;  // 10
  uint8_t m[kIter];
  int pl[kIter];
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int i = tid + k * kThreads;
    int yy, xx;
    position_of(i < kPositions ? i : tid, yy, xx);
    const int ty = oy + yy, tx = ox + xx;
    const bool inside = ty >= 0 && ty < h && tx >= 0 && tx < w;
    pl[k] = i < kPositions ? plane_index(yy, xx) : -1;
    m[k] = inside ? (reduce_mean ? (uint8_t)254 : ld_prod(amap + ty * w + tx)) : (uint8_t)255;
  }
#pragma unroll
  for (int k = 0; k < kIter; ++k)
    if (pl[k] >= 0) arg[pl[k]] = m[k];
}

// zeroes the one-pixel ring (rows 0 and kHH-1, columns 0 and kHW-1) of the N planes b0 .. b0+N-1: threads 0..65
// take the two rows, 18 threads of the last warp the two columns -- no index arithmetic beyond the thread id
template <int N>
__device__ __forceinline__ void zero_ring(float* planes, int b0, int tid) {
  static_assert(kHW <= 96 && 96 + kHH <= kThreads, "thread ranges of zero_ring");
  if (tid < kHW) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      planes[(b0 + k) * kPlane + plane_index(0, tid)] = 0.0f;
      planes[(b0 + k) * kPlane + plane_index(kHH - 1, tid)] = 0.0f;
    }
  } else if (tid >= 96 && tid < 96 + kHH) {
    const int yy = tid - 96;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      planes[(b0 + k) * kPlane + plane_index(yy, 0)] = 0.0f;
      planes[(b0 + k) * kPlane + plane_index(yy, kHW - 1)] = 0.0f;
    }
  }
}

// The same plane, four bytes per load: possible when the image column of plane index 0 (ox - kColOff) and the row
// pitch of the map are multiples of 4 -- then every aligned word of a plane row lies entirely inside or entirely
// outside the image.  18 x 17 word loads per tile instead of 1188 byte loads.
__device__ __forceinline__ void stage_arg_words(uint8_t* arg, const uint8_t* __restrict__ amap, int oy, int ox, int h, int w,
                                                bool reduce_mean, int tid) {
  static_assert(kPitch % 4 == 0, "plane rows are whole words");
  constexpr int kWords = kPitch / 4, kTot = kHH * kWords;
  constexpr int kIter = (kTot + kThreads - 1) / kThreads;
  const int cx0 = ox - kColOff;
  unsigned v[kIter];
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int i = tid + k * kThreads;
    const int yy = i / kWords, wd = i - yy * kWords;
    const int ty = oy + yy, tx = cx0 + 4 * wd;
    const bool inside = i < kTot && ty >= 0 && ty < h && tx >= 0 && tx < w;
    v[k] = inside ? (reduce_mean ? 0xfefefefeu : ld_prod(reinterpret_cast<const unsigned*>(amap + ty * w + tx))) : 0xffffffffu;
  }
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const int i = tid + k * kThreads;
    if (i < kTot) reinterpret_cast<unsigned*>(arg)[i] = v[k];
  }
}

}  // namespace sde
