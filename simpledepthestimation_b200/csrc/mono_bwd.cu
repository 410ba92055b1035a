// Fused backward of the MonoDepth2 self-supervised loss, all scales in one launch (sm_100a).
//
// Nothing but the uint8 argmin maps and O(B) per-image statistics is kept from the forward pass:
// the warp is recomputed.  One CTA owns a 64x16 block Q of SSIM window centres of one
// (scale, sample) and emits gradients for Q's 62x14 interior P (a pixel's gradient collects the
// 3x3 windows around it).  Per source frame j:
//   phase 1  as in the forward kernel: project + bilinear gather on Q + 1-pixel halo -> planes S;
//            target A, 1/depth and the argmin bytes are staged on the first source.
//   per colour channel:
//     phase 2  window sums carried down the rows in registers (two pixels per lane, packed f2);
//              for every centre q whose argmin is this source's warped candidate, the three
//              coefficients of  d ssim_q / d S_p = a_q + S_p b_q + A_p c_q  (SURVEY.md A.6),
//              scaled by the upstream gradient, go to shared planes (zero elsewhere).
//     phase 3  adjoint of reflect-pad + 3x3 box: each pixel of P gathers the coefficients of the
//              windows that contain it (border multiplicities 2 where the pad mirrors onto it),
//              adds the L1 term, and stores gS_c(p).
//   phase 4  per pixel of P: re-project, re-read the four taps, gX = sum_c gS_c dS_c/dx (gated as
//            nan_to_num/clamp gate the reference's autograd), then the camera-space gradient
//            K^T g_p, accumulated into 12 pose sums per source (9 for R, 3 for t) and into
//            d loss / d depth.
// The smoothness gradient uses the saved per-image mean inverse depth and loss (1-homogeneity,
// SURVEY.md A.5).  Pose sums: per-CTA slots, added by the last CTA in a fixed order in fp64
// (deterministic; no float atomics anywhere).
#include "mono_device.cuh"

namespace sde {

constexpr int kBwdPlanes = 13;  // A[3], S[3], 1/d, coef a/b/c, gS[3]
constexpr int kBA = 0, kBS = 3, kBInv = 6, kBCoef = 7, kBG = 10;
constexpr int kPosPerThread = (kBwdW * kBwdH + kThreads - 1) / kThreads;  // 7

struct BwdShared {
  Cam cam;
  Proj proj[SDE_MAX_SOURCES];
  float red[12][kThreads / 32];
  unsigned ticket;
  uint8_t arg[kPlane];
};

__global__ void __launch_bounds__(kThreads, 3) mono_bwd_kernel(const __grid_constant__ MonoParams p) {
  extern __shared__ __align__(16) float planes[];  // [kBwdPlanes][kPlane]
  __shared__ BwdShared sh;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const TileCoord tc = decode_btile(p, blockIdx.x);
  const int s = tc.s, b = tc.b, h = p.h[s], w = p.w[s], hw = h * w;
  const bool automask = (p.flags & SDE_MONO_AUTOMASK) != 0;
  const bool reduce_mean = (p.flags & SDE_MONO_REDUCE_MEAN) != 0;
  const bool use_ssim = p.ssim_w > 0.0f;
  // plane (yy, xx) <-> image (oy + yy, ox + xx); Q = plane [1..16]x[1..64]; P = plane [2..15]x[2..63]
  const int ox = tc.x0 - 2, oy = tc.y0 - 2;

  if (tid < p.S) {
    Cam cam;
    float k[9];
    load_cam(cam, k, p.K, b, p.sx[s], p.sy[s]);
    if (tid == 0) sh.cam = cam;
    load_proj(sh.proj[tid], k, p.pose[tid], b);
  }
  // coefficient planes: the 1-pixel border is never written and must read as zero
  for (int i = tid; i < 3 * kPlane; i += kThreads) planes[kBCoef * kPlane + i] = 0.0f;
  __syncthreads();

  const float* __restrict__ depth = p.depth[s] + (size_t)b * hw;
  const float* __restrict__ target = p.target[s] + (size_t)b * 3 * hw;
  const uint8_t* __restrict__ amap = reduce_mean ? nullptr : p.argmin[s] + (size_t)b * hw;

  const float g_rec = __ldg(p.grad_losses), g_smooth = __ldg(p.grad_losses + 1);
  const int ncand_total = (automask ? 2 : 1) * p.S;
  // d rec_loss / d pe_q for a selected pixel (MonoDepth2.py:116-124)
  const float g_pe = g_rec / ((float)p.n_scales * (float)p.B * (float)h * (float)w) /
                     (reduce_mean ? (float)ncand_total : 1.0f);
  const float g_l1 = g_pe * p.l1_w * (1.0f / 3.0f);
  const float g_ss = g_pe * p.ssim_w * (1.0f / 3.0f) * -0.5f;  // d pe / d ssim (ssim_loss.py:53)

  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;
  const f2 C1 = bc2(81.0f * p.c1), C2 = bc2(81.0f * p.c2);

  // image coordinates of this lane's pixel pair and its four rows (as centre q and as pixel p)
  const int px0 = ox + c0 + 1, px1 = px0 + 1;
  // reflect-pad adjoint weights: a border-adjacent pixel also receives the mirrored pad position
  const f2 wL = mk2(px0 == 1 ? 2.0f : 1.0f, px1 == 1 ? 2.0f : 1.0f);
  const f2 wR = mk2(px0 == w - 2 ? 2.0f : 1.0f, px1 == w - 2 ? 2.0f : 1.0f);

  float gd[kPosPerThread];
#pragma unroll
  for (int k = 0; k < kPosPerThread; ++k) gd[k] = 0.0f;

  for (int j = 0; j < p.S; ++j) {
    const int cand = automask ? 2 * j : j;
    // ---------------------------------------------------------------- phase 1
    {
      const Cam cam = sh.cam;
      const Proj pj = sh.proj[j];
      const float* __restrict__ src = p.source[s][j] + (size_t)b * 3 * hw;
      for (int i = tid; i < kPlane; i += kThreads) {
        const int yy = i / kHW, xx = i - yy * kHW;
        const int ty = oy + yy, tx = ox + xx;
        const int gy = reflect_clamp(ty, h), gx = reflect_clamp(tx, w);
        const int pix = gy * w + gx;
        const float d = __ldg(depth + pix);
        float X, Y, sv[3];
        project_px(cam, pj, (float)gx, (float)gy, d, X, Y);
        bilinear3(src, hw, w, h, X, Y, sv);
#pragma unroll
        for (int c = 0; c < 3; ++c) planes[(kBS + c) * kPlane + i] = sv[c];
        if (j == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) planes[(kBA + c) * kPlane + i] = __ldg(target + c * hw + pix);
          planes[kBInv * kPlane + i] = 1.0f / (d < 1e-6f ? 1e-6f : d);  // clamp(min=1e-6) keeps NaN, as torch.clamp does
          const bool inside = ty >= 0 && ty < h && tx >= 0 && tx < w;
          // 255 never matches a candidate: windows centred outside the image do not exist
          sh.arg[i] = inside ? (reduce_mean ? (uint8_t)254 : amap[pix]) : (uint8_t)255;
        }
      }
    }
    __syncthreads();

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      // -------------------------------------------------------------- phase 2: SSIM coefficients on Q
      if (use_ssim) {
        const float* pa = planes + (kBA + c) * kPlane + r0 * kHW + c0;
        const float* px = planes + (kBS + c) * kPlane + r0 * kHW + c0;
        f2 hA[2], hAA[2], hX[2], hXX[2], hXA[2];
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
          const f2 alo = ld2(pa + rr * kHW), ahi = ld2(pa + rr * kHW + 2);
          const f2 xlo = ld2(px + rr * kHW), xhi = ld2(px + rr * kHW + 2);
          const f2 aC = mk2(alo.y, ahi.x), aO = mk2(alo.x, ahi.y);
          const f2 xC = mk2(xlo.y, xhi.x), xO = mk2(xlo.x, xhi.y);
          const f2 aa = aC * aC, xx = xC * xC, xa = xC * aC;
          const f2 nA = aC + swp(aC) + aO, nAA = fma2(aO, aO, aa + swp(aa));
          const f2 nX = xC + swp(xC) + xO, nXX = fma2(xO, xO, xx + swp(xx)), nXA = fma2(xO, aO, xa + swp(xa));
          if (rr >= 2) {
            const int row = r0 + rr - 1;  // plane row of the window centre
            const uint8_t* pm = sh.arg + row * kHW + c0 + 1;
            const int m0 = pm[0], m1 = pm[1];
            const bool sel0 = reduce_mean ? (m0 != 255) : (m0 == cand);
            const bool sel1 = reduce_mean ? (m1 != 255) : (m1 == cand);
            f2 ca = bc2(0.0f), cb = bc2(0.0f), cc = bc2(0.0f);
            if (__any_sync(0xffffffffu, sel0 || sel1)) {
              const f2 sA = hA[0] + hA[1] + nA, sAA = hAA[0] + hAA[1] + nAA;
              const f2 sX = hX[0] + hX[1] + nX, sXX = hXX[0] + hXX[1] + nXX, sXA = hXA[0] + hXA[1] + nXA;
              const f2 aa2 = sA * sA, xx2 = sX * sX, t = sX * sA;
              const f2 n1 = fma2(bc2(2.0f), t, C1);
              const f2 n2 = fma2(bc2(2.0f), fma2(bc2(9.0f), sXA, neg2(t)), C2);
              const f2 d1 = (xx2 + aa2) + C1;   // same operation order as the forward kernel
              const f2 d2 = (fma2(bc2(9.0f), sXX, neg2(xx2)) + fma2(bc2(9.0f), sAA, neg2(aa2))) + C2;
              const f2 D = d1 * d2;
              const f2 invD = mk2(1.0f / D.x, 1.0f / D.y);
              const f2 ssim = n1 * n2 * invD;
              // torch.clamp passes the gradient on the closed interval 0 <= (1-ssim)/2 <= 1
              const f2 half = fma2(ssim, bc2(-0.5f), bc2(0.5f));
              const f2 g = mk2((sel0 && half.x >= 0.0f && half.x <= 1.0f) ? g_ss : 0.0f,
                               (sel1 && half.y >= 0.0f && half.y <= 1.0f) ? g_ss : 0.0f);
              const f2 gi = g * invD;
              // d ssim / d(sum S), d(sum S^2), d(sum S A)
              ca = gi * fma2(bc2(2.0f) * sA, n2 - n1, neg2(bc2(2.0f) * sX * ssim * (d2 - d1)));
              cb = gi * bc2(-18.0f) * ssim * d1;   // 2 * (-9 ssim / d2)
              cc = gi * bc2(18.0f) * n1;
            }
            float* pc = planes + kBCoef * kPlane + row * kHW + c0 + 1;
            pc[0] = ca.x; pc[1] = ca.y;
            pc[kPlane] = cb.x; pc[kPlane + 1] = cb.y;
            pc[2 * kPlane] = cc.x; pc[2 * kPlane + 1] = cc.y;
          }
          hA[0] = hA[1]; hA[1] = nA; hAA[0] = hAA[1]; hAA[1] = nAA;
          hX[0] = hX[1]; hX[1] = nX; hXX[0] = hXX[1]; hXX[1] = nXX; hXA[0] = hXA[1]; hXA[1] = nXA;
        }
      }
      __syncthreads();
      // -------------------------------------------------------------- phase 3: adjoint gather -> gS_c on P
      {
        f2 hq[3][2];  // horizontal (weighted) 3-sums of a, b, c for the two previous coefficient rows
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
          f2 nq[3];
          if (use_ssim) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const float* pc = planes + (kBCoef + k) * kPlane + (r0 + rr) * kHW + c0;
              const f2 lo = ld2(pc), hi = ld2(pc + 2);
              nq[k] = fma2(wL, lo, fma2(wR, hi, mk2(lo.y, hi.x)));
            }
          }
          if (rr >= 2) {
            const int row = r0 + rr - 1;        // plane row of pixel p
            const int py = oy + row;
            const float wu = py == 1 ? 2.0f : 1.0f, wd = py == h - 2 ? 2.0f : 1.0f;
            const f2 Sp = mk2(planes[(kBS + c) * kPlane + row * kHW + c0 + 1],
                              planes[(kBS + c) * kPlane + row * kHW + c0 + 2]);
            const f2 Ap = mk2(planes[(kBA + c) * kPlane + row * kHW + c0 + 1],
                              planes[(kBA + c) * kPlane + row * kHW + c0 + 2]);
            f2 gS = bc2(0.0f);
            if (use_ssim) {
              const f2 va = fma2(bc2(wu), hq[0][0], fma2(bc2(wd), nq[0], hq[0][1]));
              const f2 vb = fma2(bc2(wu), hq[1][0], fma2(bc2(wd), nq[1], hq[1][1]));
              const f2 vc = fma2(bc2(wu), hq[2][0], fma2(bc2(wd), nq[2], hq[2][1]));
              gS = fma2(Sp, vb, fma2(Ap, vc, va));
            }
            // L1 term on the pixel itself: sign(S - A) where this candidate was selected
            const uint8_t* pm = sh.arg + row * kHW + c0 + 1;
            const int m0 = pm[0], m1 = pm[1];
            const bool sel0 = reduce_mean ? (m0 != 255) : (m0 == cand);
            const bool sel1 = reduce_mean ? (m1 != 255) : (m1 == cand);
            const f2 df = Sp - Ap;
            if (sel0) gS.x += df.x > 0.0f ? g_l1 : (df.x < 0.0f ? -g_l1 : 0.0f);
            if (sel1) gS.y += df.y > 0.0f ? g_l1 : (df.y < 0.0f ? -g_l1 : 0.0f);
            float* pg = planes + (kBG + c) * kPlane + row * kHW + c0 + 1;
            pg[0] = gS.x; pg[1] = gS.y;
          }
          if (use_ssim) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { hq[k][0] = hq[k][1]; hq[k][1] = nq[k]; }
          }
        }
      }
      __syncthreads();
    }

    // ---------------------------------------------------------------- phase 4: warp backward on P
    {
      const Cam cam = sh.cam;
      const Proj pj = sh.proj[j];
      const float* __restrict__ src = p.source[s][j] + (size_t)b * 3 * hw;
      float acc[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
#pragma unroll
      for (int it = 0; it < kPosPerThread; ++it) {
        const int i = tid + it * kThreads;
        const int ly = i / kBwdW, lx = i - ly * kBwdW;
        const int gy = tc.y0 + ly, gx = tc.x0 + lx;
        if (i < kBwdW * kBwdH && gy < h && gx < w) {
          const int pl = (ly + 2) * kHW + lx + 2;
          const float g0 = planes[kBG * kPlane + pl], g1 = planes[(kBG + 1) * kPlane + pl],
                      g2 = planes[(kBG + 2) * kPlane + pl];
          if (g0 != 0.0f || g1 != 0.0f || g2 != 0.0f) {
            const float d = __ldg(depth + gy * w + gx);
            const float fxp = (float)gx, fyp = (float)gy;
            // ray = K^-1 [x,y,1], P = K^-1 [x d, y d, d]  (camera.py:125-138)
            const float xd = fxp * d, yd = fyp * d;
            const float Px = cam.ki[0] * xd + cam.ki[1] * yd + cam.ki[2] * d;
            const float Py = cam.ki[3] * xd + cam.ki[4] * yd + cam.ki[5] * d;
            const float Pz = cam.ki[6] * xd + cam.ki[7] * yd + cam.ki[8] * d;
            const float p0 = pj.m[0] * Px + pj.m[1] * Py + pj.m[2] * Pz + pj.tau[0];
            const float p1 = pj.m[3] * Px + pj.m[4] * Py + pj.m[5] * Pz + pj.tau[1];
            const float p2 = pj.m[6] * Px + pj.m[7] * Py + pj.m[8] * Pz + pj.tau[2];
            const float den = p2 + 1e-6f;
            const float X = p0 / den, Y = p1 / den;
            const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
            // gradient gates of nan_to_num and clamp (closed interval), camera.py:184-188
            const bool gate_x = (X >= 0.0f) && (X <= wm1);   // false for NaN / +-inf
            const bool gate_y = (Y >= 0.0f) && (Y <= hm1);
            if (gate_x || gate_y) {
              const float ix = fminf(fmaxf(X, 0.0f), wm1), iy = fminf(fmaxf(Y, 0.0f), hm1);
              const float x0f = floorf(ix), y0f = floorf(iy);
              const int x0 = (int)x0f, y0 = (int)y0f;
              const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
              const bool inx = x0 + 1 <= w - 1, iny = y0 + 1 <= h - 1;  // out-of-range taps contribute 0
              const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
              const int o00 = y0 * w + x0, o01 = y0 * w + x1, o10 = y1 * w + x0, o11 = y1 * w + x1;
              float gX = 0.0f, gY = 0.0f;
              const float gs[3] = {g0, g1, g2};
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float* plc = src + c * hw;
                const float v00 = __ldg(plc + o00);
                const float v01 = inx ? __ldg(plc + o01) : 0.0f;
                const float v10 = iny ? __ldg(plc + o10) : 0.0f;
                const float v11 = (inx && iny) ? __ldg(plc + o11) : 0.0f;
                gX += gs[c] * ((v01 - v00) * wy0 + (v11 - v10) * wy1);
                gY += gs[c] * ((v10 - v00) * wx0 + (v11 - v01) * wx1);
              }
              if (!gate_x) gX = 0.0f;
              if (!gate_y) gY = 0.0f;
              const float q = 1.0f / den;
              const float u0 = gX * q, u1 = gY * q;
              // K^T g_p, with the third row formed per pixel in camera-centred coordinates
              const float dx = gate_x ? X - cam.cx : 0.0f, dy = gate_y ? Y - cam.cy : 0.0f;
              const float gc0 = cam.fx * u0;
              const float gc1 = cam.sk * u0 + cam.fy * u1;
              const float gc2 = -(gX * dx + gY * dy) * q;
              acc[0] += gc0 * Px; acc[1] += gc0 * Py; acc[2] += gc0 * Pz; acc[3] += gc0;
              acc[4] += gc1 * Px; acc[5] += gc1 * Py; acc[6] += gc1 * Pz; acc[7] += gc1;
              acc[8] += gc2 * Px; acc[9] += gc2 * Py; acc[10] += gc2 * Pz; acc[11] += gc2;
              // d/d depth: (R^T K^T g_p) . K^-1 [x,y,1]
              const float gP0 = pj.r[0] * gc0 + pj.r[3] * gc1 + pj.r[6] * gc2;
              const float gP1 = pj.r[1] * gc0 + pj.r[4] * gc1 + pj.r[7] * gc2;
              const float gP2 = pj.r[2] * gc0 + pj.r[5] * gc1 + pj.r[8] * gc2;
              const float rx = cam.ki[0] * fxp + cam.ki[1] * fyp + cam.ki[2];
              const float ry = cam.ki[3] * fxp + cam.ki[4] * fyp + cam.ki[5];
              const float rz = cam.ki[6] * fxp + cam.ki[7] * fyp + cam.ki[8];
              gd[it] += gP0 * rx + gP1 * ry + gP2 * rz;
            }
          }
        }
      }
      // CTA reduction of the 12 pose sums of this source -> per-CTA slot
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) sh.red[k][wid] = v;
      }
      __syncthreads();
      if (tid < 12) {
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) v += sh.red[tid][k];
        p.pose_partials[((size_t)blockIdx.x * p.S + j) * 12 + tid] = v;
      }
      // the __syncthreads above also orders phase 4's plane reads before the next phase 1 / 3
    }
  }

  // ------------------------------------------------------------------ smoothness gradient + store
  {
    const float sscale = p.smooth_scale[s];
    const float mbar = p.stats[(s * p.B + b) * 2], Lb = p.stats[(s * p.B + b) * 2 + 1];
    const float inx = 1.0f / ((float)p.B * (float)h * (float)(w - 1));
    const float iny = 1.0f / ((float)p.B * (float)(h - 1) * (float)w);
    const float homog = mbar > 1e-6f ? Lb / ((float)h * (float)w * mbar) : 0.0f;
    float* __restrict__ gout = p.grad_depth[s] + (size_t)b * hw;
#pragma unroll
    for (int it = 0; it < kPosPerThread; ++it) {
      const int i = tid + it * kThreads;
      const int ly = i / kBwdW, lx = i - ly * kBwdW;
      const int gy = tc.y0 + ly, gx = tc.x0 + lx;
      if (i < kBwdW * kBwdH && gy < h && gx < w) {
        float g = gd[it];
        if (sscale > 0.0f) {
          const int pl = (ly + 2) * kHW + lx + 2;
          const float* pi = planes + kBInv * kPlane + pl;
          const float ic = pi[0];
          float el = 0.0f, er = 0.0f, eu = 0.0f, edn = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float* pa = planes + (kBA + c) * kPlane + pl;
            const float a = pa[0];
            el += fabsf(pa[-1] - a); er += fabsf(a - pa[1]);
            eu += fabsf(pa[-kHW] - a); edn += fabsf(a - pa[kHW]);
          }
          float G = 0.0f;
          auto sgn = [](float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); };
          if (gx + 1 < w) G += sgn(ic - pi[1]) * expf(-er * (1.0f / 3.0f)) * inx;
          if (gx >= 1) G -= sgn(pi[-1] - ic) * expf(-el * (1.0f / 3.0f)) * inx;
          if (gy + 1 < h) G += sgn(ic - pi[kHW]) * expf(-edn * (1.0f / 3.0f)) * iny;
          if (gy >= 1) G -= sgn(pi[-kHW] - ic) * expf(-eu * (1.0f / 3.0f)) * iny;
          const float d = __ldg(depth + gy * w + gx);
          const float g_inv = G / mbar - homog;
          if (d >= 1e-6f) g += -ic * ic * g_inv * (g_smooth * sscale);
        }
        gout[gy * w + gx] = g;
      }
    }
  }

  // ------------------------------------------------------------------ last CTA: pose gradients
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.ticket = atomicAdd(p.counter_bwd, 1u);
  __syncthreads();
  if (sh.ticket != gridDim.x - 1) return;
  __threadfence();
  for (int task = wid; task < p.B * p.S; task += kThreads / 32) {  // one warp per (sample, source)
    const int tb = task / p.S, tj = task - tb * p.S;
    double a[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) a[k] = 0.0;
    for (int ss = 0; ss < p.n_scales; ++ss) {
      const int per = p.btiles_x[ss] * p.btiles_y[ss];
      const size_t first = (size_t)p.btile_start[ss] + (size_t)tb * per;
      for (int t = lane; t < per; t += 32) {
        const float4* part = reinterpret_cast<const float4*>(p.pose_partials + ((first + t) * p.S + tj) * 12);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 v = __ldcg(part + k);
          a[4 * k] += (double)v.x; a[4 * k + 1] += (double)v.y; a[4 * k + 2] += (double)v.z; a[4 * k + 3] += (double)v.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) {
      float* gp = p.grad_pose[tj] + tb * 16;
#pragma unroll
      for (int k = 0; k < 12; ++k) gp[k] = (float)a[k];   // rows 0..2 = [dR | dt]
      gp[12] = gp[13] = gp[14] = gp[15] = 0.0f;
    }
  }
  if (tid == 0) *p.counter_bwd = 0u;
}

size_t mono_bwd_smem_bytes() { return (size_t)kBwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_mono_bwd(const MonoParams& p, cudaStream_t stream) {
  // 61.8 KB of dynamic shared memory needs the opt-in attribute (per device; cheap and idempotent)
  cudaError_t e = cudaFuncSetAttribute(mono_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)mono_bwd_smem_bytes());
  if (e != cudaSuccess) return e;
  mono_bwd_kernel<<<p.btile_start[p.n_scales], kThreads, mono_bwd_smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace sde
