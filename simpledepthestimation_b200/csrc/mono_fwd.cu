// Fused forward of the MonoDepth2 self-supervised loss, all scales in one launch (sm_100a).
//
// One CTA = one 64x16 tile of one (scale, sample).  Per source frame j:
//   phase 1  project every pixel of the tile + 1-pixel halo through K^-1, (R_j, t_j), K
//            (camera.py:125-163,172-178), clamp (camera.py:184-188), 4-tap bilinear gather from
//            the source (F.grid_sample, camera.py:196) -> shared-memory planes S (warped) and
//            I (unwarped source = identity/automask candidate, MonoDepth2.py:96-101); the target
//            A and 1/depth are staged on the first source.  Halo positions outside the image take
//            the reflected pixel (nn.ReflectionPad2d(1), ssim_loss.py:32,35).
//   phase 2  each lane owns two adjacent columns (one packed f2), each warp four rows; 3x3 window
//            sums of X, X^2, X*A are built from horizontal 3-sums carried down the rows in
//            registers, then SSIM (ssim_loss.py:36-53) and the 0.85/0.15 mix (MonoDepth2.py:137-144).
//            The running min / argmin over candidates stays in registers across sources
//            (MonoDepth2.py:116-119; first index wins ties).
// Edge-aware smoothness (smoothness_loss.py:62-80) is evaluated from the same staged planes using
// the 1-homogeneity of the loss in the inverse depth: the tile accumulates sum(1/d) and the
// un-normalised |d(1/d)|*exp(-|dI|) sums, the division by the per-image mean happens in the final
// reduction.  Partial sums go to a per-CTA slot; the last CTA to finish adds them in a fixed order
// (deterministic, no float atomics) and writes rec_loss / smooth_loss.
#include "mono_device.cuh"

namespace sde {

constexpr int kFwdPlanes = 10;      // A[3], S[3], I[3], 1/d
constexpr int kPlA = 0, kPlS = 3, kPlI = 6, kPlInv = 9;

struct FwdShared {
  Cam cam;
  Proj proj[SDE_MAX_SOURCES];
  float red[4][kThreads / 32];
  unsigned ticket;
};

__global__ void __launch_bounds__(kThreads, 4) mono_fwd_kernel(const __grid_constant__ MonoParams p) {
  extern __shared__ __align__(16) float planes[];  // [kFwdPlanes][kPlane]
  __shared__ FwdShared sh;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const TileCoord tc = decode_tile(p, blockIdx.x);
  const int s = tc.s, b = tc.b, h = p.h[s], w = p.w[s], hw = h * w;
  const bool automask = (p.flags & SDE_MONO_AUTOMASK) != 0;
  const bool reduce_mean = (p.flags & SDE_MONO_REDUCE_MEAN) != 0;

  if (tid < p.S) {
    Cam cam;
    float k[9];
    load_cam(cam, k, p.K, b, p.sx[s], p.sy[s]);
    if (tid == 0) sh.cam = cam;
    load_proj(sh.proj[tid], k, p.pose[tid], b);
  }
  __syncthreads();

  const float* __restrict__ depth = p.depth[s] + (size_t)b * hw;
  const float* __restrict__ target = p.target[s] + (size_t)b * 3 * hw;

  // phase-2 ownership: rows r0..r0+3 of the tile, columns 2*lane, 2*lane+1
  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;  // plane column of the left halo of this lane's pair
  const f2 wS = bc2(p.ssim_w * (1.0f / 3.0f)), wL = bc2(p.l1_w * (1.0f / 3.0f));
  const f2 C1 = bc2(81.0f * p.c1), C2 = bc2(81.0f * p.c2);
  const bool use_ssim = p.ssim_w > 0.0f;

  f2 best[kRowsPerWarp];
  int arg[kRowsPerWarp][2];
#pragma unroll
  for (int o = 0; o < kRowsPerWarp; ++o) {
    best[o] = reduce_mean ? bc2(0.0f) : bc2(__int_as_float(0x7f800000));
    arg[o][0] = arg[o][1] = 0;
  }

  for (int j = 0; j < p.S; ++j) {
    // ---------------------------------------------------------------- phase 1
    {
      const Cam cam = sh.cam;
      const Proj pj = sh.proj[j];
      const float* __restrict__ src = p.source[s][j] + (size_t)b * 3 * hw;
      for (int i = tid; i < kPlane; i += kThreads) {
        const int yy = i / kHW, xx = i - yy * kHW;
        const int gy = reflect_clamp(tc.y0 - 1 + yy, h), gx = reflect_clamp(tc.x0 - 1 + xx, w);
        const int pix = gy * w + gx;
        const float d = __ldg(depth + pix);
        float X, Y, sv[3];
        project_px(cam, pj, (float)gx, (float)gy, d, X, Y);
        bilinear3(src, hw, w, h, X, Y, sv);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          planes[(kPlS + c) * kPlane + i] = sv[c];
          planes[(kPlI + c) * kPlane + i] = __ldg(src + c * hw + pix);
        }
        if (j == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) planes[(kPlA + c) * kPlane + i] = __ldg(target + c * hw + pix);
          planes[kPlInv * kPlane + i] = 1.0f / (d < 1e-6f ? 1e-6f : d);  // clamp(min=1e-6) keeps NaN, as torch.clamp does
        }
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase 2
    f2 acc[2][kRowsPerWarp];
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) acc[0][o] = acc[1][o] = bc2(0.0f);
    const int ncand = automask ? 2 : 1;

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      const float* pa = planes + (kPlA + c) * kPlane + r0 * kHW + c0;
      f2 hA[2], hAA[2];                 // horizontal 3-sums of the two previous rows
      f2 hX[2][2], hXX[2][2], hXA[2][2];
      f2 l1p[2];                        // |X - A| of the previous row's centre pair
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
        const f2 alo = ld2(pa + rr * kHW), ahi = ld2(pa + rr * kHW + 2);
        const f2 aC = mk2(alo.y, ahi.x), aO = mk2(alo.x, ahi.y);
        const f2 aa = aC * aC;
        const f2 nA = aC + swp(aC) + aO;
        const f2 nAA = fma2(aO, aO, aa + swp(aa));
        f2 nX[2], nXX[2], nXA[2], l1n[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < ncand) {
            const float* px = planes + ((k == 0 ? kPlS : kPlI) + c) * kPlane + (r0 + rr) * kHW + c0;
            const f2 xlo = ld2(px), xhi = ld2(px + 2);
            const f2 xC = mk2(xlo.y, xhi.x), xO = mk2(xlo.x, xhi.y);
            const f2 xx = xC * xC, xa = xC * aC;
            nX[k] = xC + swp(xC) + xO;
            nXX[k] = fma2(xO, xO, xx + swp(xx));
            nXA[k] = fma2(xO, aO, xa + swp(xa));
            l1n[k] = abs2(xC - aC);
          }
        }
        if (rr >= 2) {
          // window sums (not divided by 9): every SSIM factor below is the reference's times 81,
          // which cancels in the ratio and avoids a rounded 1/9 constant
          const int o = rr - 2;
          const f2 sA = hA[0] + hA[1] + nA;
          const f2 sAA = hAA[0] + hAA[1] + nAA;
          const f2 aa2 = sA * sA;
          const f2 vA = fma2(bc2(9.0f), sAA, neg2(aa2));     // 81 sigma_y
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < ncand) {
              f2 a = acc[k][o];
              if (use_ssim) {
                const f2 sX = hX[k][0] + hX[k][1] + nX[k];
                const f2 sXX = hXX[k][0] + hXX[k][1] + nXX[k];
                const f2 sXA = hXA[k][0] + hXA[k][1] + nXA[k];
                // numerator and denominator are formed with mirrored operation orders so that
                // X == A gives ssim == 1 exactly (identity candidate of identical frames is 0)
                const f2 t = sX * sA, xx2 = sX * sX;
                const f2 n1 = fma2(bc2(2.0f), t, C1);
                const f2 n2 = fma2(bc2(2.0f), fma2(bc2(9.0f), sXA, neg2(t)), C2);
                const f2 d1 = (xx2 + aa2) + C1;
                const f2 d2 = (fma2(bc2(9.0f), sXX, neg2(xx2)) + vA) + C2;
                const f2 ssim = div2(n1 * n2, d1 * d2);
                a = fma2(sat2(fma2(ssim, bc2(-0.5f), bc2(0.5f))), wS, a);
              }
              acc[k][o] = fma2(l1p[k], wL, a);
            }
          }
        }
        hA[0] = hA[1]; hA[1] = nA;
        hAA[0] = hAA[1]; hAA[1] = nAA;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < ncand) {
            hX[k][0] = hX[k][1]; hX[k][1] = nX[k];
            hXX[k][0] = hXX[k][1]; hXX[k][1] = nXX[k];
            hXA[k][0] = hXA[k][1]; hXA[k][1] = nXA[k];
            l1p[k] = l1n[k];
          }
        }
      }
    }
    // candidates of this source: 2j (warp), 2j+1 (identity) with automask, else j
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k < ncand) {
        const int idx = automask ? 2 * j + k : j;
#pragma unroll
        for (int o = 0; o < kRowsPerWarp; ++o) {
          if (reduce_mean) {
            best[o] = best[o] + acc[k][o];
          } else {
            // strict '<' keeps the first index on ties; a NaN candidate sticks (torch.min propagates NaN)
            const f2 v = acc[k][o];
            if (v.x < best[o].x || v.x != v.x) { best[o].x = v.x; arg[o][0] = idx; }
            if (v.y < best[o].y || v.y != v.y) { best[o].y = v.y; arg[o][1] = idx; }
          }
        }
      }
    }
    if (j + 1 < p.S) __syncthreads();  // S/I planes are overwritten by the next source
  }

  // ------------------------------------------------------------------ per-thread sums, argmin, smoothness
  float rec = 0.0f, smx = 0.0f, smy = 0.0f, sinv = 0.0f;
  const int gx0 = tc.x0 + c0;  // image column of this lane's first pixel
  uint8_t* __restrict__ amap = p.argmin[s] ? p.argmin[s] + (size_t)b * hw : nullptr;
#pragma unroll
  for (int o = 0; o < kRowsPerWarp; ++o) {
    const int gy = tc.y0 + r0 + o;
    if (gy < h) {
      if (gx0 < w) { rec += best[o].x; if (amap) amap[gy * w + gx0] = (uint8_t)arg[o][0]; }
      if (gx0 + 1 < w) { rec += best[o].y; if (amap) amap[gy * w + gx0 + 1] = (uint8_t)arg[o][1]; }
    }
  }
  if (p.smooth_scale[s] > 0.0f) {
    // plane coordinates of pixel (row o, first column): (r0 + o + 1, c0 + 1)
    const float* pinv = planes + kPlInv * kPlane + (r0 + 1) * kHW + c0;
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) {
      const int gy = tc.y0 + r0 + o;
      // columns c0+1, c0+2 (own pair), c0+3 (right neighbour of the second pixel)
      const f2 lo = ld2(pinv + o * kHW), hi = ld2(pinv + o * kHW + 2);
      const f2 dn = ld2(pinv + (o + 1) * kHW + 2);
      const float i0 = lo.y, i1 = hi.x, i2 = hi.y;
      const float below0 = pinv[(o + 1) * kHW + 1], below1 = dn.x;
      float ex0 = 0.0f, ex1 = 0.0f, ey0 = 0.0f, ey1 = 0.0f;  // mean_c |dI|
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* pa = planes + (kPlA + c) * kPlane + (r0 + 1 + o) * kHW + c0;
        const f2 alo = ld2(pa), ahi = ld2(pa + 2);
        const f2 adn = ld2(pa + kHW + 2);
        const float a0 = alo.y, a1 = ahi.x, a2 = ahi.y;
        ex0 += fabsf(a0 - a1);
        ex1 += fabsf(a1 - a2);
        ey0 += fabsf(a0 - pa[kHW + 1]);
        ey1 += fabsf(a1 - adn.x);
      }
      if (gy < h) {
        const bool v0 = gx0 < w, v1 = gx0 + 1 < w, v2 = gx0 + 2 < w, vy = gy + 1 < h;
        if (v0) sinv += i0;
        if (v1) sinv += i1;
        if (v1) smx += fabsf(i0 - i1) * expf(-ex0 * (1.0f / 3.0f));
        if (v2) smx += fabsf(i1 - i2) * expf(-ex1 * (1.0f / 3.0f));
        if (v0 && vy) smy += fabsf(i0 - below0) * expf(-ey0 * (1.0f / 3.0f));
        if (v1 && vy) smy += fabsf(i1 - below1) * expf(-ey1 * (1.0f / 3.0f));
      }
    }
  }

  // ------------------------------------------------------------------ CTA reduction -> partial slot
  rec = warp_sum(rec); smx = warp_sum(smx); smy = warp_sum(smy); sinv = warp_sum(sinv);
  if (lane == 0) { sh.red[0][wid] = rec; sh.red[1][wid] = smx; sh.red[2][wid] = smy; sh.red[3][wid] = sinv; }
  __syncthreads();
  if (tid < 4) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) v += sh.red[tid][k];
    p.partials[(size_t)blockIdx.x * 4 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sh.ticket = atomicAdd(p.counter, 1u);
  __syncthreads();
  if (sh.ticket != gridDim.x - 1) return;

  // ------------------------------------------------------------------ last CTA: fixed-order final reduction
  __threadfence();
  double* fin = p.fin;  // [n_scales*B][2] staging
  const int pairs = p.n_scales * p.B;
  for (int q = wid; q < pairs; q += kThreads / 32) {   // one warp per (scale, image)
    const int qs = q / p.B, qb = q - qs * p.B;
    const int per = p.tiles_x[qs] * p.tiles_y[qs];
    const float* part = p.partials + ((size_t)p.tile_start[qs] + (size_t)qb * per) * 4;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = lane; t < per; t += 32) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + t);
      a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) {
      const double qh = p.h[qs], qw = p.w[qs], nB = p.B;
      const int ncand = (automask ? 2 : 1) * p.S;
      double rec_q = a[0] / (nB * qh * qw) / p.n_scales;
      if (reduce_mean) rec_q /= ncand;
      const double mbar = fmax(a[3] / (qh * qw), 1e-6);
      const double Lb = (a[1] / (nB * qh * (qw - 1.0)) + a[2] / (nB * (qh - 1.0) * qw)) / mbar;
      p.stats[q * 2 + 0] = (float)mbar;
      p.stats[q * 2 + 1] = (float)Lb;
      fin[q * 2 + 0] = rec_q;
      fin[q * 2 + 1] = Lb * (double)p.smooth_scale[qs];
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    double r = 0.0, sm = 0.0;
    for (int q = 0; q < pairs; ++q) { r += __ldcg(fin + q * 2); sm += __ldcg(fin + q * 2 + 1); }
    p.losses[0] = (float)r;
    p.losses[1] = (float)sm;
    *p.counter = 0u;  // leave the workspace zeroed for the next call
  }
}

size_t mono_fwd_smem_bytes() { return (size_t)kFwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_mono_fwd(const MonoParams& p, cudaStream_t stream) {
  // 47.5 KB of dynamic shared memory: below the 48 KB default limit, no opt-in attribute needed
  mono_fwd_kernel<<<p.tile_start[p.n_scales], kThreads, mono_fwd_smem_bytes(), stream>>>(p);
  return cudaGetLastError();
}

}  // namespace sde
