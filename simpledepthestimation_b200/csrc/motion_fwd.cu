// Forward of the MotionLearning two-frame loss (sm_100a): statistics pre-pass + fused loss kernel.
//
// motion_stats_kernel   per (direction, sample): sum(occlusion) and sum(depth_error * occlusion)
//                       -> depth_err_2nd_mom (MotionLearning.py:262,276).  The proximity weight of
//                       every pixel depends on this per-sample scalar, so it needs its own pass:
//                       project + 1-channel gather of depth_B, nothing else.
// motion_fwd_kernel     one CTA = one 64x16 tile of one (direction, sample):
//   phase 1  project every pixel of the tile + 1-pixel halo with its own translation, gather
//            rgb + depth of frame B (camera.py:166-202), occlusion, proximity weight; planes
//            S (warped rgb), A, U = weight + 0.01, WZ = weight with zero padding, depth_A.
//            rgb L1 (MotionLearning.py:269-271) and the optional maps are produced here.
//   phase 2  WeightedSSIM (ssim_loss.py:84-111): window sums of U*x, U*y, U*x^2, U*y^2, U*x*y on
//            the reflect-padded products, avg_w with zero padding; two pixels per lane (packed
//            f2), sums carried down the rows in registers; accumulates ssim * avg_w.
// plus smoothness_loss(depth_A, frame_A) from the same planes (1-homogeneity).
// Deterministic: per-CTA partial slots, last CTA adds them in a fixed order in fp64.
// plane layout of the tile kernel in this file: 72-float rows, column xx at index xx + 3, so that index 0 is
// image column tile_x0 - 4 (TMA boxes must start 16-byte aligned)
#define SDE_PITCH 72
#define SDE_COL_OFF 3
#include <type_traits>

#include "motion_device.cuh"

namespace sde {

constexpr int kMA = 0, kMS = 3, kMU = 6, kMW = 7, kMD = 8;
constexpr int kMotionFwdPlanes = 9;

// ------------------------------------------------------------------------------------------------
// Warp mode only: frame B and depth B of every (direction, sample) interleaved per pixel into the scratch planes of
// the `warped` buffer, so that the gather of the statistics / warp pass is one 16-byte load per tap.  Coalesced
// planar reads, coalesced 16-byte stores; 32 bytes of traffic per pixel.
__global__ void __launch_bounds__(kStatThreads) motion_pack_kernel(const __grid_constant__ MotionParams p) {
  const int img = blockIdx.x / p.stat_blocks, chunk = blockIdx.x - img * p.stat_blocks;
  const int dir = img / p.B, b = img - dir * p.B;
  const int hw = p.h * p.w;
  const float* __restrict__ fb = p.frame_b[dir] + (size_t)b * 3 * hw;
  const float* __restrict__ db = p.depth_b[dir] + (size_t)b * hw;
  float4* __restrict__ out = reinterpret_cast<float4*>(p.warped[dir] + ((size_t)b * kMotionSaved + kMotionPacked) * hw);
#pragma unroll 4
  for (int k = 0; k < kStatPixPerThread; ++k) {
    const int pix = chunk * kStatPix + k * kStatThreads + threadIdx.x;
    if (pix < hw) out[pix] = make_float4(__ldg(fb + pix), __ldg(fb + hw + pix), __ldg(fb + 2 * hw + pix), __ldg(db + pix));
  }
}

__global__ void __launch_bounds__(kStatThreads, 3) motion_stats_kernel(const __grid_constant__ MotionParams p) {
  __shared__ MCam s_cam;
  __shared__ float red[2][kStatThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int img = blockIdx.x / p.stat_blocks, chunk = blockIdx.x - img * p.stat_blocks;
  const int dir = img / p.B, b = img - dir * p.B;
  const int hw = p.h * p.w;
  pdl_wait();   // (plain launch unless bit 1 of SDE_PDL_MASK's motion half is set: see launch_motion_fwd)
  if (tid == 0) load_mcam(s_cam, p.K, p.pose[dir], b, p.sx, p.sy);
  __syncthreads();
  const MCam mc = s_cam;
  MotionStage st;
  st.depth_a = p.depth_a[dir] + (size_t)b * hw;
  st.depth_b = p.depth_b[dir] + (size_t)b * hw;
  st.frame_a = nullptr; st.frame_b = nullptr;
  st.field = p.field[dir] ? p.field[dir] + (size_t)b * 3 * hw : nullptr;
  st.planes = nullptr; st.oy = 0; st.ox = 0; st.h = p.h; st.w = p.w; st.hw = hw; st.m2 = 1.0f;
  st.frame_b = p.frame_b[dir] + (size_t)b * 3 * hw;
  // warp mode: this pass also gathers the rgb channels and leaves the planes the loss kernels stage by TMA
  float* __restrict__ wout = p.warped[dir] ? p.warped[dir] + (size_t)b * kMotionSaved * hw : nullptr;
  if (wout) st.packed = reinterpret_cast<const float4*>(wout + (size_t)kMotionPacked * hw);   // motion_pack_kernel ran
  float* __restrict__ occ_out = (wout && p.occ[dir]) ? p.occ[dir] + (size_t)b * hw : nullptr;
  float* __restrict__ crd_out = (wout && p.coords[dir]) ? p.coords[dir] + (size_t)b * hw * 2 : nullptr;
  float socc = 0.0f, serr = 0.0f;
  // depth of A and the residual translation of the NEXT pixel are in flight while the current one is projected,
  // gathered and stored (two dependent round trips per pixel otherwise)
  float nd = 0.0f, nf0 = 0.0f, nf1 = 0.0f, nf2 = 0.0f;
  auto prefetch = [&](int k) {
    const int pix = chunk * kStatPix + k * kStatThreads + tid;
    if (k < kStatPixPerThread && pix < hw) {
      nd = __ldg(st.depth_a + pix);
      if (st.field) { nf0 = __ldg(st.field + pix); nf1 = __ldg(st.field + pix + hw); nf2 = __ldg(st.field + pix + 2 * hw); }
    }
  };
  prefetch(0);
#pragma unroll 1
  for (int k = 0; k < kStatPixPerThread; ++k) {
    const int pix = chunk * kStatPix + k * kStatThreads + tid;
    const float cd = nd, cf0 = nf0, cf1 = nf1, cf2 = nf2;
    prefetch(k + 1);
    if (pix < hw) {
      const int gy = pix / p.w, gx = pix - gy * p.w;
      MotionSample sm;
      motion_sample_pre(st, mc, gy, gx, pix, cd, cf0, cf1, cf2, wout != nullptr, sm, false, wout != nullptr);
      const float e = sm.Zc - sm.Sd;
      const float derr = e * e;
      socc += sm.occ;
      serr += derr * sm.occ;
      if (wout) {
        wout[pix] = sm.S[0]; wout[hw + pix] = sm.S[1]; wout[2 * hw + pix] = sm.S[2];
        wout[3 * hw + pix] = derr;
        wout[4 * hw + pix] = sm.valid + 2.0f * sm.occ;
#pragma unroll
        for (int c = 0; c < 3; ++c) {   // derivative planes for the backward pass
          wout[(5 + c) * hw + pix] = sm.dSx[c];
          wout[(8 + c) * hw + pix] = sm.dSy[c];
        }
        if (occ_out) occ_out[pix] = sm.occ;
        if (crd_out) {
          float2 cn;
          cn.x = __fdiv_rn(2.0f * sm.Xs, (float)(p.w - 1)) - 1.0f;
          cn.y = __fdiv_rn(2.0f * sm.Ys, (float)(p.h - 1)) - 1.0f;
          *reinterpret_cast<float2*>(crd_out + (size_t)pix * 2) = cn;
        }
      }
    }
  }
  pdl_launch_dependents();   // the heavy part of this block is done (mono_fwd.cu)
  socc = warp_sum(socc); serr = warp_sum(serr);
  if (lane == 0) { red[0][wid] = socc; red[1][wid] = serr; }
  __syncthreads();
  if (tid < 2) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kStatThreads / 32; ++k) v += red[tid][k];
    p.stat_partials[(size_t)blockIdx.x * 2 + tid] = v;
    publish_fence();   // only the two threads that publish the slot fence (the planes are consumed by later kernels)
  }
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(p.counters, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  publish_fence();
  for (int q = wid; q < p.n_dirs * p.B; q += kStatThreads / 32) {   // one warp per (direction, sample)
    double a0 = 0.0, a1 = 0.0;
    for (int t = lane; t < p.stat_blocks; t += 32) {
      const float2 v = __ldcg(reinterpret_cast<const float2*>(p.stat_partials) + (size_t)q * p.stat_blocks + t);
      a0 += (double)v.x; a1 += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (lane == 0) {
      // depth_err_2nd_mom = sum(err * occ) / (sum(occ) + 1) + 1e-4   (MotionLearning.py:262,276)
      p.stats[q * 4 + 0] = (float)(a1 / (a0 + 1.0) + 1e-4);
      p.stats[q * 4 + 1] = (float)a0;
    }
  }
  if (tid == 0) p.counters[0] = 0u;
}

// ------------------------------------------------------------------------------------------------
struct MotionFwdShared {
  MCam cam;
  float red[5][kThreads / 32];
  double dred[5][kThreads / 32];
  unsigned ticket;
  __align__(8) uint64_t bar;   // TMA completion barrier
};

__global__ void __launch_bounds__(kThreads, 4) motion_fwd_kernel(const __grid_constant__ MotionParams p,
                                                                 const __grid_constant__ MotionTma maps) {
  extern __shared__ __align__(128) float planes[];  // [kMotionFwdPlanes][kPlane]
  __shared__ MotionFwdShared sh;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int dir, b, tx0, ty0;
  // tiles in reverse list order: the planes the statistics / warp pass wrote last are still in L2 (cf. mono_fwd.cu)
  const int vbid = (int)(gridDim.x - 1 - blockIdx.x);
  decode_motion_tile(vbid, p.tiles_per_dir, p.tiles_x, p.tiles_y, kTileW, kTileH, dir, b, tx0, ty0);
  const int h = p.h, w = p.w, hw = h * w;
  const bool tma = p.tma != 0;
  if (tma && tid == 0) {
    // warp mode: the statistics pass left warped rgb, depth error and valid/occlusion planes; the copy engine
    // brings them in together with frame A and depth A (U and WZ are computed in place of the last two)
    mbar_init(&sh.bar, 1);
    mbar_init_fence();
  }
  pdl_wait();   // the statistics pass's planes and per-sample moments (and every other tensor) are complete from here on
  if (tma && tid == 0) {
    mbar_arrive_expect_tx(&sh.bar, 9 * kPlaneBytesTma);
    const int bx = tx0 - 1 - kColOff, by = ty0 - 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tma_load_plane(planes + (kMA + c) * kPlane, &maps.frame_a[dir], &sh.bar, bx, by, b * 3 + c);
      tma_load_plane(planes + (kMS + c) * kPlane, &maps.warped[dir], &sh.bar, bx, by, b * kMotionSaved + c);
    }
    tma_load_plane(planes + kMU * kPlane, &maps.warped[dir], &sh.bar, bx, by, b * kMotionSaved + 3);
    tma_load_plane(planes + kMW * kPlane, &maps.warped[dir], &sh.bar, bx, by, b * kMotionSaved + 4);
    tma_load_plane(planes + kMD * kPlane, &maps.depth_a[dir], &sh.bar, bx, by, b);
  }
  if (tid == 0) load_mcam(sh.cam, p.K, p.pose[dir], b, p.sx, p.sy);
  __syncthreads();

  MotionStage st;
  st.depth_a = pinned(p.depth_a[dir] + (size_t)b * hw);
  st.depth_b = pinned(p.depth_b[dir] + (size_t)b * hw);
  st.frame_a = pinned(p.frame_a[dir] + (size_t)b * 3 * hw);
  st.frame_b = pinned(p.frame_b[dir] + (size_t)b * 3 * hw);
  st.field = p.field[dir] ? pinned(p.field[dir] + (size_t)b * 3 * hw) : nullptr;
  st.planes = planes; st.oy = ty0 - 1; st.ox = tx0 - 1; st.h = h; st.w = w; st.hw = hw;
  st.m2 = __ldg(p.stats + (dir * p.B + b) * 4);

  float* __restrict__ occ_out = p.occ[dir] ? p.occ[dir] + (size_t)b * hw : nullptr;
  float* __restrict__ wgt_out = p.weight[dir] ? p.weight[dir] + (size_t)b * hw : nullptr;
  float* __restrict__ crd_out = p.coords[dir] ? p.coords[dir] + (size_t)b * hw * 2 : nullptr;

  // ------------------------------------------------------------------ phase 1
  float l1 = 0.0f;
  if (tma) {
    mbar_wait(&sh.bar, 0);
    const bool interior = tx0 >= 1 && ty0 >= 1 && tx0 + kTileW + 1 <= w && ty0 + kTileH + 1 <= h;
    if (!interior) {
      reflect_fixup(planes, kMA, 6, st.oy, st.ox, h, w, tid);   // A and S (in-image sources only)
      reflect_fixup(planes, kMD, 1, st.oy, st.ox, h, w, tid);
      __syncthreads();
    }
    // depth error, valid/occlusion -> U = weight + 0.01 and WZ = weight (zero outside the image: valid = 0 there)
    // two stored positions (an aligned pair of the plane row, pad columns included) per iteration
    constexpr int kPairs = kPitch / 2;
    for (int i = tid; i < kHH * kPairs; i += kThreads) {
      const int yy = i / kPairs, j = i - yy * kPairs;
      const int pl = yy * kPitch + 2 * j;
      const f2 derr = ld2(planes + kMU * kPlane + pl), vo = ld2(planes + kMW * kPlane + pl);
      const float wg[2] = {proximity_weight(lo(derr), lo(vo), st.m2), proximity_weight(hi(derr), hi(vo), st.m2)};
      const float vv[2] = {lo(vo), hi(vo)};
      *reinterpret_cast<unsigned long long*>(planes + kMU * kPlane + pl) = mk2(wg[0] + 1e-2f, wg[1] + 1e-2f).v;
      *reinterpret_cast<unsigned long long*>(planes + kMW * kPlane + pl) = mk2(wg[0], wg[1]).v;
      const int ty = st.oy + yy;
      if (ty >= 0 && ty < h && yy >= 1 && yy <= kTileH) {
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int xx = 2 * j + e2 - kColOff, tx = st.ox + xx;
          if (tx >= 0 && tx < w && xx >= 1 && xx <= kTileW) {
            const float occ = vv[e2] >= 2.0f ? 1.0f : 0.0f;
            float e = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf(planes[(kMS + c) * kPlane + pl + e2] - planes[(kMA + c) * kPlane + pl + e2]);
            l1 += e * occ;
            if (wgt_out) wgt_out[ty * w + tx] = wg[e2];
          }
        }
      }
    }
    if (!interior) {
      __syncthreads();
      reflect_fixup(planes, kMU, 1, st.oy, st.ox, h, w, tid);   // the reflect-padded product x * (w + 0.01)
    }
  } else {
    const MCam mc = sh.cam;
#pragma unroll 1
    for (int i = tid; i < kPositions; i += kThreads) {
      int yy, xx;
      position_of(i, yy, xx);
      const int ty = st.oy + yy, tx = st.ox + xx;
      const bool inside = ty >= 0 && ty < h && tx >= 0 && tx < w;
      const int gy = reflect_clamp(ty, h), gx = reflect_clamp(tx, w);
      const int pix = gy * w + gx;
      MotionSample sm;
      motion_sample(st, mc, gy, gx, pix, true, sm);
      const int pl = plane_index(yy, xx);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        planes[(kMS + c) * kPlane + pl] = sm.S[c];
        planes[(kMA + c) * kPlane + pl] = sm.A[c];
      }
      planes[kMU * kPlane + pl] = sm.wgt + 1e-2f;            // w = w + 1e-2 (ssim_loss.py:89)
      planes[kMW * kPlane + pl] = inside ? sm.wgt : 0.0f;    // avg_pool2d(w, padding=1) pads with zeros
      planes[kMD * kPlane + pl] = sm.d;
      const bool owned = inside && yy >= 1 && yy <= kTileH && xx >= 1 && xx <= kTileW;
      if (owned) {
        // rgb_l1_loss = mean(|S - A| * occlusion)   (MotionLearning.py:269-271)
        l1 += (fabsf(sm.S[0] - sm.A[0]) + fabsf(sm.S[1] - sm.A[1]) + fabsf(sm.S[2] - sm.A[2])) * sm.occ;
        if (occ_out) occ_out[pix] = sm.occ;
        if (wgt_out) wgt_out[pix] = sm.wgt;
        if (crd_out) {
          // coords = 2 * clamp(X) / (w - 1) - 1   (camera.py:190-193)
          float2 cn;
          cn.x = __fdiv_rn(2.0f * sm.Xs, (float)(w - 1)) - 1.0f;
          cn.y = __fdiv_rn(2.0f * sm.Ys, (float)(h - 1)) - 1.0f;
          *reinterpret_cast<float2*>(crd_out + (size_t)pix * 2) = cn;
        }
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase 2: WeightedSSIM
  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;
  float ssim_sum = 0.0f;
  if (p.ssim_w > 0.0f) {
    const f2 C1 = bc2(p.c1), C2 = bc2(p.c2);
    const f2 ninth = bc2(1.0f / 9.0f);
    f2 avgw[kRowsPerWarp], q9[kRowsPerWarp], acc[kRowsPerWarp];
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) acc[o] = bc2(0.0f);
    {
      // avg_w = avg_pool2d(w, 3, 1, padding=1); inverse_avg_w = 1 / (avg_w + 1e-2)   (ssim_loss.py:88-90)
      const float* pw = planes + kMW * kPlane + plane_index(r0, c0);
      f2 hW[2];
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
        const Row4 wz = ld_row(pw + rr * kPitch);
        const f2 nW = (wz.c + swp(wz.c)) + wz.o;
        if (rr >= 2) {
          const int o = rr - 2;
          avgw[o] = ((hW[0] + hW[1]) + nW) * ninth;
          const f2 den = avgw[o] + bc2(1e-2f);
          q9[o] = div2(ninth, den);   // (1/9) * inverse_avg_w: turns window sums into weighted means
        }
        hW[0] = hW[1]; hW[1] = nW;
      }
    }
    // instantiated per C1/C2 mode (ssim_loss.py:97-105), as in the backward kernel: a run-time switch inside the unrolled
    // rows would put every row's formula into its own basic block and keep the scheduler from interleaving the rows
    auto channels = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      const float* pa = planes + (kMA + c) * kPlane + plane_index(r0, c0);
      const float* px = planes + (kMS + c) * kPlane + plane_index(r0, c0);
      const float* pu = planes + kMU * kPlane + plane_index(r0, c0);
      f2 hX[2], hA[2], hXX[2], hAA[2], hXA[2];
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
        const Row4 u = ld_row(pu + rr * kPitch), x = ld_row(px + rr * kPitch), a = ld_row(pa + rr * kPitch);
        const f2 uxc = u.c * x.c, uxo = u.o * x.o, uac = u.c * a.c, uao = u.o * a.o;
        const f2 xxc = uxc * x.c, aac = uac * a.c, xac = uxc * a.c;
        const f2 nX = (uxc + swp(uxc)) + uxo;
        const f2 nA = (uac + swp(uac)) + uao;
        const f2 nXX = fma2(uxo, x.o, xxc + swp(xxc));
        const f2 nAA = fma2(uao, a.o, aac + swp(aac));
        const f2 nXA = fma2(uxo, a.o, xac + swp(xac));
        if (rr >= 2) {
          const int o = rr - 2;
          const f2 k = q9[o];
          const f2 mx = ((hX[0] + hX[1]) + nX) * k, my = ((hA[0] + hA[1]) + nA) * k;
          const f2 exx = ((hXX[0] + hXX[1]) + nXX) * k, eaa = ((hAA[0] + hAA[1]) + nAA) * k;
          const f2 exa = ((hXA[0] + hXA[1]) + nXA) * k;
          const f2 sx = fma2(mx * bc2(-1.0f), mx, exx), sy = fma2(my * bc2(-1.0f), my, eaa);
          const f2 sxy = fma2(mx * bc2(-1.0f), my, exa);
          f2 n, d;
          if (MODE == 1) {            // C1 == inf
            n = fma2(bc2(2.0f), sxy, C2);
            d = (sx + sy) + C2;
          } else if (MODE == 2) {     // C2 == inf
            n = fma2(bc2(2.0f), mx * my, C1);
            d = fma2(mx, mx, my * my) + C1;
          } else {
            n = fma2(bc2(2.0f), sxy, C2) * fma2(bc2(2.0f), mx * my, C1);
            d = ((sx + sy) + C2) * (fma2(mx, mx, my * my) + C1);
          }
          const f2 nssim = ndiv2(n, d);   // -ssim
          const f2 l = mk2(__saturatef(fmaf(lo(nssim), 0.5f, 0.5f)), __saturatef(fmaf(hi(nssim), 0.5f, 0.5f)));
          acc[o] = fma2(l, avgw[o], acc[o]);
        }
        hX[0] = hX[1]; hX[1] = nX; hA[0] = hA[1]; hA[1] = nA;
        hXX[0] = hXX[1]; hXX[1] = nXX; hAA[0] = hAA[1]; hAA[1] = nAA; hXA[0] = hXA[1]; hXA[1] = nXA;
      }
    }
    };
    if (p.mode == 1) channels(std::integral_constant<int, 1>{});
    else if (p.mode == 2) channels(std::integral_constant<int, 2>{});
    else channels(std::integral_constant<int, 0>{});
    const int gx0 = tx0 + c0;
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) {
      const int gy = ty0 + r0 + o;
      if (gy < h) {
        if (gx0 < w) ssim_sum += lo(acc[o]);
        if (gx0 + 1 < w) ssim_sum += hi(acc[o]);
      }
    }
  }

  pdl_launch_dependents();
  // ------------------------------------------------------------------ smoothness(depth_A, frame_A)
  float smx = 0.0f, smy = 0.0f, sinv = 0.0f;
  // warp mode: the local smoothness gradient goes to plane 11 of `warped` for the backward pass
  tile_smoothness(planes, kMD, kMA, r0, c0, tx0, ty0, h, w, p.B,
                  (tma && p.warped[dir]) ? p.warped[dir] + ((size_t)b * kMotionSaved + 11) * hw : nullptr, smx, smy, sinv);

  // ------------------------------------------------------------------ CTA reduction -> partial slot
  l1 = warp_sum(l1); ssim_sum = warp_sum(ssim_sum); smx = warp_sum(smx); smy = warp_sum(smy); sinv = warp_sum(sinv);
  if (lane == 0) {
    sh.red[0][wid] = l1; sh.red[1][wid] = ssim_sum; sh.red[2][wid] = smx; sh.red[3][wid] = smy; sh.red[4][wid] = sinv;
  }
  __syncthreads();
  if (tid < 5) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) v += sh.red[tid][k];
    p.partials[(size_t)vbid * 8 + tid] = v;
    publish_fence();   // only the publishing threads fence
  }
  __syncthreads();
  // ------------------------------------------------------------------ two-level fixed-order reduction
  // The last tile of a (direction, sample) image adds that image's partial slots while other images are still being
  // computed; the last image to finish adds the per-image results.  (One last CTA adding all 19 200 slots of cfg4
  // was a serial tail of ~60 us.)  Same order every run.
  const int per_img = p.tiles_x * p.tiles_y, img = dir * p.B + b, n_img = p.n_dirs * p.B;
  if (tid == 0) sh.ticket = atomicAdd(p.img_counter_f + img, 1u);
  __syncthreads();
  if (sh.ticket != (unsigned)(per_img - 1)) return;
  publish_fence();
  {
    const float* part = p.partials + ((size_t)dir * p.tiles_per_dir + (size_t)b * per_img) * 8;
    double a[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int t = tid; t < per_img; t += kThreads) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part + (size_t)t * 8));
      const float v4 = __ldcg(part + (size_t)t * 8 + 4);
      a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w; a[4] += (double)v4;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) sh.dred[k][wid] = a[k];
    }
    __syncthreads();
    if (tid == 0) {
      double t5[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) t5[k] = ((sh.dred[k][0] + sh.dred[k][1]) + sh.dred[k][2]) + sh.dred[k][3];
      const double qh = h, qw = w, nB = p.B;
      const double mbar = fmax(t5[4] / (qh * qw), 1e-6);
      const double Lb = (t5[2] / (nB * qh * (qw - 1.0)) + t5[3] / (nB * (qh - 1.0) * qw)) / mbar;
      p.stats[img * 4 + 2] = (float)mbar;
      p.stats[img * 4 + 3] = (float)Lb;
      p.fin[img * 4 + 0] = t5[0]; p.fin[img * 4 + 1] = t5[1]; p.fin[img * 4 + 2] = Lb;
      p.img_counter_f[img] = 0u;   // leave the workspace zeroed for the next call
      publish_fence();
      sh.ticket = atomicAdd(p.counters + 1, 1u);
    }
  }
  __syncthreads();
  if (sh.ticket != (unsigned)(n_img - 1)) return;
  publish_fence();
  if (tid < p.n_dirs) {   // one thread per direction; samples in order
    const int qd = tid;
    double L1 = 0.0, LS = 0.0, LSm = 0.0;
    for (int qb = 0; qb < p.B; ++qb) {
      const double* f = p.fin + (size_t)(qd * p.B + qb) * 4;
      L1 += __ldcg(f); LS += __ldcg(f + 1); LSm += __ldcg(f + 2);
    }
    const double n = (double)p.B * 3.0 * h * w;
    p.losses[qd * 4 + 0] = (float)(L1 / n);
    p.losses[qd * 4 + 1] = (float)(LS / n * (double)p.ssim_w * 0.5);   // MotionLearning.py:286-289
    p.losses[qd * 4 + 2] = (float)LSm;
    p.losses[qd * 4 + 3] = 0.0f;
  }
  if (tid == 0) p.counters[1] = 0u;
}

size_t motion_fwd_smem_bytes() { return (size_t)kMotionFwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_motion_fwd(const MotionParams& p, const MotionTma& t, cudaStream_t stream) {
  if (p.warped[0]) {   // warp mode (all directions or none, motion_tma): interleave frame B + depth B for the gather
    motion_pack_kernel<<<p.n_dirs * p.B * p.stat_blocks, kStatThreads, 0, stream>>>(p);
    cudaError_t e0 = cudaGetLastError();
    if (e0 != cudaSuccess) return e0;
  }
  // the statistics pass is a plain launch (a gather kernel whose blocks start in lock-step behind a chained launch loses more
  // than the launch latency, cf. mono_warp.cu); the loss kernel is chained behind it (bit 1 of the launch mask)
  motion_stats_kernel<<<p.n_dirs * p.B * p.stat_blocks, kStatThreads, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(motion_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)motion_fwd_smem_bytes());
  if (e != cudaSuccess) return e;
  return launch_chained(1, motion_fwd_kernel, (unsigned)(p.n_dirs * p.tiles_per_dir), kThreads, motion_fwd_smem_bytes(), stream, p, t);
}

}  // namespace sde
