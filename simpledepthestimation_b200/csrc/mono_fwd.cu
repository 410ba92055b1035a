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
// plane layout of this kernel: 72-float rows, column xx at index xx + 3, so that index 0 is image column
// tile_x0 - 4 (TMA boxes must start 16-byte aligned)
#define SDE_PITCH 72
#define SDE_COL_OFF 3
#include "mono_device.cuh"

#ifndef SDE_FWD_OCC
#define SDE_FWD_OCC 4
#endif
#ifndef SDE_FWD_CARVEOUT
#define SDE_FWD_CARVEOUT -1
#endif

namespace sde {

constexpr int kFwdPlanes = 10;      // A[3], S[3], I[3], depth
constexpr int kPlA = 0, kPlS = 3, kPlI = 6, kPlD = 9;
#ifndef SDE_NB
#define SDE_NB 2
#endif

struct FwdShared {
  Cam cam;
  Proj proj[SDE_MAX_SOURCES];
  float red[4][kThreads / 32];
  double dred[4][kThreads / 32];
  unsigned ticket;
  __align__(8) uint64_t bar;   // TMA completion barrier
};

template <bool AUTOMASK>
__global__ void __launch_bounds__(kThreads, SDE_FWD_OCC) mono_fwd_kernel(const __grid_constant__ MonoParams p,
                                                                       const __grid_constant__ MonoTma maps) {
  extern __shared__ __align__(128) float planes[];  // [kFwdPlanes][kPlane]
  __shared__ FwdShared sh;
  __shared__ int next_item;
  constexpr int NC = AUTOMASK ? 2 : 1;             // candidates per source: warp [+ identity]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // Tiles are taken in REVERSE list order: the warp kernel ahead of this one wrote the planes of the coarse scales and
  // of the last samples last, so they are still in L2 when this kernel starts; the backward kernel then runs in list
  // order and starts with what this kernel touched last.  Measured at cfg2: step 271.2 -> 266.8 us (backward kernel
  // reversed instead: 269.8, both: 268.0).
  // flow (tile-level dependencies, mono_params.cuh): list order instead -- the chunks of the first tiles were written
  // first, and the images the backward kernel starts with complete first.
  const bool flow_w = (p.flow & kFlowWarp) != 0, flow_i = (p.flow & kFlowImage) != 0;
  const bool reduce_mean = (p.flags & SDE_MONO_REDUCE_MEAN) != 0;
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    mbar_init_fence();
  }
  // without flow: the warp kernel's planes (and every other tensor) are complete and visible from here on.  With flow
  // the warp kernel may still be running: a tile's inputs are complete (the warp kernel itself was serialised behind
  // whatever produced them), the warped planes of its rows once their chunk flags are set.
  SDE_TRACE_BEGIN(p, 1);
  if (!flow_w) pdl_wait();
  if (flow_i) pdl_launch_dependents();   // backward tiles wait for image flags, not for this grid
  const int total = p.tile_start[p.n_scales];
  int item = 0;
  bool first = true;
  unsigned tma_phase = 0;   // parity of the TMA barrier: carried from tile to tile
  // work items: tiles (mono_params.cuh: work_next)
  while (work_next(p, 1, total, &next_item, item, first)) {
  const int vbid = flow_w ? item : total - 1 - item;
  const TileCoord tc = decode_tile(p, vbid);
  const int s = tc.s, b = tc.b, h = p.h[s], w = p.w[s], hw = h * w;
  // TMA path (row pitch a multiple of 16 bytes): target, depth and the first source's unwarped frame (the
  // identity candidate) are handed to the copy engine before anything else happens for the tile
  const bool tma = p.tma[s] != 0;
  const bool prewarp = p.prewarp[s] != 0;   // the warped sources come from the warp kernel (mono_warp.cu)
  const int ox = tc.x0 - 1, oy = tc.y0 - 1;
  if (tma && tid == 0 && ((p.persist >> 1) & 1)) proxy_fence();   // the planes were last written by this CTA's own stores
  if (tma && tid == 0) {
    mbar_arrive_expect_tx(&sh.bar, (4 + (AUTOMASK ? 3 : 0) + (p.prewarp[s] ? 3 : 0)) * kPlaneBytesTma);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tma_load_plane(planes + (kPlA + c) * kPlane, &maps.target[s], &sh.bar, ox - kColOff, oy, b * 3 + c);
      if (AUTOMASK) tma_load_plane(planes + (kPlI + c) * kPlane, &maps.source[s][0], &sh.bar, ox - kColOff, oy, b * 3 + c);
      if (p.prewarp[s] && !flow_w) tma_load_plane(planes + (kPlS + c) * kPlane, &maps.warped[s][0], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
    }
    tma_load_plane(planes + kPlD * kPlane, &maps.depth[s], &sh.bar, ox - kColOff, oy, b);
  }
  if (flow_w) {
    // the chunks of the warp kernel that hold rows y0 - 1 .. y0 + kTileH of this image (every tile waits, whether or not
    // it reads the planes itself: an image's flag below then implies that all of the image's chunks are complete)
    const int chunks = (hw + kWarpChunk - 1) / kWarpChunk;
    const int row_lo = max(tc.y0 - 1, 0), row_hi = min(tc.y0 + kTileH, h - 1);
    const int c_lo = (row_lo * w) / kWarpChunk, c_hi = (row_hi * w + w - 1) / kWarpChunk;
    const unsigned* wf = p.warp_flag + p.warp_start[s] + b * chunks;
    for (int c = c_lo + tid; c <= c_hi; c += kThreads) flag_wait(wf + c);
  }
  // camera terms: only the in-kernel gather needs them (with the warp kernel's planes the CTA never projects, and
  // nobody waits at this barrier for two threads' global-memory round trip)
  if (tid < p.S && !(tma && prewarp)) {
    Cam cam;
    float k[9];
    load_cam(cam, k, p.K, b, p.sx[s], p.sy[s]);
    if (tid == 0) sh.cam = cam;
    load_proj(sh.proj[tid], k, p.pose[tid], b);
  }
  __syncthreads();
  SDE_TRACE_MARK(p, 1, 3);
  if (flow_w && tma && p.prewarp[s] && tid == 0) {
    flag_proxy_fence();   // the chunks were written with ordinary stores; the copy engine reads them
#pragma unroll
    for (int c = 0; c < 3; ++c)
      tma_load_plane(planes + (kPlS + c) * kPlane, &maps.warped[s][0], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
  }

  const float* __restrict__ depth = p.depth[s] + (size_t)b * hw;
  const float* __restrict__ tg0 = p.target[s] + (size_t)b * 3 * hw;
  // tiles whose halo lies inside the image need no reflection / clamping of the staged positions
  const bool interior = tc.x0 >= 1 && tc.y0 >= 1 && tc.x0 + kTileW + 1 <= w && tc.y0 + kTileH + 1 <= h;

  // phase-2 ownership: rows r0..r0+3 of the tile, columns 2*lane, 2*lane+1
  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;  // halo'd column of the left neighbour of this lane's pair
  const float wS = p.ssim_w * (1.0f / 3.0f), wL = p.l1_w * (1.0f / 3.0f);
  const f2 C1 = bc2(81.0f * p.c1), C2 = bc2(81.0f * p.c2);
  const bool use_ssim = p.ssim_w > 0.0f;

  float best[kRowsPerWarp][2];
  int arg[kRowsPerWarp][2];
#pragma unroll
  for (int o = 0; o < kRowsPerWarp; ++o) {
    best[o][0] = best[o][1] = reduce_mean ? 0.0f : __int_as_float(0x7f800000);
    arg[o][0] = arg[o][1] = 0;
  }

  StageArgs sa;
  sa.depth_mode = p.depth_mode; sa.min_disp = p.min_disp; sa.disp_range = p.disp_range;
  sa.depth = depth; sa.src = nullptr; sa.tgt = tg0; sa.amap = nullptr;
  sa.planes = planes; sa.arg = nullptr; sa.oy = tc.y0 - 1; sa.ox = tc.x0 - 1; sa.h = h; sa.w = w; sa.hw = hw;
  sa.plS = kPlS; sa.plI = kPlI; sa.plA = kPlA; sa.plD = kPlD;
  // ------------------------------------------------------------------ phase 0: depth + target
  if (tma) {
    mbar_wait(&sh.bar, tma_phase);
    tma_phase ^= 1u;
    decode_depth_plane(planes + kPlD * kPlane, p.depth_mode, p.min_disp, p.disp_range, tid);
    if (p.depth_mode != SDE_DEPTH_IS_DEPTH && !interior) __syncthreads();   // the fix-up copies decoded values
    if (!interior) {
      reflect_fixup(planes, kPlA, 3, oy, ox, h, w, tid);
      reflect_fixup(planes, kPlD, 1, oy, ox, h, w, tid);
      if (AUTOMASK) reflect_fixup(planes, kPlI, 3, oy, ox, h, w, tid);
      if (prewarp) reflect_fixup(planes, kPlS, 3, oy, ox, h, w, tid);
    }
  } else {
    if (interior) stage_target<true, false>(sa, tid, false);
    else          stage_target<false, false>(sa, tid, false);
  }
  __syncthreads();

  for (int j = 0; j < p.S; ++j) {
    // ---------------------------------------------------------------- phase 1
    {
      const Cam cam = sh.cam;
      const Proj pj = sh.proj[j];
      sa.src = p.source[s][j] + (size_t)b * 3 * hw;
      if (tma) {
        // the identity planes come from the copy engine; so do the warped ones when the warp kernel ran,
        // otherwise gather here
        if (!prewarp) {
          if (interior) stage_source<true, false, SDE_NB>(sa, cam, pj, tid);
          else          stage_source<false, false, SDE_NB>(sa, cam, pj, tid);
        }
        if ((AUTOMASK || prewarp) && j > 0) {
          mbar_wait(&sh.bar, tma_phase);
          tma_phase ^= 1u;
          if (!interior) {
            if (AUTOMASK) reflect_fixup(planes, kPlI, 3, oy, ox, h, w, tid);
            if (prewarp) reflect_fixup(planes, kPlS, 3, oy, ox, h, w, tid);
          }
        }
      } else {
        if (interior) stage_source<true, AUTOMASK, SDE_NB>(sa, cam, pj, tid);
        else          stage_source<false, AUTOMASK, SDE_NB>(sa, cam, pj, tid);
      }
    }
    __syncthreads();

    // ---------------------------------------------------------------- phase 2
    f2 acc[NC][kRowsPerWarp];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
      for (int o = 0; o < kRowsPerWarp; ++o) acc[k][o] = bc2(0.0f);

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      const float* pa = planes + (kPlA + c) * kPlane + plane_index(r0, c0);
      f2 hA[2], hAA[2];                 // horizontal 3-sums of the two previous rows
      f2 hX[NC][2], hXX[NC][2], hXA[NC][2];
      f2 dXA[NC];                       // X - A of the previous row's centre pair
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
        const Row4 a = ld_row(pa + rr * kPitch);
        const f2 aa = a.c * a.c;
        const f2 nA = (a.c + swp(a.c)) + a.o;
        const f2 nAA = fma2(a.o, a.o, aa + swp(aa));
        f2 nX[NC], nXX[NC], nXA[NC], dn[NC];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const Row4 x = ld_row(pa + ((k == 0 ? kPlS : kPlI) - kPlA) * kPlane + rr * kPitch);
          const f2 xx2 = x.c * x.c, xa = x.c * a.c;
          nX[k] = (x.c + swp(x.c)) + x.o;
          nXX[k] = fma2(x.o, x.o, xx2 + swp(xx2));
          nXA[k] = fma2(x.o, a.o, xa + swp(xa));
          dn[k] = x.c - a.c;
        }
        if (rr >= 2) {
          // window sums (not divided by 9): every SSIM factor below is the reference's times 81,
          // which cancels in the ratio and avoids a rounded 1/9 constant.  Numerator and
          // denominator use mirrored operation orders so that X == A gives ssim == 1 exactly.
          const int o = rr - 2;
          const f2 sA = (hA[0] + hA[1]) + nA;
          const f2 sAA = (hAA[0] + hAA[1]) + nAA;
          const f2 aa2 = sA * sA;
          const f2 vA = fma2(aa2, bc2(-1.0f), sAA * bc2(9.0f));    // 81 sigma_y
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            f2 a2 = acc[k][o];
            if (use_ssim) {
              const f2 sX = (hX[k][0] + hX[k][1]) + nX[k];
              const f2 sXX = (hXX[k][0] + hXX[k][1]) + nXX[k];
              const f2 sXA = (hXA[k][0] + hXA[k][1]) + nXA[k];
              const f2 t = sX * sA, xs = sX * sX;
              const f2 n1 = fma2(bc2(2.0f), t, C1);
              const f2 n2 = fma2(bc2(2.0f), fma2(t, bc2(-1.0f), sXA * bc2(9.0f)), C2);
              const f2 d1 = (xs + aa2) + C1;
              const f2 d2 = (fma2(xs, bc2(-1.0f), sXX * bc2(9.0f)) + vA) + C2;
              const f2 nssim = ndiv2(n1 * n2, d1 * d2);   // -ssim
              // clamp((1 - ssim) / 2, 0, 1), ssim_loss.py:53
              const f2 l = mk2(__saturatef(fmaf(lo(nssim), 0.5f, 0.5f)), __saturatef(fmaf(hi(nssim), 0.5f, 0.5f)));
              a2 = fma2(l, bc2(wS), a2);
            }
            acc[k][o] = mk2(fmaf(fabsf(lo(dXA[k])), wL, lo(a2)), fmaf(fabsf(hi(dXA[k])), wL, hi(a2)));
          }
        }
        hA[0] = hA[1]; hA[1] = nA;
        hAA[0] = hAA[1]; hAA[1] = nAA;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          hX[k][0] = hX[k][1]; hX[k][1] = nX[k];
          hXX[k][0] = hXX[k][1]; hXX[k][1] = nXX[k];
          hXA[k][0] = hXA[k][1]; hXA[k][1] = nXA[k];
          dXA[k] = dn[k];
        }
      }
    }
    // candidates of this source: 2j (warp), 2j+1 (identity) with automask, else j
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int idx = AUTOMASK ? 2 * j + k : j;
#pragma unroll
      for (int o = 0; o < kRowsPerWarp; ++o) {
        const float v0 = lo(acc[k][o]), v1 = hi(acc[k][o]);
        if (reduce_mean) {
          best[o][0] += v0;
          best[o][1] += v1;
        } else {
          // strict '<' keeps the first index on ties; a NaN candidate sticks (torch.min propagates NaN)
          if (v0 < best[o][0] || v0 != v0) { best[o][0] = v0; arg[o][0] = idx; }
          if (v1 < best[o][1] || v1 != v1) { best[o][1] = v1; arg[o][1] = idx; }
        }
      }
    }
    if (j + 1 < p.S) {
      __syncthreads();  // S/I planes are overwritten by the next source
      if (tma && (AUTOMASK || prewarp) && tid == 0) {   // next source's planes (fetched while a gather, if any, runs)
        proxy_fence();
        mbar_arrive_expect_tx(&sh.bar, ((AUTOMASK ? 3 : 0) + (prewarp ? 3 : 0)) * kPlaneBytesTma);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (AUTOMASK) tma_load_plane(planes + (kPlI + c) * kPlane, &maps.source[s][j + 1], &sh.bar, ox - kColOff, oy, b * 3 + c);
          if (prewarp) tma_load_plane(planes + (kPlS + c) * kPlane, &maps.warped[s][j + 1], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
        }
      }
    }
  }

  // the heavy part of this tile is done: once that holds for every tile, the next kernel of the stream may be
  // scheduled (it waits for this grid to complete before it touches memory)
  if (!flow_i && !((p.persist >> 1) & 1)) pdl_launch_dependents();
  // ------------------------------------------------------------------ per-thread sums, argmin, smoothness
  float rec = 0.0f, smx = 0.0f, smy = 0.0f, sinv = 0.0f;
  const int gx0 = tc.x0 + c0;  // image column of this lane's first pixel
  uint8_t* __restrict__ amap = p.argmin[s] ? p.argmin[s] + (size_t)b * hw : nullptr;
#pragma unroll
  for (int o = 0; o < kRowsPerWarp; ++o) {
    const int gy = tc.y0 + r0 + o;
    if (gy < h) {
      if (gx0 < w) { rec += best[o][0]; if (amap) amap[gy * w + gx0] = (uint8_t)arg[o][0]; }
      if (gx0 + 1 < w) { rec += best[o][1]; if (amap) amap[gy * w + gx0 + 1] = (uint8_t)arg[o][1]; }
    }
  }
  if (p.smooth_scale[s] > 0.0f)
    tile_smoothness(planes, kPlD, kPlA, r0, c0, tc.x0, tc.y0, h, w, p.NB, p.smooth_g[s] ? p.smooth_g[s] + (size_t)b * hw : nullptr,
                    smx, smy, sinv);

  // ------------------------------------------------------------------ CTA reduction -> partial slot
  rec = warp_sum(rec); smx = warp_sum(smx); smy = warp_sum(smy); sinv = warp_sum(sinv);
  if (lane == 0) { sh.red[0][wid] = rec; sh.red[1][wid] = smx; sh.red[2][wid] = smy; sh.red[3][wid] = sinv; }
  __syncthreads();
  // ------------------------------------------------------------------ two-level fixed-order reduction
  // The last tile of a (scale, image) pair adds that image's partial slots (while other images are still
  // being computed); the last pair to finish adds the per-image results.  Same order every run.
  // One thread publishes the slot, fences it and takes the ticket: the other threads' stores (argmin bytes,
  // smoothness gradient) are consumed by later kernels only and need no fence here.
  const int q = s * p.B + b, per = p.tiles_x[s] * p.tiles_y[s];
  SDE_TRACE_MARK(p, 1, 1);
  if (tid == 0) {
    float4 v;
    v.x = ((sh.red[0][0] + sh.red[0][1]) + sh.red[0][2]) + sh.red[0][3];
    v.y = ((sh.red[1][0] + sh.red[1][1]) + sh.red[1][2]) + sh.red[1][3];
    v.z = ((sh.red[2][0] + sh.red[2][1]) + sh.red[2][2]) + sh.red[2][3];
    v.w = ((sh.red[3][0] + sh.red[3][1]) + sh.red[3][2]) + sh.red[3][3];
    *reinterpret_cast<float4*>(p.partials + (size_t)vbid * 4) = v;
    publish_fence();
    sh.ticket = atomicAdd(p.img_counter + q, 1u);
  }
  __syncthreads();
  if (sh.ticket != (unsigned)(per - 1)) continue;
  publish_fence();
  {
    const float* part = p.partials + ((size_t)p.tile_start[s] + (size_t)b * per) * 4;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    for (int t = tid; t < per; t += kThreads) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(part) + t);
      a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    }
    if (flow_w) {
      // every tile of this image has passed its chunk flags: leave them cleared for the next call
      const int chunks = (hw + kWarpChunk - 1) / kWarpChunk;
      unsigned* wf = p.warp_flag + p.warp_start[s] + b * chunks;
      for (int c = tid; c < chunks; c += kThreads) wf[c] = 0u;
    }
    __syncthreads();   // sh.red is reused below
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) sh.dred[k][wid] = a[k];
    }
    __syncthreads();
    if (tid == 0) {
      double t4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) t4[k] = ((sh.dred[k][0] + sh.dred[k][1]) + sh.dred[k][2]) + sh.dred[k][3];
      const double qh = h, qw = w, nB = p.NB;
      const int ncand = NC * p.S;
      double rec_q = t4[0] / (nB * qh * qw) / p.n_scales;
      if (reduce_mean) rec_q /= ncand;
      const double mbar = fmax(t4[3] / (qh * qw), 1e-6);
      const double Lb = (t4[1] / (nB * qh * (qw - 1.0)) + t4[2] / (nB * (qh - 1.0) * qw)) / mbar;
      p.stats[q * 2 + 0] = (float)mbar;
      p.stats[q * 2 + 1] = (float)Lb;
      p.fin[q * 2 + 0] = rec_q;
      p.fin[q * 2 + 1] = Lb * (double)p.smooth_scale[s];
      p.img_counter[q] = 0u;   // leave the workspace zeroed for the next call
      publish_fence();
      // flow: this image's statistics are final, and so are the argmin bytes and smoothness gradients of its tiles (each
      // tile released them with its ticket, this thread acquired them with the last one): backward tiles may start
      if (flow_i) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.img_flag + q), "r"(1u) : "memory");
      sh.ticket = atomicAdd(p.counter, 1u);
    }
  }
  __syncthreads();
  const int pairs = p.n_scales * p.B;
  if (sh.ticket != (unsigned)(pairs - 1)) continue;
  publish_fence();
  // flow: this grid did not wait for the warp kernel as a grid; the last CTA does (the warp kernel has long finished),
  // so that the completion of this grid implies the completion of the one before it
  if (flow_w) pdl_wait();
  if (wid == 0) {
    double r = 0.0, sm = 0.0;
    for (int k = lane; k < pairs; k += 32) { r += __ldcg(p.fin + k * 2); sm += __ldcg(p.fin + k * 2 + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); sm += __shfl_xor_sync(0xffffffffu, sm, o); }
    if (lane == 0) {
      p.losses[0] = (float)r;
      p.losses[1] = (float)sm;
      *p.counter = 0u;
    }
  }
  }   // work items
  if (!flow_i && ((p.persist >> 1) & 1)) pdl_launch_dependents();
  work_leave(p, 1);
}

size_t mono_fwd_smem_bytes() { return (size_t)kFwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_mono_fwd(const MonoParams& p, const MonoTma& t, cudaStream_t stream) {
  const bool automask = (p.flags & SDE_MONO_AUTOMASK) != 0;
  auto kernel = automask ? mono_fwd_kernel<true> : mono_fwd_kernel<false>;
  // 48.9 KB of dynamic shared memory needs the opt-in attribute (per device; cheap and idempotent)
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mono_fwd_smem_bytes());
  if (e != cudaSuccess) return e;
  if (SDE_FWD_CARVEOUT >= 0) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, SDE_FWD_CARVEOUT);
  unsigned grid = (unsigned)p.tile_start[p.n_scales];
  if (((p.persist >> 1) & 1)) {
    static unsigned slots[2] = {0, 0};
    unsigned& sl = slots[automask ? 1 : 0];
    if (!sl) sl = resident_ctas(kernel, kThreads, mono_fwd_smem_bytes());
    if (sl && grid > sl) grid = sl;
  }
  return launch_chained(1, kernel, grid, kThreads, mono_fwd_smem_bytes(), stream, p, t);
}

}  // namespace sde
