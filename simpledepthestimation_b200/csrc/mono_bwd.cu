// Fused backward of the MonoDepth2 self-supervised loss, all scales in one launch (sm_100a).
//
// Two instantiations.  SAVED = false: nothing but the uint8 argmin maps and O(B) per-image statistics is kept
// from the forward pass and the warp is recomputed (phase 1 and phase 4 gather from the source frames).
// SAVED = true (default of the host side): the forward pass kept, per source, the warped planes and their
// derivatives w.r.t. the sample coordinate (mono_warp.cu) and the local smoothness gradient (mono_fwd.cu); phase 1
// is a TMA box load, phase 4 reads the derivative planes and the smoothness tail is a multiply-add.
// One CTA owns a 64x16 block Q of SSIM window centres of one
// (scale, sample) and emits gradients for Q's 60x14 interior P (a pixel's gradient collects the
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
//   phase 4  warp-autonomous (no CTA barrier): every warp compacts the pixels of ITS rows of P that carry a
//            gradient into a dense fixed-order list, then per listed pixel: re-project, gX = sum_c gS_c dS_c/dX
//            (derivative planes, or the four taps re-read; gated as nan_to_num/clamp gate the reference's
//            autograd), then the camera-space gradient K^T g_p, accumulated into 12 pose sums per source (9 for
//            R, 3 for t) and into d loss / d depth.
// The smoothness gradient uses the saved per-image mean inverse depth and loss (1-homogeneity,
// SURVEY.md A.5).  Pose sums: per-warp slots, added by the last CTA of a sample in a fixed order in fp64
// (deterministic; no float atomics anywhere).
#include <type_traits>

#include "mono_device.cuh"

namespace sde {

// Ten planes: A[3], depth and two 3-plane regions X, Y.  For an even source the warped planes S sit in X and the
// coefficient planes in Y, for an odd one the other way round: gS_c overwrites S_c in place (S_c is dead once the
// adjoint pass of channel c has read its own pixel), the next source's S planes are prefetched into the region whose
// coefficients were just consumed, and 10 planes (53 KB) instead of 13 let a fourth CTA fit on an SM.
constexpr int kBwdPlanes = 10;
constexpr int kBA = 0, kBD = 3, kBX = 4, kBY = 7;
#ifndef SDE_NB
#define SDE_NB 2
#endif
// the gradient block P starts at plane column 3 (not 2): with 60-wide tiles that puts plane index 0 at image
// column tile_x0 - 4, a multiple of 4, which the TMA box start needs
constexpr int kBwdColOff = 3;
constexpr int kWarps = kThreads / 32;
constexpr int kListPerWarp = kRowsPerWarp * kTileW;   // a warp lists pixels of its own four rows only

struct BwdShared {
  Cam cam;
  Proj proj[SDE_MAX_SOURCES];
  // SAVED: d depth-gradient / d (a0, a1, a2) as an affine function of the (centred) pixel coordinates, per source:
  // rows of diag-combined K^T-weights times R K^-1 (see phase 4); row 2 is stored negated
  float wmat[SDE_MAX_SOURCES][9];
  // SAVED: the projection as p_i = depth * (N_i . [x', y', 1]) + tau_i with N = K R K^-1 re-centred on (w/2, h/2):
  // [0..8] N, [9..11] tau = K t, [12] cx, [13] cy -- phase 4 recomputes the sample coordinate with it (packed fp32)
  float nmat[SDE_MAX_SOURCES][14];
  double dred[12][kWarps];
  unsigned ticket;
  __align__(8) uint64_t bar;                     // TMA completion barrier
  unsigned short list[kWarps][kListPerWarp];     // per warp: dense list ((plane row << 7) | stored column) of its pixels of P that carry a gradient
  __align__(8) uint8_t arg[kPlane];
};

#ifndef SDE_BWD_OCC
#define SDE_BWD_OCC 4
#endif

template <bool SAVED>
__global__ void __launch_bounds__(kThreads, SDE_BWD_OCC) mono_bwd_kernel(const __grid_constant__ MonoParams p,
                                                               const __grid_constant__ MonoTma maps) {
  extern __shared__ __align__(128) float planes[];  // [kBwdPlanes][kPlane]
  __shared__ BwdShared sh;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int vbid = (int)blockIdx.x;   // list order (the forward kernel runs in reverse order, see mono_fwd.cu)
  const TileCoord tc = decode_btile(p, vbid);
  const int s = tc.s, b = tc.b, h = p.h[s], w = p.w[s], hw = h * w;
  const bool automask = (p.flags & SDE_MONO_AUTOMASK) != 0;
  const bool reduce_mean = (p.flags & SDE_MONO_REDUCE_MEAN) != 0;
  const bool use_ssim = p.ssim_w > 0.0f;
  // plane (yy, xx) <-> image (oy + yy, ox + xx); Q = plane [1..16]x[1..64]; P = plane [2..15]x[3..62]
  const int ox = tc.x0 - kBwdColOff, oy = tc.y0 - 2;
  const bool interior = ox >= 0 && oy >= 0 && ox + kHW <= w && oy + kHH <= h;
  // reflect-pad adjoint: pixels next to the image border also receive the mirrored pad position
  const bool lr_border = ox + 2 <= 1 || ox + kHW - 3 >= w - 2;

  // Camera terms (needed from phase 4 on; SAVED = false: from phase 1 on): threads 32 .. 32+S-1 issue the raw loads
  // of K and the pose now and form the terms after the argmin bytes are staged, so that nobody waits at the first
  // barrier for two threads' global-memory round trip.
  const bool cam_thread = tid >= 32 && tid < 32 + p.S;
  // coefficient planes: the coefficient pass writes every window centre of the block, the one-pixel ring around it
  // is never written and must read as zero (shared memory only: before the dependency wait)
  zero_ring<3>(planes, kBY, tid);
  const bool tma = SAVED && p.tma[s] != 0;
  if (tma && tid == 0) {
    mbar_init(&sh.bar, 1);
    mbar_init_fence();
  }
  // without flow: the forward pass's outputs (and every other tensor) are complete and visible from here on.  With flow
  // (mono_params.cuh) the forward kernel may still be running: this tile waits for its image's flag -- the image's
  // statistics, argmin bytes, smoothness gradients and (through the forward tiles' own waits) warped planes are final.
  const bool flow_i = (p.flow & kFlowImage) != 0;
  SDE_TRACE_BEGIN(p, 2);
  if (!flow_i) pdl_wait();
  else if (tid == 0) {
    flag_wait(p.img_flag + s * p.B + b);
    flag_proxy_fence();   // the warped planes were written with ordinary stores; the copy engine reads them
  }
  CamRaw raw;
  if (cam_thread) load_cam_raw(raw, p.K, p.pose[tid - 32], b);
  auto make_camera = [&]() {
    if (cam_thread) {
      Cam cam;
      float k[9];
      load_cam(cam, k, raw.k, 0, p.sx[s], p.sy[s]);
      if (tid == 32) sh.cam = cam;
      Proj& pj = sh.proj[tid - 32];
      load_proj(pj, k, raw.T, 0);
      if (SAVED) {
        // d (R P + t) / d depth = R K^-1 [x, y, 1]^T =: V [x, y, 1]^T; d loss / d depth = (K^T g_p) . V [x, y, 1]^T with
        // K^T g_p = (fx a0, sk a0 + fy a1, a2): rows of W = (fx V0 + sk V1, fy V1, V2), re-centred on (w/2, h/2)
        float V[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) V[r * 3 + c] = pj.r[r * 3] * cam.ki[c] + pj.r[r * 3 + 1] * cam.ki[3 + c] + pj.r[r * 3 + 2] * cam.ki[6 + c];
        const float x0c = (float)(w >> 1), y0c = (float)(h >> 1);
        float* wm = sh.wmat[tid - 32];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wm[c] = cam.fx * V[c] + cam.sk * V[3 + c];
          wm[3 + c] = cam.fy * V[3 + c];
          wm[6 + c] = -V[6 + c];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) wm[r * 3 + 2] += wm[r * 3] * x0c + wm[r * 3 + 1] * y0c;
        float* nm = sh.nmat[tid - 32];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) nm[r * 3 + c] = pj.m[r * 3] * cam.ki[c] + pj.m[r * 3 + 1] * cam.ki[3 + c] + pj.m[r * 3 + 2] * cam.ki[6 + c];
          nm[r * 3 + 2] += nm[r * 3] * x0c + nm[r * 3 + 1] * y0c;
          nm[9 + r] = pj.tau[r];
        }
        nm[12] = cam.cx; nm[13] = cam.cy;
      }
    }
  };
  if (!SAVED) make_camera();
  // TMA path (saved warps, row pitch a multiple of 16 bytes): one thread hands the tile planes -- target,
  // depth and the first source's warp -- to the copy engine before anything else happens in the CTA
  if (tma && tid == 0) {
    mbar_arrive_expect_tx(&sh.bar, 7 * kPlaneBytesTma);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tma_load_plane(planes + (kBA + c) * kPlane, &maps.target[s], &sh.bar, ox - kColOff, oy, b * 3 + c);
      tma_load_plane(planes + (kBX + c) * kPlane, &maps.warped[s][0], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
    }
    tma_load_plane(planes + kBD * kPlane, &maps.depth[s], &sh.bar, ox - kColOff, oy, b);
  }
  __syncthreads();   // (flow: thread 0 arrives here after the image's flag)
  const unsigned behind_flag = launder_zero();   // see ld_plane (sde_common.cuh)
  SDE_TRACE_MARK(p, 2, 3);

  const float* __restrict__ depth = p.depth[s] + (size_t)b * hw;
  const float* __restrict__ tg0 = p.target[s] + (size_t)b * 3 * hw;
  const uint8_t* __restrict__ amap = reduce_mean ? nullptr : p.argmin[s] + (size_t)b * hw;

  const float g_rec = __ldg(p.grad_losses), g_smooth = __ldg(p.grad_losses + 1);
  // d rec_loss / d pe_q for a selected pixel (MonoDepth2.py:116-124); the normaliser comes from the host
  const float g_pe = g_rec * p.inv_norm[s];
  const float g_l1 = g_pe * p.l1_w * (1.0f / 3.0f);
  const float g_ss = g_pe * p.ssim_w * (1.0f / 3.0f) * -0.5f;  // d pe / d ssim (ssim_loss.py:53)

  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;
  const f2 C1 = bc2(81.0f * p.c1), C2 = bc2(81.0f * p.c2);

  // image columns of this lane's pixel pair; extra multiplicities of the reflect-pad adjoint
  const int px0 = ox + c0 + 1, px1 = px0 + 1;
  const float eL0 = px0 == 1 ? 1.0f : 0.0f, eL1 = px1 == 1 ? 1.0f : 0.0f;
  const float eR0 = px0 == w - 2 ? 1.0f : 0.0f, eR1 = px1 == w - 2 ? 1.0f : 0.0f;

  // d loss / d depth of this lane's pixel pairs (rows r0+1 .. r0+4, columns c0+1, c0+2 of the plane), summed over sources
  f2 gd[kRowsPerWarp];
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) gd[k] = bc2(0.0f);
  // this lane's pair belongs to the gradient block P (plane columns 3..62) and to the image
  f2 Gs[kRowsPerWarp];   // SAVED: local smoothness gradient of the same pairs (loaded before the last phase 4)
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) Gs[k] = bc2(0.0f);
  const int gxp = ox + c0 + 1;
  const bool colP = lane >= 1 && lane <= 30;
  const bool col_ok0 = colP && gxp < w, col_ok1 = colP && gxp + 1 < w;

  StageArgs sa;
  sa.depth_mode = p.depth_mode; sa.min_disp = p.min_disp; sa.disp_range = p.disp_range;
  sa.depth = depth; sa.src = nullptr; sa.tgt = tg0; sa.amap = amap;
  sa.planes = planes; sa.arg = sh.arg; sa.oy = oy; sa.ox = ox; sa.h = h; sa.w = w; sa.hw = hw;
  sa.plS = kBX; sa.plI = 0; sa.plA = kBA; sa.plD = kBD;
  // ------------------------------------------------------------------ phase 0: depth + target + argmin
  unsigned tma_phase = 0;
  if (tma) {
    // argmin bytes: plain loads while the copy engine works (the TMA path implies w % 4 == 0 and ox - kColOff is a
    // multiple of 4 by construction of the tiles, so whole words can be moved if the map itself is word-aligned)
    if ((reinterpret_cast<uintptr_t>(amap) & 3) == 0) stage_arg_words(sh.arg, amap, oy, ox, h, w, reduce_mean, tid);
    else stage_arg(sh.arg, amap, oy, ox, h, w, reduce_mean, tid);
  } else {
    if (interior) stage_target<true, true>(sa, tid, reduce_mean);
    else          stage_target<false, true>(sa, tid, reduce_mean);
  }
  if (SAVED) make_camera();
  __syncthreads();

  for (int j = 0; j < p.S; ++j) {
    // byte of the argmin plane that selects this source's warped candidate ('mean': every staged pixel holds 254)
    const int cand = reduce_mean ? 254 : (automask ? 2 * j : j);
    // plane regions of this source: S (later gS, in place) and the coefficients
    const int bS = (j & 1) ? kBY : kBX, bC = (j & 1) ? kBX : kBY;
    sa.plS = bS;
    if (j > 0) {
      // the coefficient region held the previous source's S / gS planes: its one-pixel border ring (never written
      // by the coefficient pass, never read by phase 4) must read as zero
      zero_ring<3>(planes, bC, tid);
    }
    const Cam cam = sh.cam;
    const Proj pj = sh.proj[j];
    const float* __restrict__ sc0 = p.source[s][j] + (size_t)b * 3 * hw;
    const float* __restrict__ sc1 = sc0 + hw;
    const float* __restrict__ sc2 = sc1 + hw;
    // ---------------------------------------------------------------- phase 1
    sa.src = sc0;
    if (tma) {
      mbar_wait(&sh.bar, tma_phase);
      tma_phase ^= 1u;
      if (j == 0 && p.depth_mode != SDE_DEPTH_IS_DEPTH) {
        decode_depth_plane(planes + kBD * kPlane, p.depth_mode, p.min_disp, p.disp_range, tid);
        if (!interior) __syncthreads();   // the fix-up copies decoded values
      }
      if (!interior) {
        if (j == 0) reflect_fixup(planes, kBA, 3, oy, ox, h, w, tid), reflect_fixup(planes, kBD, 1, oy, ox, h, w, tid);
        reflect_fixup(planes, bS, 3, oy, ox, h, w, tid);
      }
    } else if (SAVED) {
      const float* wsrc = p.warped[s][j] + (size_t)b * kSavedPlanes * hw;
      if (interior) stage_saved<true>(sa, wsrc, tid);
      else          stage_saved<false>(sa, wsrc, tid);
    } else {
      if (interior) stage_source<true, false, SDE_NB>(sa, cam, pj, tid);
      else          stage_source<false, false, SDE_NB>(sa, cam, pj, tid);
    }
    __syncthreads();

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      // -------------------------------------------------------------- phase 2: SSIM coefficients on Q
      if (use_ssim) {
        const float* pa = planes + (kBA + c) * kPlane + plane_index(r0, c0);
        f2 hA[2], hAA[2], hX[2], hXX[2], hXA[2];
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
          const Row4 a = ld_row(pa + rr * kPitch);
          const Row4 x = ld_row(pa + (bS - kBA) * kPlane + rr * kPitch);
          const f2 aa = a.c * a.c, xx2 = x.c * x.c, xa = x.c * a.c;
          const f2 nA = (a.c + swp(a.c)) + a.o, nAA = fma2(a.o, a.o, aa + swp(aa));
          const f2 nX = (x.c + swp(x.c)) + x.o, nXX = fma2(x.o, x.o, xx2 + swp(xx2));
          const f2 nXA = fma2(x.o, a.o, xa + swp(xa));
          if (rr >= 2) {
            const int row = r0 + rr - 1;  // plane row of the window centre
            const uchar2 m = *reinterpret_cast<const uchar2*>(sh.arg + plane_index(row, c0 + 1));
            const bool sel0 = m.x == cand, sel1 = m.y == cand;
            f2 ca = bc2(0.0f), cb = bc2(0.0f), cc = bc2(0.0f);
            {   // (no skip of rows of 64 unselected windows: the branch kept the scheduler from interleaving the unrolled rows,
                //  measured in mono_bwd_pair.cu: -1.6 %; unselected windows get zero coefficients through g)
              const f2 sA = (hA[0] + hA[1]) + nA, sAA = (hAA[0] + hAA[1]) + nAA;
              const f2 sX = (hX[0] + hX[1]) + nX, sXX = (hXX[0] + hXX[1]) + nXX, sXA = (hXA[0] + hXA[1]) + nXA;
              // same operation order as the forward kernel
              const f2 aa2 = sA * sA, xs = sX * sX, t = sX * sA;
              const f2 vA = fma2(aa2, bc2(-1.0f), sAA * bc2(9.0f));
              const f2 n1 = fma2(bc2(2.0f), t, C1);
              const f2 n2 = fma2(bc2(2.0f), fma2(t, bc2(-1.0f), sXA * bc2(9.0f)), C2);
              const f2 d1 = (xs + aa2) + C1;
              const f2 d2 = (fma2(xs, bc2(-1.0f), sXX * bc2(9.0f)) + vA) + C2;
              const f2 N = n1 * n2, D = d1 * d2;
              // both quotients negated (ndiv2: one packed instruction less each); the signs fold into the constants below
              const f2 nssim = ndiv2(N, D);
              const f2 ninvD = ndiv2(bc2(1.0f), D);
              // torch.clamp passes the gradient on the closed interval 0 <= (1-ssim)/2 <= 1
              const float h0 = fmaf(lo(nssim), 0.5f, 0.5f), h1 = fmaf(hi(nssim), 0.5f, 0.5f);
              const f2 g = mk2((sel0 && h0 >= 0.0f && h0 <= 1.0f) ? g_ss : 0.0f,
                               (sel1 && h1 >= 0.0f && h1 <= 1.0f) ? g_ss : 0.0f);
              const f2 ngi = g * ninvD;
              // d ssim / d(sum S), d(sum S^2), d(sum S A)   (SURVEY.md A.6 in window-sum form)
              const f2 u = fma2(sX * nssim, d2 - d1, sA * (n2 - n1));
              ca = (ngi * bc2(-2.0f)) * u;
              cb = (ngi * bc2(-18.0f)) * (nssim * d1);
              cc = (ngi * bc2(-18.0f)) * n1;
            }
            float* pc = planes + bC * kPlane + plane_index(row, c0 + 1);
            *reinterpret_cast<unsigned long long*>(pc) = ca.v;
            *reinterpret_cast<unsigned long long*>(pc + kPlane) = cb.v;
            *reinterpret_cast<unsigned long long*>(pc + 2 * kPlane) = cc.v;
          }
          hA[0] = hA[1]; hA[1] = nA; hAA[0] = hAA[1]; hAA[1] = nAA;
          hX[0] = hX[1]; hX[1] = nX; hXX[0] = hXX[1]; hXX[1] = nXX; hXA[0] = hXA[1]; hXA[1] = nXA;
        }
      }
      __syncthreads();
      // -------------------------------------------------------------- phase 3: adjoint gather -> gS_c on P
      // instantiated twice: tiles on the left / right image border add the mirrored pad column (block-uniform)
      auto phase3 = [&](auto lr_tag, auto ssim_tag) {
        constexpr bool LR = decltype(lr_tag)::value;
        constexpr bool SSIM = decltype(ssim_tag)::value;   // compile-time: no control-flow joins inside the unrolled rows
        f2 hq[3][2];  // horizontal 3-sums of a, b, c for the two previous coefficient rows
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
          f2 nq[3];
          if (SSIM) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const Row4 q = ld_row(planes + (bC + k) * kPlane + plane_index(r0 + rr, c0));
              f2 hsum = (q.c + swp(q.c)) + q.o;
              if (LR) hsum = hsum + mk2(eL0 * lo(q.o) + eR0 * hi(q.c), eL1 * lo(q.c) + eR1 * hi(q.o));
              nq[k] = hsum;
            }
          }
          if (rr >= 2) {
            const int row = r0 + rr - 1;        // plane row of pixel p
            const int py = oy + row;
            const f2 wu = bc2(py == 1 ? 2.0f : 1.0f), wd = bc2(py == h - 2 ? 2.0f : 1.0f);
            const f2 Sp = ld2(planes + (bS + c) * kPlane + plane_index(row, c0 + 1));
            const f2 Ap = ld2(planes + (kBA + c) * kPlane + plane_index(row, c0 + 1));
            f2 gS = bc2(0.0f);
            if (SSIM) {
              const f2 va = fma2(wu, hq[0][0], fma2(wd, nq[0], hq[0][1]));
              const f2 vb = fma2(wu, hq[1][0], fma2(wd, nq[1], hq[1][1]));
              const f2 vc = fma2(wu, hq[2][0], fma2(wd, nq[2], hq[2][1]));
              gS = fma2(Sp, vb, fma2(Ap, vc, va));
            }
            // L1 term on the pixel itself: g_l1 * sign(S - A) where this candidate was selected (branch-free)
            const uchar2 m = *reinterpret_cast<const uchar2*>(sh.arg + plane_index(row, c0 + 1));
            const f2 df = Sp - Ap;
            const float d0 = lo(df), d1 = hi(df);
            // g_l1 * sign(d): the sign bit of d flips g_l1 (one LOP3), sign(0) = 0 and unselected pixels through the select
            float l0 = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d0) & 0x80000000));
            float l1 = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d1) & 0x80000000));
            l0 = (m.x == cand && d0 != 0.0f) ? l0 : 0.0f;
            l1 = (m.y == cand && d1 != 0.0f) ? l1 : 0.0f;
            gS = gS + mk2(l0, l1);
            *reinterpret_cast<unsigned long long*>(planes + (bS + c) * kPlane + plane_index(row, c0 + 1)) = gS.v;   // in place of S_c
          }
          if (SSIM) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { hq[k][0] = hq[k][1]; hq[k][1] = nq[k]; }
          }
        }
      };
      if (use_ssim) {
        if (lr_border) phase3(std::true_type{}, std::true_type{});
        else           phase3(std::false_type{}, std::true_type{});
      } else {
        if (lr_border) phase3(std::true_type{}, std::false_type{});
        else           phase3(std::false_type{}, std::false_type{});
      }
      __syncthreads();
    }

    // this source's coefficient planes are dead from here on: let the copy engine put the next source's warp there
    // during phase 4 (the regions swap roles)
    if (tma && j + 1 < p.S && tid == 0) {
      proxy_fence();
      mbar_arrive_expect_tx(&sh.bar, 3 * kPlaneBytesTma);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        tma_load_plane(planes + (bC + c) * kPlane, &maps.warped[s][j + 1], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
    }
    // ---------------------------------------------------------------- phase 4: warp backward on P
    // Only pixels whose 3x3 neighbourhood holds a window that selected this source's warped candidate carry a
    // gradient, so each warp first compacts those of its own rows into a dense list -- in a fixed order, so the
    // pose sums stay deterministic -- and the expensive part runs on full warps.  The gS planes of a warp's rows are
    // read and written by that warp only: no CTA barrier until the next source's planes are needed.
    if (SAVED && j == p.S - 1 && p.smooth_scale[s] > 0.0f) {
      // local smoothness gradient of this lane's pairs (kept by the forward kernel): the loads overlap phase 4
      const float* __restrict__ sg = p.smooth_g[s] + (size_t)b * hw;
      const bool pair = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(sg) & 7) == 0;   // gxp is even
#pragma unroll
      for (int o = 0; o < kRowsPerWarp; ++o) {
        const int row = r0 + 1 + o, gy = oy + row;
        float G0 = 0.0f, G1 = 0.0f;
        if (row >= 2 && row <= kBwdH + 1 && gy < h && col_ok0) {
          if (pair) {
            const float2 G = ld_prod(reinterpret_cast<const float2*>(sg + gy * w + gxp));
            G0 = G.x; G1 = G.y;
          } else {
            G0 = ld_prod(sg + gy * w + gxp);
            if (col_ok1) G1 = ld_prod(sg + gy * w + gxp + 1);
          }
        }
        Gs[o] = mk2(G0, G1);
      }
    }
    if constexpr (SAVED) {
      // Dense form: every lane processes its own pixel pairs in packed fp32; nothing is projected or divided here --
      // the warp kernel left q dS_c/dX and q dS_c/dY (gated); (ex, ey) = (X - cx, Y - cy) is re-projected in packed fp32:
      //   a0 = sum_c gS_c (q dS_c/dX), a1 = sum_c gS_c (q dS_c/dY), a2 = -(a0 ex + a1 ey)      [= g_p in camera units]
      //   K^T g_p = (fx a0, sk a0 + fy a1, a2);  d loss / d depth = a0 w0 + a1 w1 + a2 w2 with w = W [x, y, 1]^T (wmat)
      //   d loss / d [R | t] = K^T g_p (x) [P, 1] with P = depth K^-1 [x, y, 1]^T, i.e. linear in the twelve sums of
      //   a_i depth x, a_i depth y, a_i depth, a_i -- the linear map is applied once per (sample, scale) in fp64 by the
      //   tile that finishes the sample.  x is fixed per lane: sum(b x) = x sum(b), formed after the rows.
      // Pairs outside the gradient block / the image load zeros, so they add nothing.
      const float* __restrict__ dwp = p.warped[s][j] + (size_t)b * kSavedPlanes * hw + behind_flag;
      const float* wm = sh.wmat[j];
      const float* nm = sh.nmat[j];
      const bool pair = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(dwp) & 7) == 0;   // gxp is even
      const float xc0 = (float)(gxp - (w >> 1));
      const f2 xc = mk2(xc0, xc0 + 1.0f);
      f2 Sa[3], Sb[3], Sy[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) Sa[k] = Sb[k] = Sy[k] = bc2(0.0f);
      // all loads of the source's rows are issued before the first multiply-add: one L2 round trip per (warp, source)
      // (instantiated on the alignment switch, so that the loads sit in one basic block)
      auto rows = [&](auto pair_tag) {
        constexpr bool PAIR = decltype(pair_tag)::value;
        f2 dq[kRowsPerWarp][6];
        bool ok[kRowsPerWarp];
#pragma unroll
        for (int o = 0; o < kRowsPerWarp; ++o) {
          const int row = r0 + 1 + o, gy = oy + row;
          const bool row_ok = row >= 2 && row <= kBwdH + 1 && gy < h;   // warp-uniform
          ok[o] = row_ok && col_ok0;
          if (PAIR) {
            // branch-free: pairs outside the block / the image read the sample's first pixel instead (always a valid
            // address) and are switched off through gS below
            SDE_CHECK(!ok[o] || (gy >= 0 && gy < h && gxp >= 0 && gxp + 1 < w));
            const float* q = dwp + (ok[o] ? gy * w + gxp : 0) + 3 * hw;
#pragma unroll
            for (int i = 0; i < 6; ++i) dq[o][i].v = ld_plane(reinterpret_cast<const unsigned long long*>(q + i * hw));
          } else {
            const float* q = dwp + (gy * w + gxp) + 3 * hw;
#pragma unroll
            for (int i = 0; i < 6; ++i)
              dq[o][i] = mk2(ok[o] ? ld_plane(q + i * hw) : 0.0f, row_ok && col_ok1 ? ld_plane(q + i * hw + 1) : 0.0f);
          }
        }
        f2 wb[3], nb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          wb[k] = fma2(bc2(wm[3 * k]), xc, bc2(wm[3 * k + 2]));
          nb[k] = fma2(bc2(nm[3 * k]), xc, bc2(nm[3 * k + 2]));
        }
        const float wy0 = wm[1], wy1 = wm[4], wy2 = wm[7];
        const f2 ncx = bc2(-nm[12]), ncy = bc2(-nm[13]);
#pragma unroll
        for (int o = 0; o < kRowsPerWarp; ++o) {
          const int row = r0 + 1 + o, gy = oy + row;
          const int pl = plane_index(row, c0 + 1);
          f2 g0 = ld2(planes + bS * kPlane + pl), g1 = ld2(planes + (bS + 1) * kPlane + pl),
             g2 = ld2(planes + (bS + 2) * kPlane + pl);
          if (PAIR && !ok[o]) g0 = g1 = g2 = bc2(0.0f);
          const f2 dd = ld2(planes + kBD * kPlane + pl);
          const f2 a0 = fma2(g2, dq[o][2], fma2(g1, dq[o][1], g0 * dq[o][0]));
          const f2 a1 = fma2(g2, dq[o][5], fma2(g1, dq[o][4], g0 * dq[o][3]));
          const float yc = (float)(gy - (h >> 1));
          // the sample coordinate again, relative to the principal point: p_i = depth n_i + tau_i, X = p0 / (p2 + 1e-6)
          // (one Newton-refined reciprocal: 2e-7 relative, the planes' own resolution).  Only pixels whose derivative
          // planes are non-zero use it, and those project inside the image; the clamp keeps the rest finite (a NaN
          // clamps to a bound), so 0 * (X - cx) stays 0.
          f2 ex, ey;
          {
            const f2 ycc2 = bc2(yc);
            const f2 n0 = fma2(bc2(nm[1]), ycc2, nb[0]), n1 = fma2(bc2(nm[4]), ycc2, nb[1]), n2 = fma2(bc2(nm[7]), ycc2, nb[2]);
            const f2 p0 = fma2(dd, n0, bc2(nm[9])), p1 = fma2(dd, n1, bc2(nm[10]));
            const f2 den = fma2(dd, n2, bc2(nm[11])) + bc2(1e-6f);
            f2 r = mk2(rcp_approx(lo(den)), rcp_approx(hi(den)));
            r = fma2(fma2(den * bc2(-1.0f), r, bc2(1.0f)), r, r);
            const f2 X = p0 * r, Y = p1 * r;
            const float big = 16777216.0f;
            ex = mk2(fminf(fmaxf(lo(X), -big), big), fminf(fmaxf(hi(X), -big), big)) + ncx;
            ey = mk2(fminf(fmaxf(lo(Y), -big), big), fminf(fmaxf(hi(Y), -big), big)) + ncy;
          }
          const f2 a2 = fma2(a1, ey, a0 * ex);   // -a2 (the sign lives in wmat row 2 and in the final map)
          const f2 b0 = a0 * dd, b1 = a1 * dd, b2 = a2 * dd;
          Sa[0] = Sa[0] + a0; Sa[1] = Sa[1] + a1; Sa[2] = Sa[2] + a2;
          Sb[0] = Sb[0] + b0; Sb[1] = Sb[1] + b1; Sb[2] = Sb[2] + b2;
          const f2 ycc = bc2(yc);
          Sy[0] = fma2(b0, ycc, Sy[0]); Sy[1] = fma2(b1, ycc, Sy[1]); Sy[2] = fma2(b2, ycc, Sy[2]);
          const f2 w0 = wb[0] + bc2(wy0 * yc), w1 = wb[1] + bc2(wy1 * yc), w2 = wb[2] + bc2(wy2 * yc);
          gd[o] = fma2(a2, w2, fma2(a1, w1, fma2(a0, w0, gd[o])));
        }
      };
      if (pair) rows(std::true_type{});
      else      rows(std::false_type{});
      // the 12 sums of this (warp, source) -> per-warp slot: [i][x', y', 1 (all times depth), plain]
      {
        float v16[16];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const f2 sx = Sb[i] * xc;
          v16[4 * i] = lo(sx) + hi(sx);
          v16[4 * i + 1] = lo(Sy[i]) + hi(Sy[i]);
          v16[4 * i + 2] = lo(Sb[i]) + hi(Sb[i]);
          v16[4 * i + 3] = lo(Sa[i]) + hi(Sa[i]);
        }
#pragma unroll
        for (int k = 12; k < 16; ++k) v16[k] = 0.0f;
        const float mine = warp_sum16(v16, lane);
        const int slot = warp_slot(lane);
        if ((lane & 1) == 0 && slot < 12) p.pose_partials[(((size_t)vbid * kWarps + wid) * p.S + j) * 12 + slot] = mine;
      }
    } else
    {
      unsigned short* const wlist = sh.list[wid];
      const unsigned lt = (1u << lane) - 1u;
      int total = 0;
#pragma unroll
      for (int o = 0; o < kRowsPerWarp; ++o) {
        const int row = r0 + 1 + o;
        const int pl = plane_index(row, c0 + 1);
        const bool row_ok = row >= 2 && row <= kBwdH + 1 && oy + row < h;
        const f2 g0 = ld2(planes + bS * kPlane + pl), g1 = ld2(planes + (bS + 1) * kPlane + pl),
                 g2 = ld2(planes + (bS + 2) * kPlane + pl);
        const bool s0 = row_ok && col_ok0 && (lo(g0) != 0.0f || lo(g1) != 0.0f || lo(g2) != 0.0f);
        const bool s1 = row_ok && col_ok1 && (hi(g0) != 0.0f || hi(g1) != 0.0f || hi(g2) != 0.0f);
        const unsigned b0 = __ballot_sync(0xffffffffu, s0), b1 = __ballot_sync(0xffffffffu, s1);
        SDE_CHECK(total + __popc(b0) + __popc(b1) <= kListPerWarp);
        if (s0) wlist[total + __popc(b0 & lt)] = (unsigned short)((row << 7) | (c0 + 1 + kColOff));
        total += __popc(b0);
        if (s1) wlist[total + __popc(b1 & lt)] = (unsigned short)((row << 7) | (c0 + 2 + kColOff));
        total += __popc(b1);
      }
      __syncwarp();

      float acc[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
      const float* __restrict__ dw = SAVED ? p.warped[s][j] + ((size_t)b * kSavedPlanes + 3) * hw + behind_flag : nullptr;
      const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
      // the derivative-plane loads of the NEXT listed pixel are in flight while the current one is processed
      float nd[6];
      int npl = 0;
      auto fetch = [&](int k) {
        npl = wlist[k];
        if (SAVED) {
          const int row = npl >> 7, col = (npl & 127) - kColOff;
          const float* q = dw + ((oy + row) * w + (ox + col));
#pragma unroll
          for (int i = 0; i < 6; ++i) nd[i] = ld_plane(q + i * hw);
        }
      };
      if (lane < total) fetch(lane);
#pragma unroll 1
      for (int k = lane; k < total; k += 32) {
        const int ent = npl;   // (plane row << 7) | stored column index
        float dv[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) dv[i] = nd[i];
        if (k + 32 < total) fetch(k + 32);
        const int row = ent >> 7, ci = ent & 127, pl = row * kPitch + ci;
        const int gy = oy + row, gx = ox + ci - kColOff;
        float* const pg = planes + bS * kPlane + pl;
        const float g0 = pg[0], g1 = pg[kPlane], g2 = pg[2 * kPlane];
        const float d = planes[kBD * kPlane + pl];
        const float fxp = (float)gx, fyp = (float)gy;
        float P[3], den, X, Y;
        project_full(cam, pj, fxp, fyp, d, P, den, X, Y);
        // gradient gates of nan_to_num and clamp (closed interval), camera.py:184-188
        const bool gate_x = (X >= 0.0f) && (X <= wm1);   // false for NaN / +-inf
        const bool gate_y = (Y >= 0.0f) && (Y <= hm1);
        float gdep = 0.0f;
        if (gate_x || gate_y) {
          float gX, gY;
          if (SAVED) {
            // derivative planes of the warp kernel (already gated)
            gX = g0 * dv[0]; gX += g1 * dv[1]; gX += g2 * dv[2];
            gY = g0 * dv[3]; gY += g1 * dv[4]; gY += g2 * dv[5];
          } else {
            const Cell cell = bilinear_cell(X, Y, w, h);
            const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
            gX = 0.0f; gY = 0.0f;
            const float gs[3] = {g0, g1, g2};
            const float* pls[3] = {sc0, sc1, sc2};
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              const float* q0 = pls[cc] + cell.off;
              const float* q1 = q0 + w;
              const float v00 = __ldg(q0), v01 = __ldg(q0 + 1), v10 = __ldg(q1), v11 = __ldg(q1 + 1);
              gX += gs[cc] * ((v01 - v00) * by + (v11 - v10) * cell.ay);
              gY += gs[cc] * ((v10 - v00) * bx + (v11 - v01) * cell.ax);
            }
            if (!gate_x) gX = 0.0f;
            if (!gate_y) gY = 0.0f;
          }
          const float q = 1.0f / den;
          const float u0 = gX * q, u1 = gY * q;
          // K^T g_p, with the third row formed per pixel in camera-centred coordinates
          const float dx = gate_x ? X - cam.cx : 0.0f, dy = gate_y ? Y - cam.cy : 0.0f;
          const float gc0 = cam.fx * u0;
          const float gc1 = cam.sk * u0 + cam.fy * u1;
          const float gc2 = -(gX * dx + gY * dy) * q;
          acc[0] += gc0 * P[0]; acc[1] += gc0 * P[1]; acc[2] += gc0 * P[2]; acc[3] += gc0;
          acc[4] += gc1 * P[0]; acc[5] += gc1 * P[1]; acc[6] += gc1 * P[2]; acc[7] += gc1;
          acc[8] += gc2 * P[0]; acc[9] += gc2 * P[1]; acc[10] += gc2 * P[2]; acc[11] += gc2;
          // d/d depth: (R^T K^T g_p) . K^-1 [x,y,1]
          const float gP0 = pj.r[0] * gc0 + pj.r[3] * gc1 + pj.r[6] * gc2;
          const float gP1 = pj.r[1] * gc0 + pj.r[4] * gc1 + pj.r[7] * gc2;
          const float gP2 = pj.r[2] * gc0 + pj.r[5] * gc1 + pj.r[8] * gc2;
          const float rx = cam.ki[0] * fxp + cam.ki[1] * fyp + cam.ki[2];
          const float ry = cam.ki[3] * fxp + cam.ki[4] * fyp + cam.ki[5];
          const float rz = cam.ki[6] * fxp + cam.ki[7] * fyp + cam.ki[8];
          gdep = gP0 * rx + gP1 * ry + gP2 * rz;
        }
        pg[0] = gdep;   // in place of gS_0: unlisted pixels of P hold gS_0 == 0 there
      }
      // the 12 pose sums of this (warp, source) -> per-warp slot
      {
        float v16[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v16[k] = k < 12 ? acc[k] : 0.0f;
        const float mine = warp_sum16(v16, lane);
        const int slot = warp_slot(lane);
        if ((lane & 1) == 0 && slot < 12) p.pose_partials[(((size_t)vbid * kWarps + wid) * p.S + j) * 12 + slot] = mine;
      }
      __syncwarp();
      // pick up the depth gradients of this lane's own pixels (pairs outside P / the image hold garbage that is never stored)
#pragma unroll
      for (int o = 0; o < kRowsPerWarp; ++o) gd[o] = gd[o] + ld2(planes + bS * kPlane + plane_index(r0 + 1 + o, c0 + 1));
    }
  }

  pdl_launch_dependents();   // see mono_fwd.cu
#ifdef SDE_BWD_EARLY_TICKET
  // publish the pose slots and take the sample's ticket BEFORE the gradient stores: the fence then waits for the few
  // slot stores only, not for this warp's share of the gradient planes
  if ((lane & 1) == 0 && warp_slot(lane) < 12) publish_fence();
  __syncthreads();
  if (tid == 0) sh.ticket = atomicAdd(p.smp_counter + b, 1u);
  __syncthreads();
#endif
  // ------------------------------------------------------------------ smoothness gradient + store
  {
    const float sscale = p.smooth_scale[s];
    const float mbar = ld_prod(p.stats + (s * p.B + b) * 2), Lb = ld_prod(p.stats + (s * p.B + b) * 2 + 1);
    const float inx = 1.0f / ((float)p.NB * (float)h * (float)(w - 1));
    const float iny = 1.0f / ((float)p.NB * (float)(h - 1) * (float)w);
    const float homog = mbar > 1e-6f ? Lb / ((float)h * (float)w * mbar) : 0.0f;
    const float rmbar = 1.0f / mbar;
    const float gsm = g_smooth * sscale;
    float* __restrict__ gout = p.grad_depth[s] + (size_t)b * hw;
    // the pair (gxp, gxp + 1) is 8-byte aligned in global memory (gxp is even) if the rows and the buffers are
    const bool even = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(gout) & 7) == 0;
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) {
      const int row = r0 + 1 + o;
      const int gy = oy + row;
      if (row >= 2 && row <= kBwdH + 1 && gy < h && col_ok0) {
        float g0 = lo(gd[o]), g1 = hi(gd[o]);
        if (sscale > 0.0f) {
          const int pl = plane_index(row, c0 + 1);
          const float* pd = planes + kBD * kPlane + pl;
          const float d0 = pd[0], d1 = pd[1];
          const float ic0 = inv_depth(d0), ic1 = inv_depth(d1);
          float G0, G1;
          if (SAVED) {
            G0 = lo(Gs[o]); G1 = hi(Gs[o]);
          } else {
            G0 = smooth_grad_local(pd, planes + kBA * kPlane + pl, ic0, gxp, gy, w, h, inx, iny);
            G1 = smooth_grad_local(pd + 1, planes + kBA * kPlane + pl + 1, ic1, gxp + 1, gy, w, h, inx, iny);
          }
          if (d0 >= 1e-6f) g0 += -ic0 * ic0 * (G0 * rmbar - homog) * gsm;
          if (d1 >= 1e-6f) g1 += -ic1 * ic1 * (G1 * rmbar - homog) * gsm;
        }
        if (p.depth_mode != SDE_DEPTH_IS_DEPTH) {
          // chain rule through disp_to_depth (and softplus): the gradient leaves w.r.t. what depth[] holds
          const int pl = plane_index(row, c0 + 1);
          const float* praw = depth + gy * w + gxp;
          const float raw0 = p.depth_mode == SDE_DEPTH_IS_LOGIT ? __ldg(praw) : 0.0f;
          const float raw1 = (p.depth_mode == SDE_DEPTH_IS_LOGIT && col_ok1) ? __ldg(praw + 1) : 0.0f;
          g0 *= decode_depth_grad(planes[kBD * kPlane + pl], raw0, p.depth_mode, p.disp_range);
          g1 *= decode_depth_grad(planes[kBD * kPlane + pl + 1], raw1, p.depth_mode, p.disp_range);
        }
        SDE_CHECK(gy >= 0 && gy < h && gxp >= 0 && gxp < w && (!col_ok1 || gxp + 1 < w));
        float* po = gout + gy * w + gxp;
        if (even && col_ok1) {
          *reinterpret_cast<float2*>(po) = make_float2(g0, g1);
        } else {
          po[0] = g0;
          if (col_ok1) po[1] = g1;
        }
      }
    }
  }

  // ------------------------------------------------------------------ last tile of a sample: pose gradients
  // The tile that finishes a sample last (over all scales) adds that sample's per-CTA slots in a fixed
  // order in fp64, while other samples are still being computed.  Only the lanes that wrote pose slots fence.
  SDE_TRACE_MARK(p, 2, 1);
  int total_b = 0;
  for (int ss = 0; ss < p.n_scales; ++ss) total_b += p.btiles_x[ss] * p.btiles_y[ss];
#ifndef SDE_BWD_EARLY_TICKET
  if ((lane & 1) == 0 && warp_slot(lane) < 12) publish_fence();
  __syncthreads();
  if (tid == 0) sh.ticket = atomicAdd(p.smp_counter + b, 1u);
  __syncthreads();
#endif
  if (sh.ticket != (unsigned)(total_b - 1)) return;
  publish_fence();
  if (flow_i) {
    // every tile of this sample has passed its image flag: leave the flags cleared for the next call.  This grid did
    // not wait for the forward kernel as a grid; the CTAs that finish a sample do (it has long finished), so that the
    // completion of this grid implies the completion of the grids before it.
    if (tid < p.n_scales) p.img_flag[tid * p.B + b] = 0u;
    pdl_wait();
  }
  for (int tj = 0; tj < p.S; ++tj) {
    double a[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) a[k] = 0.0;
    for (int ss = 0; ss < p.n_scales; ++ss) {
      const int per = p.btiles_x[ss] * p.btiles_y[ss] * kWarps;   // one slot per warp of every tile of the sample
      const size_t first = ((size_t)p.btile_start[ss] + (size_t)b * p.btiles_x[ss] * p.btiles_y[ss]) * kWarps;
      double q[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) q[k] = 0.0;
      for (int t = tid; t < per; t += kThreads) {
        const float4* part = reinterpret_cast<const float4*>(p.pose_partials + ((first + t) * p.S + tj) * 12);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 v = __ldcg(part + k);
          q[4 * k] += (double)v.x; q[4 * k + 1] += (double)v.y; q[4 * k + 2] += (double)v.z; q[4 * k + 3] += (double)v.w;
        }
      }
      if (SAVED) {
        // the slots hold, per i, the sums of a_i depth x', a_i depth y', a_i depth, a_i over the scale's pixels (x', y'
        // centred on (w/2, h/2); i = 2 with the opposite sign): the linear map to d loss / d [R | t] uses the
        // intrinsics of this scale (applied to this thread's partial sums; the map is linear)
        Cam cam;
        float kk[9];
        load_cam(cam, kk, p.K, b, p.sx[ss], p.sy[ss]);
        const double x0 = (double)(p.w[ss] >> 1), y0 = (double)(p.h[ss] >> 1);
        double A[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const double sb = q[4 * i + 2], sx = q[4 * i] + x0 * sb, sy = q[4 * i + 1] + y0 * sb;
#pragma unroll
          for (int c = 0; c < 3; ++c) A[i][c] = (double)cam.ki[3 * c] * sx + (double)cam.ki[3 * c + 1] * sy + (double)cam.ki[3 * c + 2] * sb;
          A[i][3] = q[4 * i + 3];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          a[c] += (double)cam.fx * A[0][c];
          a[4 + c] += (double)cam.sk * A[0][c] + (double)cam.fy * A[1][c];
          a[8 + c] -= A[2][c];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] += q[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) sh.dred[k][wid] = a[k];
    }
    __syncthreads();
    if (tid < 12) {
      float* gp = p.grad_pose[tj] + b * 16;
      gp[tid] = (float)(((sh.dred[tid][0] + sh.dred[tid][1]) + sh.dred[tid][2]) + sh.dred[tid][3]);   // rows 0..2 = [dR | dt]
      if (tid < 4) gp[12 + tid] = 0.0f;
    }
  }
  if (tid == 0) p.smp_counter[b] = 0u;   // leave the workspace zeroed for the next call
}

size_t mono_bwd_smem_bytes() { return (size_t)kBwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_mono_bwd_pair(const MonoParams& p, const MonoTma& t, cudaStream_t stream);   // mono_bwd_pair.cu
bool bwd_pair_enabled();                                                                          // sde_api.cu

cudaError_t launch_mono_bwd(const MonoParams& p, const MonoTma& t, cudaStream_t stream) {
  const bool saved = p.warped[0][0] != nullptr;   // all-or-nothing, checked by the caller
  // min-reprojection over an even number of sources with kept warps: two sources per pass (mono_bwd_pair.cu)
  if (saved && (p.S & 1) == 0 && !(p.flags & SDE_MONO_REDUCE_MEAN) && bwd_pair_enabled()) return launch_mono_bwd_pair(p, t, stream);
  auto kernel = saved ? mono_bwd_kernel<true> : mono_bwd_kernel<false>;
  // 47.8 KB of dynamic shared memory (+ 4 KB static): four CTAs per SM; the opt-in attribute is per device, cheap and idempotent
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mono_bwd_smem_bytes());
  if (e != cudaSuccess) return e;
  return launch_chained(2, kernel, (unsigned)p.btile_start[p.n_scales], kThreads, mono_bwd_smem_bytes(), stream, p, t);
}

}  // namespace sde
