// Fused backward of the MonoDepth2 loss for min-reprojection with an EVEN number of sources and kept warps: two
// sources per pass (sm_100a).  Same tiles, same outputs and same reductions as mono_bwd_kernel<true> (mono_bwd.cu, which
// keeps serving 'mean' reduction, odd source counts and the recompute mode); what differs is the work per tile:
//
//   * torch.min selects ONE candidate per window (MonoDepth2.py:116-119), so at most one of the two sources of a pair
//     receives a window's SSIM gradient.  Both sources' warped planes are resident (14 planes, three CTAs per SM); the
//     coefficient pass forms the window sums of both, picks per window the sums of the selected source and evaluates
//     d ssim_q / d S_p = a_q + S_p b_q + A_p c_q ONCE -- the division-heavy part of the pass -- instead of once per
//     source with half of the results masked away.  The target-side sums (A, A^2) are formed once per pair.
//   * the coefficients go to ONE set of planes plus a 0/1 plane m_q (1: the first source of the pair was selected).
//     The adjoint pass forms, per coefficient plane, the box sums T of v_q and T0 of m_q v_q from one set of loads and
//     T1 = T - T0 (the coefficients are zero where neither source was selected): gS^0_c = T0a + S^0_p T0b + A_p T0c,
//     gS^1_c likewise with T1.
//   * per colour channel two CTA barriers cover both sources (six per tile instead of twelve), and every pass carries
//     two independent dependency chains per lane.
// Instructions per cfg2 launch 63.6 M -> see profiles/r2_notes.md.  Phase 4 (dense multiply-add warp backward), the
// smoothness tail and the fixed-order pose reduction are those of mono_bwd.cu.
#include <type_traits>

#include "mono_device.cuh"

namespace sde {

constexpr int kPairPlanes = 14;
constexpr int kPA = 0, kPD = 3, kPS0 = 4, kPS1 = 7, kPC = 10, kPM = 13;
constexpr int kPairColOff = 3;   // as in mono_bwd.cu: plane index 0 = image column tile_x0 - 4 (TMA box alignment)
constexpr int kPairWarps = kThreads / 32;

struct PairShared {
  float wmat[SDE_MAX_SOURCES][9];    // see mono_bwd.cu: d depth-gradient / d (a0, a1, a2), affine in the centred pixel coordinates
  float nmat[SDE_MAX_SOURCES][14];   // the projection p_i = depth (N_i . [x', y', 1]) + tau_i, and (cx, cy)
  double dred[12][kPairWarps];
  unsigned ticket;
  __align__(8) uint64_t bar;
  __align__(8) uint8_t arg[kPlane];
};

#ifndef SDE_PAIR_OCC
#define SDE_PAIR_OCC 3
#endif
#ifndef SDE_PAIR_UNROLL
#define SDE_PAIR_UNROLL 1
#endif
#define SDE_PRAGMA_(x) _Pragma(#x)
#define SDE_PRAGMA_UNROLL(n) SDE_PRAGMA_(unroll n)

__global__ void __launch_bounds__(kThreads, SDE_PAIR_OCC) mono_bwd_pair_kernel(const __grid_constant__ MonoParams p,
                                                                              const __grid_constant__ MonoTma maps) {
  extern __shared__ __align__(128) float planes[];  // [kPairPlanes][kPlane]
  __shared__ PairShared sh;
  __shared__ int next_item;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool automask = (p.flags & SDE_MONO_AUTOMASK) != 0;
  const bool use_ssim = p.ssim_w > 0.0f;
  const bool cam_thread = tid >= 32 && tid < 32 + p.S;
  // coefficient and mask planes: the coefficient pass writes every window centre, the ring around the block reads as
  // zero (no tile ever writes it: once per CTA)
  zero_ring<4>(planes, kPC, tid);
  if (tid == 0) {
    mbar_init(&sh.bar, 1);
    mbar_init_fence();
  }
  const bool flow_i = (p.flow & kFlowImage) != 0;   // see mono_bwd.cu
  SDE_TRACE_BEGIN(p, 2);
  if (!flow_i) pdl_wait();
  const int total = p.btile_start[p.n_scales];
  int total_b = 0;   // tiles of one sample, all scales
  for (int ss = 0; ss < p.n_scales; ++ss) total_b += p.btiles_x[ss] * p.btiles_y[ss];
  int vbid = 0;
  bool first = true;
  unsigned tma_phase = 0;   // parity of the TMA barrier: carried from tile to tile
  // work items: gradient tiles in list order (mono_params.cuh: work_next)
  while (work_next(p, 2, total, &next_item, vbid, first)) {
  const TileCoord tc = decode_btile(p, vbid);
  const int s = tc.s, b = tc.b, h = p.h[s], w = p.w[s], hw = h * w;
  // plane (yy, xx) <-> image (oy + yy, ox + xx); Q = plane [1..16]x[1..64]; P = plane [2..15]x[3..62]
  const int ox = tc.x0 - kPairColOff, oy = tc.y0 - 2;
  const bool interior = ox >= 0 && oy >= 0 && ox + kHW <= w && oy + kHH <= h;
  const bool lr_border = ox + 2 <= 1 || ox + kHW - 3 >= w - 2;
  const bool tb_border = oy + 2 <= 1 || oy + kBwdH + 1 >= h - 2;   // a row of P is image row 1 or h - 2 (doubled pad row)
  const bool tma = p.tma[s] != 0;
  if (flow_i && tid == 0) {
    flag_wait(p.img_flag + s * p.B + b);
    flag_proxy_fence();
  }
  if (tma && tid == 0 && ((p.persist >> 2) & 1)) proxy_fence();   // the planes were last written by this CTA's own stores
  CamRaw raw;
  if (cam_thread) load_cam_raw(raw, p.K, p.pose[tid - 32], b);
  auto load_pair = [&](int j0, bool first) {   // thread 0: target + depth (first pair only) and the pair's warped planes
    mbar_arrive_expect_tx(&sh.bar, (first ? 10 : 6) * kPlaneBytesTma);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (first) tma_load_plane(planes + (kPA + c) * kPlane, &maps.target[s], &sh.bar, ox - kColOff, oy, b * 3 + c);
      tma_load_plane(planes + (kPS0 + c) * kPlane, &maps.warped[s][j0], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
      tma_load_plane(planes + (kPS1 + c) * kPlane, &maps.warped[s][j0 + 1], &sh.bar, ox - kColOff, oy, b * kSavedPlanes + c);
    }
    if (first) tma_load_plane(planes + kPD * kPlane, &maps.depth[s], &sh.bar, ox - kColOff, oy, b);
  };
  if (tma && tid == 0) load_pair(0, true);
  __syncthreads();   // (flow: thread 0 arrives here after the image's flag)
  SDE_TRACE_MARK(p, 2, 3);
  const unsigned behind_flag = launder_zero();   // see ld_plane (sde_common.cuh)

  const float* __restrict__ depth = p.depth[s] + (size_t)b * hw;
  const float* __restrict__ tg0 = p.target[s] + (size_t)b * 3 * hw;
  const uint8_t* __restrict__ amap = p.argmin[s] + (size_t)b * hw;

  const float g_rec = __ldg(p.grad_losses), g_smooth = __ldg(p.grad_losses + 1);
  const float g_pe = g_rec * p.inv_norm[s];
  const float g_l1 = g_pe * p.l1_w * (1.0f / 3.0f);
  const float g_ss = g_pe * p.ssim_w * (1.0f / 3.0f) * -0.5f;  // d pe / d ssim (ssim_loss.py:53)

  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;
  const f2 C1 = bc2(81.0f * p.c1), C2 = bc2(81.0f * p.c2);

  const int px0 = ox + c0 + 1, px1 = px0 + 1;
  const float eL0 = px0 == 1 ? 1.0f : 0.0f, eL1 = px1 == 1 ? 1.0f : 0.0f;
  const float eR0 = px0 == w - 2 ? 1.0f : 0.0f, eR1 = px1 == w - 2 ? 1.0f : 0.0f;

  f2 gd[kRowsPerWarp], Gs[kRowsPerWarp];
#pragma unroll
  for (int k = 0; k < kRowsPerWarp; ++k) gd[k] = Gs[k] = bc2(0.0f);
  const int gxp = ox + c0 + 1;
  const bool colP = lane >= 1 && lane <= 30;
  const bool col_ok0 = colP && gxp < w, col_ok1 = colP && gxp + 1 < w;

  StageArgs sa;
  sa.depth_mode = p.depth_mode; sa.min_disp = p.min_disp; sa.disp_range = p.disp_range;
  sa.depth = depth; sa.src = nullptr; sa.tgt = tg0; sa.amap = amap;
  sa.planes = planes; sa.arg = sh.arg; sa.oy = oy; sa.ox = ox; sa.h = h; sa.w = w; sa.hw = hw;
  sa.plS = kPS0; sa.plI = 0; sa.plA = kPA; sa.plD = kPD;
  // ------------------------------------------------------------------ phase 0: argmin bytes (+ depth, target without TMA)
  if (tma) {
    if ((reinterpret_cast<uintptr_t>(amap) & 3) == 0) stage_arg_words(sh.arg, amap, oy, ox, h, w, false, tid);
    else stage_arg(sh.arg, amap, oy, ox, h, w, false, tid);
  } else {
    if (interior) stage_target<true, true>(sa, tid, false);
    else          stage_target<false, true>(sa, tid, false);
  }
  if (cam_thread) {
    // camera terms of phase 4 (mono_bwd.cu: make_camera)
    Cam cam;
    float k[9];
    load_cam(cam, k, raw.k, 0, p.sx[s], p.sy[s]);
    Proj pj;
    load_proj(pj, k, raw.T, 0);
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
  __syncthreads();

  for (int j0 = 0; j0 < p.S; j0 += 2) {
    // bytes of the argmin plane that select the pair's warped candidates
    const int cand0 = automask ? 2 * j0 : j0, cand1 = automask ? 2 * j0 + 2 : j0 + 1;
    // ---------------------------------------------------------------- phase 1: the pair's warped planes
    if (j0 > 0) {
      __syncthreads();   // every warp is done with the previous pair's gS planes
      if (tma && tid == 0) {
        proxy_fence();
        load_pair(j0, false);
      }
    }
    if (tma) {
      mbar_wait(&sh.bar, tma_phase);
      tma_phase ^= 1u;
      if (j0 == 0 && p.depth_mode != SDE_DEPTH_IS_DEPTH) {
        decode_depth_plane(planes + kPD * kPlane, p.depth_mode, p.min_disp, p.disp_range, tid);
        if (!interior) __syncthreads();   // the fix-up copies decoded values
      }
      if (!interior) {
        if (j0 == 0) reflect_fixup(planes, kPA, 3, oy, ox, h, w, tid), reflect_fixup(planes, kPD, 1, oy, ox, h, w, tid);
        reflect_fixup(planes, kPS0, 6, oy, ox, h, w, tid);
      }
    } else {
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        sa.plS = jj ? kPS1 : kPS0;
        const float* wsrc = p.warped[s][j0 + jj] + (size_t)b * kSavedPlanes * hw;
        if (interior) stage_saved<true>(sa, wsrc, tid);
        else          stage_saved<false>(sa, wsrc, tid);
      }
    }
    __syncthreads();

#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      // -------------------------------------------------------------- phase 2: SSIM coefficients of the selected source on Q
      if (use_ssim) {
        const float* pa = planes + (kPA + c) * kPlane + plane_index(r0, c0);
        // horizontal 3-sums of a row: target side (A, A^2) and, per source, (X, X^2, X A)
        struct HRow { f2 A, AA, X[2], XX[2], XA[2]; };
        auto hsum = [&](int rr) {
          HRow n;
          const Row4 a = ld_row(pa + rr * kPitch);
          const f2 aa = a.c * a.c;
          n.A = (a.c + swp(a.c)) + a.o;
          n.AA = fma2(a.o, a.o, aa + swp(aa));
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const Row4 x = ld_row(pa + ((k ? kPS1 : kPS0) - kPA) * kPlane + rr * kPitch);
            const f2 xx2 = x.c * x.c, xa = x.c * a.c;
            n.X[k] = (x.c + swp(x.c)) + x.o;
            n.XX[k] = fma2(x.o, x.o, xx2 + swp(xx2));
            n.XA[k] = fma2(x.o, a.o, xa + swp(xa));
          }
          return n;
        };
        // coefficients of the windows centred on plane row `row` from pp = the 3-sums of the two rows above it, added up,
        // and n = the 3-sums of the row below it
        auto coef_row = [&](int row, const HRow& pp, const HRow& n) {
          const uchar2 m = *reinterpret_cast<const uchar2*>(sh.arg + plane_index(row, c0 + 1));
          const bool f0 = m.x == cand0, s0 = m.x == cand1, f1 = m.y == cand0, s1 = m.y == cand1;
          f2 ca = bc2(0.0f), cb = bc2(0.0f), cc = bc2(0.0f);
#ifdef SDE_PAIR_SKIP
          if (__any_sync(0xffffffffu, f0 || s0 || f1 || s1))   // skip rows of 64 unselected windows: rare, and the branch
#endif                                                         // keeps the scheduler from interleaving the two rows of a trip
          {
            const f2 sA = pp.A + n.A, sAA = pp.AA + n.AA;
            const f2 sXa = pp.X[0] + n.X[0], sXXa = pp.XX[0] + n.XX[0], sXAa = pp.XA[0] + n.XA[0];
            const f2 sXb = pp.X[1] + n.X[1], sXXb = pp.XX[1] + n.XX[1], sXAb = pp.XA[1] + n.XA[1];
            // the window sums of the source this window selected (either, where it selected neither: its coefficients are zeroed)
            const f2 sX = mk2(s0 ? lo(sXb) : lo(sXa), s1 ? hi(sXb) : hi(sXa));
            const f2 sXX = mk2(s0 ? lo(sXXb) : lo(sXXa), s1 ? hi(sXXb) : hi(sXXa));
            const f2 sXA = mk2(s0 ? lo(sXAb) : lo(sXAa), s1 ? hi(sXAb) : hi(sXAa));
            // same operation order as the forward kernel
            const f2 aa2 = sA * sA, xs = sX * sX, t = sX * sA;
            const f2 vA = fma2(aa2, bc2(-1.0f), sAA * bc2(9.0f));
            const f2 n1 = fma2(bc2(2.0f), t, C1);
            const f2 n2 = fma2(bc2(2.0f), fma2(t, bc2(-1.0f), sXA * bc2(9.0f)), C2);
            const f2 d1 = (xs + aa2) + C1;
            const f2 d2 = (fma2(xs, bc2(-1.0f), sXX * bc2(9.0f)) + vA) + C2;
            const f2 N = n1 * n2, D = d1 * d2;
            const f2 nssim = ndiv2(N, D);
            const f2 ninvD = ndiv2(bc2(1.0f), D);
            // torch.clamp passes the gradient on the closed interval 0 <= (1-ssim)/2 <= 1
            const float h0 = fmaf(lo(nssim), 0.5f, 0.5f), h1 = fmaf(hi(nssim), 0.5f, 0.5f);
            const f2 g = mk2(((f0 || s0) && h0 >= 0.0f && h0 <= 1.0f) ? g_ss : 0.0f,
                             ((f1 || s1) && h1 >= 0.0f && h1 <= 1.0f) ? g_ss : 0.0f);
            const f2 ngi = g * ninvD;
            const f2 uu = fma2(sX * nssim, d2 - d1, sA * (n2 - n1));
            ca = (ngi * bc2(-2.0f)) * uu;
            cb = (ngi * bc2(-18.0f)) * (nssim * d1);
            cc = (ngi * bc2(-18.0f)) * n1;
          }
          float* pc = planes + kPC * kPlane + plane_index(row, c0 + 1);
          *reinterpret_cast<unsigned long long*>(pc) = ca.v;
          *reinterpret_cast<unsigned long long*>(pc + kPlane) = cb.v;
          *reinterpret_cast<unsigned long long*>(pc + 2 * kPlane) = cc.v;
          // which source the window selected does not depend on the channel
          if (c == 0) *reinterpret_cast<unsigned long long*>(pc + 3 * kPlane) = mk2(f0 ? 1.0f : 0.0f, f1 ? 1.0f : 0.0f).v;
        };
        auto add_rows = [](const HRow& u, const HRow& v) {
          HRow r;
          r.A = u.A + v.A; r.AA = u.AA + v.AA;
#pragma unroll
          for (int k = 0; k < 2; ++k) { r.X[k] = u.X[k] + v.X[k]; r.XX[k] = u.XX[k] + v.XX[k]; r.XA[k] = u.XA[k] + v.XA[k]; }
          return r;
        };
        // Sliding window over the rows with two row slots and the running pair sum pp = (row r-2) + (row r-1): the new
        // row replaces the slot of row r-2 (already folded into pp), the window sum is pp + new -- the association
        // (r-2 + r-1) + r of the forward kernel -- and pp becomes (row r-1) + new.  The slots swap roles after one row
        // and are back after two, so a loop over row PAIRS carries no register moves and can stay rolled
        // (SDE_PAIR_UNROLL 1): a third of the unrolled code -- three CTAs in different phases share the instruction cache.
        HRow hx = hsum(0), hy = hsum(1);
        HRow pp = add_rows(hx, hy);
SDE_PRAGMA_UNROLL(SDE_PAIR_UNROLL)
        for (int it = 0; it < kRowsPerWarp / 2; ++it) {
          const int rr = 2 + 2 * it;
          hx = hsum(rr);
          coef_row(r0 + rr - 1, pp, hx);
          pp = add_rows(hy, hx);
          hy = hsum(rr + 1);
          coef_row(r0 + rr, pp, hy);
          pp = add_rows(hx, hy);
        }
      }
      __syncthreads();
      // -------------------------------------------------------------- phase 3: adjoint gather -> gS^0_c, gS^1_c on P
      auto phase3 = [&](auto lr_tag, auto ssim_tag) {
        constexpr bool LR = decltype(lr_tag)::value;
        constexpr bool SSIM = decltype(ssim_tag)::value;
        f2 hq[3][2], hq0[3][2];  // horizontal 3-sums of a, b, c (all windows / windows of the first source), two previous rows
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
          f2 nq[3], nq0[3];
          if (SSIM) {
            const Row4 qm = ld_row(planes + kPM * kPlane + plane_index(r0 + rr, c0));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              const Row4 q = ld_row(planes + (kPC + k) * kPlane + plane_index(r0 + rr, c0));
              const f2 mc = q.c * qm.c, mo = q.o * qm.o;
              f2 hs = (q.c + swp(q.c)) + q.o, hs0 = (mc + swp(mc)) + mo;
              if (LR) {
                hs = hs + mk2(eL0 * lo(q.o) + eR0 * hi(q.c), eL1 * lo(q.c) + eR1 * hi(q.o));
                hs0 = hs0 + mk2(eL0 * lo(mo) + eR0 * hi(mc), eL1 * lo(mc) + eR1 * hi(mo));
              }
              nq[k] = hs; nq0[k] = hs0;
            }
          }
          if (rr >= 2) {
            const int row = r0 + rr - 1;        // plane row of pixel p
            const int py = oy + row;
            const f2 wu = bc2(py == 1 ? 2.0f : 1.0f), wd = bc2(py == h - 2 ? 2.0f : 1.0f);
            const int pl = plane_index(row, c0 + 1);
            const f2 Ap = ld2(planes + (kPA + c) * kPlane + pl);
            const f2 Sp0 = ld2(planes + (kPS0 + c) * kPlane + pl), Sp1 = ld2(planes + (kPS1 + c) * kPlane + pl);
            f2 gS0 = bc2(0.0f), gS1 = bc2(0.0f);
            if (SSIM) {
              const f2 va = fma2(wu, hq[0][0], fma2(wd, nq[0], hq[0][1])), va0 = fma2(wu, hq0[0][0], fma2(wd, nq0[0], hq0[0][1]));
              const f2 vb = fma2(wu, hq[1][0], fma2(wd, nq[1], hq[1][1])), vb0 = fma2(wu, hq0[1][0], fma2(wd, nq0[1], hq0[1][1]));
              const f2 vc = fma2(wu, hq[2][0], fma2(wd, nq[2], hq[2][1])), vc0 = fma2(wu, hq0[2][0], fma2(wd, nq0[2], hq0[2][1]));
              gS0 = fma2(Sp0, vb0, fma2(Ap, vc0, va0));
              gS1 = fma2(Sp1, vb - vb0, fma2(Ap, vc - vc0, va - va0));
            }
            // L1 term on the pixel itself: g_l1 * sign(S - A) where this candidate was selected (branch-free, mono_bwd.cu)
            const uchar2 m = *reinterpret_cast<const uchar2*>(sh.arg + pl);
            auto l1 = [&](f2 Sp, int cand) {
              const f2 df = Sp - Ap;
              const float d0 = lo(df), d1 = hi(df);
              float l0 = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d0) & 0x80000000));
              float l1v = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d1) & 0x80000000));
              l0 = (m.x == cand && d0 != 0.0f) ? l0 : 0.0f;
              l1v = (m.y == cand && d1 != 0.0f) ? l1v : 0.0f;
              return mk2(l0, l1v);
            };
            gS0 = gS0 + l1(Sp0, cand0);
            gS1 = gS1 + l1(Sp1, cand1);
            *reinterpret_cast<unsigned long long*>(planes + (kPS0 + c) * kPlane + pl) = gS0.v;   // in place of S_c
            *reinterpret_cast<unsigned long long*>(planes + (kPS1 + c) * kPlane + pl) = gS1.v;
          }
          if (SSIM) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { hq[k][0] = hq[k][1]; hq[k][1] = nq[k]; hq0[k][0] = hq0[k][1]; hq0[k][1] = nq0[k]; }
          }
        }
      };
      // Tiles away from the image border (no mirrored pad column, no doubled pad row: every weight of the adjoint is 1)
      // take the same pass as a rolled loop over row pairs with running pair sums, as the coefficient pass above.
      auto phase3_interior = [&]() {
        struct QRow { f2 t[3], t0[3]; };   // horizontal 3-sums of a, b, c: all windows / windows of the first source
        auto qsum = [&](int rr) {
          QRow n;
          const Row4 qm = ld_row(planes + kPM * kPlane + plane_index(r0 + rr, c0));
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const Row4 q = ld_row(planes + (kPC + k) * kPlane + plane_index(r0 + rr, c0));
            const f2 mc = q.c * qm.c, mo = q.o * qm.o;
            n.t[k] = (q.c + swp(q.c)) + q.o;
            n.t0[k] = (mc + swp(mc)) + mo;
          }
          return n;
        };
        auto add_q = [](const QRow& u, const QRow& v) {
          QRow r;
#pragma unroll
          for (int k = 0; k < 3; ++k) { r.t[k] = u.t[k] + v.t[k]; r.t0[k] = u.t0[k] + v.t0[k]; }
          return r;
        };
        auto out_row = [&](int row, const QRow& pp, const QRow& n) {
          const int pl = plane_index(row, c0 + 1);
          const f2 Ap = ld2(planes + (kPA + c) * kPlane + pl);
          const f2 Sp0 = ld2(planes + (kPS0 + c) * kPlane + pl), Sp1 = ld2(planes + (kPS1 + c) * kPlane + pl);
          const f2 va = pp.t[0] + n.t[0], vb = pp.t[1] + n.t[1], vc = pp.t[2] + n.t[2];
          const f2 va0 = pp.t0[0] + n.t0[0], vb0 = pp.t0[1] + n.t0[1], vc0 = pp.t0[2] + n.t0[2];
          f2 gS0 = fma2(Sp0, vb0, fma2(Ap, vc0, va0));
          f2 gS1 = fma2(Sp1, vb - vb0, fma2(Ap, vc - vc0, va - va0));
          const uchar2 m = *reinterpret_cast<const uchar2*>(sh.arg + pl);
          auto l1 = [&](f2 Sp, int cand) {
            const f2 df = Sp - Ap;
            const float d0 = lo(df), d1 = hi(df);
            float l0 = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d0) & 0x80000000));
            float l1v = __int_as_float(__float_as_int(g_l1) ^ (__float_as_int(d1) & 0x80000000));
            l0 = (m.x == cand && d0 != 0.0f) ? l0 : 0.0f;
            l1v = (m.y == cand && d1 != 0.0f) ? l1v : 0.0f;
            return mk2(l0, l1v);
          };
          gS0 = gS0 + l1(Sp0, cand0);
          gS1 = gS1 + l1(Sp1, cand1);
          *reinterpret_cast<unsigned long long*>(planes + (kPS0 + c) * kPlane + pl) = gS0.v;   // in place of S_c
          *reinterpret_cast<unsigned long long*>(planes + (kPS1 + c) * kPlane + pl) = gS1.v;
        };
        QRow qx = qsum(0), qy = qsum(1);
        QRow pp = add_q(qx, qy);
SDE_PRAGMA_UNROLL(SDE_PAIR_UNROLL)
        for (int it = 0; it < kRowsPerWarp / 2; ++it) {
          const int rr = 2 + 2 * it;
          qx = qsum(rr);
          out_row(r0 + rr - 1, pp, qx);
          pp = add_q(qy, qx);
          qy = qsum(rr + 1);
          out_row(r0 + rr, pp, qy);
          pp = add_q(qx, qy);
        }
      };
      if (use_ssim) {
        if (lr_border || tb_border) phase3(std::true_type{}, std::true_type{});
        else                        phase3_interior();
      } else {
        phase3(std::false_type{}, std::false_type{});
      }
      __syncthreads();
    }

    // ---------------------------------------------------------------- phase 4: warp backward on P, both sources
    if (j0 + 2 >= p.S && p.smooth_scale[s] > 0.0f) {
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
#pragma unroll 1
    for (int jj = 0; jj < 2; ++jj) {
      // dense stream of packed multiply-adds from the derivative planes of the warp kernel (mono_bwd.cu, phase 4)
      const int j = j0 + jj, bS = jj ? kPS1 : kPS0;
      const float* __restrict__ dwp = p.warped[s][j] + (size_t)b * kSavedPlanes * hw + behind_flag;
      const float* wm = sh.wmat[j];
      const float* nm = sh.nmat[j];
      const bool pair = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(dwp) & 7) == 0;   // gxp is even
      const float xc0 = (float)(gxp - (w >> 1));
      const f2 xc = mk2(xc0, xc0 + 1.0f);
      f2 Sa[3], Sb[3], Sy[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) Sa[k] = Sb[k] = Sy[k] = bc2(0.0f);
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
          const f2 dd = ld2(planes + kPD * kPlane + pl);
          const f2 a0 = fma2(g2, dq[o][2], fma2(g1, dq[o][1], g0 * dq[o][0]));
          const f2 a1 = fma2(g2, dq[o][5], fma2(g1, dq[o][4], g0 * dq[o][3]));
          const float yc = (float)(gy - (h >> 1));
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
        if ((lane & 1) == 0 && slot < 12) p.pose_partials[(((size_t)vbid * kPairWarps + wid) * p.S + j) * 12 + slot] = mine;
      }
    }
  }

  if (!((p.persist >> 2) & 1)) pdl_launch_dependents();
  // ------------------------------------------------------------------ smoothness gradient + store (mono_bwd.cu)
  {
    const float sscale = p.smooth_scale[s];
    const float mbar = ld_prod(p.stats + (s * p.B + b) * 2), Lb = ld_prod(p.stats + (s * p.B + b) * 2 + 1);
    const float homog = mbar > 1e-6f ? Lb / ((float)h * (float)w * mbar) : 0.0f;
    const float rmbar = 1.0f / mbar;
    const float gsm = g_smooth * sscale;
    float* __restrict__ gout = p.grad_depth[s] + (size_t)b * hw;
    const bool even = (w & 1) == 0 && (reinterpret_cast<uintptr_t>(gout) & 7) == 0;
#pragma unroll
    for (int o = 0; o < kRowsPerWarp; ++o) {
      const int row = r0 + 1 + o;
      const int gy = oy + row;
      if (row >= 2 && row <= kBwdH + 1 && gy < h && col_ok0) {
        float g0 = lo(gd[o]), g1 = hi(gd[o]);
        const int pl = plane_index(row, c0 + 1);
        if (sscale > 0.0f) {
          const float* pd = planes + kPD * kPlane + pl;
          const float d0 = pd[0], d1 = pd[1];
          const float ic0 = inv_depth(d0), ic1 = inv_depth(d1);
          if (d0 >= 1e-6f) g0 += -ic0 * ic0 * (lo(Gs[o]) * rmbar - homog) * gsm;
          if (d1 >= 1e-6f) g1 += -ic1 * ic1 * (hi(Gs[o]) * rmbar - homog) * gsm;
        }
        if (p.depth_mode != SDE_DEPTH_IS_DEPTH) {
          // chain rule through disp_to_depth (and softplus): the gradient leaves w.r.t. what depth[] holds
          const float* praw = depth + gy * w + gxp;
          const float raw0 = p.depth_mode == SDE_DEPTH_IS_LOGIT ? __ldg(praw) : 0.0f;
          const float raw1 = (p.depth_mode == SDE_DEPTH_IS_LOGIT && col_ok1) ? __ldg(praw + 1) : 0.0f;
          g0 *= decode_depth_grad(planes[kPD * kPlane + pl], raw0, p.depth_mode, p.disp_range);
          g1 *= decode_depth_grad(planes[kPD * kPlane + pl + 1], raw1, p.depth_mode, p.disp_range);
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

  // ------------------------------------------------------------------ last tile of a sample: pose gradients (mono_bwd.cu)
  SDE_TRACE_MARK(p, 2, 1);
  if ((lane & 1) == 0 && warp_slot(lane) < 12) publish_fence();
  __syncthreads();
  if (tid == 0) sh.ticket = atomicAdd(p.smp_counter + b, 1u);
  __syncthreads();
  if (sh.ticket != (unsigned)(total_b - 1)) continue;
  publish_fence();
  if (flow_i) {
    if (tid < p.n_scales) p.img_flag[tid * p.B + b] = 0u;
    pdl_wait();
  }
  for (int tj = 0; tj < p.S; ++tj) {
    double a[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) a[k] = 0.0;
    for (int ss = 0; ss < p.n_scales; ++ss) {
      const int per = p.btiles_x[ss] * p.btiles_y[ss] * kPairWarps;
      const size_t first = ((size_t)p.btile_start[ss] + (size_t)b * p.btiles_x[ss] * p.btiles_y[ss]) * kPairWarps;
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
  }   // work items
  if (((p.persist >> 2) & 1)) pdl_launch_dependents();
  work_leave(p, 2);
}

size_t mono_bwd_pair_smem_bytes() { return (size_t)kPairPlanes * kPlane * sizeof(float); }

cudaError_t launch_mono_bwd_pair(const MonoParams& p, const MonoTma& t, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(mono_bwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mono_bwd_pair_smem_bytes());
  if (e != cudaSuccess) return e;
  unsigned grid = (unsigned)p.btile_start[p.n_scales];
  if (((p.persist >> 2) & 1)) {
    static unsigned slots = 0;
    if (!slots) slots = resident_ctas(mono_bwd_pair_kernel, kThreads, mono_bwd_pair_smem_bytes());
    if (slots && grid > slots) grid = slots;
  }
  return launch_chained(2, mono_bwd_pair_kernel, grid, kThreads, mono_bwd_pair_smem_bytes(), stream, p, t);
}

}  // namespace sde
