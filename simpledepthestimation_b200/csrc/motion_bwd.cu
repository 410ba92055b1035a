// Backward of the MotionLearning two-frame loss (sm_100a).
//
// Gradients flow through the warped rgb only: the occlusion mask is a comparison, the proximity
// weight is detached and DEPTH_L1_WEIGHT is 0 in every shipped config (MotionLearning.py:259,
// 264-267,283).  One CTA owns a 64x16 block Q of WeightedSSIM window centres of one
// (direction, sample) and emits gradients for Q's 60x14 interior P:
//   phase 1  as the forward kernel: project + gather on Q + 1-pixel halo -> planes S, A, U, WZ, depth
//   per channel:
//     phase 2  window sums -> coefficients of  d ssim_q / d S_p = U_p (a_q + S_p b_q + A_p c_q)
//              scaled by the upstream gradient and avg_w_q, to shared planes
//     phase 3  adjoint of reflect-pad + 3x3 box, times U_p, plus the occlusion-masked L1 term
//   phase 4  warp backward per pixel of P: d/d depth_A, d/d pose (12 sums) and the per-pixel
//            d/d translation field  K^T g_p.
// Smoothness gradient from the saved per-image mean and loss.  Deterministic reductions.
#include <type_traits>

#include "motion_device.cuh"

namespace sde {

// Eleven planes so that four CTAs fit on an SM: gS_c overwrites S_c in place (S_c is dead once the adjoint pass of
// channel c has read its own pixel); WZ (zero-padded weight) is only read before the channel loop (avg_w), so it
// shares coefficient plane 0 (whose border ring is zeroed afterwards).
constexpr int kNA = 0, kNS = 3, kNU = 6, kND = 7, kNCoef = 8, kNG = kNS, kNW = kNCoef;
constexpr int kMotionBwdPlanes = 11;
// the gradient block P starts at plane column 3: with 60-wide tiles plane index 0 is image column tile_x0 - 4 (TMA)
constexpr int kMBwdColOff = 3;
constexpr int kMPosPerThread = (kBwdW * kBwdH + kThreads - 1) / kThreads;  // 7

struct MotionBwdShared {
  MCam cam;
  // dense phase 4 (warp mode): the projection as p_i = depth (N_i . [x, y, 1]) + tau_i with N = K R K^-1 and
  // tau = K (t_pose + field), and d (R P + t) / d depth = V [x, y, 1]^T with V = R K^-1
  float nmat[9], vmat[9], kt[3];
  float red[12][kThreads / 32];
  double dred[12][kThreads / 32];
  unsigned ticket;
  __align__(8) uint64_t bar;             // TMA completion barrier
  __align__(8) uint8_t flag[kPlane];     // bit 0: occlusion mask of the staged position, bit 1: it lies in the image
};

__global__ void __launch_bounds__(kThreads, 4) motion_bwd_kernel(const __grid_constant__ MotionParams p,
                                                                 const __grid_constant__ MotionTma maps) {
  extern __shared__ __align__(128) float planes[];  // [kMotionBwdPlanes][kPlane]
  __shared__ MotionBwdShared sh;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int dir, b, tx0, ty0;
  decode_motion_tile(blockIdx.x, p.btiles_per_dir, p.btiles_x, p.btiles_y, kBwdW, kBwdH, dir, b, tx0, ty0);
  const int h = p.h, w = p.w, hw = h * w;
  const int ox = tx0 - kMBwdColOff, oy = ty0 - 2;
  const bool lr_border = ox + 2 <= 1 || ox + kHW - 3 >= w - 2;

  const bool tma = p.tma != 0;
  if (tma && tid == 0) {
    // warp mode: warped rgb, depth error, valid/occlusion from the forward pass, frame A and depth A by TMA
    mbar_init(&sh.bar, 1);
    mbar_init_fence();
  }
  pdl_wait();   // the forward pass's planes and statistics (and every other tensor) are complete from here on
  if (tma && tid == 0) {
    mbar_arrive_expect_tx(&sh.bar, 9 * kPlaneBytesTma);
    const int bx = ox - kColOff;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tma_load_plane(planes + (kNA + c) * kPlane, &maps.frame_a[dir], &sh.bar, bx, oy, b * 3 + c);
      tma_load_plane(planes + (kNS + c) * kPlane, &maps.warped[dir], &sh.bar, bx, oy, b * kMotionSaved + c);
    }
    tma_load_plane(planes + kNU * kPlane, &maps.warped[dir], &sh.bar, bx, oy, b * kMotionSaved + 3);
    tma_load_plane(planes + kNW * kPlane, &maps.warped[dir], &sh.bar, bx, oy, b * kMotionSaved + 4);
    tma_load_plane(planes + kND * kPlane, &maps.depth_a[dir], &sh.bar, bx, oy, b);
  }
  if (tid == 0) {
    load_mcam(sh.cam, p.K, p.pose[dir], b, p.sx, p.sy);
    const MCam& c = sh.cam;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        sh.nmat[i * 3 + j] = c.m[i * 3] * c.cam.ki[j] + c.m[i * 3 + 1] * c.cam.ki[3 + j] + c.m[i * 3 + 2] * c.cam.ki[6 + j];
        sh.vmat[i * 3 + j] = c.r[i * 3] * c.cam.ki[j] + c.r[i * 3 + 1] * c.cam.ki[3 + j] + c.r[i * 3 + 2] * c.cam.ki[6 + j];
      }
      sh.kt[i] = c.k[i * 3] * c.t[0] + c.k[i * 3 + 1] * c.t[1] + c.k[i * 3 + 2] * c.t[2];
    }
  }
  // coefficient planes 1, 2: the border ring is never written and must read as zero (plane 0 holds WZ for now)
  zero_ring<2>(planes, kNCoef + 1, tid);
  __syncthreads();

  MotionStage st;
  st.depth_a = pinned(p.depth_a[dir] + (size_t)b * hw);
  st.depth_b = pinned(p.depth_b[dir] + (size_t)b * hw);
  st.frame_a = pinned(p.frame_a[dir] + (size_t)b * 3 * hw);
  st.frame_b = pinned(p.frame_b[dir] + (size_t)b * 3 * hw);
  st.field = p.field[dir] ? pinned(p.field[dir] + (size_t)b * 3 * hw) : nullptr;
  st.planes = planes; st.oy = oy; st.ox = ox; st.h = h; st.w = w; st.hw = hw;
  st.m2 = __ldg(p.stats + (dir * p.B + b) * 4);

  const float* gl = p.grad_losses + dir * 4;
  const float g_l1v = __ldg(gl), g_ssv = __ldg(gl + 1), g_smooth = __ldg(gl + 2);
  const float inv_n = 1.0f / ((float)p.B * 3.0f * (float)h * (float)w);
  const float g_l1 = g_l1v * inv_n;
  const float g_ss = p.ssim_w > 0.0f ? g_ssv * inv_n * p.ssim_w * 0.5f * -0.5f : 0.0f;   // d loss / d ssim_q / avg_w_q
  const bool use_ssim = p.ssim_w > 0.0f;

  // ------------------------------------------------------------------ phase 1
  if (tma) {
    mbar_wait(&sh.bar, 0);
    const bool interior = ox >= 0 && oy >= 0 && ox + kHW <= w && oy + kHH <= h;
    if (!interior) {
      reflect_fixup(planes, kNA, 6, oy, ox, h, w, tid);   // A and S
      reflect_fixup(planes, kND, 1, oy, ox, h, w, tid);
      __syncthreads();
    }
    // depth error, valid/occlusion -> U = weight + 0.01, WZ = weight and the flag bytes, two stored positions (an
    // aligned pair of the plane row, pad columns included) per iteration
    constexpr int kPairs = kPitch / 2;
    for (int i = tid; i < kHH * kPairs; i += kThreads) {
      const int yy = i / kPairs, j = i - yy * kPairs;
      const int pl = yy * kPitch + 2 * j;
      const f2 derr = ld2(planes + kNU * kPlane + pl), vo = ld2(planes + kNW * kPlane + pl);
      const float w0 = proximity_weight(lo(derr), lo(vo), st.m2), w1 = proximity_weight(hi(derr), hi(vo), st.m2);
      *reinterpret_cast<unsigned long long*>(planes + kNU * kPlane + pl) = mk2(w0 + 1e-2f, w1 + 1e-2f).v;
      *reinterpret_cast<unsigned long long*>(planes + kNW * kPlane + pl) = mk2(w0, w1).v;
      const int ty = oy + yy, tx = ox + 2 * j - kColOff;   // image position of the pair's first element
      const bool row_in = ty >= 0 && ty < h;
      uchar2 f;
      f.x = (uint8_t)((lo(vo) >= 2.0f ? 1 : 0) | ((row_in && tx >= 0 && tx < w) ? 2 : 0));
      f.y = (uint8_t)((hi(vo) >= 2.0f ? 1 : 0) | ((row_in && tx + 1 >= 0 && tx + 1 < w) ? 2 : 0));
      *reinterpret_cast<uchar2*>(sh.flag + pl) = f;
    }
    if (!interior) {
      __syncthreads();
      reflect_fixup(planes, kNU, 1, oy, ox, h, w, tid);
    }
  } else {
    const MCam mc = sh.cam;
#pragma unroll 1
    for (int i = tid; i < kPositions; i += kThreads) {
      int yy, xx;
      position_of(i, yy, xx);
      const int ty = oy + yy, tx = ox + xx;
      const bool inside = ty >= 0 && ty < h && tx >= 0 && tx < w;
      const int gy = reflect_clamp(ty, h), gx = reflect_clamp(tx, w);
      const int pix = gy * w + gx;
      MotionSample sm;
      motion_sample(st, mc, gy, gx, pix, true, sm);
      const int pl = plane_index(yy, xx);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        planes[(kNS + c) * kPlane + pl] = sm.S[c];
        planes[(kNA + c) * kPlane + pl] = sm.A[c];
      }
      planes[kNU * kPlane + pl] = sm.wgt + 1e-2f;
      planes[kNW * kPlane + pl] = inside ? sm.wgt : 0.0f;
      planes[kND * kPlane + pl] = sm.d;
      sh.flag[pl] = (sm.occ != 0.0f ? 1 : 0) | (inside ? 2 : 0);
    }
  }
  __syncthreads();

  const int r0 = wid * kRowsPerWarp;
  const int c0 = 2 * lane;
  const int px0 = ox + c0 + 1, px1 = px0 + 1;
  const float eL0 = px0 == 1 ? 1.0f : 0.0f, eL1 = px1 == 1 ? 1.0f : 0.0f;
  const float eR0 = px0 == w - 2 ? 1.0f : 0.0f, eR1 = px1 == w - 2 ? 1.0f : 0.0f;
  const f2 C1 = bc2(p.c1), C2 = bc2(p.c2);
  const f2 ninth = bc2(1.0f / 9.0f);

  // avg_w and (1/9) inverse_avg_w of this lane's window centres (rows r0+1 .. r0+4 of the plane)
  f2 avgw[kRowsPerWarp], q9[kRowsPerWarp];
  if (use_ssim) {
    const float* pw = planes + kNW * kPlane + plane_index(r0, c0);
    f2 hW[2];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
      const Row4 wz = ld_row(pw + rr * kPitch);
      const f2 nW = (wz.c + swp(wz.c)) + wz.o;
      if (rr >= 2) {
        const int o = rr - 2;
        avgw[o] = ((hW[0] + hW[1]) + nW) * ninth;
        q9[o] = div2(ninth, avgw[o] + bc2(1e-2f));
      }
      hW[0] = hW[1]; hW[1] = nW;
    }
  }
  // WZ is dead: its plane becomes coefficient plane 0, whose border ring must read as zero (the coefficient pass
  // rewrites the interior of all three planes every channel)
  __syncthreads();
  zero_ring<1>(planes, kNCoef, tid);

#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    // ---------------------------------------------------------------- phase 2: coefficients on Q
    // instantiated per C1/C2 mode (ssim_loss.py:97-105): the unused SSIM factor and the mode switches are compiled out
    auto phase2 = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      const float* pa = planes + (kNA + c) * kPlane + plane_index(r0, c0);
      const float* px = planes + (kNS + c) * kPlane + plane_index(r0, c0);
      const float* pu = planes + kNU * kPlane + plane_index(r0, c0);
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
          const int row = r0 + rr - 1;   // plane row of the window centre
          const f2 k = q9[o];
          const f2 mx = ((hX[0] + hX[1]) + nX) * k, my = ((hA[0] + hA[1]) + nA) * k;
          const f2 exx = ((hXX[0] + hXX[1]) + nXX) * k, eaa = ((hAA[0] + hAA[1]) + nAA) * k;
          const f2 exa = ((hXA[0] + hXA[1]) + nXA) * k;
          const f2 sx = fma2(mx * bc2(-1.0f), mx, exx), sy = fma2(my * bc2(-1.0f), my, eaa);
          const f2 sxy = fma2(mx * bc2(-1.0f), my, exa);
          const f2 n2 = fma2(bc2(2.0f), sxy, C2), d2 = (sx + sy) + C2;
          const f2 n1 = fma2(bc2(2.0f), mx * my, C1), d1 = fma2(mx, mx, my * my) + C1;
          f2 N, D;
          if (MODE == 1) { N = n2; D = d2; }
          else if (MODE == 2) { N = n1; D = d1; }
          else { N = n1 * n2; D = d1 * d2; }
          const f2 nssim = ndiv2(N, D);   // -ssim: one packed instruction less, and the minus signs below disappear
          const uchar2 fl = *reinterpret_cast<const uchar2*>(sh.flag + plane_index(row, c0 + 1));
          const bool in_x = (fl.x & 2) != 0, in_y = (fl.y & 2) != 0;
          const float h0 = fmaf(lo(nssim), 0.5f, 0.5f), h1 = fmaf(hi(nssim), 0.5f, 0.5f);
          // torch.clamp passes the gradient on the closed interval; windows centred outside the image do not exist
          const f2 g = mk2((in_x && h0 >= 0.0f && h0 <= 1.0f) ? g_ss : 0.0f, (in_y && h1 >= 0.0f && h1 <= 1.0f) ? g_ss : 0.0f);
          // d ssim / d x_p = U_p * base * [ ... ]  with  base = 2 g avg_w (inverse_avg_w / 9) / D
          const f2 base = div2((g * avgw[o]) * (k * bc2(2.0f)), D);
          f2 ca, cb, cc;
          if (MODE == 1) {
            cc = base;
            cb = base * nssim;
            ca = (cb * mx + cc * my) * bc2(-1.0f);
          } else if (MODE == 2) {
            cc = bc2(0.0f); cb = bc2(0.0f);
            ca = base * fma2(nssim, mx, my);
          } else {
            cc = base * n1;
            const f2 bns = base * nssim;
            cb = bns * d1;
            ca = fma2(bns * mx, d2, (base * my) * n2) - (cb * mx + cc * my);
          }
          float* pc = planes + kNCoef * kPlane + plane_index(row, c0 + 1);
          *reinterpret_cast<unsigned long long*>(pc) = ca.v;
          *reinterpret_cast<unsigned long long*>(pc + kPlane) = cb.v;
          *reinterpret_cast<unsigned long long*>(pc + 2 * kPlane) = cc.v;
        }
        hX[0] = hX[1]; hX[1] = nX; hA[0] = hA[1]; hA[1] = nA;
        hXX[0] = hXX[1]; hXX[1] = nXX; hAA[0] = hAA[1]; hAA[1] = nAA; hXA[0] = hXA[1]; hXA[1] = nXA;
      }
    };
    if (use_ssim) {
      if (p.mode == 1) phase2(std::integral_constant<int, 1>{});
      else if (p.mode == 2) phase2(std::integral_constant<int, 2>{});
      else phase2(std::integral_constant<int, 0>{});
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 3: adjoint -> gS_c on P
    // instantiated twice: tiles on the left / right image border add the mirrored pad column (block-uniform)
    auto phase3 = [&](auto lr_tag, auto ssim_tag) {
      constexpr bool LR = decltype(lr_tag)::value;
      constexpr bool SSIM = decltype(ssim_tag)::value;   // compile-time: no control-flow joins inside the unrolled rows
      f2 hq[3][2];
#pragma unroll
      for (int rr = 0; rr < kRowsPerWarp + 2; ++rr) {
        f2 nq[3];
        if (SSIM) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const Row4 q = ld_row(planes + (kNCoef + k) * kPlane + plane_index(r0 + rr, c0));
            f2 hsum = (q.c + swp(q.c)) + q.o;
            if (LR) hsum = hsum + mk2(eL0 * lo(q.o) + eR0 * hi(q.c), eL1 * lo(q.c) + eR1 * hi(q.o));
            nq[k] = hsum;
          }
        }
        if (rr >= 2) {
          const int row = r0 + rr - 1;
          const int py = oy + row;
          const f2 wu = bc2(py == 1 ? 2.0f : 1.0f), wd = bc2(py == h - 2 ? 2.0f : 1.0f);
          const int pl = plane_index(row, c0 + 1);
          const f2 Sp = ld2(planes + (kNS + c) * kPlane + pl);
          const f2 Ap = ld2(planes + (kNA + c) * kPlane + pl);
          f2 gS = bc2(0.0f);
          if (SSIM) {
            const f2 va = fma2(wu, hq[0][0], fma2(wd, nq[0], hq[0][1]));
            const f2 vb = fma2(wu, hq[1][0], fma2(wd, nq[1], hq[1][1]));
            const f2 vc = fma2(wu, hq[2][0], fma2(wd, nq[2], hq[2][1]));
            const f2 Up = ld2(planes + kNU * kPlane + pl);
            gS = Up * fma2(Sp, vb, fma2(Ap, vc, va));
          }
          // rgb L1 on the pixel itself: occlusion * sign(S - A) (branch-free)
          const uchar2 m = *reinterpret_cast<const uchar2*>(sh.flag + pl);
          const f2 df = Sp - Ap;
          const float d0 = lo(df), d1v = hi(df);
          float l0 = d0 > 0.0f ? g_l1 : -g_l1, l1 = d1v > 0.0f ? g_l1 : -g_l1;
          l0 = ((m.x & 1) && d0 != 0.0f) ? l0 : 0.0f;
          l1 = ((m.y & 1) && d1v != 0.0f) ? l1 : 0.0f;
          gS = gS + mk2(l0, l1);
          *reinterpret_cast<unsigned long long*>(planes + (kNG + c) * kPlane + pl) = gS.v;
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

  // ------------------------------------------------------------------ phase 4: warp backward on P + smoothness
  {
    const MCam& mc = sh.cam;   // camera terms are read from shared memory where they are used
    const float* __restrict__ sc0 = st.frame_b;
    const float* __restrict__ saved = tma ? p.warped[dir] + (size_t)b * kMotionSaved * hw : nullptr;
    float* __restrict__ gout = p.grad_depth[dir] + (size_t)b * hw;
    float* __restrict__ gfield = p.grad_field[dir] ? p.grad_field[dir] + (size_t)b * 3 * hw : nullptr;
    const float mbar = p.stats[(dir * p.B + b) * 4 + 2], Lb = p.stats[(dir * p.B + b) * 4 + 3];
    const float inx = 1.0f / ((float)p.B * (float)h * (float)(w - 1));
    const float iny = 1.0f / ((float)p.B * (float)(h - 1) * (float)w);
    const float homog = mbar > 1e-6f ? Lb / ((float)h * (float)w * mbar) : 0.0f;
    const float rmbar = 1.0f / mbar;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
    // Warp mode with 8-byte aligned gradient / field tensors: dense form, as phase 4 of the MonoDepth2 kernel
    // (mono_bwd.cu).  Every lane takes its own pixel pairs (the rows of its warp, two columns) in packed fp32; the
    // global operands of a row -- residual translation, the six derivative planes of the statistics pass (already gated
    // as nan_to_num / clamp gate the reference's gradient), the local smoothness gradient -- are coalesced 8-byte loads,
    // issued for two rows at a time; the sample coordinate is re-projected in packed arithmetic with one refined
    // reciprocal, which is also the factor 1 / (p2 + 1e-6) of d (X, Y) / d p.  Pairs outside the gradient block / the
    // image read the sample's first pixel and are switched off through gS.  (The per-pixel loop below it -- one
    // position per thread and iteration, scalar projection with IEEE divisions -- was 42 % of this kernel's instructions.)
    const bool dense = tma && ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(st.field) |
                                reinterpret_cast<uintptr_t>(gfield) | reinterpret_cast<uintptr_t>(saved)) & 7) == 0;
    if (dense) {
      const int gxp = ox + c0 + 1;   // even, and w % 4 == 0 in warp mode: a pair is inside the image or outside it
      const bool col_ok = lane >= 1 && lane <= 30 && gxp < w;
      const f2 xs = mk2((float)gxp, (float)gxp + 1.0f);
      const float* nm = sh.nmat;
      const float* vm = sh.vmat;
      f2 nb[3], vb[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        nb[i] = fma2(bc2(nm[3 * i]), xs, bc2(nm[3 * i + 2]));
        vb[i] = fma2(bc2(vm[3 * i]), xs, bc2(vm[3 * i + 2]));
      }
      const f2 ncx = bc2(-mc.cam.cx), ncy = bc2(-mc.cam.cy);
      const bool smooth = g_smooth != 0.0f;
      f2 Sa[3], Sb[3], Sy[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) Sa[k] = Sb[k] = Sy[k] = bc2(0.0f);
#pragma unroll
      for (int o2 = 0; o2 < kRowsPerWarp; o2 += 2) {
        f2 fq[2][3], dq[2][6], Gq[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = r0 + 1 + o2 + u, gy = oy + row;
          ok[u] = row >= 2 && row <= kBwdH + 1 && gy < h && col_ok;
          SDE_CHECK(!ok[u] || (gy >= 0 && gy < h && gxp >= 0 && gxp + 1 < w));
          const int pix = ok[u] ? gy * w + gxp : 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) fq[u][k] = st.field ? ldg2(st.field + pix + k * hw) : bc2(0.0f);
#pragma unroll
          for (int k = 0; k < 6; ++k) dq[u][k] = ldg2(saved + (5 + k) * hw + pix);
          Gq[u] = smooth ? ldg2(saved + 11 * hw + pix) : bc2(0.0f);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = r0 + 1 + o2 + u, gy = oy + row;
          const int pl = plane_index(row, c0 + 1);
          f2 g0 = ld2(planes + kNG * kPlane + pl), g1 = ld2(planes + (kNG + 1) * kPlane + pl), g2 = ld2(planes + (kNG + 2) * kPlane + pl);
          if (!ok[u]) g0 = g1 = g2 = bc2(0.0f);
          const f2 dd = ld2(planes + kND * kPlane + pl);
          const f2 gX = fma2(g2, dq[u][2], fma2(g1, dq[u][1], g0 * dq[u][0]));
          const f2 gY = fma2(g2, dq[u][5], fma2(g1, dq[u][4], g0 * dq[u][3]));
          const f2 yf = bc2((float)gy);
          // p_i = depth n_i + tau_i, tau = K t_pose + K field
          f2 pi[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const f2 ni = fma2(bc2(nm[3 * i + 1]), yf, nb[i]);
            const f2 tau = fma2(bc2(mc.k[3 * i + 2]), fq[u][2], fma2(bc2(mc.k[3 * i + 1]), fq[u][1], fma2(bc2(mc.k[3 * i]), fq[u][0], bc2(sh.kt[i]))));
            pi[i] = fma2(dd, ni, tau);
          }
          const f2 den = pi[2] + bc2(1e-6f);
          f2 r = mk2(rcp_approx(lo(den)), rcp_approx(hi(den)));
          r = fma2(fma2(den * bc2(-1.0f), r, bc2(1.0f)), r, r);
          // pixels whose derivative planes are non-zero project inside the image (finite r, X, Y); the clamps keep the
          // rest finite (a NaN clamps to a bound), so that their zero gradients stay zero
          const float huge = 3.0e38f, big = 16777216.0f;
          r = mk2(fminf(fmaxf(lo(r), -huge), huge), fminf(fmaxf(hi(r), -huge), huge));
          const f2 X = pi[0] * r, Y = pi[1] * r;
          const f2 ex = mk2(fminf(fmaxf(lo(X), -big), big), fminf(fmaxf(hi(X), -big), big)) + ncx;
          const f2 ey = mk2(fminf(fmaxf(lo(Y), -big), big), fminf(fmaxf(hi(Y), -big), big)) + ncy;
          const f2 u0 = gX * r, u1 = gY * r;
          // K^T g_p = d loss / d (t_pose + field) of the pixels
          const f2 gt0 = bc2(mc.cam.fx) * u0;
          const f2 gt1 = fma2(bc2(mc.cam.sk), u0, bc2(mc.cam.fy) * u1);
          const f2 gt2 = (fma2(gY, ey, gX * ex) * r) * bc2(-1.0f);
          const int pix = gy * w + gxp;
          if (gfield && ok[u]) {
            *reinterpret_cast<unsigned long long*>(gfield + pix) = gt0.v;
            *reinterpret_cast<unsigned long long*>(gfield + pix + hw) = gt1.v;
            *reinterpret_cast<unsigned long long*>(gfield + pix + 2 * hw) = gt2.v;
          }
          // pose sums: sum gt_i [P, 1] with P = depth K^-1 [x, y, 1]^T is linear in sum gt_i depth {x, y, 1} and sum gt_i;
          // x is fixed per lane half and applied after the rows
          const f2 b0 = gt0 * dd, b1 = gt1 * dd, b2 = gt2 * dd;
          Sa[0] = Sa[0] + gt0; Sa[1] = Sa[1] + gt1; Sa[2] = Sa[2] + gt2;
          Sb[0] = Sb[0] + b0; Sb[1] = Sb[1] + b1; Sb[2] = Sb[2] + b2;
          Sy[0] = fma2(b0, yf, Sy[0]); Sy[1] = fma2(b1, yf, Sy[1]); Sy[2] = fma2(b2, yf, Sy[2]);
          // d/d depth: g_t . (R K^-1 [x, y, 1]^T)
          const f2 v0 = fma2(bc2(vm[1]), yf, vb[0]), v1 = fma2(bc2(vm[4]), yf, vb[1]), v2 = fma2(bc2(vm[7]), yf, vb[2]);
          f2 gd = fma2(gt2, v2, fma2(gt1, v1, gt0 * v0));
          if (smooth) {
            const float d0 = lo(dd), d1 = hi(dd);
            const float ic0 = inv_depth(d0), ic1 = inv_depth(d1);
            const float s0 = d0 >= 1e-6f ? -ic0 * ic0 * (lo(Gq[u]) * rmbar - homog) * g_smooth : 0.0f;
            const float s1 = d1 >= 1e-6f ? -ic1 * ic1 * (hi(Gq[u]) * rmbar - homog) * g_smooth : 0.0f;
            gd = gd + mk2(s0, s1);
          }
          if (ok[u]) *reinterpret_cast<unsigned long long*>(gout + pix) = gd.v;
        }
      }
      // the twelve sums sum gt_i P_j, sum gt_i of this lane: P_j = depth (ki_j0 x + ki_j1 y + ki_j2)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const f2 sxv = Sb[i] * xs;
        const float sx = lo(sxv) + hi(sxv), sy = lo(Sy[i]) + hi(Sy[i]), sb = lo(Sb[i]) + hi(Sb[i]);
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[4 * i + j] = mc.cam.ki[3 * j] * sx + mc.cam.ki[3 * j + 1] * sy + mc.cam.ki[3 * j + 2] * sb;
        acc[4 * i + 3] = lo(Sa[i]) + hi(Sa[i]);
      }
    } else {
    // warp mode: the global operands of a pixel (translation field, derivative planes, smoothness gradient; all
    // coalesced) are fetched one iteration ahead
    float nx[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) nx[k] = 0.0f;
    auto fetch = [&](int it) {
      const int i = tid + it * kThreads;
      const int ly = i / kBwdW, lx = i - ly * kBwdW;
      const int gy = ty0 + ly, gx = tx0 + lx;
      if (tma && i < kBwdW * kBwdH && gy < h && gx < w) {
        const int pix = gy * w + gx;
        if (st.field) { nx[0] = __ldg(st.field + pix); nx[1] = __ldg(st.field + pix + hw); nx[2] = __ldg(st.field + pix + 2 * hw); }
#pragma unroll
        for (int k = 0; k < 6; ++k) nx[3 + k] = __ldg(saved + (5 + k) * hw + pix);
        if (g_smooth != 0.0f) nx[9] = __ldg(saved + 11 * hw + pix);
      }
    };
    fetch(0);
#pragma unroll 1
    for (int it = 0; it < kMPosPerThread; ++it) {
      const int i = tid + it * kThreads;
      const int ly = i / kBwdW, lx = i - ly * kBwdW;
      const int gy = ty0 + ly, gx = tx0 + lx;
      float cv[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) cv[k] = nx[k];
      if (it + 1 < kMPosPerThread) fetch(it + 1);
      if (i < kBwdW * kBwdH && gy < h && gx < w) {
        const int pl = plane_index(ly + 2, lx + kMBwdColOff);
        const int pix = gy * w + gx;
        const float g0 = planes[kNG * kPlane + pl], g1 = planes[(kNG + 1) * kPlane + pl], g2 = planes[(kNG + 2) * kPlane + pl];
        const float d = planes[kND * kPlane + pl];
        float gd = 0.0f, gt0 = 0.0f, gt1 = 0.0f, gt2 = 0.0f;
        if (g0 != 0.0f || g1 != 0.0f || g2 != 0.0f) {
          float f0 = cv[0], f1 = cv[1], f2v = cv[2];
          if (!tma && st.field) { f0 = __ldg(st.field + pix); f1 = __ldg(st.field + pix + hw); f2v = __ldg(st.field + pix + 2 * hw); }
          const float fxp = (float)gx, fyp = (float)gy;
          float P[3], den, X, Y, Z;
          mproject(mc, fxp, fyp, d, f0, f1, f2v, P, den, X, Y, Z);
          const float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
          const bool gate_x = (X >= 0.0f) && (X <= wm1);   // nan_to_num + clamp gates (camera.py:184-188)
          const bool gate_y = (Y >= 0.0f) && (Y <= hm1);
          if (gate_x || gate_y) {
            float gX = 0.0f, gY = 0.0f;
            if (tma) {
              // warp mode: derivative planes of the statistics pass (already gated), coalesced loads
              gX = g0 * cv[3] + g1 * cv[4] + g2 * cv[5];
              gY = g0 * cv[6] + g1 * cv[7] + g2 * cv[8];
            } else {
              const Cell cell = bilinear_cell(X, Y, w, h);
              const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
              const float gs[3] = {g0, g1, g2};
#pragma unroll
              for (int cc = 0; cc < 3; ++cc) {
                const float* q0 = sc0 + cc * hw + cell.off;
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
            const float dx = gate_x ? X - mc.cam.cx : 0.0f, dy = gate_y ? Y - mc.cam.cy : 0.0f;
            // K^T g_p = d loss / d (t_pose + field) of this pixel
            gt0 = mc.cam.fx * u0;
            gt1 = mc.cam.sk * u0 + mc.cam.fy * u1;
            gt2 = -(gX * dx + gY * dy) * q;
            acc[0] += gt0 * P[0]; acc[1] += gt0 * P[1]; acc[2] += gt0 * P[2]; acc[3] += gt0;
            acc[4] += gt1 * P[0]; acc[5] += gt1 * P[1]; acc[6] += gt1 * P[2]; acc[7] += gt1;
            acc[8] += gt2 * P[0]; acc[9] += gt2 * P[1]; acc[10] += gt2 * P[2]; acc[11] += gt2;
            const float gP0 = mc.r[0] * gt0 + mc.r[3] * gt1 + mc.r[6] * gt2;
            const float gP1 = mc.r[1] * gt0 + mc.r[4] * gt1 + mc.r[7] * gt2;
            const float gP2 = mc.r[2] * gt0 + mc.r[5] * gt1 + mc.r[8] * gt2;
            const float rx = mc.cam.ki[0] * fxp + mc.cam.ki[1] * fyp + mc.cam.ki[2];
            const float ry = mc.cam.ki[3] * fxp + mc.cam.ki[4] * fyp + mc.cam.ki[5];
            const float rz = mc.cam.ki[6] * fxp + mc.cam.ki[7] * fyp + mc.cam.ki[8];
            gd = gP0 * rx + gP1 * ry + gP2 * rz;
          }
        }
        if (g_smooth != 0.0f) {
          const float* pd = planes + kND * kPlane + pl;
          const float ic = inv_depth(d);
          const float G = tma ? cv[9]
                              : smooth_grad_local(pd, planes + kNA * kPlane + pl, ic, gx, gy, w, h, inx, iny);
          const float g_inv = G * rmbar - homog;
          if (d >= 1e-6f) gd += -ic * ic * g_inv * g_smooth;
        }
        gout[pix] = gd;
        if (gfield) { gfield[pix] = gt0; gfield[pix + hw] = gt1; gfield[pix + 2 * hw] = gt2; }
      }
    }
    }
    pdl_launch_dependents();
    {
      float v16[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v16[k] = k < 12 ? acc[k] : 0.0f;
      const float mine = warp_sum16(v16, lane);   // 16 shuffles instead of 60
      const int slot = warp_slot(lane);
      if ((lane & 1) == 0 && slot < 12) sh.red[slot][wid] = mine;
    }
    __syncthreads();
    if (tid < 12) {
      float v = 0.0f;
#pragma unroll
      for (int k = 0; k < kThreads / 32; ++k) v += sh.red[tid][k];
      p.pose_partials[(size_t)blockIdx.x * 12 + tid] = v;
      publish_fence();   // only the publishing threads fence
    }
  }

  // ------------------------------------------------------------------ last tile of an image: its pose gradient
  // Every (direction, sample) has its own grad_pose, so the tile that finishes an image last adds that image's slots
  // in a fixed order in fp64 while the other images are still being computed (no serial last-CTA tail).
  __syncthreads();
  const int per_img = p.btiles_x * p.btiles_y, img = dir * p.B + b;
  if (tid == 0) sh.ticket = atomicAdd(p.img_counter_b + img, 1u);
  __syncthreads();
  if (sh.ticket != (unsigned)(per_img - 1)) return;
  publish_fence();
  {
    double a[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) a[k] = 0.0;
    const size_t first = (size_t)dir * p.btiles_per_dir + (size_t)b * per_img;
    for (int t = tid; t < per_img; t += kThreads) {
      const float4* part = reinterpret_cast<const float4*>(p.pose_partials + (first + t) * 12);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float4 v = __ldcg(part + k);
        a[4 * k] += (double)v.x; a[4 * k + 1] += (double)v.y; a[4 * k + 2] += (double)v.z; a[4 * k + 3] += (double)v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 12; ++k) sh.dred[k][wid] = a[k];
    }
    __syncthreads();
    float* gp = p.grad_pose[dir] + b * 16;
    if (tid < 12) gp[tid] = (float)(((sh.dred[tid][0] + sh.dred[tid][1]) + sh.dred[tid][2]) + sh.dred[tid][3]);
    if (tid < 4) gp[12 + tid] = 0.0f;
    if (tid == 0) p.img_counter_b[img] = 0u;   // leave the workspace zeroed for the next call
  }
}

size_t motion_bwd_smem_bytes() { return (size_t)kMotionBwdPlanes * kPlane * sizeof(float); }

cudaError_t launch_motion_bwd(const MotionParams& p, const MotionTma& t, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(motion_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)motion_bwd_smem_bytes());
  if (e != cudaSuccess) return e;
  return launch_chained(2, motion_bwd_kernel, (unsigned)(p.n_dirs * p.btiles_per_dir), kThreads, motion_bwd_smem_bytes(), stream, p, t);
}

}  // namespace sde
