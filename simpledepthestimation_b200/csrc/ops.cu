// Stand-alone operators of the view-synthesis loss path (sm_100a), one entry point per reference
// function, for callers that use the pieces outside the fused losses:
//   view_synthesis        geometry/camera.py:166-202   (forward; backward w.r.t. depth, R, t and the image)
//   SSIM / WeightedSSIM   modeling/losses/ssim_loss.py:34-53, 84-111   (forward; backward w.r.t. x and y)
//   smoothness_loss       modeling/losses/smoothness_loss.py:42-80     (forward; backward w.r.t. depth)
//   resize_img            geometry/camera.py:40-46     (bilinear, align_corners=True)
// One thread per pixel, direct (L1-cached) global loads: these kernels are bandwidth-trivial utility
// operators, the fused kernels are the hot path.  Reductions are deterministic (per-CTA slots added in a
// fixed order); the bilinear scatter of the image gradient accumulates in 64-bit fixed point with integer
// atomics, which is order-independent, so it is deterministic too (no floating-point atomics).
#include "motion_device.cuh"
#include "ops_params.cuh"

namespace sde {

constexpr int kOpThreads = 256;

__device__ __forceinline__ int reflect1(int i, int n) {   // nn.ReflectionPad2d(1): -1 -> 1, n -> n-2
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}
// number of pad positions q-1, q, q+1 of the window centred at q that mirror onto pixel p
__device__ __forceinline__ int reflect_count(int q, int p, int n) {
  return (reflect1(q - 1, n) == p) + (q == p) + (reflect1(q + 1, n) == p);
}

// =================================================================================================
// view_synthesis
// =================================================================================================
struct VsCam {
  float k[9], ki[9], m[9], r[9];
};

__device__ __forceinline__ void load_vscam(VsCam& c, const float* __restrict__ K, const float* __restrict__ R, int b) {
#pragma unroll
  for (int i = 0; i < 9; ++i) { c.k[i] = K[b * 9 + i]; c.ki[i] = c.k[i]; c.r[i] = R[b * 9 + i]; }
  // inv_intrinsics, camera.py:25-37
  c.ki[0] = 1.0f / c.k[0];
  c.ki[4] = 1.0f / c.k[4];
  c.ki[2] = -1.0f * c.k[2] / c.k[0];
  c.ki[5] = -1.0f * c.k[5] / c.k[4];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      c.m[i * 3 + j] = c.k[i * 3] * c.r[j] + c.k[i * 3 + 1] * c.r[3 + j] + c.k[i * 3 + 2] * c.r[6 + j];
}

__device__ __forceinline__ void vs_project(const VsCam& c, float gx, float gy, float d, float t0, float t1, float t2,
                                           float P[3], float& den, float& X, float& Y, float& Z) {
  const float xd = gx * d, yd = gy * d;
  P[0] = c.ki[0] * xd + c.ki[1] * yd + c.ki[2] * d;
  P[1] = c.ki[3] * xd + c.ki[4] * yd + c.ki[5] * d;
  P[2] = c.ki[6] * xd + c.ki[7] * yd + c.ki[8] * d;
  const float tau0 = c.k[0] * t0 + c.k[1] * t1 + c.k[2] * t2;
  const float tau1 = c.k[3] * t0 + c.k[4] * t1 + c.k[5] * t2;
  const float tau2 = c.k[6] * t0 + c.k[7] * t1 + c.k[8] * t2;
  const float p0 = c.m[0] * P[0] + c.m[1] * P[1] + c.m[2] * P[2] + tau0;
  const float p1 = c.m[3] * P[0] + c.m[4] * P[1] + c.m[5] * P[2] + tau1;
  Z = c.m[6] * P[0] + c.m[7] * P[1] + c.m[8] * P[2] + tau2;
  den = Z + 1e-6f;
  divide2(p0, p1, den, X, Y);
}

__device__ __forceinline__ void vs_translation(const VsParams& p, int b, int pix, int hw, float& t0, float& t1, float& t2) {
  if (p.t_per_pixel) {
    const float* t = p.t + (size_t)b * 3 * hw + pix;
    t0 = __ldg(t); t1 = __ldg(t + hw); t2 = __ldg(t + 2 * hw);
  } else {
    t0 = __ldg(p.t + b * 3); t1 = __ldg(p.t + b * 3 + 1); t2 = __ldg(p.t + b * 3 + 2);
  }
}

__global__ void __launch_bounds__(kOpThreads) vs_fwd_kernel(const __grid_constant__ VsParams p) {
  __shared__ VsCam s_cam;
  const int b = blockIdx.y, hw = p.h * p.w;
  if (threadIdx.x == 0) load_vscam(s_cam, p.K, p.R, b);
  __syncthreads();
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  if (pix >= hw) return;
  const int gy = pix / p.w, gx = pix - gy * p.w;
  const float d = __ldg(p.depth + (size_t)b * hw + pix);
  float t0, t1, t2;
  vs_translation(p, b, pix, hw, t0, t1, t2);
  float P[3], den, X, Y, Z;
  vs_project(s_cam, (float)gx, (float)gy, d, t0, t1, t2, P, den, X, Y, Z);
  const Cell cell = bilinear_cell(X, Y, p.w, p.h);
  const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
  const float w00 = bx * by, w01 = cell.ax * by, w10 = bx * cell.ay, w11 = cell.ax * cell.ay;
  for (int c = 0; c < p.C; ++c)
    p.sampled[((size_t)b * p.C + c) * hw + pix] = bilinear4(p.image + ((size_t)b * p.C + c) * hw + cell.off, p.w, w00, w01, w10, w11);
  if (p.depth_in_b) p.depth_in_b[(size_t)b * hw + pix] = clamp_depth(Z);
  if (p.valid) p.valid[(size_t)b * hw + pix] = valid_mask(X, Y, Z, p.w, p.h) != 0.0f ? 1 : 0;
  if (p.coords) {
    const float Xs = fminf(fmaxf(X, 0.0f), (float)(p.w - 1)), Ys = fminf(fmaxf(Y, 0.0f), (float)(p.h - 1));
    float2 cn;
    cn.x = __fdiv_rn(2.0f * Xs, (float)(p.w - 1)) - 1.0f;
    cn.y = __fdiv_rn(2.0f * Ys, (float)(p.h - 1)) - 1.0f;
    *reinterpret_cast<float2*>(p.coords + ((size_t)b * hw + pix) * 2) = cn;
  }
}

// backward: one thread per pixel; per-CTA slots for d/dR (9) and rigid d/dt (3), added by the last CTA of a sample
__global__ void __launch_bounds__(kOpThreads) vs_bwd_kernel(const __grid_constant__ VsParams p) {
  __shared__ VsCam s_cam;
  __shared__ float red[12][kOpThreads / 32];
  __shared__ double dred[12][kOpThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, hw = p.h * p.w;
  if (tid == 0) load_vscam(s_cam, p.K, p.R, b);
  __syncthreads();
  const VsCam& cam = s_cam;
  const int pix = blockIdx.x * kOpThreads + tid;
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0f;
  if (pix < hw) {
    const int gy = pix / p.w, gx = pix - gy * p.w;
    const float d = __ldg(p.depth + (size_t)b * hw + pix);
    float t0, t1, t2;
    vs_translation(p, b, pix, hw, t0, t1, t2);
    const float fxp = (float)gx, fyp = (float)gy;
    float P[3], den, X, Y, Z;
    vs_project(cam, fxp, fyp, d, t0, t1, t2, P, den, X, Y, Z);
    const float wm1 = (float)(p.w - 1), hm1 = (float)(p.h - 1);
    const bool gate_x = (X >= 0.0f) && (X <= wm1), gate_y = (Y >= 0.0f) && (Y <= hm1);   // nan_to_num + clamp
    const Cell cell = bilinear_cell(X, Y, p.w, p.h);
    const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
    float gX = 0.0f, gY = 0.0f;
    for (int c = 0; c < p.C; ++c) {
      const size_t plane = ((size_t)b * p.C + c) * hw;
      const float g = __ldg(p.g_sampled + plane + pix);
      const float* q0 = p.image + plane + cell.off;
      const float* q1 = q0 + p.w;
      const float v00 = __ldg(q0), v01 = __ldg(q0 + 1), v10 = __ldg(q1), v11 = __ldg(q1 + 1);
      gX += g * ((v01 - v00) * by + (v11 - v10) * cell.ay);
      gY += g * ((v10 - v00) * bx + (v11 - v01) * cell.ax);
      if (p.g_image_fix && g != 0.0f) {
        long long* o = p.g_image_fix + plane + cell.off;
        fix_add(o, g * (bx * by)); fix_add(o + 1, g * (cell.ax * by));
        fix_add(o + p.w, g * (bx * cell.ay)); fix_add(o + p.w + 1, g * (cell.ax * cell.ay));
      }
    }
    if (p.g_coords) {   // coords = 2 clamp(X) / (w-1) - 1
      const float2 gc = *reinterpret_cast<const float2*>(p.g_coords + ((size_t)b * hw + pix) * 2);
      gX += gc.x * (2.0f / wm1);
      gY += gc.y * (2.0f / hm1);
    }
    if (!gate_x) gX = 0.0f;
    if (!gate_y) gY = 0.0f;
    float gZ = 0.0f;
    if (p.g_depth_in_b) {
      const float g = __ldg(p.g_depth_in_b + (size_t)b * hw + pix);
      gZ = (Z >= 1e-5f) ? g : 0.0f;   // clamp(min=1e-5) passes the gradient on the closed side
    }
    const float q = 1.0f / den;
    // g_p (d loss / d projected point) and K^T g_p = d loss / d t of this pixel
    const float g0 = gX * q, g1 = gY * q;
    const float g2 = -(gX * (gate_x ? X : 0.0f) + gY * (gate_y ? Y : 0.0f)) * q + gZ;
    const float gt0 = cam.k[0] * g0 + cam.k[3] * g1 + cam.k[6] * g2;
    const float gt1 = cam.k[1] * g0 + cam.k[4] * g1 + cam.k[7] * g2;
    const float gt2 = cam.k[2] * g0 + cam.k[5] * g1 + cam.k[8] * g2;
    acc[0] = gt0 * P[0]; acc[1] = gt0 * P[1]; acc[2] = gt0 * P[2]; acc[3] = gt0;
    acc[4] = gt1 * P[0]; acc[5] = gt1 * P[1]; acc[6] = gt1 * P[2]; acc[7] = gt1;
    acc[8] = gt2 * P[0]; acc[9] = gt2 * P[1]; acc[10] = gt2 * P[2]; acc[11] = gt2;
    const float gP0 = cam.r[0] * gt0 + cam.r[3] * gt1 + cam.r[6] * gt2;
    const float gP1 = cam.r[1] * gt0 + cam.r[4] * gt1 + cam.r[7] * gt2;
    const float gP2 = cam.r[2] * gt0 + cam.r[5] * gt1 + cam.r[8] * gt2;
    const float rx = cam.ki[0] * fxp + cam.ki[1] * fyp + cam.ki[2];
    const float ry = cam.ki[3] * fxp + cam.ki[4] * fyp + cam.ki[5];
    const float rz = cam.ki[6] * fxp + cam.ki[7] * fyp + cam.ki[8];
    p.g_depth[(size_t)b * hw + pix] = gP0 * rx + gP1 * ry + gP2 * rz;
    if (p.t_per_pixel) {
      float* gt = p.g_t + (size_t)b * 3 * hw + pix;
      gt[0] = gt0; gt[hw] = gt1; gt[2 * hw] = gt2;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[k][wid] = v;
  }
  __syncthreads();
  if (tid < 12) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kOpThreads / 32; ++k) v += red[tid][k];
    p.partials[((size_t)b * gridDim.x + blockIdx.x) * 12 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(p.counters + b, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  double a[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) a[k] = 0.0;
  for (int t = tid; t < (int)gridDim.x; t += kOpThreads) {
    const float4* part = reinterpret_cast<const float4*>(p.partials + ((size_t)b * gridDim.x + t) * 12);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 v = __ldcg(part + k);
      a[4 * k] += (double)v.x; a[4 * k + 1] += (double)v.y; a[4 * k + 2] += (double)v.z; a[4 * k + 3] += (double)v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) dred[k][wid] = a[k];
  }
  __syncthreads();
  if (tid < 12) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kOpThreads / 32; ++k) v += dred[tid][k];
    const int row = tid >> 2, col = tid & 3;
    if (col < 3) p.g_R[b * 9 + row * 3 + col] = (float)v;
    else if (!p.t_per_pixel) p.g_t[b * 3 + row] = (float)v;
  }
  if (tid == 0) p.counters[b] = 0u;
}

// fixed point -> float, and re-zero the accumulator for the next call
__global__ void __launch_bounds__(kOpThreads) fix_to_float_kernel(long long* __restrict__ acc, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * kOpThreads + threadIdx.x;
  if (i < n) {
    out[i] = fix_to_float(acc[i]);
    acc[i] = 0;
  }
}

cudaError_t launch_vs_fwd(const VsParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B);
  vs_fwd_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_vs_bwd(const VsParams& p, float* g_image, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B);
  vs_bwd_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !g_image) return e;
  const size_t n = (size_t)p.B * p.C * p.h * p.w;
  fix_to_float_kernel<<<(unsigned)((n + kOpThreads - 1) / kOpThreads), kOpThreads, 0, stream>>>(p.g_image_fix, g_image, n);
  return cudaGetLastError();
}

// =================================================================================================
// SSIM / WeightedSSIM
// =================================================================================================
struct SsimTerms {
  float mx, my, n1, d1, n2, d2, N, D, ssim, k;   // k = 1/9 (plain) or inverse_avg_w / 9 (weighted)
  float avg_w;
};

// window sums around (qy, qx) of plane `c` of sample `b` on the reflect-padded products
__device__ __forceinline__ void ssim_window(const SsimParams& p, size_t plane, size_t wplane, int qy, int qx, SsimTerms& t) {
  float sx = 0.0f, sy = 0.0f, sxx = 0.0f, syy = 0.0f, sxy = 0.0f, sw = 0.0f;
#pragma unroll
  for (int ey = -1; ey <= 1; ++ey) {
    const int ty = qy + ey, ry = reflect1(ty, p.h);
#pragma unroll
    for (int ex = -1; ex <= 1; ++ex) {
      const int tx = qx + ex, rx = reflect1(tx, p.w);
      const int pix = ry * p.w + rx;
      const float x = __ldg(p.x + plane + pix), y = __ldg(p.y + plane + pix);
      float u = 1.0f;
      if (p.weight) {
        const float wv = __ldg(p.weight + wplane + pix);
        u = wv + 1e-2f;                                                           // ssim_loss.py:89
        if (ty >= 0 && ty < p.h && tx >= 0 && tx < p.w) sw += wv;                 // avg_pool2d zero padding
      }
      const float ux = u * x, uy = u * y;
      sx += ux; sy += uy; sxx += ux * x; syy += uy * y; sxy += ux * y;
    }
  }
  t.avg_w = sw * (1.0f / 9.0f);
  t.k = p.weight ? fdiv(1.0f / 9.0f, t.avg_w + 1e-2f) : (1.0f / 9.0f);
  t.mx = sx * t.k; t.my = sy * t.k;
  const float vx = fmaf(-t.mx, t.mx, sxx * t.k), vy = fmaf(-t.my, t.my, syy * t.k), vxy = fmaf(-t.mx, t.my, sxy * t.k);
  t.n1 = 2.0f * t.mx * t.my + p.c1; t.d1 = t.mx * t.mx + t.my * t.my + p.c1;
  t.n2 = 2.0f * vxy + p.c2; t.d2 = vx + vy + p.c2;
  if (p.mode == 1) { t.N = t.n2; t.D = t.d2; }
  else if (p.mode == 2) { t.N = t.n1; t.D = t.d1; }
  else { t.N = t.n1 * t.n2; t.D = t.d1 * t.d2; }
  t.ssim = fdiv(t.N, t.D);
}

__global__ void __launch_bounds__(kOpThreads) ssim_fwd_kernel(const __grid_constant__ SsimParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  const int bc = blockIdx.y;
  if (pix >= hw) return;
  const int qy = pix / p.w, qx = pix - qy * p.w;
  const size_t plane = (size_t)bc * hw, wplane = (size_t)(bc / p.C) * hw;
  SsimTerms t;
  ssim_window(p, plane, wplane, qy, qx, t);
  p.out[plane + pix] = __saturatef(fmaf(t.ssim, -0.5f, 0.5f));   // clamp((1 - ssim) / 2, 0, 1)
  if (p.avg_w && bc % p.C == 0) p.avg_w[wplane + pix] = t.avg_w;
}

// backward pass 1: per window centre q the coefficients of  d out_q / d x_p = u_p (a + x_p b + y_p c)  and the
// mirrored set for y, scaled by the upstream gradient -> six planes
__global__ void __launch_bounds__(kOpThreads) ssim_coef_kernel(const __grid_constant__ SsimParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  const int bc = blockIdx.y;
  if (pix >= hw) return;
  const int qy = pix / p.w, qx = pix - qy * p.w;
  const size_t plane = (size_t)bc * hw, wplane = (size_t)(bc / p.C) * hw;
  SsimTerms t;
  ssim_window(p, plane, wplane, qy, qx, t);
  const float hv = fmaf(t.ssim, -0.5f, 0.5f);
  const float g = (hv >= 0.0f && hv <= 1.0f) ? -0.5f * __ldg(p.g_out + plane + pix) : 0.0f;   // d out / d ssim
  const float base = fdiv(2.0f * g * t.k, t.D);
  float ax, bx, cx, ay, by, cy;   // x: a + x b + y c ; y: a' + y b' + x c'
  if (p.mode == 1) {
    cx = base; bx = -base * t.ssim; ax = -(bx * t.mx + cx * t.my);
    cy = base; by = -base * t.ssim; ay = -(by * t.my + cy * t.mx);
  } else if (p.mode == 2) {
    bx = cx = by = cy = 0.0f;
    ax = base * (t.my - t.ssim * t.mx);
    ay = base * (t.mx - t.ssim * t.my);
  } else {
    cx = base * t.n1; bx = -base * t.ssim * t.d1;
    ax = base * t.my * t.n2 - base * t.ssim * t.mx * t.d2 - (bx * t.mx + cx * t.my);
    cy = cx; by = bx;
    ay = base * t.mx * t.n2 - base * t.ssim * t.my * t.d2 - (by * t.my + cy * t.mx);
  }
  const size_t n = (size_t)gridDim.y * hw;
  float* c = p.coef + plane + pix;
  c[0] = ax; c[n] = bx; c[2 * n] = cx; c[3 * n] = ay; c[4 * n] = by; c[5 * n] = cy;
}

// backward pass 2: adjoint of reflect-pad + 3x3 box applied to the coefficient planes
__global__ void __launch_bounds__(kOpThreads) ssim_adjoint_kernel(const __grid_constant__ SsimParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  const int bc = blockIdx.y;
  if (pix >= hw) return;
  const int py = pix / p.w, px = pix - py * p.w;
  const size_t plane = (size_t)bc * hw, wplane = (size_t)(bc / p.C) * hw;
  const size_t n = (size_t)gridDim.y * hw;
  float s[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
  for (int qy = max(py - 1, 0); qy <= min(py + 1, p.h - 1); ++qy) {
    const int my = reflect_count(qy, py, p.h);
    for (int qx = max(px - 1, 0); qx <= min(px + 1, p.w - 1); ++qx) {
      const float m = (float)(my * reflect_count(qx, px, p.w));
      const float* c = p.coef + plane + qy * p.w + qx;
#pragma unroll
      for (int k = 0; k < 6; ++k) s[k] += m * __ldcg(c + k * n);
    }
  }
  const float x = __ldg(p.x + plane + pix), y = __ldg(p.y + plane + pix);
  const float u = p.weight ? __ldg(p.weight + wplane + pix) + 1e-2f : 1.0f;
  if (p.g_x) p.g_x[plane + pix] = u * (s[0] + x * s[1] + y * s[2]);
  if (p.g_y) p.g_y[plane + pix] = u * (s[3] + y * s[4] + x * s[5]);
}

cudaError_t launch_ssim_fwd(const SsimParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B * p.C);
  ssim_fwd_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_ssim_bwd(const SsimParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B * p.C);
  ssim_coef_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  ssim_adjoint_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// =================================================================================================
// smoothness_loss
// =================================================================================================
__global__ void __launch_bounds__(kOpThreads) smooth_fwd_kernel(const __grid_constant__ SmoothParams p) {
  __shared__ float red[3][kOpThreads / 32];
  __shared__ double dred[3][kOpThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, hw = p.h * p.w;
  const int pix = blockIdx.x * kOpThreads + tid;
  float sinv = 0.0f, smx = 0.0f, smy = 0.0f;
  if (pix < hw) {
    const int gy = pix / p.w, gx = pix - gy * p.w;
    const float* dp = p.depth + (size_t)b * hw + pix;
    const float* ip = p.image + (size_t)b * p.C * hw + pix;
    auto inv = [](float d) { return 1.0f / (d < 1e-6f ? 1e-6f : d); };
    const float ic = inv(__ldg(dp));
    sinv = ic;
    if (gx + 1 < p.w) {
      float e = 0.0f;
      for (int c = 0; c < p.C; ++c) e += fabsf(__ldg(ip + c * hw) - __ldg(ip + c * hw + 1));
      smx = fabsf(ic - inv(__ldg(dp + 1))) * expf(-e / (float)p.C);
    }
    if (gy + 1 < p.h) {
      float e = 0.0f;
      for (int c = 0; c < p.C; ++c) e += fabsf(__ldg(ip + c * hw) - __ldg(ip + c * hw + p.w));
      smy = fabsf(ic - inv(__ldg(dp + p.w))) * expf(-e / (float)p.C);
    }
  }
  sinv = warp_sum(sinv); smx = warp_sum(smx); smy = warp_sum(smy);
  if (lane == 0) { red[0][wid] = sinv; red[1][wid] = smx; red[2][wid] = smy; }
  __syncthreads();
  if (tid < 3) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kOpThreads / 32; ++k) v += red[tid][k];
    p.partials[((size_t)b * gridDim.x + blockIdx.x) * 4 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(p.counters + 1 + b, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  double a[3] = {0.0, 0.0, 0.0};
  for (int t = tid; t < (int)gridDim.x; t += kOpThreads) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p.partials + ((size_t)b * gridDim.x + t) * 4));
    a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  }
  if (lane == 0) { dred[0][wid] = a[0]; dred[1][wid] = a[1]; dred[2][wid] = a[2]; }
  __syncthreads();
  if (tid == 0) {
    double t3[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < kOpThreads / 32; ++k) { t3[0] += dred[0][k]; t3[1] += dred[1][k]; t3[2] += dred[2][k]; }
    const double qh = p.h, qw = p.w, nB = p.B;
    const double mbar = fmax(t3[0] / (qh * qw), 1e-6);
    const double Lb = (t3[1] / (nB * qh * (qw - 1.0)) + t3[2] / (nB * (qh - 1.0) * qw)) / mbar;
    p.stats[b * 2] = (float)mbar;
    p.stats[b * 2 + 1] = (float)Lb;
    p.fin[b] = Lb;
    p.counters[1 + b] = 0u;
    __threadfence();
    ticket = atomicAdd(p.counters, 1u);
  }
  __syncthreads();
  if (ticket != (unsigned)(p.B - 1)) return;
  __threadfence();
  if (tid == 0) {
    double L = 0.0;
    for (int k = 0; k < p.B; ++k) L += __ldcg(p.fin + k);
    p.loss[0] = (float)L;
    p.counters[0] = 0u;
  }
}

__global__ void __launch_bounds__(kOpThreads) smooth_bwd_kernel(const __grid_constant__ SmoothParams p) {
  const int b = blockIdx.y, hw = p.h * p.w;
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  if (pix >= hw) return;
  const int gy = pix / p.w, gx = pix - gy * p.w;
  const float* dp = p.depth + (size_t)b * hw + pix;
  const float* ip = p.image + (size_t)b * p.C * hw + pix;
  const float mbar = p.stats[b * 2], Lb = p.stats[b * 2 + 1];
  const float inx = 1.0f / ((float)p.B * (float)p.h * (float)(p.w - 1));
  const float iny = 1.0f / ((float)p.B * (float)(p.h - 1) * (float)p.w);
  const float homog = mbar > 1e-6f ? Lb / ((float)p.h * (float)p.w * mbar) : 0.0f;
  const float d = __ldg(dp);
  const float ic = inv_depth(d);
  auto sgn = [](float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); };
  auto edge = [&](int off) {   // exp(-mean_c |I(p) - I(p + off)|)
    float e = 0.0f;
    for (int c = 0; c < p.C; ++c) e += fabsf(__ldg(ip + c * hw) - __ldg(ip + c * hw + off));
    return __expf(-e / (float)p.C);
  };
  float G = 0.0f;
  if (gx + 1 < p.w) G += sgn(ic - inv_depth(__ldg(dp + 1))) * edge(1) * inx;
  if (gx >= 1) G -= sgn(inv_depth(__ldg(dp - 1)) - ic) * edge(-1) * inx;
  if (gy + 1 < p.h) G += sgn(ic - inv_depth(__ldg(dp + p.w))) * edge(p.w) * iny;
  if (gy >= 1) G -= sgn(inv_depth(__ldg(dp - p.w)) - ic) * edge(-p.w) * iny;
  const float g_inv = G / mbar - homog;
  p.g_depth[(size_t)b * hw + pix] = d >= 1e-6f ? -ic * ic * g_inv * __ldg(p.g_loss) : 0.0f;
}

cudaError_t launch_smooth_fwd(const SmoothParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B);
  smooth_fwd_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_smooth_bwd(const SmoothParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kOpThreads - 1) / kOpThreads, p.B);
  smooth_bwd_kernel<<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// =================================================================================================
// resize_img: F.interpolate(mode='bilinear', align_corners=True)
// =================================================================================================
__global__ void __launch_bounds__(kOpThreads) resize_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                     int planes, int sh, int sw, int dh, int dw,
                                                                     float rh, float rw) {
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  if (pix >= dh * dw) return;
  const int y = pix / dw, x = pix - y * dw;
  // ATen's area_pixel_compute_source_index with align_corners: src = scale * dst
  const float fy = rh * (float)y, fx = rw * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int yp = y0 < sh - 1 ? 1 : 0, xp = x0 < sw - 1 ? 1 : 0;
  const float ly1 = fy - (float)y0, ly0 = 1.0f - ly1, lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
  for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
    const float* s = src + (size_t)pl * sh * sw + (size_t)y0 * sw + x0;
    const float v00 = __ldg(s), v01 = __ldg(s + xp), v10 = __ldg(s + yp * sw), v11 = __ldg(s + yp * sw + xp);
    dst[(size_t)pl * dh * dw + pix] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
}

cudaError_t launch_resize_bilinear(const float* src, float* dst, int planes, int sh, int sw, int dh, int dw, cudaStream_t stream) {
  const float rh = dh > 1 ? (float)(sh - 1) / (float)(dh - 1) : 0.0f;
  const float rw = dw > 1 ? (float)(sw - 1) / (float)(dw - 1) : 0.0f;
  const dim3 grid((dh * dw + kOpThreads - 1) / kOpThreads, planes < 65535 ? planes : 65535);
  resize_bilinear_kernel<<<grid, kOpThreads, 0, stream>>>(src, dst, planes, sh, sw, dh, dw, rh, rw);
  return cudaGetLastError();
}

// resize_img_avgpool(image, dst_size) = F.adaptive_avg_pool2d (camera.py:49-54; MotionLearning.py:126-144 at
// NUM_SCALES > 1): output cell (i, j) averages the source rows floor(i sh / dh) .. ceil((i + 1) sh / dh) - 1 and the
// columns likewise.  One thread per output element; the backward kernel is the adjoint, one thread per source
// element gathering from the (at most two per axis when shrinking) cells that contain it -- no atomics.
__device__ __forceinline__ int pool_start(int i, int src, int dst) { return (int)(((long long)i * src) / dst); }
__device__ __forceinline__ int pool_end(int i, int src, int dst) { return (int)((((long long)i + 1) * src + dst - 1) / dst); }

__global__ void __launch_bounds__(kOpThreads) avgpool_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst, int sh,
                                                                  int sw, int dh, int dw) {
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  if (pix >= dh * dw) return;
  const int i = pix / dw, j = pix - i * dw;
  const int y0 = pool_start(i, sh, dh), y1 = pool_end(i, sh, dh), x0 = pool_start(j, sw, dw), x1 = pool_end(j, sw, dw);
  const float* s = src + (size_t)blockIdx.y * sh * sw;
  float sum = 0.0f;
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) sum += __ldg(s + (size_t)y * sw + x);
  dst[(size_t)blockIdx.y * dh * dw + pix] = sum / (float)((y1 - y0) * (x1 - x0));
}

__global__ void __launch_bounds__(kOpThreads) avgpool_bwd_kernel(const float* __restrict__ g_dst, float* __restrict__ g_src, int sh,
                                                                  int sw, int dh, int dw) {
  const int pix = blockIdx.x * kOpThreads + threadIdx.x;
  if (pix >= sh * sw) return;
  const int y = pix / sw, x = pix - y * sw;
  const float* g = g_dst + (size_t)blockIdx.y * dh * dw;
  // cells whose row range contains y: a superset is floor(y dh / sh) - 1 .. ceil((y + 1) dh / sh)
  const int i_lo = max(0, (int)(((long long)y * dh) / sh) - 1), i_hi = min(dh - 1, (int)((((long long)y + 1) * dh + sh - 1) / sh));
  const int j_lo = max(0, (int)(((long long)x * dw) / sw) - 1), j_hi = min(dw - 1, (int)((((long long)x + 1) * dw + sw - 1) / sw));
  float sum = 0.0f;
  for (int i = i_lo; i <= i_hi; ++i) {
    const int y0 = pool_start(i, sh, dh), y1 = pool_end(i, sh, dh);
    if (y < y0 || y >= y1) continue;
    for (int j = j_lo; j <= j_hi; ++j) {
      const int x0 = pool_start(j, sw, dw), x1 = pool_end(j, sw, dw);
      if (x < x0 || x >= x1) continue;
      sum += __ldg(g + (size_t)i * dw + j) / (float)((y1 - y0) * (x1 - x0));
    }
  }
  g_src[(size_t)blockIdx.y * sh * sw + pix] = sum;
}

cudaError_t launch_avgpool(bool backward, const float* in, float* out, int planes, int sh, int sw, int dh, int dw, cudaStream_t stream) {
  const int n = backward ? sh * sw : dh * dw;
  for (int p0 = 0; p0 < planes; p0 += 65535) {   // gridDim.y limit
    const int np = planes - p0 < 65535 ? planes - p0 : 65535;
    const dim3 grid((n + kOpThreads - 1) / kOpThreads, np);
    if (backward) avgpool_bwd_kernel<<<grid, kOpThreads, 0, stream>>>(in + (size_t)p0 * dh * dw, out + (size_t)p0 * sh * sw, sh, sw, dh, dw);
    else          avgpool_fwd_kernel<<<grid, kOpThreads, 0, stream>>>(in + (size_t)p0 * sh * sw, out + (size_t)p0 * dh * dw, sh, sw, dh, dw);
  }
  return cudaGetLastError();
}

// The image pyramid of a step (MonoDepth2.py:82,88: the target and every source frame resized to every coarser
// prediction size) in ONE launch: blockIdx.x walks the pixel blocks of all levels, .y the planes, .z the frames.
// U8: the frames are the decoded uint8 images; a tap is byte / 255 in fp32 -- the bits torchvision's ToTensor produces
// (kitti_v2.py:207-208) -- so the pyramid equals the one built from the converted frames, and a level of the source
// size is the conversion itself.
template <bool U8>
__global__ void __launch_bounds__(kOpThreads) resize_pyramid_kernel(const __grid_constant__ PyramidParams p) {
  int l = 0;
  while (l + 1 < p.n_levels && (int)blockIdx.x >= p.blk_start[l + 1]) ++l;
  const int dh = p.dh[l], dw = p.dw[l], sh = p.sh, sw = p.sw;
  const int pix = ((int)blockIdx.x - p.blk_start[l]) * kOpThreads + threadIdx.x;
  if (pix >= dh * dw) return;
  const int y = pix / dw, x = pix - y * dw;
  const float fy = p.rh[l] * (float)y, fx = p.rw[l] * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int yp = y0 < sh - 1 ? 1 : 0, xp = x0 < sw - 1 ? 1 : 0;
  const float ly1 = fy - (float)y0, ly0 = 1.0f - ly1, lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
  float* __restrict__ dst = p.dst[blockIdx.z][l];
  for (int pl = blockIdx.y; pl < p.planes; pl += gridDim.y) {
    const size_t off = (size_t)pl * sh * sw + (size_t)y0 * sw + x0;
    float v00, v01, v10, v11;
    if (U8) {
      const uint8_t* s = static_cast<const uint8_t*>(p.src[blockIdx.z]) + off;
      v00 = (float)__ldg(s) / 255.0f; v01 = (float)__ldg(s + xp) / 255.0f;
      v10 = (float)__ldg(s + yp * sw) / 255.0f; v11 = (float)__ldg(s + yp * sw + xp) / 255.0f;
    } else {
      const float* s = static_cast<const float*>(p.src[blockIdx.z]) + off;
      v00 = __ldg(s); v01 = __ldg(s + xp); v10 = __ldg(s + yp * sw); v11 = __ldg(s + yp * sw + xp);
    }
    dst[(size_t)pl * dh * dw + pix] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);   // same order as above
  }
}

cudaError_t launch_resize_pyramid(PyramidParams& p, cudaStream_t stream) {
  int start = 0;
  for (int l = 0; l < p.n_levels; ++l) {
    p.blk_start[l] = start;
    start += (p.dh[l] * p.dw[l] + kOpThreads - 1) / kOpThreads;
    p.rh[l] = p.dh[l] > 1 ? (float)(p.sh - 1) / (float)(p.dh[l] - 1) : 0.0f;
    p.rw[l] = p.dw[l] > 1 ? (float)(p.sw - 1) / (float)(p.dw[l] - 1) : 0.0f;
  }
  p.blk_start[p.n_levels] = start;
  const dim3 grid(start, p.planes < 65535 ? p.planes : 65535, p.n_frames);
  if (p.src_u8) resize_pyramid_kernel<true><<<grid, kOpThreads, 0, stream>>>(p);
  else          resize_pyramid_kernel<false><<<grid, kOpThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace sde
