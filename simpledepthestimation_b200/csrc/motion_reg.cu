// MotionLearning regularisers (sm_100a), SURVEY.md row N1 -- detectron2/modeling/losses/motion_loss.py:
//   motion_consistency_loss   :7-48   translation cycle term (the rotation term is [B,3,3] arithmetic and stays
//                                      on the host side); backward scatters into t_B2A through the bilinear taps
//                                      in 64-bit fixed point (integer atomics: deterministic)
//   motion_smoothness_loss_fn :51-55
//   motion_sparsity_loss_fn   :58-64
// One thread per pixel; sums go through per-block slots added in a fixed order in fp64 by the last block.
#include "motion_device.cuh"
#include "ops_params.cuh"

namespace sde {

constexpr int kRegThreads = 256;
constexpr int kMcPix = 4;   // motion-consistency kernels: pixels per thread (one block epilogue per 1024 pixels)

// block sum of `v` -> slot; the last block adds all slots of its group (fixed order, fp64) and returns true on
// thread 0 with the total in `total`
__device__ __forceinline__ bool group_sum(float v, float* slots, int slot, int n_slots, unsigned* counter, double& total) {
  __shared__ float red[kRegThreads / 32];
  __shared__ double dred[kRegThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (tid == 0) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kRegThreads / 32; ++k) s += red[k];
    slots[slot] = s;
    __threadfence();
    ticket = atomicAdd(counter, 1u);
  }
  __syncthreads();
  if (ticket != (unsigned)(n_slots - 1)) return false;
  __threadfence();
  double a = 0.0;
  for (int t = tid; t < n_slots; t += kRegThreads) a += (double)__ldcg(slots + t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) dred[wid] = a;
  __syncthreads();
  if (tid != 0) return false;
  total = 0.0;
#pragma unroll
  for (int k = 0; k < kRegThreads / 32; ++k) total += dred[k];
  *counter = 0u;
  return true;
}

// ------------------------------------------------------------------------------------------------ consistency
struct TapSet {
  int off[4];
  float wgt[4];
};

// F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) footprint of normalised (u, v)
__device__ __forceinline__ TapSet grid_taps(float u, float v, int w, int h) {
  const float ix = (u + 1.0f) * 0.5f * (float)(w - 1), iy = (v + 1.0f) * 0.5f * (float)(h - 1);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = ix - fx, ay = iy - fy;
  TapSet t;
  const float wx[2] = {1.0f - ax, ax}, wy[2] = {1.0f - ay, ay};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
    const bool in = xx >= 0 && xx < w && yy >= 0 && yy < h;   // NaN coordinates fall outside
    t.off[k] = in ? yy * w + xx : -1;
    t.wgt[k] = wx[k & 1] * wy[k >> 1];
  }
  return t;
}

struct McTerms {
  float th[3], z[3], ta[3], n, dn, m;
};

// t_ab(p) = [pose_ab translation] + [field t_ab(p)]; t_hat(p) = grid_sample(t_B2A, coords(p)) with zero padding -- the
// constant part of t_B2A contributes its value times the weight of the in-range taps (win)
__device__ __forceinline__ void mcons_terms(const McParams& p, int b, int pix, int hw, const float* R, TapSet& taps, McTerms& t,
                                            float& win) {
  const float2 c = *reinterpret_cast<const float2*>(p.coords + ((size_t)b * hw + pix) * 2);
  taps = grid_taps(c.x, c.y, p.w, p.h);
  win = 0.0f;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (taps.off[q] >= 0) win += taps.wgt[q];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float s = p.pose_ba ? __ldg(p.pose_ba + b * 16 + k * 4 + 3) * win : 0.0f;
    if (p.t_ba) {
      const float* plane = p.t_ba + ((size_t)b * 3 + k) * hw;
      float f = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (taps.off[q] >= 0) f += __ldg(plane + taps.off[q]) * taps.wgt[q];
      s += f;
    }
    t.th[k] = s;
    t.ta[k] = (p.pose_ab ? __ldg(p.pose_ab + b * 16 + k * 4 + 3) : 0.0f) +
              (p.t_ab ? __ldg(p.t_ab + ((size_t)b * 3 + k) * hw + pix) : 0.0f);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) t.z[k] = R[k * 3] * t.th[0] + R[k * 3 + 1] * t.th[1] + R[k * 3 + 2] * t.th[2] + t.ta[k];
  t.n = t.z[0] * t.z[0] + t.z[1] * t.z[1] + t.z[2] * t.z[2];
  t.dn = (t.ta[0] * t.ta[0] + t.ta[1] * t.ta[1] + t.ta[2] * t.ta[2]) +
         (t.th[0] * t.th[0] + t.th[1] * t.th[1] + t.th[2] * t.th[2]) + 1e-24f;
  t.m = __ldg(p.mask + (size_t)b * hw + pix);
}

__global__ void __launch_bounds__(kRegThreads) mcons_fwd_kernel(const __grid_constant__ McParams p) {
  const int b = blockIdx.y, hw = p.h * p.w;
  float R[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = __ldg(p.R + b * 9 + k);
  float e = 0.0f;
#pragma unroll 1
  for (int k = 0; k < kMcPix; ++k) {
    const int pix = (blockIdx.x * kMcPix + k) * kRegThreads + threadIdx.x;
    if (pix >= hw) break;
    TapSet taps;
    McTerms t;
    float win;
    mcons_terms(p, b, pix, hw, R, taps, t, win);
    e += t.m * fdiv(t.n, t.dn);
  }
  double total;
  if (group_sum(e, p.slots, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, p.counters, total))
    p.loss[0] = (float)(total / ((double)p.B * p.h * p.w));
}

__global__ void __launch_bounds__(kRegThreads) mcons_bwd_kernel(const __grid_constant__ McParams p) {
  constexpr int NS = 15;   // d / d R [9], sum d / d t_ab [3], sum d / d t_hat over the in-range taps [3]
  __shared__ float red[NS][kRegThreads / 32];
  __shared__ double dred[NS][kRegThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int b = blockIdx.y, hw = p.h * p.w;
  float R[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = __ldg(p.R + b * 9 + k);
  const float g = __ldg(p.g_loss) / ((float)p.B * (float)p.h * (float)p.w);
  float gs[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) gs[k] = 0.0f;
#pragma unroll 1
  for (int kk = 0; kk < kMcPix; ++kk) {
    const int pix = (blockIdx.x * kMcPix + kk) * kRegThreads + tid;
    if (pix >= hw) break;
    TapSet taps;
    McTerms t;
    float win;
    mcons_terms(p, b, pix, hw, R, taps, t, win);
    const float inv = fdiv(1.0f, t.dn);
    const float ge = g * t.m;                       // d loss / d (n / dn)
    const float cz = 2.0f * ge * inv;               // d / d z = cz * z
    const float cd = -2.0f * ge * t.n * inv * inv;  // d / d t_ab (through dn) = cd * t_ab ; d / d t_hat likewise
    float gth[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float gab = cz * t.z[k] + cd * t.ta[k];
      if (p.g_t_ab) p.g_t_ab[((size_t)b * 3 + k) * hw + pix] = gab;
      gs[9 + k] += gab;
      gth[k] = cz * (R[k] * t.z[0] + R[3 + k] * t.z[1] + R[6 + k] * t.z[2]) + cd * t.th[k];   // R^T z
      gs[12 + k] += gth[k] * win;
#pragma unroll
      for (int c = 0; c < 3; ++c) gs[k * 3 + c] += cz * t.z[k] * t.th[c];
    }
    if (p.scatter) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        long long* plane = p.g_t_ba_fix + ((size_t)b * 3 + k) * hw;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (taps.off[q] >= 0 && gth[k] != 0.0f) fix_add(plane + taps.off[q], gth[k] * taps.wgt[q]);
      }
    }
  }
  // per-sample sums: per-block slots, last block of the sample adds them
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const float v = warp_sum(gs[k]);
    if (lane == 0) red[k][wid] = v;
  }
  __syncthreads();
  if (tid < NS) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < kRegThreads / 32; ++k) v += red[tid][k];
    p.slots[((size_t)b * gridDim.x + blockIdx.x) * 16 + tid] = v;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(p.counters + 1 + b, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  double a[NS];
#pragma unroll
  for (int k = 0; k < NS; ++k) a[k] = 0.0;
  for (int t = tid; t < (int)gridDim.x; t += kRegThreads) {
    const float* s = p.slots + ((size_t)b * gridDim.x + t) * 16;
#pragma unroll
    for (int k = 0; k < NS; ++k) a[k] += (double)__ldcg(s + k);
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NS; ++k) dred[k][wid] = a[k];
  }
  __syncthreads();
  if (tid < NS) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kRegThreads / 32; ++k) v += dred[tid][k];
    if (tid < 9) p.g_R[b * 9 + tid] = (float)v;
    else if (tid < 12) { if (p.g_pose_t_ab) p.g_pose_t_ab[b * 3 + tid - 9] = (float)v; }
    else if (p.g_pose_t_ba) p.g_pose_t_ba[b * 3 + tid - 12] = (float)v;
  }
  if (tid == 0) p.counters[1 + b] = 0u;
}

__global__ void __launch_bounds__(kRegThreads) reg_fix_to_float_kernel(long long* __restrict__ acc, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * kRegThreads + threadIdx.x;
  if (i < n) {
    out[i] = fix_to_float(acc[i]);
    acc[i] = 0;
  }
}

cudaError_t launch_mcons_fwd(const McParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kRegThreads * kMcPix - 1) / (kRegThreads * kMcPix), p.B);
  mcons_fwd_kernel<<<grid, kRegThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_mcons_bwd(const McParams& p, float* g_t_ba, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kRegThreads * kMcPix - 1) / (kRegThreads * kMcPix), p.B);
  mcons_bwd_kernel<<<grid, kRegThreads, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (!p.scatter) return cudaSuccess;
  const size_t n = (size_t)p.B * 3 * p.h * p.w;
  reg_fix_to_float_kernel<<<(unsigned)((n + kRegThreads - 1) / kRegThreads), kRegThreads, 0, stream>>>(p.g_t_ba_fix, g_t_ba, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ smoothness
// mean over [B,C,H-1,W-1] of sqrt(1e-24 + dx^2 + dy^2), dx = m(y,x) - m(y,x-1), dy = m(y,x) - m(y-1,x), y,x >= 1
__global__ void __launch_bounds__(kRegThreads) msmooth_fwd_kernel(const __grid_constant__ MregParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  const float* m = p.field + (size_t)blockIdx.y * hw;
  float s = 0.0f;
  if (pix < hw) {
    const int y = pix / p.w, x = pix - y * p.w;
    if (y >= 1 && x >= 1) {
      const float c = __ldg(m + pix), dx = c - __ldg(m + pix - 1), dy = c - __ldg(m + pix - p.w);
      s = sqrtf(1e-24f + dx * dx + dy * dy);
    }
  }
  double total;
  if (group_sum(s, p.slots, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, p.counters, total))
    p.loss[0] = (float)(total / ((double)p.B * p.C * (p.h - 1) * (p.w - 1)));
}

__global__ void __launch_bounds__(kRegThreads) msmooth_bwd_kernel(const __grid_constant__ MregParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  if (pix >= hw) return;
  const float* m = p.field + (size_t)blockIdx.y * hw;
  const int y = pix / p.w, x = pix - y * p.w;
  const float g = __ldg(p.g_loss) / ((float)p.B * (float)p.C * (float)(p.h - 1) * (float)(p.w - 1));
  auto term = [&](int yy, int xx, float& dx, float& dy, float& rs) {   // differences and 1/s at (yy, xx), yy, xx >= 1
    const int q = yy * p.w + xx;
    const float c = __ldg(m + q);
    dx = c - __ldg(m + q - 1);
    dy = c - __ldg(m + q - p.w);
    rs = rsqrtf(1e-24f + dx * dx + dy * dy);
  };
  float out = 0.0f, dx, dy, rs;
  if (y >= 1 && x >= 1) { term(y, x, dx, dy, rs); out += (dx + dy) * rs; }
  if (y >= 1 && x + 1 < p.w) { term(y, x + 1, dx, dy, rs); out -= dx * rs; }
  if (x >= 1 && y + 1 < p.h) { term(y + 1, x, dx, dy, rs); out -= dy * rs; }
  p.g_field[(size_t)blockIdx.y * hw + pix] = out * g;
}

// ------------------------------------------------------------------------------------------------ sparsity
// pass 1: mean |m| per (b, c) plane -> stats; pass 2: mean(2 a_mean sqrt(|m| / (a_mean + 1e-24) + 1))
__global__ void __launch_bounds__(kRegThreads) msparse_mean_kernel(const __grid_constant__ MregParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  const float v = pix < hw ? fabsf(__ldg(p.field + (size_t)blockIdx.y * hw + pix)) : 0.0f;
  double total;
  if (group_sum(v, p.slots + (size_t)blockIdx.y * gridDim.x, blockIdx.x, gridDim.x, p.counters + 1 + blockIdx.y, total))
    p.stats[blockIdx.y] = (float)(total / (double)hw);
}

__global__ void __launch_bounds__(kRegThreads) msparse_fwd_kernel(const __grid_constant__ MregParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  const float am = p.stats[blockIdx.y];
  float s = 0.0f;
  if (pix < hw) {
    const float a = fabsf(__ldg(p.field + (size_t)blockIdx.y * hw + pix));
    s = 2.0f * am * sqrtf(fdiv(a, am + 1e-24f) + 1.0f);
  }
  double total;
  if (group_sum(s, p.slots, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, p.counters, total))
    p.loss[0] = (float)(total / ((double)p.B * p.C * p.h * p.w));
}

__global__ void __launch_bounds__(kRegThreads) msparse_bwd_kernel(const __grid_constant__ MregParams p) {
  const int hw = p.h * p.w;
  const int pix = blockIdx.x * kRegThreads + threadIdx.x;
  if (pix >= hw) return;
  const float am = p.stats[blockIdx.y];   // detached (motion_loss.py:60)
  const float v = __ldg(p.field + (size_t)blockIdx.y * hw + pix);
  const float a = fabsf(v);
  const float g = __ldg(p.g_loss) / ((float)p.B * (float)p.C * (float)p.h * (float)p.w);
  // d/da [2 am sqrt(a / (am + eps) + 1)] = am / ((am + eps) sqrt(a / (am + eps) + 1))
  const float r = fdiv(am, am + 1e-24f) * rsqrtf(fdiv(a, am + 1e-24f) + 1.0f);
  const float sg = v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f);
  p.g_field[(size_t)blockIdx.y * hw + pix] = g * r * sg;
}

// ------------------------------------------------------------------------------------------------ variance_loss
// variance_loss(depth) = 1 / mean((depth / mean(depth) - 1)^2) over the whole tensor (losses.py:16-18); two passes
// (mean, then the centred second moment) so that a nearly constant depth map does not cancel catastrophically
__global__ void __launch_bounds__(kRegThreads) var_mean_kernel(const __grid_constant__ VarParams p) {
  const long long i = (long long)blockIdx.x * kRegThreads + threadIdx.x;
  const float v = i < p.n ? __ldg(p.depth + i) : 0.0f;
  double total;
  if (group_sum(v, p.slots, blockIdx.x, gridDim.x, p.counters, total)) p.stats[0] = (float)(total / (double)p.n);
}

__global__ void __launch_bounds__(kRegThreads) var_sq_kernel(const __grid_constant__ VarParams p) {
  const long long i = (long long)blockIdx.x * kRegThreads + threadIdx.x;
  const float m = p.stats[0];
  float v = 0.0f;
  if (i < p.n) {
    const float u = fdiv(__ldg(p.depth + i), m) - 1.0f;
    v = u * u;
  }
  double total;
  if (group_sum(v, p.slots, blockIdx.x, gridDim.x, p.counters, total)) {
    const double V = total / (double)p.n;
    p.stats[1] = (float)V;
    p.loss[0] = (float)(1.0 / V);
  }
}

__global__ void __launch_bounds__(kRegThreads) var_bwd_kernel(const __grid_constant__ VarParams p) {
  const long long i = (long long)blockIdx.x * kRegThreads + threadIdx.x;
  if (i >= p.n) return;
  const float m = p.stats[0], V = p.stats[1];
  const float u = fdiv(__ldg(p.depth + i), m) - 1.0f;
  // d(1/V)/d depth_i = -(1/V^2) * 2 (u_i - V) / (N m)      (mean(u) = 0)
  p.g_depth[i] = -__ldg(p.g_loss) * 2.0f * (u - V) / (V * V * (float)p.n * m);
}

cudaError_t launch_var(bool backward, const VarParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)((p.n + kRegThreads - 1) / kRegThreads);
  if (backward) {
    var_bwd_kernel<<<grid, kRegThreads, 0, stream>>>(p);
  } else {
    var_mean_kernel<<<grid, kRegThreads, 0, stream>>>(p);
    var_sq_kernel<<<grid, kRegThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_mreg(int which, const MregParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kRegThreads - 1) / kRegThreads, p.B * p.C);
  switch (which) {
    case 0: msmooth_fwd_kernel<<<grid, kRegThreads, 0, stream>>>(p); break;
    case 1: msmooth_bwd_kernel<<<grid, kRegThreads, 0, stream>>>(p); break;
    case 2:
      msparse_mean_kernel<<<grid, kRegThreads, 0, stream>>>(p);
      msparse_fwd_kernel<<<grid, kRegThreads, 0, stream>>>(p);
      break;
    default: msparse_bwd_kernel<<<grid, kRegThreads, 0, stream>>>(p); break;
  }
  return cudaGetLastError();
}

}  // namespace sde
