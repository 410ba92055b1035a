// Regularisers of the MotionLearning residual translation field, fused (sm_100a).  Replaces, per
// direction, the block of MotionLearningModel.forward that normalises the field and applies both regularisers
// (detectron2/modeling/meta_arch/MotionLearning.py:203-220; modeling/losses/motion_loss.py:51-64):
//   t = pose[:, :3, 3] + m                     overall translation field (:143-147), never materialised here
//   s = 1 / sqrt(3 mean_{c,h,w}(t^2) + 1e-12)  per sample (:205-208; NOT detached: gradients reach m and the pose)
//   mn = m s
//   motion_smoothness_loss_fn(mn) = mean_{B,3,h-1,w-1} sqrt(1e-24 + dx^2 + dy^2)
//   motion_sparsity_loss_fn(mn)   = mean_{B,3,h,w} 2 abar sqrt(|mn| / (abar + 1e-24) + 1), abar = mean_{h,w}|mn| detached
// Forward: a statistics pass (sum t^2, sum t_c, sum |m_c| per sample) and a loss pass that also accumulates
// E = sum_p (d loss / d mn_p) mn_p per sample and term -- which is all the backward pass needs to route the gradient
// through s: d loss / d s = E / s.  Backward: one element-wise pass, no reduction (d / d pose_t = -kappa sum_p t_c uses
// the saved sums).  Per-block slots, fixed-order fp64 sums by the last block of a sample: deterministic.
#include "motion_device.cuh"
#include "ops_params.cuh"

namespace sde {

constexpr int kFieldThreads = 256;
constexpr int kFieldPix = 8;      // pixels per thread: the block epilogue (reduction, fence, ticket) is paid once per 2048 pixels
constexpr int kFieldChunk = kFieldThreads * kFieldPix;
constexpr int kFieldStats = 12;   // per sample: s, sum t_c [3], abar_c [3], E_smooth, E_sparse, (3 spare)

// Sums N values over the block and publishes them in the block's slot; returns true in EVERY thread of the last block
// of the group (per-sample ticket), which then owns the group's slots.
template <int N>
__device__ __forceinline__ bool publish_sums(const float (&v)[N], float* slot, unsigned* ticket_ctr, int group_size) {
  __shared__ float red[N][kFieldThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float s = warp_sum(v[k]);
    if (lane == 0) red[k][wid] = s;
  }
  __syncthreads();
  if (tid < N) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kFieldThreads / 32; ++k) s += red[tid][k];
    slot[tid] = s;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(ticket_ctr, 1u);
  __syncthreads();
  if (ticket != (unsigned)(group_size - 1)) return false;
  __threadfence();
  return true;
}

// adds the group's slots (stride N floats) in a fixed order in fp64; the totals arrive in thread 0
template <int N>
__device__ __forceinline__ void collect_sums(const float* slots, int group_size, double (&total)[N]) {
  __shared__ double dred[N][kFieldThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double a[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] = 0.0;
  for (int t = tid; t < group_size; t += kFieldThreads) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] += (double)__ldcg(slots + (size_t)t * N + k);
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) dred[k][wid] = a[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    total[k] = 0.0;
#pragma unroll
    for (int q = 0; q < kFieldThreads / 32; ++q) total[k] += dred[k][q];
  }
}

__device__ __forceinline__ float pose_t(const float* pose, int b, int c) { return pose ? __ldg(pose + b * 16 + c * 4 + 3) : 0.0f; }

__global__ void __launch_bounds__(kFieldThreads) mfield_stats_kernel(const __grid_constant__ MfieldParams p) {
  const int b = blockIdx.y, hw = p.h * p.w;
  float v[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};   // sum t^2, sum t_c, sum |m_c|, spare
  const float T[3] = {pose_t(p.pose, b, 0), pose_t(p.pose, b, 1), pose_t(p.pose, b, 2)};
#pragma unroll
  for (int k = 0; k < kFieldPix; ++k) {
    const int pix = blockIdx.x * kFieldChunk + k * kFieldThreads + threadIdx.x;
    if (pix < hw) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float m = __ldg(p.field + ((size_t)b * 3 + c) * hw + pix);
        const float t = T[c] + m;
        v[0] += t * t;
        v[1 + c] += t;
        v[4 + c] += fabsf(m);
      }
    }
  }
  float* slots = p.slots + (size_t)b * gridDim.x * 8;
  if (!publish_sums<8>(v, slots + (size_t)blockIdx.x * 8, p.counters + 1 + b, gridDim.x)) return;
  double tot[8];
  collect_sums<8>(slots, gridDim.x, tot);
  if (threadIdx.x == 0) {
    // t_scale = mean_{c,h,w}(t^2) * 3 = sum t^2 / (h w)
    const float t_scale = (float)(tot[0] / (double)hw);
    const float s = 1.0f / sqrtf(t_scale + 1e-12f);
    float* st = p.stats + b * kFieldStats;
    st[0] = s;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      st[1 + c] = (float)tot[1 + c];
      st[4 + c] = (float)(tot[4 + c] / (double)hw) * s;   // mean |mn_c|
    }
    p.counters[1 + b] = 0u;
  }
}

__device__ __forceinline__ float sparsity_ratio(float a, float am) {   // d/da [2 am sqrt(a / (am + eps) + 1)]
  return fdiv(am, am + 1e-24f) * rsqrtf(fdiv(a, am + 1e-24f) + 1.0f);
}

__global__ void __launch_bounds__(kFieldThreads) mfield_loss_kernel(const __grid_constant__ MfieldParams p) {
  const int b = blockIdx.y, hw = p.h * p.w;
  const float* st = p.stats + b * kFieldStats;
  const float s = st[0];
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // smoothness, E_smooth, sparsity, E_sparse
#pragma unroll 2
  for (int k = 0; k < kFieldPix; ++k) {
    const int pix = blockIdx.x * kFieldChunk + k * kFieldThreads + threadIdx.x;
    if (pix >= hw) break;
    const int y = pix / p.w, x = pix - y * p.w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* m = p.field + ((size_t)b * 3 + c) * hw;
      const float mc = __ldg(m + pix) * s;
      if (y >= 1 && x >= 1) {
        const float dx = mc - __ldg(m + pix - 1) * s, dy = mc - __ldg(m + pix - p.w) * s;
        const float q = dx * dx + dy * dy, f = sqrtf(1e-24f + q);
        v[0] += f;
        v[1] += q / f;
      }
      const float a = fabsf(mc), am = st[4 + c];
      v[2] += 2.0f * am * sqrtf(fdiv(a, am + 1e-24f) + 1.0f);
      v[3] += a * sparsity_ratio(a, am);
    }
  }
  float* slots = p.slots + (size_t)b * gridDim.x * 8;
  if (!publish_sums<4>(v, slots + (size_t)blockIdx.x * 4, p.counters + 1 + b, gridDim.x)) return;
  double tot[4];
  collect_sums<4>(slots, gridDim.x, tot);
  if (threadIdx.x == 0) {
    const double n_sm = (double)p.B * 3.0 * (p.h - 1) * (p.w - 1), n_sp = (double)p.B * 3.0 * p.h * p.w;
    float* stw = p.stats + b * kFieldStats;
    stw[7] = (float)(tot[1] / n_sm);
    stw[8] = (float)(tot[3] / n_sp);
    p.fin[b * 2] = tot[0] / n_sm;
    p.fin[b * 2 + 1] = tot[2] / n_sp;
    p.counters[1 + b] = 0u;
    __threadfence();
    if (atomicAdd(p.counters, 1u) == (unsigned)(p.B - 1)) {
      __threadfence();
      double sm = 0.0, sp = 0.0;
      for (int k = 0; k < p.B; ++k) { sm += __ldcg(p.fin + k * 2); sp += __ldcg(p.fin + k * 2 + 1); }
      p.losses[0] = (float)sm;
      p.losses[1] = (float)sp;
      p.counters[0] = 0u;
    }
  }
}

__global__ void __launch_bounds__(kFieldThreads) mfield_bwd_kernel(const __grid_constant__ MfieldParams p) {
  const int b = blockIdx.y, hw = p.h * p.w;
  const float* st = p.stats + b * kFieldStats;
  const float s = st[0];
  const float g_sm = __ldg(p.g_losses), g_sp = __ldg(p.g_losses + 1);
  // d loss / d t_{c,p} through the normaliser: -(g_sm E_sm + g_sp E_sp) s^2 t / (h w)
  const float kappa = (g_sm * st[7] + g_sp * st[8]) * s * s / (float)hw;
  if (blockIdx.x == 0 && threadIdx.x < 3 && p.g_pose_t) p.g_pose_t[b * 3 + threadIdx.x] = -kappa * st[1 + threadIdx.x];
  const float c_sm = g_sm / ((float)p.B * 3.0f * (float)(p.h - 1) * (float)(p.w - 1));
  const float c_sp = g_sp / ((float)p.B * 3.0f * (float)p.h * (float)p.w);
#pragma unroll 2
  for (int k = 0; k < kFieldPix; ++k) {
    const int pix = blockIdx.x * kFieldChunk + k * kFieldThreads + threadIdx.x;
    if (pix >= hw) break;
    const int y = pix / p.w, x = pix - y * p.w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* m = p.field + ((size_t)b * 3 + c) * hw;
      auto term = [&](int q, float& dx, float& dy, float& rs) {   // differences of mn and 1 / f at pixel q (row, column >= 1)
        const float cc = __ldg(m + q) * s;
        dx = cc - __ldg(m + q - 1) * s;
        dy = cc - __ldg(m + q - p.w) * s;
        rs = rsqrtf(1e-24f + dx * dx + dy * dy);
      };
      float G = 0.0f, dx, dy, rs;
      if (y >= 1 && x >= 1) { term(pix, dx, dy, rs); G += (dx + dy) * rs; }
      if (y >= 1 && x + 1 < p.w) { term(pix + 1, dx, dy, rs); G -= dx * rs; }
      if (x >= 1 && y + 1 < p.h) { term(pix + p.w, dx, dy, rs); G -= dy * rs; }
      G *= c_sm;
      const float mv = __ldg(m + pix);
      const float a = fabsf(mv) * s;
      const float sg = mv > 0.0f ? 1.0f : (mv < 0.0f ? -1.0f : 0.0f);
      G += c_sp * sparsity_ratio(a, st[4 + c]) * sg;
      const float t = pose_t(p.pose, b, c) + mv;
      p.g_field[((size_t)b * 3 + c) * hw + pix] = s * G - kappa * t;
    }
  }
}

cudaError_t launch_mfield(bool backward, const MfieldParams& p, cudaStream_t stream) {
  const dim3 grid((p.h * p.w + kFieldChunk - 1) / kFieldChunk, p.B);
  if (backward) {
    mfield_bwd_kernel<<<grid, kFieldThreads, 0, stream>>>(p);
  } else {
    mfield_stats_kernel<<<grid, kFieldThreads, 0, stream>>>(p);
    mfield_loss_kernel<<<grid, kFieldThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace sde
