// Producers and supervised terms either side of the fused loss, sm_100a:
//   silog_loss      detectron2/modeling/losses/losses.py:5-13   (caller MonoDepth2.py:107-110)
//   disp_to_depth   detectron2/layers/depth_decoder.py:9-18     (callers DepthResNet.py:41,57, PackNet01.py:109)
//   pose_vec2mat    detectron2/geometry/pose_utils.py:98-137    (callers PoseNet.py:63, GooglePoseNet.py:85,206)
// Elementwise / small-reduction kernels; reductions are fixed-order (per-block slots, last block adds them in fp64).
#include "sde_common.cuh"
#include "ops_params.cuh"

namespace sde {

constexpr int kDepthThreads = 256;

// ------------------------------------------------------------------------------------------------ silog_loss
// mask = gt > 1; d = log(est[mask]) - log(gt[mask]); loss = sqrt(mean(d^2) - vf * mean(d)^2) * 10
__device__ __forceinline__ bool silog_term(const SilogParams& p, long long i, float& d) {
  if (i >= p.n) return false;
  const float g = __ldg(p.gt + i);
  if (!(g > 1.0f)) return false;
  d = logf(__ldg(p.est + i)) - logf(g);
  return true;
}

__global__ void __launch_bounds__(kDepthThreads) silog_fwd_kernel(const __grid_constant__ SilogParams p) {
  __shared__ float red[3][kDepthThreads / 32];
  __shared__ double dred[3][kDepthThreads / 32];
  __shared__ unsigned ticket;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  float d = 0.0f;
  const bool m = silog_term(p, (long long)blockIdx.x * kDepthThreads + tid, d);
  float v[3] = {m ? d : 0.0f, m ? d * d : 0.0f, m ? 1.0f : 0.0f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) red[k][wid] = v[k];
  }
  __syncthreads();
  if (tid < 3) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kDepthThreads / 32; ++k) s += red[tid][k];
    p.slots[(size_t)blockIdx.x * 3 + tid] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(p.counter, 1u);
  __syncthreads();
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  double a[3] = {0.0, 0.0, 0.0};
  for (unsigned t = tid; t < gridDim.x; t += kDepthThreads) {
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] += (double)__ldcg(p.slots + (size_t)t * 3 + k);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if (lane == 0) dred[k][wid] = a[k];
  }
  __syncthreads();
  if (tid != 0) return;
  double t3[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    t3[k] = 0.0;
#pragma unroll
    for (int q = 0; q < kDepthThreads / 32; ++q) t3[k] += dred[k][q];
  }
  // an empty mask gives 0 / 0 = NaN, as the mean of an empty tensor does in the reference
  const double m1 = t3[0] / t3[2], m2 = t3[1] / t3[2];
  const double root = sqrt(m2 - (double)p.vf * m1 * m1);
  p.loss[0] = (float)(root * 10.0);
  p.stats[0] = (float)m1;
  p.stats[1] = (float)root;
  p.stats[2] = (float)t3[2];
  *p.counter = 0u;
}

__global__ void __launch_bounds__(kDepthThreads) silog_bwd_kernel(const __grid_constant__ SilogParams p) {
  const long long i = (long long)blockIdx.x * kDepthThreads + threadIdx.x;
  if (i >= p.n) return;
  float d = 0.0f, g = 0.0f;
  if (silog_term(p, i, d)) {
    // d sqrt(m2 - vf m1^2) / d d_i = (d_i - vf m1) / (n root);  d d_i / d est_i = 1 / est_i
    const float m1 = p.stats[0], root = p.stats[1], cnt = p.stats[2];
    g = __ldg(p.g_loss) * 10.0f * (d - p.vf * m1) / (cnt * root * __ldg(p.est + i));
  }
  p.g_est[i] = g;
}

cudaError_t launch_silog(bool backward, const SilogParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)((p.n + kDepthThreads - 1) / kDepthThreads);
  if (backward) silog_bwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  else          silog_fwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ disp_to_depth
// scaled = min_disp + (max_disp - min_disp) * disp;  depth = 1 / scaled   (min_disp = 1 / max_depth, max_disp = 1 / min_depth)
__global__ void __launch_bounds__(kDepthThreads) disp_fwd_kernel(const __grid_constant__ DispParams p) {
  const long long i = (long long)blockIdx.x * kDepthThreads + threadIdx.x;
  if (i >= p.n) return;
  const float sd = p.min_disp + p.range * __ldg(p.disp + i);
  if (p.scaled) p.scaled[i] = sd;
  p.depth[i] = 1.0f / sd;
}

__global__ void __launch_bounds__(kDepthThreads) disp_bwd_kernel(const __grid_constant__ DispParams p) {
  const long long i = (long long)blockIdx.x * kDepthThreads + threadIdx.x;
  if (i >= p.n) return;
  const float sd = p.min_disp + p.range * __ldg(p.disp + i);
  const float inv = 1.0f / sd;
  float g = 0.0f;
  if (p.g_depth) g = -__ldg(p.g_depth + i) * inv * inv;   // d (1 / sd) / d sd
  if (p.g_scaled) g += __ldg(p.g_scaled + i);
  p.g_disp[i] = g * p.range;
}

cudaError_t launch_disp(bool backward, const DispParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)((p.n + kDepthThreads - 1) / kDepthThreads);
  if (backward) disp_bwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  else          disp_fwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ pose_vec2mat
// vec = [tx, ty, tz, rx, ry, rz];  R = Rx(rx) Ry(ry) Rz(rz) (euler2mat, pose_utils.py:98-127);  T = [[R, t], [0, 1]]
__device__ __forceinline__ void mat3_mul(const float a[9], const float b[9], float c[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

struct Euler {
  float X[9], Y[9], Z[9];      // the three factors
  float dX[9], dY[9], dZ[9];   // their derivatives w.r.t. the angle
};

__device__ __forceinline__ void euler_factors(float x, float y, float z, Euler& e) {
  float sx, cx, sy, cy, sz, cz;
  sincosf(x, &sx, &cx);
  sincosf(y, &sy, &cy);
  sincosf(z, &sz, &cz);
  const float Z[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1}, dZ[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
  const float Y[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, dY[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
  const float X[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, dX[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
#pragma unroll
  for (int k = 0; k < 9; ++k) { e.X[k] = X[k]; e.Y[k] = Y[k]; e.Z[k] = Z[k]; e.dX[k] = dX[k]; e.dY[k] = dY[k]; e.dZ[k] = dZ[k]; }
}

__global__ void __launch_bounds__(kDepthThreads) posevec_fwd_kernel(const __grid_constant__ PoseVecParams p) {
  const int b = blockIdx.x * kDepthThreads + threadIdx.x;
  if (b >= p.B) return;
  const float* v = p.vec + b * 6;
  Euler e;
  euler_factors(v[3], v[4], v[5], e);
  float XY[9], R[9];
  mat3_mul(e.X, e.Y, XY);
  mat3_mul(XY, e.Z, R);
  float* T = p.mat + b * 16;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j];
    T[i * 4 + 3] = v[i];
  }
  T[12] = 0.0f; T[13] = 0.0f; T[14] = 0.0f; T[15] = 1.0f;
}

__global__ void __launch_bounds__(kDepthThreads) posevec_bwd_kernel(const __grid_constant__ PoseVecParams p) {
  const int b = blockIdx.x * kDepthThreads + threadIdx.x;
  if (b >= p.B) return;
  const float* v = p.vec + b * 6;
  const float* G = p.g_mat + b * 16;
  Euler e;
  euler_factors(v[3], v[4], v[5], e);
  float gR[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) gR[i * 3 + j] = G[i * 4 + j];
  auto dot9 = [](const float a[9], const float c[9]) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) s += a[k] * c[k];
    return s;
  };
  float t1[9], t2[9];
  float* gv = p.g_vec + b * 6;
  gv[0] = G[3]; gv[1] = G[7]; gv[2] = G[11];
  mat3_mul(e.dX, e.Y, t1); mat3_mul(t1, e.Z, t2); gv[3] = dot9(gR, t2);   // dR/drx = dX Y Z
  mat3_mul(e.X, e.dY, t1); mat3_mul(t1, e.Z, t2); gv[4] = dot9(gR, t2);   // dR/dry = X dY Z
  mat3_mul(e.X, e.Y, t1); mat3_mul(t1, e.dZ, t2); gv[5] = dot9(gR, t2);   // dR/drz = X Y dZ
}

cudaError_t launch_posevec(bool backward, const PoseVecParams& p, cudaStream_t stream) {
  const unsigned grid = (unsigned)((p.B + kDepthThreads - 1) / kDepthThreads);
  if (backward) posevec_bwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  else          posevec_fwd_kernel<<<grid, kDepthThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace sde
