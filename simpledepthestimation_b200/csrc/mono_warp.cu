// Warp kernel of the MonoDepth2 loss (sm_100a): project + bilinear gather of every source at every scale,
// one thread per target pixel, written as [B,9,h,w] planes (the `warped` buffers of sde_mono_buffers):
// the warped source (planes 0..2) and, from the same four taps, its derivatives w.r.t. the sample coordinate, times
// q = 1 / (p2 + 1e-6) (planes 3..5: q d/dX, 6..8: q d/dY; zero where nan_to_num / clamp gate the reference's
// gradient, camera.py:184-188) -- so that the backward kernel never touches the source frames again and its warp
// backward is a stream of multiply-adds.  (Two more planes with the sample coordinate were measured: every plane
// costs this kernel ~4 us at cfg2 -- it writes 140 MB per step -- which is more than re-projecting in the backward
// kernel's packed arithmetic.)
//
// The fused loss kernels are stencil kernels with fat threads (128-160 registers, 12-16 warps per SM),
// which is the wrong shape for the gather: it is latency-bound and wants many thin threads.  Run on its own
// with 64-register threads the gather gets 4x the warps per SM and no halo redundancy (the fused forward
// kernel re-gathers 16 % halo positions, the backward kernel 37 %); the loss kernels then receive the warped
// planes through TMA like every other input.  Same device functions, same operation order as the in-kernel
// gather (mono_device.cuh: project_full, bilinear_cell), hence bit-identical planes.
#include "mono_device.cuh"

namespace sde {


struct WarpShared {
  Cam cam;
  Proj proj[SDE_MAX_SOURCES];
};

__global__ void __launch_bounds__(kWarpThreads, 1024 / kWarpThreads) mono_warp_kernel(const __grid_constant__ MonoParams p) {
  __shared__ WarpShared sh;
  __shared__ int next_item;
  const int tid = threadIdx.x;
  SDE_TRACE_BEGIN(p, 0);
  pdl_wait();   // K, pose and depth may come from the kernel ahead; the planes written below may still be read by it
  // flow: the forward kernel may be scheduled as soon as every block of this grid is resident -- its tiles wait for the
  // chunk flags published below, not for the grid
  const bool flow = (p.flow & kFlowWarp) != 0;
  if (flow) pdl_launch_dependents();
  const int total = p.warp_start[p.n_scales];
  int item = 0;
  bool first = true;
  // work items: (scale, sample, chunk of kWarpChunk pixels), largest scale first (mono_params.cuh: work_next)
  while (work_next(p, 0, total, &next_item, item, first)) {
    int s = 0, bid = item;
    while (s + 1 < p.n_scales && bid >= p.warp_start[s + 1]) ++s;
    bid -= p.warp_start[s];
    const int h = p.h[s], w = p.w[s], hw = h * w;
    const int chunks = (hw + kWarpChunk - 1) / kWarpChunk;
    const int b = bid / chunks, chunk = bid - b * chunks;
    if (tid < p.S) {
      Cam cam;
      float k[9];
      load_cam(cam, k, p.K, b, p.sx[s], p.sy[s]);
      if (tid == 0) sh.cam = cam;
      load_proj(sh.proj[tid], k, p.pose[tid], b);
    }
    // this thread's depths are loaded before the barrier, so that their latency overlaps the camera threads' loads
    const int pix0 = chunk * kWarpChunk + tid;
    float dv[kWarpPixPerThread];
#pragma unroll
    for (int it = 0; it < kWarpPixPerThread; ++it) {
      const int pix = pix0 + it * kWarpThreads;
      dv[it] = pix < hw ? decode_depth(__ldg(p.depth[s] + (size_t)b * hw + pix), p.depth_mode, p.min_disp, p.disp_range) : 0.0f;
    }
    // pixel coordinates: one integer division per thread, the other pixels follow by stepping kWarpThreads columns
    int gy = pix0 / w, gx = pix0 - gy * w;
    __syncthreads();
    // camera-space points of this thread's pixels (independent of the source); K^-1 stays in registers meanwhile
    float P[kWarpPixPerThread][3];
    {
      const Cam cam = sh.cam;
#pragma unroll
      for (int it = 0; it < kWarpPixPerThread; ++it) {
        backproject(cam, (float)gx, (float)gy, dv[it], P[it]);
        gx += kWarpThreads;
        while (gx >= w) { gx -= w; ++gy; }
      }
    }
#pragma unroll 1
    for (int j = 0; j < p.S; ++j) {
      // K R and K t of this source: registers for the thread's four pixels (they were 21 shared-memory loads per
      // pixel and source, 15 % of the kernel's L1 wavefronts)
      Proj pj;
#pragma unroll
      for (int k = 0; k < 9; ++k) pj.m[k] = sh.proj[j].m[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) pj.tau[k] = sh.proj[j].tau[k];
      const float* __restrict__ srcb = p.source[s][j] + (size_t)b * 3 * hw;
      float* __restrict__ dstb = p.warped[s][j] + (size_t)b * kSavedPlanes * hw;
#pragma unroll
      for (int it = 0; it < kWarpPixPerThread; ++it) {
        // one item per block: once every block is on its last pixel, the next kernel of the stream may be scheduled (it
        // waits for this grid to complete before it touches memory): its launch latency overlaps this kernel's tail
        if (!flow && !((p.persist >> 0) & 1) && j == p.S - 1 && it == kWarpPixPerThread - 1) pdl_launch_dependents();
        const int pix = pix0 + it * kWarpThreads;
        if (pix >= hw) break;
        float den, X, Y, q;
        project_point(pj, P[it], den, X, Y, q);
        const Cell cell = bilinear_cell(X, Y, w, h);
        const float bx = 1.0f - cell.ax, by = 1.0f - cell.ay;
        const float w00 = bx * by, w01 = cell.ax * by, w10 = bx * cell.ay, w11 = cell.ax * cell.ay;
        const float* src = srcb + cell.off;
        float* dst = dstb + pix;
        // gradient gates of nan_to_num and clamp (closed interval): false for NaN / +-inf
        const bool gate_x = X >= 0.0f && X <= (float)(w - 1);
        const bool gate_y = Y >= 0.0f && Y <= (float)(h - 1);
        // d (X, Y) / d p carries the factor q = 1 / (p2 + 1e-6): folded into the derivative planes here, where the
        // issue slots are free (this kernel waits on L1), so that the backward kernel is a stream of multiply-adds
        float t[3][4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // scalar taps: a 16-byte load per lane is served quarter-warp by quarter-warp and touched MORE L1
          // sectors on the scattered addresses of this gather (measured: 100 us against 62 us)
          const float* q4 = src + c * hw;
          t[c][0] = __ldg(q4); t[c][1] = __ldg(q4 + 1); t[c][2] = __ldg(q4 + w); t[c][3] = __ldg(q4 + w + 1);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // ATen's accumulation order: nw, ne, sw, se
          dst[c * hw] = t[c][0] * w00 + t[c][1] * w01 + t[c][2] * w10 + t[c][3] * w11;
          const float ddx = (t[c][1] - t[c][0]) * by + (t[c][3] - t[c][2]) * cell.ay;
          const float ddy = (t[c][2] - t[c][0]) * bx + (t[c][3] - t[c][1]) * cell.ax;
          dst[(3 + c) * hw] = gate_x ? ddx * q : 0.0f;
          dst[(6 + c) * hw] = gate_y ? ddy * q : 0.0f;
        }
      }
    }
    if (flow) {
      __syncthreads();
      if (tid == 0) flag_publish(p.warp_flag + item);
    }
  }
  if (!flow && ((p.persist >> 0) & 1)) pdl_launch_dependents();
  work_leave(p, 0);
  SDE_TRACE_MARK(p, 0, 1);
}

cudaError_t launch_mono_warp(const MonoParams& p, cudaStream_t stream) {
  int grid = p.warp_start[p.n_scales];
  if (grid == 0) return cudaSuccess;
  if (((p.persist >> 0) & 1)) {
    static unsigned slots = 0;
    if (!slots) slots = resident_ctas(mono_warp_kernel, kWarpThreads, 0);
    if (slots && (unsigned)grid > slots) grid = (int)slots;
  }
  return launch_chained(0, mono_warp_kernel, (unsigned)grid, kWarpThreads, 0, stream, p);
}

}  // namespace sde
