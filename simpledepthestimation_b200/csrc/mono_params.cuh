// Kernel-side parameter block of the fused MonoDepth2 loss (one launch covers every scale).
#pragma once
#include <cuda.h>

#include "sde_common.cuh"

namespace sde {

constexpr int kTileW = 64;   // 32 lanes x 2 pixels (one f2 per lane)
constexpr int kTileH = 16;   // 4 warps x 4 rows
constexpr int kThreads = 128;
constexpr int kRowsPerWarp = 4;
// planes per sample of a `warped` buffer: warp[3], q d/dX[3], q d/dY[3] with q = 1 / (p2 + 1e-6) (include/sde_loss.h)
constexpr int kSavedPlanes = SDE_MONO_SAVED_PLANES;
// backward: a CTA recomputes SSIM on a kTileW x kTileH block of window centres and emits
// gradients for its interior (the 3x3 adjoint needs one ring of neighbours)
// warp kernel (mono_warp.cu): threads per block, pixels per thread, pixels per block.  Measured at cfg2 (full step,
// same box): 128 x 4: 274.7 us, 256 x 4: 273.8, 512 x 4: 270.9, 1024 x 4: 273.9, 512 x 2: 274.0, 512 x 8: 281.7
#ifndef SDE_WARP_THREADS
#define SDE_WARP_THREADS 512
#endif
constexpr int kWarpThreads = SDE_WARP_THREADS;
#ifndef SDE_WARP_PIX
#define SDE_WARP_PIX 4
#endif
constexpr int kWarpPixPerThread = SDE_WARP_PIX;
constexpr int kWarpChunk = kWarpThreads * kWarpPixPerThread;
constexpr int kBwdW = kTileW - 4;   // 60: a multiple of 4, so that TMA boxes of the backward tiles start 16-byte aligned
constexpr int kBwdH = kTileH - 2;   // 14

struct MonoParams {
  int B, n_scales, S;
  int NB;                 // batch size the means are taken over (>= B: sde_mono_desc.norm_batch)
  int h[SDE_MAX_SCALES], w[SDE_MAX_SCALES];
  int tiles_x[SDE_MAX_SCALES], tiles_y[SDE_MAX_SCALES];
  int tile_start[SDE_MAX_SCALES + 1];  // CTA index where scale s starts; [n_scales] = grid size
  float sx[SDE_MAX_SCALES], sy[SDE_MAX_SCALES];  // w_i / W, h_i / H
  const float* target[SDE_MAX_SCALES];
  const float* source[SDE_MAX_SCALES][SDE_MAX_SOURCES];
  const float* depth[SDE_MAX_SCALES];
  const float* K;
  const float* pose[SDE_MAX_SOURCES];
  uint8_t* argmin[SDE_MAX_SCALES];
  float* warped[SDE_MAX_SCALES][SDE_MAX_SOURCES];   // optional [B,kSavedPlanes,h,w] (nullptr = recompute)
  float* smooth_g[SDE_MAX_SCALES];                  // optional [B,h,w]: local smoothness gradient kept by the forward pass
  float* losses;
  float* stats;           // [n_scales*B][2] = (mean inverse depth, per-image smoothness)
  float ssim_w, l1_w, c1, c2;
  float smooth_scale[SDE_MAX_SCALES];  // scale_w * SMOOTHNESS_WEIGHT / n_scales (MonoDepth2.py:80,103-105)
  // reciprocals of the tile counts (forward: x, y; backward: x, y) for the division-free tile decoding
  float rtiles[SDE_MAX_SCALES][4];
  float inv_norm[SDE_MAX_SCALES];      // 1 / (n_scales * B * h * w [* candidates for 'mean']): d rec_loss / d pe of a selected pixel
  unsigned flags;
  // what depth[] holds (SDE_DEPTH_IS_*): disparity / logits are decoded where a kernel stages them
  int depth_mode;
  float min_disp, disp_range;   // 1 / max_depth, 1 / min_depth - 1 / max_depth
  // forward workspace
  float* partials;        // [grid][4]
  double* fin;            // [n_scales*B][2]
  unsigned* counter;       // global ticket: (scale, image) pairs finished
  unsigned* img_counter;   // [n_scales*B] tiles finished per (scale, image)
  // backward
  const float* grad_losses;
  float* grad_depth[SDE_MAX_SCALES];
  float* grad_pose[SDE_MAX_SOURCES];
  float* pose_partials;   // [bwd grid][warps per CTA][S][12]
  unsigned* smp_counter;   // [B] backward tiles finished per sample (all scales)
  int btiles_x[SDE_MAX_SCALES], btiles_y[SDE_MAX_SCALES];
  int btile_start[SDE_MAX_SCALES + 1];
  int tma[SDE_MAX_SCALES];   // tile planes of this scale are staged by TMA (tensor maps in MonoTma are valid)
  int prewarp[SDE_MAX_SCALES];            // forward: the loss kernel takes warped[s][*] (filled by the warp kernel) through TMA
  int warp_start[SDE_MAX_SCALES + 1];     // warp kernel: first block of scale s (blocks of kWarpChunk pixels per sample)
  // Tile-level dependencies between the kernels of a step (sde_api.cu: "flow"), instead of whole-grid ones:
  //   kFlowWarp   the warp kernel publishes one flag per block (= chunk of kWarpChunk pixels of one image); a forward
  //               tile waits for the chunks that hold its rows only, so forward tiles run while the warp kernel drains
  //   kFlowImage  the forward tile that completes an image publishes the image's flag (its statistics, argmin bytes and
  //               smoothness gradients are final); a backward tile waits for its image only
  // Flags live in the workspace, are zero between calls, and are cleared by the consumer side (the forward tile that
  // completes an image clears the image's chunk flags, the backward tile that completes a sample its image flags).
  unsigned flow;
  unsigned* warp_flag;    // [warp grid]
  unsigned* img_flag;     // [n_scales*B]
  // Persistent CTAs (bit k of `persist`: 0 warp, 1 forward, 2 backward kernel): the grid is one CTA per resident slot
  // (occupancy x SMs) and every CTA takes work items -- chunks of the warp kernel, tiles of the loss kernels, in list
  // order -- from a queue in the workspace until it is empty.  queue[2k] = next item, queue[2k + 1] = CTAs that have
  // left (the last one clears both).  Bit clear: one item per CTA, grid = number of items.
  // Measured (profiles/r2_notes.md): per-CTA time stamps show a slot empty for 2.1-2.3 us between a CTA's last
  // instruction and its successor's first, but most of that is the CTA's other warps finishing and its stores draining,
  // which a loop pays as well -- persistent CTAs gain 3.6 % for the backward kernel at batch 96, nothing at batch 12,
  // and cost the warp / forward kernels 2 % (queue atomics and a barrier per item).  Default: backward only.
  int persist;
  unsigned* queue;        // [6]
#ifdef SDE_TRACE
  unsigned long long* trace;
#endif
};
constexpr unsigned kFlowWarp = 1u, kFlowImage = 2u;

// Work-item loop of the three kernels: `persist` -- take the next item of queue k (thread 0 fetches, `slot` broadcasts
// it; the barrier also separates the previous item's shared-memory traffic from the next one's), else the one item
// blockIdx.x.  Returns false when there is nothing left; the caller then calls work_leave().
__device__ __forceinline__ bool work_next(const MonoParams& p, int k, int total, int* slot, int& item, bool& first) {
  if ((p.persist >> k) & 1) {
    if (threadIdx.x == 0) *slot = (int)atomicAdd(p.queue + 2 * k, 1u);
    __syncthreads();
    item = *slot;
  } else {
    if (!first) return false;
    item = (int)blockIdx.x;
  }
  first = false;
  return item < total;
}
__device__ __forceinline__ void work_leave(const MonoParams& p, int k) {
  if (((p.persist >> k) & 1) && threadIdx.x == 0) {
    if (atomicAdd(p.queue + 2 * k + 1, 1u) == gridDim.x - 1) {   // every CTA has fetched its last (out-of-range) item
      p.queue[2 * k] = 0u;
      p.queue[2 * k + 1] = 0u;
    }
  }
}

// Tensor maps over the [planes, h, w] inputs of every scale, box {68, 18, 1} (tma.cuh).  Second kernel
// parameter: the copy engine reads the descriptors straight from the parameter space.
struct alignas(64) MonoTma {
  CUtensorMap target[SDE_MAX_SCALES];
  CUtensorMap depth[SDE_MAX_SCALES];
  CUtensorMap source[SDE_MAX_SCALES][SDE_MAX_SOURCES];
  CUtensorMap warped[SDE_MAX_SCALES][SDE_MAX_SOURCES];
};

// depth of one element of depth[]: depth itself, disp_to_depth of a disparity
// (depth_decoder.py:9-18: scaled = min_disp + (max_disp - min_disp) * disp, depth = 1 / scaled -- separately rounded
// multiply, add and IEEE division, as the reference's three ATen operations), or of softplus(logit) (depth_decoder.py:108)
__device__ __forceinline__ float decode_depth(float v, int mode, float min_disp, float range) {
  if (mode == SDE_DEPTH_IS_DEPTH) return v;
  if (mode == SDE_DEPTH_IS_LOGIT) v = v > 20.0f ? v : log1pf(expf(v));   // nn.Softplus(beta=1, threshold=20)
  return __fdiv_rn(1.0f, __fadd_rn(min_disp, __fmul_rn(range, v)));
}
// d depth / d (what depth[] holds), given the decoded depth and the raw element
__device__ __forceinline__ float decode_depth_grad(float depth, float raw, int mode, float range) {
  if (mode == SDE_DEPTH_IS_DEPTH) return 1.0f;
  float g = -range * depth * depth;
  if (mode == SDE_DEPTH_IS_LOGIT && raw <= 20.0f) g *= 1.0f / (1.0f + expf(-raw));   // softplus' = sigmoid
  return g;
}

struct TileCoord {
  int s, b, x0, y0;
};

// t / d and t % d for 0 <= t < 2^22 through the host-rounded reciprocal rd = 1.0f / d: (t + 0.5) / d is at least
// 0.5 / d away from an integer while the rounding error of (t + 0.5) * rd stays below (t / d) * 2^-23, so the
// truncation is exact (the host rejects descriptors with 2^22 or more tiles per scale).
__device__ __forceinline__ void small_divmod(int t, int d, float rd, int& q, int& r) {
  q = (int)(((float)t + 0.5f) * rd);
  r = t - q * d;
}

__device__ __forceinline__ TileCoord decode_tile(const MonoParams& p, int bid) {
  int s = 0;
  while (s + 1 < p.n_scales && bid >= p.tile_start[s + 1]) ++s;
  int t = bid - p.tile_start[s], tx, ty, b;
  small_divmod(t, p.tiles_x[s], p.rtiles[s][0], t, tx);
  small_divmod(t, p.tiles_y[s], p.rtiles[s][1], b, ty);
  TileCoord c;
  c.s = s;
  c.b = b;
  c.x0 = tx * kTileW;
  c.y0 = ty * kTileH;
  return c;
}

// backward tiles: (x0, y0) is the first pixel of the gradient (interior) region
__device__ __forceinline__ TileCoord decode_btile(const MonoParams& p, int bid) {
  int s = 0;
  while (s + 1 < p.n_scales && bid >= p.btile_start[s + 1]) ++s;
  int t = bid - p.btile_start[s], tx, ty, b;
  small_divmod(t, p.btiles_x[s], p.rtiles[s][2], t, tx);
  small_divmod(t, p.btiles_y[s], p.rtiles[s][3], b, ty);
  TileCoord c;
  c.s = s;
  c.b = b;
  c.x0 = tx * kBwdW;
  c.y0 = ty * kBwdH;
  return c;
}

}  // namespace sde
