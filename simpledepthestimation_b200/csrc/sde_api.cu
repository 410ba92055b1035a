// C ABI of libsde_loss.so (see include/sde_loss.h): argument validation, parameter-block
// construction and kernel launches.  No torch types, no allocation, no synchronisation.
#include <stdlib.h>
#include <string.h>

#include <math.h>

#include "motion_device.cuh"
#include "ops_params.cuh"

namespace sde {
cudaError_t launch_mono_fwd(const MonoParams& p, const MonoTma& t, cudaStream_t stream);
cudaError_t launch_mono_bwd(const MonoParams& p, const MonoTma& t, cudaStream_t stream);
cudaError_t launch_mono_warp(const MonoParams& p, cudaStream_t stream);
cudaError_t launch_motion_fwd(const MotionParams& p, const MotionTma& t, cudaStream_t stream);
cudaError_t launch_motion_bwd(const MotionParams& p, const MotionTma& t, cudaStream_t stream);
cudaError_t launch_vs_fwd(const VsParams& p, cudaStream_t stream);
cudaError_t launch_vs_bwd(const VsParams& p, float* g_image, cudaStream_t stream);
cudaError_t launch_ssim_fwd(const SsimParams& p, cudaStream_t stream);
cudaError_t launch_ssim_bwd(const SsimParams& p, cudaStream_t stream);
cudaError_t launch_smooth_fwd(const SmoothParams& p, cudaStream_t stream);
cudaError_t launch_smooth_bwd(const SmoothParams& p, cudaStream_t stream);
cudaError_t launch_resize_bilinear(const float* src, float* dst, int planes, int sh, int sw, int dh, int dw, cudaStream_t stream);
cudaError_t launch_resize_pyramid(PyramidParams& p, cudaStream_t stream);
cudaError_t launch_avgpool(bool backward, const float* in, float* out, int planes, int sh, int sw, int dh, int dw, cudaStream_t stream);
cudaError_t launch_mcons_fwd(const McParams& p, cudaStream_t stream);
cudaError_t launch_mcons_bwd(const McParams& p, float* g_t_ba, cudaStream_t stream);
cudaError_t launch_mreg(int which, const MregParams& p, cudaStream_t stream);
cudaError_t launch_mfield(bool backward, const MfieldParams& p, cudaStream_t stream);
cudaError_t launch_var(bool backward, const VarParams& p, cudaStream_t stream);
cudaError_t launch_silog(bool backward, const SilogParams& p, cudaStream_t stream);
cudaError_t launch_disp(bool backward, const DispParams& p, cudaStream_t stream);
cudaError_t launch_posevec(bool backward, const PoseVecParams& p, cudaStream_t stream);
constexpr int kOpBlock = 256;

static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e) {
  strncpy(g_cuda_err, cudaGetErrorString(e), sizeof(g_cuda_err) - 1);
  g_cuda_err[sizeof(g_cuda_err) - 1] = 0;
  return SDE_ERR_CUDA;
}

static size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

struct MonoLayout {
  int grid, bgrid, wgrid;
  size_t off_imgc, off_smpc, off_fin, off_partials, off_pose, off_queue, off_wflag, off_iflag, total;
};

static int mono_check(const sde_mono_desc* d) {
  if (!d) return SDE_ERR_INVALID_ARG;
  if (d->batch < 1 || d->n_scales < 1 || d->n_scales > SDE_MAX_SCALES) return SDE_ERR_INVALID_ARG;
  if (d->n_sources < 1 || d->n_sources > SDE_MAX_SOURCES) return SDE_ERR_INVALID_ARG;
  if (d->full_height < 2 || d->full_width < 2) return SDE_ERR_INVALID_ARG;
  for (int i = 0; i < d->n_scales; ++i)
    if (d->height[i] < 2 || d->width[i] < 2) return SDE_ERR_INVALID_ARG;  // reflect pad needs >= 2
  if (!(d->ssim_weight >= 0.0f) || !(d->smooth_weight >= 0.0f)) return SDE_ERR_INVALID_ARG;
  if (d->depth_mode < SDE_DEPTH_IS_DEPTH || d->depth_mode > SDE_DEPTH_IS_LOGIT) return SDE_ERR_INVALID_ARG;
  if (d->depth_mode != SDE_DEPTH_IS_DEPTH && !(d->min_depth > 0.0f && d->max_depth > d->min_depth)) return SDE_ERR_INVALID_ARG;
  if (d->norm_batch != 0 && d->norm_batch < d->batch) return SDE_ERR_INVALID_ARG;
  return SDE_OK;
}

static MonoLayout mono_layout(const sde_mono_desc* d) {
  MonoLayout L;
  L.grid = 0;
  L.bgrid = 0;
  L.wgrid = 0;
  for (int i = 0; i < d->n_scales; ++i) {
    L.wgrid += d->batch * ((d->width[i] * d->height[i] + kWarpChunk - 1) / kWarpChunk);
    L.grid += d->batch * ((d->width[i] + kTileW - 1) / kTileW) * ((d->height[i] + kTileH - 1) / kTileH);
    L.bgrid += d->batch * ((d->width[i] + kBwdW - 1) / kBwdW) * ((d->height[i] + kBwdH - 1) / kBwdH);
  }
  size_t off = 16;  // global ticket
  L.off_imgc = off;
  off = align16(off + (size_t)d->n_scales * d->batch * sizeof(unsigned));
  L.off_smpc = off;
  off = align16(off + (size_t)d->batch * sizeof(unsigned));
  L.off_fin = off;
  off = align16(off + (size_t)d->n_scales * d->batch * 2 * sizeof(double));
  L.off_partials = off;
  off = align16(off + (size_t)L.grid * 4 * sizeof(float));
  L.off_pose = off;
  off = align16(off + (size_t)L.bgrid * (kThreads / 32) * d->n_sources * 12 * sizeof(float));
  L.off_queue = off;   // persistent CTAs: (next item, CTAs gone) per kernel
  off = align16(off + 6 * sizeof(unsigned));
  L.off_wflag = off;   // flow: one flag per block of the warp kernel, one per (scale, image)
  off = align16(off + (size_t)L.wgrid * sizeof(unsigned));
  L.off_iflag = off;
  off = align16(off + (size_t)d->n_scales * d->batch * sizeof(unsigned));
  L.total = off;
  return L;
}

// Which kernels run persistent CTAs (mono_params.cuh: persist): bit 0 warp, 1 forward, 2 backward.  SDE_PERSIST=<bits>
// overrides the default 4; calls with SDE_MONO_NO_FLOW never do (two calls on two streams would take the machine in
// turns instead of sharing it).
static int g_persist = -1;
static int read_persist() {
  const char* m = getenv("SDE_PERSIST");
  return m ? (atoi(m) & 7) : 4;
}
static int persist_mask() {
  if (g_persist < 0) g_persist = read_persist();
  return g_persist;
}

static int mono_params(const sde_mono_desc* d, const sde_mono_buffers* b, bool backward, MonoParams& p) {
  int st = mono_check(d);
  if (st != SDE_OK) return st;
  if (!b || !b->intrinsics || !b->workspace || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.n_scales = d->n_scales; p.S = d->n_sources;
  p.NB = d->norm_batch > 0 ? d->norm_batch : d->batch;
  int start = 0, bstart = 0;
  for (int i = 0; i < d->n_scales; ++i) {
    if (!b->target[i] || !b->depth[i]) return SDE_ERR_INVALID_ARG;
    p.h[i] = d->height[i]; p.w[i] = d->width[i];
    p.tiles_x[i] = (p.w[i] + kTileW - 1) / kTileW;
    p.tiles_y[i] = (p.h[i] + kTileH - 1) / kTileH;
    p.tile_start[i] = start;
    start += p.B * p.tiles_x[i] * p.tiles_y[i];
    p.btiles_x[i] = (p.w[i] + kBwdW - 1) / kBwdW;
    p.btiles_y[i] = (p.h[i] + kBwdH - 1) / kBwdH;
    p.btile_start[i] = bstart;
    bstart += p.B * p.btiles_x[i] * p.btiles_y[i];
    if ((long long)p.B * p.tiles_x[i] * p.tiles_y[i] >= (1 << 22) || (long long)p.B * p.btiles_x[i] * p.btiles_y[i] >= (1 << 22))
      return SDE_ERR_INVALID_ARG;   // small_divmod (mono_params.cuh) is exact below 2^22 tiles per scale
    p.rtiles[i][0] = 1.0f / (float)p.tiles_x[i]; p.rtiles[i][1] = 1.0f / (float)p.tiles_y[i];
    p.rtiles[i][2] = 1.0f / (float)p.btiles_x[i]; p.rtiles[i][3] = 1.0f / (float)p.btiles_y[i];
    // x_scale = w_i / W as the reference forms it (MonoDepth2.py:83-85), then cast to fp32 by the multiply
    p.sx[i] = (float)((double)p.w[i] / (double)d->full_width);
    p.sy[i] = (float)((double)p.h[i] / (double)d->full_height);
    p.target[i] = b->target[i]; p.depth[i] = b->depth[i];
    p.argmin[i] = b->argmin[i];
    p.smooth_g[i] = b->smooth_g[i];
    p.grad_depth[i] = b->grad_depth[i];
    for (int j = 0; j < d->n_sources; ++j) {
      if (!b->source[i][j]) return SDE_ERR_INVALID_ARG;
      p.source[i][j] = b->source[i][j];
      p.warped[i][j] = b->warped[i][j];
    }
    const double scale_w = 1.0 / (double)(1 << (d->n_scales - i - 1));
    p.smooth_scale[i] = d->smooth_weight > 0.0f ? (float)(scale_w * (double)d->smooth_weight / d->n_scales) : 0.0f;
    {
      const bool mean = (d->flags & SDE_MONO_REDUCE_MEAN) != 0, am = (d->flags & SDE_MONO_AUTOMASK) != 0;
      const double ncand = mean ? (double)((am ? 2 : 1) * d->n_sources) : 1.0;
      p.inv_norm[i] = (float)(1.0 / ((double)d->n_scales * (double)p.NB * (double)p.h[i] * (double)p.w[i] * ncand));
    }
  }
  for (int i = d->n_scales; i <= SDE_MAX_SCALES; ++i) {
    p.tile_start[i] = start;
    p.btile_start[i] = bstart;
  }
  for (int j = 0; j < d->n_sources; ++j) {
    if (!b->pose[j]) return SDE_ERR_INVALID_ARG;
    p.pose[j] = b->pose[j];
    p.grad_pose[j] = b->grad_pose[j];
  }
  p.K = b->intrinsics;
  p.losses = b->losses; p.stats = b->saved_stats;
  p.ssim_w = d->ssim_weight;
  p.l1_w = d->ssim_weight > 0.0f ? 1.0f - d->ssim_weight : 1.0f;  // MonoDepth2.py:137-144
  p.c1 = d->c1; p.c2 = d->c2;
  p.flags = d->flags;
  p.depth_mode = d->depth_mode;
  if (d->depth_mode != SDE_DEPTH_IS_DEPTH) {
    // min_disp = 1 / max_depth, max_disp - min_disp: Python floats in the reference, fp32 on the multiply / add
    p.min_disp = (float)(1.0 / (double)d->max_depth);
    p.disp_range = (float)(1.0 / (double)d->min_depth - 1.0 / (double)d->max_depth);
  }
  const MonoLayout L = mono_layout(d);
  char* ws = static_cast<char*>(b->workspace);
  p.counter = reinterpret_cast<unsigned*>(ws);
  p.img_counter = reinterpret_cast<unsigned*>(ws + L.off_imgc);
  p.smp_counter = reinterpret_cast<unsigned*>(ws + L.off_smpc);
  p.fin = reinterpret_cast<double*>(ws + L.off_fin);
  p.partials = reinterpret_cast<float*>(ws + L.off_partials);
  p.pose_partials = reinterpret_cast<float*>(ws + L.off_pose);
  p.queue = reinterpret_cast<unsigned*>(ws + L.off_queue);
  p.persist = (d->flags & SDE_MONO_NO_FLOW) ? 0 : persist_mask();
  p.warp_flag = reinterpret_cast<unsigned*>(ws + L.off_wflag);
  p.img_flag = reinterpret_cast<unsigned*>(ws + L.off_iflag);
  p.grad_losses = b->grad_losses;
  // what the forward pass keeps for the backward pass is all-or-nothing
  {
    int have = 0, total = 0;
    for (int i = 0; i < d->n_scales; ++i) {
      ++total; have += b->smooth_g[i] != nullptr;
      for (int j = 0; j < d->n_sources; ++j) { ++total; have += b->warped[i][j] != nullptr; }
    }
    if (have != 0 && have != total) return SDE_ERR_INVALID_ARG;
  }
  if (!backward) {
    if (!b->losses) return SDE_ERR_INVALID_ARG;
  } else {
    if (!b->grad_losses) return SDE_ERR_INVALID_ARG;
    for (int i = 0; i < d->n_scales; ++i)
      if (!b->grad_depth[i] || (!(d->flags & SDE_MONO_REDUCE_MEAN) && !b->argmin[i])) return SDE_ERR_INVALID_ARG;
    for (int j = 0; j < d->n_sources; ++j)
      if (!b->grad_pose[j]) return SDE_ERR_INVALID_ARG;
  }
  return SDE_OK;
}
// ------------------------------------------------------------------------------------------------ TMA
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime once, so the library
// keeps no link-time dependency on libcuda and still loads on a machine without a driver.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// tensor map over a contiguous [planes, h, w] fp32 tensor with the tile-plane box {68, 18, 1}; zero fill outside
static bool encode_planes(CUtensorMap* m, const float* base, int planes, int h, int w, int box_w) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc || (w & 3) != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
  const cuuint64_t gstride[2] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)kHH, 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estride,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Which launches carry the programmatic-dependent-launch attribute: bit 0 warp kernel, bit 1 forward, bit 2 backward.
// Measured at cfg2 (profiles/r1_notes.md): chaining the forward kernel behind the warp kernel and the backward
// kernel behind whatever precedes it saves 2.1 + 2.4 us per step; chaining the warp kernel costs 9 us (its blocks,
// parked on the SMs until the predecessor has drained, then start in lock-step and their gathers collide in L1),
// hence the default mask 6.  SDE_DISABLE_PDL=1 / SDE_PDL_MASK=<bits> override it (read once; sde_reload_env() re-reads).
static int g_pdl_mask = -1;   // cached developer switches (sde_reload_env)
static int read_pdl_mask() {
  const char* off = getenv("SDE_DISABLE_PDL");
  if (off && off[0] == '1') return 0;
  const char* m = getenv("SDE_PDL_MASK");
  return m ? (atoi(m) & 7) : 6;
}
bool pdl_enabled(int which) {
  if (g_pdl_mask < 0) g_pdl_mask = read_pdl_mask();
  return (g_pdl_mask >> which) & 1;
}
// Tile-level dependencies between the kernels of a step (mono_params.cuh: flow): bit 0 warp -> forward (any call that
// runs the warp kernel), bit 1 forward -> backward (sde_mono_loss_step only: a stand-alone backward call must not wait
// for flags that a forward call may have set for another step).  SDE_FLOW_MASK=<bits> overrides the default 3.
static int g_flow_mask = -1;
static int read_flow_mask() {
  const char* m = getenv("SDE_FLOW_MASK");
  return m ? (atoi(m) & 3) : 3;
}
// SDE_BWD_PAIR=0: the one-source-per-pass backward kernel on every configuration (A/B timing, parity tests)
static int g_bwd_pair = -1;
static int read_bwd_pair() {
  const char* m = getenv("SDE_BWD_PAIR");
  return (m && m[0] == '0') ? 0 : 1;
}
bool bwd_pair_enabled() {
  if (g_bwd_pair < 0) g_bwd_pair = read_bwd_pair();
  return g_bwd_pair != 0;
}
static unsigned flow_mask() {
  if (g_flow_mask < 0) g_flow_mask = read_flow_mask();
  return (unsigned)g_flow_mask;
}

// Decides per scale whether the tile planes are staged by TMA (row pitch a multiple of 16 bytes, encoder
// available; the backward pass also needs the saved warps) and builds the descriptors.
static void mono_tma(const sde_mono_desc* d, const sde_mono_buffers* b, bool backward, MonoParams& p, MonoTma& t) {
  memset(&t, 0, sizeof(t));
  // SDE_MONO_NO_TMA forces the thread-staged path on every shape (the parity tests cover it that way); it travels in
  // the descriptor, so the forward and the backward call of a step take the same path
  const bool disabled = (d->flags & SDE_MONO_NO_TMA) != 0;
  for (int i = 0; i < d->n_scales; ++i) {
    p.tma[i] = 0;
    if (disabled) continue;
    if (backward && !b->warped[i][0]) continue;
    const int h = d->height[i], w = d->width[i];
    const int bw = backward ? 68 : 72;   // row pitch of the kernel's shared-memory planes (mono_bwd.cu / mono_fwd.cu)
    bool ok = encode_planes(&t.target[i], b->target[i], d->batch * 3, h, w, bw) &&
              encode_planes(&t.depth[i], b->depth[i], d->batch, h, w, bw);
    // forward with `warped` buffers: the warp kernel fills them and the loss kernel takes them through TMA
    const bool prewarp = !backward && b->warped[i][0] != nullptr;
    for (int j = 0; ok && j < d->n_sources; ++j) {
      if (backward || prewarp) ok = encode_planes(&t.warped[i][j], b->warped[i][j], d->batch * kSavedPlanes, h, w, bw);
      if (!backward && ok) ok = encode_planes(&t.source[i][j], b->source[i][j], d->batch * 3, h, w, bw);
    }
    p.tma[i] = ok ? 1 : 0;
    p.prewarp[i] = (ok && prewarp) ? 1 : 0;
  }
  int start = 0;
  for (int i = 0; i <= SDE_MAX_SCALES; ++i) {
    p.warp_start[i] = start;
    // the warp kernel fills the `warped` buffers of every scale (the loss kernels of a scale that cannot take
    // the TMA path stage them with plain loads / gather themselves)
    if (i < d->n_scales && !backward && b->warped[i][0]) start += d->batch * ((d->height[i] * d->width[i] + kWarpChunk - 1) / kWarpChunk);
  }
}

// Warp mode of the MotionLearning kernels: every direction has a `warped` buffer and the planes can be boxed.
static void motion_tma(const sde_motion_desc* d, const sde_motion_buffers* b, bool backward, MotionParams& p, MotionTma& t) {
  memset(&t, 0, sizeof(t));
  p.tma = 0;
  bool ok = (d->flags & SDE_MOTION_NO_TMA) == 0;
  const int bw = backward ? 68 : 72;
  for (int k = 0; ok && k < d->n_dirs; ++k) {
    ok = b->warped[k] != nullptr &&
         encode_planes(&t.frame_a[k], b->frame_a[k], d->batch * 3, d->height, d->width, bw) &&
         encode_planes(&t.depth_a[k], b->depth_a[k], d->batch, d->height, d->width, bw) &&
         encode_planes(&t.warped[k], b->warped[k], d->batch * kMotionSaved, d->height, d->width, bw);
  }
  p.tma = ok ? 1 : 0;
  if (!ok)
    for (int k = 0; k < SDE_MAX_DIRS; ++k) p.warped[k] = nullptr;   // the statistics pass then only gathers depth
}

// ------------------------------------------------------------------------------------------------
struct MotionLayout {
  int tiles_x, tiles_y, btiles_x, btiles_y, stat_blocks, grid, bgrid;
  size_t off_imgc, off_fin, off_stat, off_partials, off_pose, total;
};

static int motion_check(const sde_motion_desc* d) {
  if (!d) return SDE_ERR_INVALID_ARG;
  if (d->batch < 1 || d->n_dirs < 1 || d->n_dirs > SDE_MAX_DIRS) return SDE_ERR_INVALID_ARG;
  if (d->height < 2 || d->width < 2) return SDE_ERR_INVALID_ARG;
  if (!(d->ssim_weight >= 0.0f) || !(d->scale_x > 0.0f) || !(d->scale_y > 0.0f)) return SDE_ERR_INVALID_ARG;
  if (isinf(d->c1) && isinf(d->c2)) return SDE_ERR_INVALID_ARG;
  return SDE_OK;
}

static MotionLayout motion_layout(const sde_motion_desc* d) {
  MotionLayout L;
  L.tiles_x = (d->width + kTileW - 1) / kTileW;
  L.tiles_y = (d->height + kTileH - 1) / kTileH;
  L.btiles_x = (d->width + kBwdW - 1) / kBwdW;
  L.btiles_y = (d->height + kBwdH - 1) / kBwdH;
  L.stat_blocks = (d->height * d->width + kStatPix - 1) / kStatPix;
  L.grid = d->n_dirs * d->batch * L.tiles_x * L.tiles_y;
  L.bgrid = d->n_dirs * d->batch * L.btiles_x * L.btiles_y;
  size_t off = 16;  // three counters
  L.off_imgc = off;   // per-image tile tickets, forward and backward
  off = align16(off + (size_t)2 * d->n_dirs * d->batch * sizeof(unsigned));
  L.off_fin = off;
  off = align16(off + (size_t)d->n_dirs * d->batch * 4 * sizeof(double));
  L.off_stat = off;
  off = align16(off + (size_t)d->n_dirs * d->batch * L.stat_blocks * 2 * sizeof(float));
  L.off_partials = off;
  off = align16(off + (size_t)L.grid * 8 * sizeof(float));
  L.off_pose = off;
  off = align16(off + (size_t)L.bgrid * 12 * sizeof(float));
  L.total = off;
  return L;
}

static int motion_params(const sde_motion_desc* d, const sde_motion_buffers* b, bool backward, MotionParams& p) {
  int st = motion_check(d);
  if (st != SDE_OK) return st;
  if (!b || !b->intrinsics || !b->workspace || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  const MotionLayout L = motion_layout(d);
  p.B = d->batch; p.n_dirs = d->n_dirs; p.h = d->height; p.w = d->width;
  p.tiles_x = L.tiles_x; p.tiles_y = L.tiles_y; p.tiles_per_dir = d->batch * L.tiles_x * L.tiles_y;
  p.btiles_x = L.btiles_x; p.btiles_y = L.btiles_y; p.btiles_per_dir = d->batch * L.btiles_x * L.btiles_y;
  p.stat_blocks = L.stat_blocks;
  p.sx = d->scale_x; p.sy = d->scale_y;
  const bool field = (d->flags & SDE_MOTION_FIELD) != 0;
  for (int k = 0; k < d->n_dirs; ++k) {
    if (!b->frame_a[k] || !b->frame_b[k] || !b->depth_a[k] || !b->depth_b[k] || !b->pose[k]) return SDE_ERR_INVALID_ARG;
    if (field && !b->field[k]) return SDE_ERR_INVALID_ARG;
    p.frame_a[k] = b->frame_a[k]; p.frame_b[k] = b->frame_b[k];
    p.depth_a[k] = b->depth_a[k]; p.depth_b[k] = b->depth_b[k];
    p.pose[k] = b->pose[k];
    p.field[k] = field ? b->field[k] : nullptr;
    p.occ[k] = b->occlusion[k]; p.weight[k] = b->weight[k]; p.coords[k] = b->coords[k];
    p.warped[k] = b->warped[k];
    p.grad_depth[k] = b->grad_depth_a[k]; p.grad_pose[k] = b->grad_pose[k];
    p.grad_field[k] = field ? b->grad_field[k] : nullptr;
    if (backward && (!b->grad_depth_a[k] || !b->grad_pose[k] || (field && !b->grad_field[k]))) return SDE_ERR_INVALID_ARG;
  }
  p.K = b->intrinsics;
  p.ssim_w = d->ssim_weight; p.c1 = d->c1; p.c2 = d->c2;
  p.mode = isinf(d->c1) ? 1 : (isinf(d->c2) ? 2 : 0);   // ssim_loss.py:97-105
  p.losses = b->losses; p.stats = b->saved_stats;
  char* ws = static_cast<char*>(b->workspace);
  p.counters = reinterpret_cast<unsigned*>(ws);
  p.img_counter_f = reinterpret_cast<unsigned*>(ws + L.off_imgc);
  p.img_counter_b = p.img_counter_f + d->n_dirs * d->batch;
  p.fin = reinterpret_cast<double*>(ws + L.off_fin);
  p.stat_partials = reinterpret_cast<float*>(ws + L.off_stat);
  p.partials = reinterpret_cast<float*>(ws + L.off_partials);
  p.pose_partials = reinterpret_cast<float*>(ws + L.off_pose);
  p.grad_losses = b->grad_losses;
  if (!backward && !b->losses) return SDE_ERR_INVALID_ARG;
  if (backward && !b->grad_losses) return SDE_ERR_INVALID_ARG;
  return SDE_OK;
}
// ------------------------------------------------------------------------------------------------ ops
static bool vs_ok(const sde_vs_desc* d) {
  return d && d->batch >= 1 && d->channels >= 1 && d->height >= 2 && d->width >= 2;
}
struct VsLayout { int blocks; size_t off_partials, off_fix, total; };
static VsLayout vs_layout(const sde_vs_desc* d) {
  VsLayout L;
  L.blocks = (d->height * d->width + kOpBlock - 1) / kOpBlock;
  size_t off = align16((size_t)d->batch * sizeof(unsigned));
  L.off_partials = off;
  off = align16(off + (size_t)d->batch * L.blocks * 12 * sizeof(float));
  L.off_fix = off;
  off = align16(off + (size_t)d->batch * d->channels * d->height * d->width * sizeof(long long));
  L.total = off;
  return L;
}
static int vs_params(const sde_vs_desc* d, const sde_vs_buffers* b, bool backward, VsParams& p) {
  if (!vs_ok(d) || !b) return SDE_ERR_INVALID_ARG;
  if (!b->image_b || !b->depth_a || !b->intrinsics || !b->rotation || !b->translation) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.C = d->channels; p.h = d->height; p.w = d->width;
  p.t_per_pixel = (d->flags & SDE_VS_T_PER_PIXEL) ? 1 : 0;
  p.image = b->image_b; p.depth = b->depth_a; p.K = b->intrinsics; p.R = b->rotation; p.t = b->translation;
  p.sampled = b->sampled; p.depth_in_b = b->depth_in_b; p.coords = b->coords; p.valid = b->valid;
  if (!backward) return b->sampled ? SDE_OK : SDE_ERR_INVALID_ARG;
  if (!b->grad_sampled || !b->grad_depth_a || !b->grad_rotation || !b->grad_translation || !b->workspace)
    return SDE_ERR_INVALID_ARG;
  const VsLayout L = vs_layout(d);
  char* ws = static_cast<char*>(b->workspace);
  p.g_sampled = b->grad_sampled; p.g_depth_in_b = b->grad_depth_in_b; p.g_coords = b->grad_coords;
  p.g_depth = b->grad_depth_a; p.g_R = b->grad_rotation; p.g_t = b->grad_translation;
  p.g_image_fix = b->grad_image_b ? reinterpret_cast<long long*>(ws + L.off_fix) : nullptr;
  p.counters = reinterpret_cast<unsigned*>(ws);
  p.partials = reinterpret_cast<float*>(ws + L.off_partials);
  return SDE_OK;
}

static bool ssim_ok(const sde_ssim_desc* d) {
  return d && d->batch >= 1 && d->channels >= 1 && d->height >= 2 && d->width >= 2 && !(isinf(d->c1) && isinf(d->c2));
}
static int ssim_params(const sde_ssim_desc* d, const sde_ssim_buffers* b, bool backward, SsimParams& p) {
  if (!ssim_ok(d) || !b || !b->x || !b->y) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.C = d->channels; p.h = d->height; p.w = d->width;
  p.c1 = d->c1; p.c2 = d->c2;
  p.mode = isinf(d->c1) ? 1 : (isinf(d->c2) ? 2 : 0);
  p.x = b->x; p.y = b->y; p.weight = b->weight; p.out = b->out; p.avg_w = b->avg_w;
  if (!backward) return b->out ? SDE_OK : SDE_ERR_INVALID_ARG;
  if (!b->grad_out || !b->workspace || (!b->grad_x && !b->grad_y)) return SDE_ERR_INVALID_ARG;
  p.g_out = b->grad_out; p.g_x = b->grad_x; p.g_y = b->grad_y;
  p.coef = static_cast<float*>(b->workspace);
  return SDE_OK;
}

static bool smooth_ok(const sde_smooth_desc* d) {
  return d && d->batch >= 1 && d->channels >= 1 && d->height >= 2 && d->width >= 2;
}
struct SmoothLayout { int blocks; size_t off_fin, off_partials, total; };
static SmoothLayout smooth_layout(const sde_smooth_desc* d) {
  SmoothLayout L;
  L.blocks = (d->height * d->width + kOpBlock - 1) / kOpBlock;
  size_t off = align16((size_t)(1 + d->batch) * sizeof(unsigned));
  L.off_fin = off;
  off = align16(off + (size_t)d->batch * sizeof(double));
  L.off_partials = off;
  off = align16(off + (size_t)d->batch * L.blocks * 4 * sizeof(float));
  L.total = off;
  return L;
}
static int smooth_params(const sde_smooth_desc* d, const sde_smooth_buffers* b, bool backward, SmoothParams& p) {
  if (!smooth_ok(d) || !b || !b->depth || !b->image || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.C = d->channels; p.h = d->height; p.w = d->width;
  p.depth = b->depth; p.image = b->image; p.stats = b->saved_stats;
  if (!backward) {
    if (!b->loss || !b->workspace) return SDE_ERR_INVALID_ARG;
    const SmoothLayout L = smooth_layout(d);
    char* ws = static_cast<char*>(b->workspace);
    p.loss = b->loss;
    p.counters = reinterpret_cast<unsigned*>(ws);
    p.fin = reinterpret_cast<double*>(ws + L.off_fin);
    p.partials = reinterpret_cast<float*>(ws + L.off_partials);
  } else {
    if (!b->grad_loss || !b->grad_depth) return SDE_ERR_INVALID_ARG;
    p.g_loss = b->grad_loss; p.g_depth = b->grad_depth;
  }
  return SDE_OK;
}
// ------------------------------------------------------------------------------------------------ motion regularisers
static bool mcons_ok(const sde_mcons_desc* d) { return d && d->batch >= 1 && d->height >= 2 && d->width >= 2; }
struct McLayout { int blocks; size_t off_slots, off_fix, total; };
static McLayout mcons_layout(const sde_mcons_desc* d) {
  McLayout L;
  L.blocks = (d->height * d->width + kOpBlock - 1) / kOpBlock;
  size_t off = align16((size_t)(1 + d->batch) * sizeof(unsigned));
  L.off_slots = off;
  off = align16(off + (size_t)d->batch * L.blocks * 16 * sizeof(float));
  L.off_fix = off;
  off = align16(off + (size_t)d->batch * 3 * d->height * d->width * sizeof(long long));
  L.total = off;
  return L;
}
static int mcons_params(const sde_mcons_desc* d, const sde_mcons_buffers* b, bool backward, McParams& p) {
  if (!mcons_ok(d) || !b || !b->coords || !b->mask || !b->rotation || !b->workspace) return SDE_ERR_INVALID_ARG;
  // either the full fields, or pose + optional residual field per direction
  if ((!b->t_ab && !b->pose_ab) || (!b->t_ba && !b->pose_ba)) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  const McLayout L = mcons_layout(d);
  char* ws = static_cast<char*>(b->workspace);
  p.B = d->batch; p.h = d->height; p.w = d->width;
  p.coords = b->coords; p.mask = b->mask; p.R = b->rotation; p.t_ab = b->t_ab; p.t_ba = b->t_ba;
  p.pose_ab = b->pose_ab; p.pose_ba = b->pose_ba;
  p.loss = b->loss;
  p.counters = reinterpret_cast<unsigned*>(ws);
  p.slots = reinterpret_cast<float*>(ws + L.off_slots);
  p.g_t_ba_fix = reinterpret_cast<long long*>(ws + L.off_fix);
  if (!backward) return b->loss ? SDE_OK : SDE_ERR_INVALID_ARG;
  if (!b->grad_loss || !b->grad_rotation) return SDE_ERR_INVALID_ARG;
  if ((b->t_ab && !b->grad_t_ab) || (b->t_ba && !b->grad_t_ba)) return SDE_ERR_INVALID_ARG;
  if ((b->pose_ab && !b->grad_pose_t_ab) || (b->pose_ba && !b->grad_pose_t_ba)) return SDE_ERR_INVALID_ARG;
  p.g_loss = b->grad_loss; p.g_t_ab = b->t_ab ? b->grad_t_ab : nullptr; p.g_R = b->grad_rotation;
  p.g_pose_t_ab = b->pose_ab ? b->grad_pose_t_ab : nullptr;
  p.g_pose_t_ba = b->pose_ba ? b->grad_pose_t_ba : nullptr;
  p.scatter = b->t_ba ? 1 : 0;
  return SDE_OK;
}

struct MfieldLayout { int blocks; size_t off_fin, off_slots, total; };
static MfieldLayout mfield_layout(const sde_mreg_desc* d) {
  MfieldLayout L;
  L.blocks = (d->height * d->width + kOpBlock - 1) / kOpBlock;   // an upper bound of the grid (2048 pixels per block)
  size_t off = align16((size_t)(1 + d->batch) * sizeof(unsigned));
  L.off_fin = off;
  off = align16(off + (size_t)d->batch * 2 * sizeof(double));
  L.off_slots = off;
  off = align16(off + (size_t)d->batch * L.blocks * 8 * sizeof(float));
  L.total = off;
  return L;
}

static bool mreg_ok(const sde_mreg_desc* d) {
  return d && d->batch >= 1 && d->channels >= 1 && d->height >= 2 && d->width >= 2;
}
struct MregLayout { int blocks; size_t off_slots, total; };
static MregLayout mreg_layout(const sde_mreg_desc* d) {
  MregLayout L;
  L.blocks = (d->height * d->width + kOpBlock - 1) / kOpBlock;
  size_t off = align16((size_t)(1 + d->batch * d->channels) * sizeof(unsigned));
  L.off_slots = off;
  off = align16(off + (size_t)d->batch * d->channels * L.blocks * sizeof(float));
  L.total = off;
  return L;
}
static int mreg_params(const sde_mreg_desc* d, const sde_mreg_buffers* b, bool backward, bool sparsity, MregParams& p) {
  if (!mreg_ok(d) || !b || !b->field || (sparsity && !b->saved_stats)) return SDE_ERR_INVALID_ARG;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.C = d->channels; p.h = d->height; p.w = d->width;
  p.field = b->field; p.stats = b->saved_stats;
  if (!backward) {
    if (!b->loss || !b->workspace) return SDE_ERR_INVALID_ARG;
    const MregLayout L = mreg_layout(d);
    char* ws = static_cast<char*>(b->workspace);
    p.loss = b->loss;
    p.counters = reinterpret_cast<unsigned*>(ws);
    p.slots = reinterpret_cast<float*>(ws + L.off_slots);
  } else {
    if (!b->grad_loss || !b->grad_field) return SDE_ERR_INVALID_ARG;
    p.g_loss = b->grad_loss; p.g_field = b->grad_field;
  }
  return SDE_OK;
}
}  // namespace sde

using namespace sde;

extern "C" {

int sde_version(void) { return SDE_ABI_VERSION; }

const char* sde_strerror(int status) {
  switch (status) {
    case SDE_OK: return "ok";
    case SDE_ERR_INVALID_ARG: return "invalid argument";
    case SDE_ERR_UNSUPPORTED: return "option not supported by the fused path";
    case SDE_ERR_CUDA: return "CUDA runtime error (see sde_last_cuda_error)";
    case SDE_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
  }
}

const char* sde_last_cuda_error(void) { return g_cuda_err; }
#ifdef SDE_TRACE
// developer build (tools/trace_step.py): per-CTA time stamps of the last MonoDepth2 launches
static unsigned long long* g_trace_buf = nullptr;
extern "C" int sde_debug_trace(unsigned long long* dst) {
  const size_t n = (size_t)3 * kTraceBlocks * 4 * sizeof(unsigned long long);
  if (!g_trace_buf) { if (cudaMalloc(&g_trace_buf, n) != cudaSuccess) return 1; cudaMemset(g_trace_buf, 0, n); return 0; }
  return dst && cudaMemcpy(dst, g_trace_buf, n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}
#endif

void sde_reload_env(void) { g_pdl_mask = read_pdl_mask(); g_flow_mask = read_flow_mask(); g_bwd_pair = read_bwd_pair(); g_persist = read_persist(); }

size_t sde_mono_workspace_bytes(const sde_mono_desc* desc) {
  if (mono_check(desc) != SDE_OK) return 0;
  return mono_layout(desc).total;
}

// warp kernel + loss forward kernel; `step`: the backward kernel follows in the same call (sde_mono_loss_step)
static int mono_forward_launches(const sde_mono_desc* desc, const sde_mono_buffers* buf, cudaStream_t stream, bool step, unsigned* flow_out) {
  MonoParams p;
  int st = mono_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  MonoTma t;
  mono_tma(desc, buf, false, p, t);
  // flow needs the forward kernel to be scheduled behind the warp kernel programmatically (bit 1 of the launch mask),
  // and the warp kernel to run at all; the image flags also need the backward kernel's launch to be programmatic
  unsigned flow = 0;
  if (p.warp_start[p.n_scales] > 0 && pdl_enabled(1)) flow |= flow_mask() & kFlowWarp;
  if (step && pdl_enabled(2)) flow |= flow_mask() & kFlowImage;
  if (desc->flags & SDE_MONO_NO_FLOW) flow = 0;
  p.flow = flow;
#ifdef SDE_TRACE
  p.trace = g_trace_buf;
#endif
  if (flow_out) *flow_out = flow;
  cudaError_t e = launch_mono_warp(p, stream);
  if (e != cudaSuccess) return cuda_fail(e);
  e = launch_mono_fwd(p, t, stream);
  return e == cudaSuccess ? SDE_OK : cuda_fail(e);
}

int sde_mono_loss_forward(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream) {
  return mono_forward_launches(desc, buf, static_cast<cudaStream_t>(stream), false, nullptr);
}

int sde_mono_loss_step(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream) {
  // validate the backward side before anything is launched
  MonoParams pb;
  int st = mono_params(desc, buf, true, pb);
  if (st != SDE_OK) return st;
  unsigned flow = 0;
  st = mono_forward_launches(desc, buf, static_cast<cudaStream_t>(stream), true, &flow);
  if (st != SDE_OK) return st;
  MonoTma tb;
  mono_tma(desc, buf, true, pb, tb);
  pb.flow = flow & kFlowImage;
#ifdef SDE_TRACE
  pb.trace = g_trace_buf;
#endif
  cudaError_t e = launch_mono_bwd(pb, tb, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SDE_OK : cuda_fail(e);
}

int sde_mono_loss_backward(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream) {
  MonoParams p;
  int st = mono_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  MonoTma t;
  mono_tma(desc, buf, true, p, t);
  cudaError_t e = launch_mono_bwd(p, t, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SDE_OK : cuda_fail(e);
}

size_t sde_motion_workspace_bytes(const sde_motion_desc* desc) {
  if (motion_check(desc) != SDE_OK) return 0;
  return motion_layout(desc).total;
}

int sde_motion_loss_forward(const sde_motion_desc* desc, const sde_motion_buffers* buf, void* stream) {
  MotionParams p;
  int st = motion_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  MotionTma t;
  motion_tma(desc, buf, false, p, t);
  cudaError_t e = launch_motion_fwd(p, t, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SDE_OK : cuda_fail(e);
}

int sde_motion_loss_backward(const sde_motion_desc* desc, const sde_motion_buffers* buf, void* stream) {
  MotionParams p;
  int st = motion_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  MotionTma t;
  motion_tma(desc, buf, true, p, t);
  cudaError_t e = launch_motion_bwd(p, t, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SDE_OK : cuda_fail(e);
}

#define SDE_LAUNCH(expr) do { cudaError_t e_ = (expr); return e_ == cudaSuccess ? SDE_OK : cuda_fail(e_); } while (0)

size_t sde_view_synthesis_workspace_bytes(const sde_vs_desc* desc) { return vs_ok(desc) ? vs_layout(desc).total : 0; }

int sde_view_synthesis_forward(const sde_vs_desc* desc, const sde_vs_buffers* buf, void* stream) {
  VsParams p;
  int st = vs_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_vs_fwd(p, static_cast<cudaStream_t>(stream)));
}

int sde_view_synthesis_backward(const sde_vs_desc* desc, const sde_vs_buffers* buf, void* stream) {
  VsParams p;
  int st = vs_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_vs_bwd(p, buf->grad_image_b, static_cast<cudaStream_t>(stream)));
}

size_t sde_ssim_workspace_bytes(const sde_ssim_desc* desc) {
  return ssim_ok(desc) ? (size_t)6 * desc->batch * desc->channels * desc->height * desc->width * sizeof(float) : 0;
}

int sde_ssim_forward(const sde_ssim_desc* desc, const sde_ssim_buffers* buf, void* stream) {
  SsimParams p;
  int st = ssim_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_ssim_fwd(p, static_cast<cudaStream_t>(stream)));
}

int sde_ssim_backward(const sde_ssim_desc* desc, const sde_ssim_buffers* buf, void* stream) {
  SsimParams p;
  int st = ssim_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_ssim_bwd(p, static_cast<cudaStream_t>(stream)));
}

size_t sde_smoothness_workspace_bytes(const sde_smooth_desc* desc) { return smooth_ok(desc) ? smooth_layout(desc).total : 0; }

int sde_smoothness_forward(const sde_smooth_desc* desc, const sde_smooth_buffers* buf, void* stream) {
  SmoothParams p;
  int st = smooth_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_smooth_fwd(p, static_cast<cudaStream_t>(stream)));
}

int sde_smoothness_backward(const sde_smooth_desc* desc, const sde_smooth_buffers* buf, void* stream) {
  SmoothParams p;
  int st = smooth_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_smooth_bwd(p, static_cast<cudaStream_t>(stream)));
}

int sde_resize_bilinear(const float* src, float* dst, int32_t planes, int32_t src_h, int32_t src_w, int32_t dst_h,
                        int32_t dst_w, void* stream) {
  if (!src || !dst || planes < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return SDE_ERR_INVALID_ARG;
  SDE_LAUNCH(launch_resize_bilinear(src, dst, planes, src_h, src_w, dst_h, dst_w, static_cast<cudaStream_t>(stream)));
}

int sde_resize_avgpool_forward(const float* src, float* dst, int32_t planes, int32_t src_h, int32_t src_w, int32_t dst_h,
                               int32_t dst_w, void* stream) {
  if (!src || !dst || planes < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return SDE_ERR_INVALID_ARG;
  SDE_LAUNCH(launch_avgpool(false, src, dst, planes, src_h, src_w, dst_h, dst_w, static_cast<cudaStream_t>(stream)));
}

int sde_resize_avgpool_backward(const float* grad_dst, float* grad_src, int32_t planes, int32_t src_h, int32_t src_w,
                                int32_t dst_h, int32_t dst_w, void* stream) {
  if (!grad_dst || !grad_src || planes < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return SDE_ERR_INVALID_ARG;
  SDE_LAUNCH(launch_avgpool(true, grad_dst, grad_src, planes, src_h, src_w, dst_h, dst_w, static_cast<cudaStream_t>(stream)));
}

static int pyramid_call(bool u8, int32_t n_frames, int32_t planes, int32_t src_h, int32_t src_w, int32_t n_levels,
                        const int32_t* dst_h, const int32_t* dst_w, const void* const* src, float* const (*dst)[SDE_MAX_SCALES],
                        void* stream) {
  if (!src || !dst || !dst_h || !dst_w || n_frames < 1 || n_frames > kPyrFrames || n_levels < 1 || n_levels > SDE_MAX_SCALES ||
      planes < 1 || src_h < 1 || src_w < 1)
    return SDE_ERR_INVALID_ARG;
  PyramidParams p;
  memset(&p, 0, sizeof(p));
  p.n_frames = n_frames; p.n_levels = n_levels; p.planes = planes; p.sh = src_h; p.sw = src_w; p.src_u8 = u8 ? 1 : 0;
  for (int l = 0; l < n_levels; ++l) {
    if (dst_h[l] < 1 || dst_w[l] < 1) return SDE_ERR_INVALID_ARG;
    p.dh[l] = dst_h[l]; p.dw[l] = dst_w[l];
  }
  for (int f = 0; f < n_frames; ++f) {
    if (!src[f]) return SDE_ERR_INVALID_ARG;
    p.src[f] = src[f];
    for (int l = 0; l < n_levels; ++l) {
      if (!dst[f][l]) return SDE_ERR_INVALID_ARG;
      p.dst[f][l] = dst[f][l];
    }
  }
  SDE_LAUNCH(launch_resize_pyramid(p, static_cast<cudaStream_t>(stream)));
}

int sde_resize_pyramid(int32_t n_frames, int32_t planes, int32_t src_h, int32_t src_w, int32_t n_levels,
                       const int32_t* dst_h, const int32_t* dst_w, const sde_pyramid_buffers* buf, void* stream) {
  if (!buf) return SDE_ERR_INVALID_ARG;
  return pyramid_call(false, n_frames, planes, src_h, src_w, n_levels, dst_h, dst_w,
                      reinterpret_cast<const void* const*>(buf->src), buf->dst, stream);
}

int sde_resize_pyramid_u8(int32_t n_frames, int32_t planes, int32_t src_h, int32_t src_w, int32_t n_levels,
                          const int32_t* dst_h, const int32_t* dst_w, const sde_pyramid_u8_buffers* buf, void* stream) {
  if (!buf) return SDE_ERR_INVALID_ARG;
  return pyramid_call(true, n_frames, planes, src_h, src_w, n_levels, dst_h, dst_w,
                      reinterpret_cast<const void* const*>(buf->src), buf->dst, stream);
}

size_t sde_motion_consistency_workspace_bytes(const sde_mcons_desc* desc) { return mcons_ok(desc) ? mcons_layout(desc).total : 0; }

int sde_motion_consistency_forward(const sde_mcons_desc* desc, const sde_mcons_buffers* buf, void* stream) {
  McParams p;
  int st = mcons_params(desc, buf, false, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_mcons_fwd(p, static_cast<cudaStream_t>(stream)));
}

int sde_motion_consistency_backward(const sde_mcons_desc* desc, const sde_mcons_buffers* buf, void* stream) {
  McParams p;
  int st = mcons_params(desc, buf, true, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_mcons_bwd(p, buf->grad_t_ba, static_cast<cudaStream_t>(stream)));
}

size_t sde_motion_reg_workspace_bytes(const sde_mreg_desc* desc) { return mreg_ok(desc) ? mreg_layout(desc).total : 0; }

static int mreg_call(int which, bool backward, bool sparsity, const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream) {
  MregParams p;
  int st = mreg_params(desc, buf, backward, sparsity, p);
  if (st != SDE_OK) return st;
  SDE_LAUNCH(launch_mreg(which, p, static_cast<cudaStream_t>(stream)));
}

int sde_motion_smoothness_forward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream) {
  return mreg_call(0, false, false, desc, buf, stream);
}
int sde_motion_smoothness_backward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream) {
  return mreg_call(1, true, false, desc, buf, stream);
}
int sde_motion_sparsity_forward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream) {
  return mreg_call(2, false, true, desc, buf, stream);
}
int sde_motion_sparsity_backward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream) {
  return mreg_call(3, true, true, desc, buf, stream);
}

size_t sde_motion_field_reg_workspace_bytes(const sde_mreg_desc* desc) {
  return (mreg_ok(desc) && desc->channels == 3) ? mfield_layout(desc).total : 0;
}

static int mfield_call(bool backward, const sde_mreg_desc* d, const sde_mfield_buffers* b, void* stream) {
  if (!mreg_ok(d) || d->channels != 3 || !b || !b->field || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  MfieldParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->batch; p.h = d->height; p.w = d->width;
  p.pose = b->pose; p.field = b->field; p.stats = b->saved_stats;
  if (!backward) {
    if (!b->losses || !b->workspace) return SDE_ERR_INVALID_ARG;
    const MfieldLayout L = mfield_layout(d);
    char* ws = static_cast<char*>(b->workspace);
    p.losses = b->losses;
    p.counters = reinterpret_cast<unsigned*>(ws);
    p.fin = reinterpret_cast<double*>(ws + L.off_fin);
    p.slots = reinterpret_cast<float*>(ws + L.off_slots);
  } else {
    if (!b->grad_losses || !b->grad_field) return SDE_ERR_INVALID_ARG;
    p.g_losses = b->grad_losses; p.g_field = b->grad_field; p.g_pose_t = b->grad_pose_t;
  }
  SDE_LAUNCH(launch_mfield(backward, p, static_cast<cudaStream_t>(stream)));
}

int sde_motion_field_reg_forward(const sde_mreg_desc* desc, const sde_mfield_buffers* buf, void* stream) {
  return mfield_call(false, desc, buf, stream);
}
int sde_motion_field_reg_backward(const sde_mreg_desc* desc, const sde_mfield_buffers* buf, void* stream) {
  return mfield_call(true, desc, buf, stream);
}

size_t sde_variance_workspace_bytes(int64_t count) {
  if (count < 1) return 0;
  return align16(16 + (size_t)((count + kOpBlock - 1) / kOpBlock) * sizeof(float));
}

static int var_call(bool backward, int64_t count, const sde_var_buffers* b, void* stream) {
  if (count < 1 || !b || !b->depth || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  VarParams p;
  memset(&p, 0, sizeof(p));
  p.n = count; p.depth = b->depth; p.stats = b->saved_stats;
  if (!backward) {
    if (!b->loss || !b->workspace) return SDE_ERR_INVALID_ARG;
    p.loss = b->loss;
    p.counters = reinterpret_cast<unsigned*>(b->workspace);
    p.slots = reinterpret_cast<float*>(static_cast<char*>(b->workspace) + 16);
  } else {
    if (!b->grad_loss || !b->grad_depth) return SDE_ERR_INVALID_ARG;
    p.g_loss = b->grad_loss; p.g_depth = b->grad_depth;
  }
  SDE_LAUNCH(launch_var(backward, p, static_cast<cudaStream_t>(stream)));
}

int sde_variance_loss_forward(int64_t count, const sde_var_buffers* buf, void* stream) { return var_call(false, count, buf, stream); }
int sde_variance_loss_backward(int64_t count, const sde_var_buffers* buf, void* stream) { return var_call(true, count, buf, stream); }

size_t sde_silog_workspace_bytes(int64_t count) {
  if (count < 1) return 0;
  return align16(16 + (size_t)((count + kOpBlock - 1) / kOpBlock) * 3 * sizeof(float));
}

static int silog_call(bool backward, int64_t count, float vf, const sde_silog_buffers* b, void* stream) {
  if (count < 1 || !b || !b->depth_est || !b->depth_gt || !b->saved_stats) return SDE_ERR_INVALID_ARG;
  SilogParams p;
  memset(&p, 0, sizeof(p));
  p.n = count; p.est = b->depth_est; p.gt = b->depth_gt; p.vf = vf; p.stats = b->saved_stats;
  if (!backward) {
    if (!b->loss || !b->workspace) return SDE_ERR_INVALID_ARG;
    p.loss = b->loss;
    p.counter = reinterpret_cast<unsigned*>(b->workspace);
    p.slots = reinterpret_cast<float*>(static_cast<char*>(b->workspace) + 16);
  } else {
    if (!b->grad_loss || !b->grad_depth_est) return SDE_ERR_INVALID_ARG;
    p.g_loss = b->grad_loss; p.g_est = b->grad_depth_est;
  }
  SDE_LAUNCH(launch_silog(backward, p, static_cast<cudaStream_t>(stream)));
}

int sde_silog_loss_forward(int64_t count, float variance_focus, const sde_silog_buffers* buf, void* stream) {
  return silog_call(false, count, variance_focus, buf, stream);
}
int sde_silog_loss_backward(int64_t count, float variance_focus, const sde_silog_buffers* buf, void* stream) {
  return silog_call(true, count, variance_focus, buf, stream);
}

static int disp_call(bool backward, int64_t count, float min_depth, float max_depth, const sde_disp_buffers* b, void* stream) {
  if (count < 1 || !b || !b->disp || !(min_depth > 0.0f) || !(max_depth > min_depth)) return SDE_ERR_INVALID_ARG;
  DispParams p;
  memset(&p, 0, sizeof(p));
  p.n = count; p.disp = b->disp;
  // min_disp = 1 / max_depth, max_disp = 1 / min_depth, in fp32 as the reference's Python floats become on the multiply
  p.min_disp = (float)(1.0 / (double)max_depth);
  p.range = (float)(1.0 / (double)min_depth - 1.0 / (double)max_depth);
  if (!backward) {
    if (!b->depth) return SDE_ERR_INVALID_ARG;
    p.scaled = b->scaled_disp; p.depth = b->depth;
  } else {
    if (!b->grad_disp || (!b->grad_depth && !b->grad_scaled_disp)) return SDE_ERR_INVALID_ARG;
    p.g_scaled = b->grad_scaled_disp; p.g_depth = b->grad_depth; p.g_disp = b->grad_disp;
  }
  SDE_LAUNCH(launch_disp(backward, p, static_cast<cudaStream_t>(stream)));
}

int sde_disp_to_depth_forward(int64_t count, float min_depth, float max_depth, const sde_disp_buffers* buf, void* stream) {
  return disp_call(false, count, min_depth, max_depth, buf, stream);
}
int sde_disp_to_depth_backward(int64_t count, float min_depth, float max_depth, const sde_disp_buffers* buf, void* stream) {
  return disp_call(true, count, min_depth, max_depth, buf, stream);
}

static int posevec_call(bool backward, int32_t batch, const sde_posevec_buffers* b, void* stream) {
  if (batch < 1 || !b || !b->vec) return SDE_ERR_INVALID_ARG;
  PoseVecParams p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.vec = b->vec;
  if (!backward) {
    if (!b->pose) return SDE_ERR_INVALID_ARG;
    p.mat = b->pose;
  } else {
    if (!b->grad_pose || !b->grad_vec) return SDE_ERR_INVALID_ARG;
    p.g_mat = b->grad_pose; p.g_vec = b->grad_vec;
  }
  SDE_LAUNCH(launch_posevec(backward, p, static_cast<cudaStream_t>(stream)));
}

int sde_pose_vec2mat_forward(int32_t batch, const sde_posevec_buffers* buf, void* stream) { return posevec_call(false, batch, buf, stream); }
int sde_pose_vec2mat_backward(int32_t batch, const sde_posevec_buffers* buf, void* stream) { return posevec_call(true, batch, buf, stream); }

}  // extern "C"
