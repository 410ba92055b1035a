/*
 * sde_loss.h -- C ABI of libsde_loss.so: the B200 (sm_100a) view-synthesis loss path
 * of SimpleDepthEstimation (MonoDepth2 / PackNet / MotionLearning self-supervision).
 *
 * The reference has no FFI: its operator API for this path is Python.  Each entry
 * point below states the reference symbol (file:line in zzzxxxttt/SimpleDepthEstimation)
 * whose arithmetic it replaces; INTEGRATION.md shows the ctypes binding a maintainer
 * adds on the reference side.
 *
 * Conventions (all entry points):
 *   - return 0 (SDE_OK) or a negative sde_status; never throws, never exits;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never
 *     allocates, frees or synchronises; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*, NULL = legacy default stream);
 *   - device pointers, fp32, contiguous NCHW, 4-byte aligned; intrinsics [B,3,3]
 *     row-major; poses [B,4,4] row-major (R = [:3,:3], t = [:3,3]);
 *   - workspaces must be zero-filled ONCE after allocation; every call leaves them
 *     zeroed again, so they can be reused call after call on the same stream.  A workspace
 *     belongs to ONE stream at a time (its ticket counters are shared by the CTAs of a call);
 *     after a call that returned an error, or a launch that faulted, zero-fill it again
 *     before reuse (the tickets may be left non-zero);
 *   - re-entrant for distinct (stream, workspace) pairs; no mutable globals;
 *   - results are deterministic: fixed-order reductions, no floating-point atomics.
 */
#ifndef SDE_LOSS_H_
#define SDE_LOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDE_ABI_VERSION 5
#define SDE_MAX_SCALES 6
#define SDE_MAX_SOURCES 4
#define SDE_MONO_SAVED_PLANES 9    /* planes per sample of sde_mono_buffers.warped[i][j] */
#define SDE_MOTION_SAVED_PLANES 16 /* planes per sample of sde_motion_buffers.warped[d] */

typedef enum sde_status {
  SDE_OK = 0,
  SDE_ERR_INVALID_ARG = -1,   /* NULL pointer, non-positive size, h or w < 2, too many scales/sources */
  SDE_ERR_UNSUPPORTED = -2,   /* option not implemented by the fused path (e.g. CLIP > 0) */
  SDE_ERR_CUDA = -3,          /* a CUDA runtime call failed; see sde_last_cuda_error() */
  SDE_ERR_NO_DEVICE = -4      /* no sm_100 device / driver */
} sde_status;

/* flags of sde_mono_desc.flags */
#define SDE_MONO_AUTOMASK 1u       /* LOSS.AUTOMASK (MonoDepth2.py:96-101) */
#define SDE_MONO_REDUCE_MEAN 2u    /* LOSS.PHOTOMETRIC_REDUCE == 'mean' (MonoDepth2.py:116-117); default 'min' */
#define SDE_MONO_NO_TMA 4u         /* stage the tile planes with plain loads even where TMA boxes are possible (testing);
                                      part of the descriptor so that the forward and the backward call of a step agree */
#define SDE_MONO_NO_FLOW 8u        /* whole-grid dependencies between the kernels of a call instead of tile-level ones
                                      (see sde_mono_loss_step); set by callers that overlap several calls on several
                                      streams themselves (MonoLossPlan(streams > 1)): measured slower when combined */
/* values of sde_mono_desc.depth_mode */
#define SDE_DEPTH_IS_DEPTH 0
#define SDE_DEPTH_IS_DISP 1
#define SDE_DEPTH_IS_LOGIT 2

int sde_version(void);
const char* sde_strerror(int status);
/* text of the last CUDA error seen by this thread ("" if none) */
const char* sde_last_cuda_error(void);
/* Developer switches (SDE_DISABLE_PDL, SDE_PDL_MASK) are read from the environment once, at the first call; this
 * re-reads them (tests that flip a switch between two plans).  Not on the hot path. */
void sde_reload_env(void);

/* ------------------------------------------------------------------------------------------
 * MonoDepth2 multi-scale loss (fused).  Replaces the loss loop of
 * MonoDepth2Model.forward, detectron2/modeling/meta_arch/MonoDepth2.py:78-124, i.e. per
 * scale and source: view_synthesis (geometry/camera.py:166-202) -> SSIM
 * (modeling/losses/ssim_loss.py:34-53) + L1 (MonoDepth2.py:130-151) -> identity automask
 * candidates -> per-pixel min (MonoDepth2.py:116-119) -> mean; plus smoothness_loss
 * (modeling/losses/smoothness_loss.py:42-80) with the scale weights of MonoDepth2.py:80,103-105.
 * The image pyramid (resize_img, camera.py:40-46) is an input: target[i], source[i][j].
 * ------------------------------------------------------------------------------------------ */
typedef struct sde_mono_desc {
  int32_t batch;                    /* B */
  int32_t n_scales;                 /* len(depth_pred), 1..SDE_MAX_SCALES, finest first */
  int32_t n_sources;                /* S = len(ctx_img), 1..SDE_MAX_SOURCES */
  int32_t height[SDE_MAX_SCALES];   /* depth_pred[i].shape[-2] */
  int32_t width[SDE_MAX_SCALES];    /* depth_pred[i].shape[-1] */
  int32_t full_height, full_width;  /* image.shape[-2:]; intrinsics are given at this size */
  float ssim_weight;                /* LOSS.SSIM_WEIGHT (0.85); 0 -> pure L1 */
  float c1, c2;                     /* LOSS.C1, LOSS.C2 */
  float smooth_weight;              /* LOSS.SMOOTHNESS_WEIGHT (1e-3); 0 -> smooth_loss = 0 */
  uint32_t flags;                   /* SDE_MONO_* */
  /* What depth[i] holds (the decoder's tail folded into the loss kernels, so that no depth map is
   * written and re-read between the network and the loss, and the backward pass emits the gradient w.r.t. what the
   * network produced):
   *   SDE_DEPTH_IS_DEPTH  depth (default);
   *   SDE_DEPTH_IS_DISP   the decoder's disparity output `disp`: depth = disp_to_depth(disp, min_depth, max_depth)[1] =
   *                       1 / (1/max_depth + (1/min_depth - 1/max_depth) disp), detectron2/layers/depth_decoder.py:9-18
   *                       (DepthResNet.py:41,57);
   *   SDE_DEPTH_IS_LOGIT  the pre-activation output of the last convolution: disp = softplus(x) (nn.Softplus(),
   *                       depth_decoder.py:93,108), then as SDE_DEPTH_IS_DISP.
   * grad_depth[i] receives d loss / d depth[i] in the same representation. */
  int32_t depth_mode;
  float min_depth, max_depth;       /* used unless depth_mode == SDE_DEPTH_IS_DEPTH; 0 < min_depth < max_depth */
  /* The batch size the means of the losses are taken over; 0 = `batch`.  A caller that splits a batch of N samples into
   * sub-batches -- one call per sub-batch, e.g. on two streams so that the tail of one call's kernels overlaps the head
   * of the other's (MonoLossPlan(streams=2)) -- passes N in every call: the losses and the gradients of the
   * sub-batches then ADD UP to those of the whole batch (every term of the reference's losses is a mean of per-sample
   * terms: MonoDepth2.py:119, smoothness_loss.py:62-80). */
  int32_t norm_batch;
} sde_mono_desc;

typedef struct sde_mono_buffers {
  /* inputs */
  const float* target[SDE_MAX_SCALES];                   /* [B,3,h_i,w_i] */
  const float* source[SDE_MAX_SCALES][SDE_MAX_SOURCES];  /* [B,3,h_i,w_i] */
  const float* depth[SDE_MAX_SCALES];                    /* [B,1,h_i,w_i] */
  const float* intrinsics;                               /* [B,3,3] at full_height x full_width */
  const float* pose[SDE_MAX_SOURCES];                    /* [B,4,4] target->source_j */
  /* forward outputs */
  float* losses;                      /* [2]: rec_loss, smooth_loss */
  uint8_t* argmin[SDE_MAX_SCALES];    /* [B,h_i,w_i] winning candidate index (warp0,ident0,warp1,ident1..) */
  float* saved_stats;                 /* [n_scales*B*2] per-image (mean inverse depth, smoothness) for backward */
  /* optional, forward output / backward input, all-or-nothing (every warped[i][j] and every smooth_g[i], or none):
   * what the forward pass keeps for the backward pass.
   *   warped[i][j]  [B,SDE_MONO_SAVED_PLANES,h_i,w_i]: planes 0..2 the warped source; planes 3..5 / 6..8 its
   *                 derivatives d warped_c / dX and d warped_c / dY w.r.t. the sample coordinate, already divided by
   *                 the projective denominator p2 + 1e-6 (camera.py:150-151) and zero where nan_to_num / clamp gate
   *                 the gradient (camera.py:184-188) -- with them the backward kernel turns d loss / d warped into
   *                 d loss / d (depth, R, t) by multiply-adds, without gathering or dividing; written by the warp kernel;
   *   smooth_g[i]   [B,1,h_i,w_i]: d smoothness / d (1/depth) before the division by the per-image mean.
   * With them the loss kernels take the warped planes through TMA and the backward kernel neither re-projects
   * the halo nor re-gathers (trades 72 B/pixel/source + 4 B/pixel of HBM traffic, which this issue-bound path
   * has to spare, for the dominant gather cost).  NULL = the backward kernel recomputes everything. */
  float* warped[SDE_MAX_SCALES][SDE_MAX_SOURCES];
  float* smooth_g[SDE_MAX_SCALES];
  /* backward inputs / outputs */
  const float* grad_losses;           /* [2] upstream d/d rec_loss, d/d smooth_loss (device) */
  float* grad_depth[SDE_MAX_SCALES];  /* [B,1,h_i,w_i] */
  float* grad_pose[SDE_MAX_SOURCES];  /* [B,4,4]; last row written as zeros */
  /* scratch */
  void* workspace;                    /* sde_mono_workspace_bytes(), zero-filled once */
} sde_mono_buffers;

size_t sde_mono_workspace_bytes(const sde_mono_desc* desc);
/* writes losses, argmin, saved_stats */
int sde_mono_loss_forward(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream);
/* reads argmin, saved_stats, grad_losses; recomputes the warp; writes grad_depth, grad_pose */
int sde_mono_loss_backward(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream);
/* forward + backward of one step in one call (value-and-grad: grad_losses is known when the losses are asked for, as for
 * a trainer that sums the loss keys, projects/MonoDepth2/train.py:91-101).  Same kernels, same results as the two calls
 * above; the three launches are chained with TILE-level dependencies instead of grid-level ones: a forward tile waits
 * for the chunks of the warp kernel that hold its rows, a backward tile for the forward tiles of its image, so each
 * kernel fills the SMs its predecessor leaves idle while it drains.  Needs every buffer of both calls. */
int sde_mono_loss_step(const sde_mono_desc* desc, const sde_mono_buffers* buf, void* stream);

/* ------------------------------------------------------------------------------------------
 * MotionLearning two-frame loss (fused).  One call covers n_dirs directions (1->2 and 2->1) of
 * one scale.  Replaces, per direction, MotionLearningModel.rgbd_consistency_loss
 * (detectron2/modeling/meta_arch/MotionLearning.py:248-291): view_synthesis of cat[frame_B,
 * depth_B] (geometry/camera.py:166-202) with the per-pixel translation t = pose[:3,3] + field,
 * occlusion mask, rgb L1, depth-proximity weight, WeightedSSIM
 * (modeling/losses/ssim_loss.py:84-111); plus smoothness_loss(depth_A, frame_A)
 * (modeling/losses/smoothness_loss.py:42-80; MotionLearning.py:231-235).
 * Direction d uses frame_a[d] as the target (A) and frame_b[d] / depth_b[d] as the source (B).
 * ------------------------------------------------------------------------------------------ */
#define SDE_MAX_DIRS 2
#define SDE_MOTION_FIELD 1u        /* field[d] given: residual translation [B,3,h,w] (motion_pred) */
#define SDE_MOTION_NO_TMA 2u       /* recompute mode even where the kept planes could be boxed (testing); see SDE_MONO_NO_TMA */
#define SDE_MOTION_N_LOSSES 4      /* per direction: rgb_l1_loss, ssim_loss, smooth_loss, reserved */

typedef struct sde_motion_desc {
  int32_t batch;                 /* B */
  int32_t n_dirs;                /* 1 or 2 */
  int32_t height, width;         /* size of this scale */
  float scale_x, scale_y;        /* scale_intrinsics factors (MotionLearning.py:131-132); 1 at full size */
  float ssim_weight;             /* LOSS.SSIM_WEIGHT (3.0): ssim_loss = mean(ssim * avg_w) * ssim_weight * 0.5 */
  float c1, c2;                  /* LOSS.C1 / LOSS.C2; INFINITY selects the reference's one-factor forms */
  uint32_t flags;                /* SDE_MOTION_* */
} sde_motion_desc;

typedef struct sde_motion_buffers {
  /* inputs, per direction */
  const float* frame_a[SDE_MAX_DIRS];   /* [B,3,h,w] target frame A */
  const float* frame_b[SDE_MAX_DIRS];   /* [B,3,h,w] source frame B */
  const float* depth_a[SDE_MAX_DIRS];   /* [B,1,h,w] */
  const float* depth_b[SDE_MAX_DIRS];   /* [B,1,h,w] */
  const float* pose[SDE_MAX_DIRS];      /* [B,4,4] A->B */
  const float* field[SDE_MAX_DIRS];     /* [B,3,h,w] residual translation or NULL */
  const float* intrinsics;              /* [B,3,3] at full size; scaled by (scale_x, scale_y) in-kernel */
  /* forward outputs */
  float* losses;                        /* [n_dirs][SDE_MOTION_N_LOSSES] */
  float* saved_stats;                   /* [n_dirs][B][4]: depth_err_2nd_mom, sum(occ), mean 1/depth, smoothness */
  float* occlusion[SDE_MAX_DIRS];       /* optional [B,1,h,w] occlusion_mask (MotionLearning.py:257-259) */
  float* weight[SDE_MAX_DIRS];          /* optional [B,1,h,w] depth_proximity_weight (MotionLearning.py:279-282) */
  float* coords[SDE_MAX_DIRS];          /* optional [B,h,w,2] coords_A_in_B, normalised (camera.py:190-193) */
  /* optional, forward output / backward input: [B,SDE_MOTION_SAVED_PLANES,h,w] per direction = warped rgb (3),
   * depth_error (1), valid + 2 * occlusion (1), d warped_c / dX (3) and d warped_c / dY (3) w.r.t. the sample
   * coordinate (zero where nan_to_num / clamp gate the gradient), the local smoothness gradient
   * d smoothness / d (1/depth_A) (1), and four planes of scratch (frame B and depth B interleaved per pixel for the
   * gather of the forward pass).  When given (and the row pitch is a multiple of 16 bytes) the statistics pre-pass
   * doubles as the warp kernel, the loss kernels take these planes through TMA instead of re-projecting and
   * re-gathering, and the backward kernel never touches frame B again.  NULL = recompute. */
  float* warped[SDE_MAX_DIRS];
  /* backward */
  const float* grad_losses;             /* [n_dirs][SDE_MOTION_N_LOSSES] upstream gradients (device) */
  float* grad_depth_a[SDE_MAX_DIRS];    /* [B,1,h,w] */
  float* grad_pose[SDE_MAX_DIRS];       /* [B,4,4] (rows 0..2 = [dR | dt]; dt = sum over pixels) */
  float* grad_field[SDE_MAX_DIRS];      /* [B,3,h,w], required iff SDE_MOTION_FIELD */
  void* workspace;                      /* sde_motion_workspace_bytes(), zero-filled once */
} sde_motion_buffers;

size_t sde_motion_workspace_bytes(const sde_motion_desc* desc);
/* writes losses, saved_stats and the optional maps (launches: [interleave frame B + depth B, warp mode only,]
 * statistics / warp pre-pass, fused loss) */
int sde_motion_loss_forward(const sde_motion_desc* desc, const sde_motion_buffers* buf, void* stream);
/* reads saved_stats, grad_losses; recomputes the warp; writes grad_depth_a, grad_pose, grad_field */
int sde_motion_loss_backward(const sde_motion_desc* desc, const sde_motion_buffers* buf, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone operators: one entry point per reference function, for callers that use the pieces
 * outside the fused losses.  Same conventions as above.
 * ------------------------------------------------------------------------------------------ */

/* view_synthesis(image_B, depth_A, intrinsics, R_A_to_B, t_A_to_B), detectron2/geometry/camera.py:166-202
 * (inv_intrinsics :25-37, img_to_points :125-138, points_to_img :141-163, nan_to_num + clamp + normalise
 * :184-193, F.grid_sample bilinear / align_corners=True :196).  `intrinsics` is used as given (already
 * scaled to this size). */
#define SDE_VS_T_PER_PIXEL 1u      /* translation is [B,3,h,w]; otherwise [B,3] */
typedef struct sde_vs_desc {
  int32_t batch, channels, height, width;
  uint32_t flags;
} sde_vs_desc;

typedef struct sde_vs_buffers {
  const float* image_b;        /* [B,C,h,w] */
  const float* depth_a;        /* [B,1,h,w] */
  const float* intrinsics;     /* [B,3,3] */
  const float* rotation;       /* [B,3,3] */
  const float* translation;    /* [B,3,h,w] or [B,3] */
  /* forward outputs */
  float* sampled;              /* [B,C,h,w] */
  float* depth_in_b;           /* [B,1,h,w], optional */
  float* coords;               /* [B,h,w,2] normalised (x,y), optional */
  uint8_t* valid;              /* [B,1,h,w] 0/1, optional */
  /* backward inputs (upstream gradients; the last two optional) and outputs */
  const float* grad_sampled;
  const float* grad_depth_in_b;
  const float* grad_coords;
  float* grad_depth_a;         /* [B,1,h,w] */
  float* grad_rotation;        /* [B,3,3] */
  float* grad_translation;     /* same shape as translation */
  float* grad_image_b;         /* [B,C,h,w], optional: bilinear scatter, accumulated in 64-bit fixed point
                                  (integer atomics: order-independent, hence deterministic).  A contribution of 2^15
                                  or more in magnitude, a sum beyond 2^16, or a NaN / inf contribution gives NaN in
                                  that element (ATen's float scatter would overflow / propagate likewise). */
  void* workspace;             /* sde_view_synthesis_workspace_bytes(), zero-filled once */
} sde_vs_buffers;

size_t sde_view_synthesis_workspace_bytes(const sde_vs_desc* desc);
int sde_view_synthesis_forward(const sde_vs_desc* desc, const sde_vs_buffers* buf, void* stream);
int sde_view_synthesis_backward(const sde_vs_desc* desc, const sde_vs_buffers* buf, void* stream);

/* SSIM(C1,C2)(x,y) -> clamp((1-ssim)/2,0,1), detectron2/modeling/losses/ssim_loss.py:34-53, and
 * WeightedSSIM(C1,C2)(x,y,w) -> (map, avg_w), ssim_loss.py:84-111 (weight != NULL).  Backward w.r.t. x
 * and y (the weight is detached at the reference's only call site, MotionLearning.py:279-285). */
typedef struct sde_ssim_desc {
  int32_t batch, channels, height, width;
  float c1, c2;                /* INFINITY selects the one-factor forms of WeightedSSIM */
} sde_ssim_desc;

typedef struct sde_ssim_buffers {
  const float* x;              /* [B,C,h,w] */
  const float* y;              /* [B,C,h,w] */
  const float* weight;         /* [B,1,h,w] or NULL (plain SSIM) */
  float* out;                  /* [B,C,h,w] */
  float* avg_w;                /* [B,1,h,w], optional (WeightedSSIM) */
  const float* grad_out;       /* [B,C,h,w] */
  float* grad_x;               /* optional */
  float* grad_y;               /* optional */
  void* workspace;             /* sde_ssim_workspace_bytes(): six coefficient planes; no initialisation needed */
} sde_ssim_buffers;

size_t sde_ssim_workspace_bytes(const sde_ssim_desc* desc);
int sde_ssim_forward(const sde_ssim_desc* desc, const sde_ssim_buffers* buf, void* stream);
int sde_ssim_backward(const sde_ssim_desc* desc, const sde_ssim_buffers* buf, void* stream);

/* smoothness_loss(depth, image), detectron2/modeling/losses/smoothness_loss.py:42-80. */
typedef struct sde_smooth_desc {
  int32_t batch, channels, height, width;   /* channels of `image` */
} sde_smooth_desc;

typedef struct sde_smooth_buffers {
  const float* depth;          /* [B,1,h,w] */
  const float* image;          /* [B,C,h,w] */
  float* loss;                 /* [1] */
  float* saved_stats;          /* [B*2] per-image (mean inverse depth, loss) for backward */
  const float* grad_loss;      /* [1] (device) */
  float* grad_depth;           /* [B,1,h,w] */
  void* workspace;             /* sde_smoothness_workspace_bytes(), zero-filled once */
} sde_smooth_buffers;

size_t sde_smoothness_workspace_bytes(const sde_smooth_desc* desc);
int sde_smoothness_forward(const sde_smooth_desc* desc, const sde_smooth_buffers* buf, void* stream);
int sde_smoothness_backward(const sde_smooth_desc* desc, const sde_smooth_buffers* buf, void* stream);

/* resize_img(image, dst_size) = F.interpolate(mode='bilinear', align_corners=True),
 * detectron2/geometry/camera.py:40-46: src [planes,sh,sw] -> dst [planes,dh,dw]. */
int sde_resize_bilinear(const float* src, float* dst, int32_t planes, int32_t src_h, int32_t src_w, int32_t dst_h,
                        int32_t dst_w, void* stream);

/* resize_img_avgpool(image, dst_size) = F.adaptive_avg_pool2d, detectron2/geometry/camera.py:49-54 (MotionLearning.py:
 * 126-144 resizes frames, depths and motion fields with it when NUM_SCALES > 1): src [planes,sh,sw] -> dst
 * [planes,dh,dw]; the backward entry is its adjoint (grad_dst [planes,dh,dw] -> grad_src [planes,sh,sw]). */
int sde_resize_avgpool_forward(const float* src, float* dst, int32_t planes, int32_t src_h, int32_t src_w, int32_t dst_h,
                               int32_t dst_w, void* stream);
int sde_resize_avgpool_backward(const float* grad_dst, float* grad_src, int32_t planes, int32_t src_h, int32_t src_w,
                                int32_t dst_h, int32_t dst_w, void* stream);

/* The image pyramid of a training step in one launch: resize_img (camera.py:40-46) of `n_frames` frames (the target
 * and every source, MonoDepth2.py:82,88) [planes,src_h,src_w] to `n_levels` sizes each; dst[f][l] = [planes,dst_h[l],dst_w[l]].
 * n_frames <= SDE_MAX_SOURCES + 1, n_levels <= SDE_MAX_SCALES.  Same arithmetic as sde_resize_bilinear. */
typedef struct sde_pyramid_buffers {
  const float* src[SDE_MAX_SOURCES + 1];
  float* dst[SDE_MAX_SOURCES + 1][SDE_MAX_SCALES];
} sde_pyramid_buffers;

int sde_resize_pyramid(int32_t n_frames, int32_t planes, int32_t src_h, int32_t src_w, int32_t n_levels,
                       const int32_t* dst_h, const int32_t* dst_w, const sde_pyramid_buffers* buf, void* stream);

/* The same from the DECODED frames: src[f] is uint8 [planes,src_h,src_w] (what the data loader reads before
 * torchvision's ToTensor, detectron2/data/datasets/kitti_v2.py:207-208); every tap is byte / 255 in fp32, so dst[f][l]
 * holds the bits resize_img gives on the converted frame, and a level of the source size is the conversion itself.
 * A host-side integration then copies a quarter of the bytes per frame. */
typedef struct sde_pyramid_u8_buffers {
  const uint8_t* src[SDE_MAX_SOURCES + 1];
  float* dst[SDE_MAX_SOURCES + 1][SDE_MAX_SCALES];
} sde_pyramid_u8_buffers;

int sde_resize_pyramid_u8(int32_t n_frames, int32_t planes, int32_t src_h, int32_t src_w, int32_t n_levels,
                          const int32_t* dst_h, const int32_t* dst_w, const sde_pyramid_u8_buffers* buf, void* stream);

/* ------------------------------------------------------------------------------------------
 * MotionLearning regularisers, detectron2/modeling/losses/motion_loss.py (callers
 * MotionLearning.py:188-220).
 * ------------------------------------------------------------------------------------------ */

/* motion_consistency_loss(coords_A_in_B, mask, R_A2B, R_B2A, t_A2B, t_B2A), motion_loss.py:7-48: the
 * translation term  mean(mask * |R_A2B t_hat + t_A2B|^2 / (|t_A2B|^2 + |t_hat|^2 + 1e-24))  with
 * t_hat = grid_sample(t_B2A, coords) (bilinear, zeros, align_corners=True; coords carry no gradient, :11).
 * The rotation term (:24,35-38) is [B,3,3] arithmetic and stays on the host side. */
typedef struct sde_mcons_desc {
  int32_t batch, height, width;
} sde_mcons_desc;

typedef struct sde_mcons_buffers {
  const float* coords;        /* [B,h,w,2] normalised (x,y) */
  const float* mask;          /* [B,1,h,w] */
  const float* rotation;      /* [B,3,3] R_A2B */
  const float* t_ab;          /* [B,3,h,w]: the translation field t_A2B -- or, with pose_ab, only its residual part */
  const float* t_ba;          /* [B,3,h,w] likewise for B->A */
  float* loss;                /* [1] trans_error */
  const float* grad_loss;     /* [1] (device) */
  float* grad_t_ab;           /* [B,3,h,w] */
  float* grad_t_ba;           /* [B,3,h,w]: bilinear scatter, 64-bit fixed point + integer atomics (deterministic) */
  float* grad_rotation;       /* [B,3,3] */
  void* workspace;            /* sde_motion_consistency_workspace_bytes(), zero-filled once */
  /* Optional split form (MotionLearning.py:143-147 builds t = pose[:, :3, [3], None] + motion field; here the sum is
   * formed per pixel, the [B,3,h,w] overall field is never materialised): pose_ab / pose_ba [B,4,4] contribute their
   * translation column; t_ab / t_ba are then the residual fields and may be NULL (rigid motion).  The backward pass
   * writes d / d pose_*[:, :3, 3] to grad_pose_t_* [B,3] and the field gradients to grad_t_* (required iff the
   * field is given). */
  const float* pose_ab;
  const float* pose_ba;
  float* grad_pose_t_ab;
  float* grad_pose_t_ba;
} sde_mcons_buffers;

size_t sde_motion_consistency_workspace_bytes(const sde_mcons_desc* desc);
int sde_motion_consistency_forward(const sde_mcons_desc* desc, const sde_mcons_buffers* buf, void* stream);
int sde_motion_consistency_backward(const sde_mcons_desc* desc, const sde_mcons_buffers* buf, void* stream);

/* motion_smoothness_loss_fn(m), motion_loss.py:51-55, and motion_sparsity_loss_fn(m), motion_loss.py:58-64. */
typedef struct sde_mreg_desc {
  int32_t batch, channels, height, width;
} sde_mreg_desc;

typedef struct sde_mreg_buffers {
  const float* field;         /* [B,C,h,w] */
  float* loss;                /* [1] */
  float* saved_stats;         /* sparsity: [B*C] mean |m| per plane (detached in the reference, :60) */
  const float* grad_loss;     /* [1] (device) */
  float* grad_field;          /* [B,C,h,w] */
  void* workspace;            /* sde_motion_reg_workspace_bytes(), zero-filled once */
} sde_mreg_buffers;

size_t sde_motion_reg_workspace_bytes(const sde_mreg_desc* desc);
int sde_motion_smoothness_forward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream);
int sde_motion_smoothness_backward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream);
int sde_motion_sparsity_forward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream);
int sde_motion_sparsity_backward(const sde_mreg_desc* desc, const sde_mreg_buffers* buf, void* stream);

/* The field regularisers as the model applies them (detectron2/modeling/meta_arch/MotionLearning.py:203-220), fused:
 * t = pose[:, :3, 3] + field; s = 1 / sqrt(3 mean_{c,h,w}(t^2) + 1e-12) per sample (not detached); mn = field * s;
 * losses[0] = motion_smoothness_loss_fn(mn), losses[1] = motion_sparsity_loss_fn(mn) (motion_loss.py:51-64).
 * Neither t nor mn is materialised; the gradient reaches the field and the pose translation (through s). */
typedef struct sde_mfield_buffers {
  const float* pose;          /* [B,4,4] A->B, or NULL (zero translation) */
  const float* field;         /* [B,3,h,w] residual translation (motion_pred, after masking / resizing) */
  float* losses;              /* [2] */
  float* saved_stats;         /* [B*12] for backward */
  const float* grad_losses;   /* [2] upstream gradients (device) */
  float* grad_field;          /* [B,3,h,w] */
  float* grad_pose_t;         /* [B,3] d / d pose[:, :3, 3], or NULL */
  void* workspace;            /* sde_motion_field_reg_workspace_bytes(), zero-filled once */
} sde_mfield_buffers;

size_t sde_motion_field_reg_workspace_bytes(const sde_mreg_desc* desc);   /* desc->channels must be 3 */
int sde_motion_field_reg_forward(const sde_mreg_desc* desc, const sde_mfield_buffers* buf, void* stream);
int sde_motion_field_reg_backward(const sde_mreg_desc* desc, const sde_mfield_buffers* buf, void* stream);

/* variance_loss(depth) = 1 / mean((depth / mean(depth) - 1)^2), detectron2/modeling/losses/losses.py:16-18
 * (PackNet config, projects/MonoDepth2/configs/packnet_1a.yaml:12; callers MonoDepth2.py:112-113,
 * MotionLearning.py:237-239).  `count` = number of elements of the depth tensor. */
typedef struct sde_var_buffers {
  const float* depth;         /* [count] */
  float* loss;                /* [1] */
  float* saved_stats;         /* [2] mean and centred second moment, for backward */
  const float* grad_loss;     /* [1] (device) */
  float* grad_depth;          /* [count] */
  void* workspace;            /* sde_variance_workspace_bytes(count), zero-filled once */
} sde_var_buffers;

size_t sde_variance_workspace_bytes(int64_t count);
int sde_variance_loss_forward(int64_t count, const sde_var_buffers* buf, void* stream);
int sde_variance_loss_backward(int64_t count, const sde_var_buffers* buf, void* stream);

/* silog_loss(variance_focus)(depth_est, depth_gt), detectron2/modeling/losses/losses.py:5-13 (supervised term,
 * caller MonoDepth2.py:107-110): mask = depth_gt > 1; d = log(est[mask]) - log(gt[mask]);
 * loss = sqrt(mean(d^2) - variance_focus * mean(d)^2) * 10.  An empty mask gives NaN, as in the reference. */
typedef struct sde_silog_buffers {
  const float* depth_est;     /* [count] */
  const float* depth_gt;      /* [count] */
  float* loss;                /* [1] */
  float* saved_stats;         /* [3] for backward */
  const float* grad_loss;     /* [1] (device) */
  float* grad_depth_est;      /* [count] */
  void* workspace;            /* sde_silog_workspace_bytes(count), zero-filled once */
} sde_silog_buffers;

size_t sde_silog_workspace_bytes(int64_t count);
int sde_silog_loss_forward(int64_t count, float variance_focus, const sde_silog_buffers* buf, void* stream);
int sde_silog_loss_backward(int64_t count, float variance_focus, const sde_silog_buffers* buf, void* stream);

/* disp_to_depth(disp, min_depth, max_depth) -> (scaled_disp, depth), detectron2/layers/depth_decoder.py:9-18
 * (DepthResNet.py:41,57; PackNet01.py:109): scaled = 1/max_depth + (1/min_depth - 1/max_depth) * disp, depth = 1/scaled.
 * `scaled_disp`, `grad_scaled_disp` and `grad_depth` are optional (NULL). */
typedef struct sde_disp_buffers {
  const float* disp;              /* [count] */
  float* scaled_disp;             /* [count] or NULL */
  float* depth;                   /* [count] */
  const float* grad_scaled_disp;  /* [count] or NULL */
  const float* grad_depth;        /* [count] or NULL */
  float* grad_disp;               /* [count] */
} sde_disp_buffers;

int sde_disp_to_depth_forward(int64_t count, float min_depth, float max_depth, const sde_disp_buffers* buf, void* stream);
int sde_disp_to_depth_backward(int64_t count, float min_depth, float max_depth, const sde_disp_buffers* buf, void* stream);

/* pose_vec2mat(vec), detectron2/geometry/pose_utils.py:98-137 (PoseNet.py:63; GooglePoseNet.py:85,206):
 * vec [B,6] = (tx, ty, tz, rx, ry, rz) -> [B,4,4] with R = Rx Ry Rz (euler2mat) and t in the last column. */
typedef struct sde_posevec_buffers {
  const float* vec;           /* [B,6] */
  float* pose;                /* [B,4,4] */
  const float* grad_pose;     /* [B,4,4] */
  float* grad_vec;            /* [B,6] */
} sde_posevec_buffers;

int sde_pose_vec2mat_forward(int32_t batch, const sde_posevec_buffers* buf, void* stream);
int sde_pose_vec2mat_backward(int32_t batch, const sde_posevec_buffers* buf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDE_LOSS_H_ */
