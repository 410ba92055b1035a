// Parameter blocks of the stand-alone operators (ops.cu).
#pragma once
#include <stdint.h>

#include "../../include/sde_loss.h"

namespace sde {

struct VsParams {
  int B, C, h, w;
  int t_per_pixel;              // t is [B,3,h,w] (1) or [B,3] (0)
  const float* image;           // [B,C,h,w]
  const float* depth;           // [B,1,h,w]
  const float* K;               // [B,3,3] at this size
  const float* R;               // [B,3,3]
  const float* t;
  // forward outputs
  float* sampled;               // [B,C,h,w]
  float* depth_in_b;            // [B,1,h,w] or null
  float* coords;                // [B,h,w,2] or null
  uint8_t* valid;               // [B,1,h,w] or null
  // backward
  const float* g_sampled;       // [B,C,h,w]
  const float* g_depth_in_b;    // or null
  const float* g_coords;        // or null
  float* g_depth;               // [B,1,h,w]
  float* g_R;                   // [B,3,3]
  float* g_t;                   // [B,3,h,w] or [B,3]
  long long* g_image_fix;       // [B,C,h,w] 64-bit fixed-point accumulator (zero on entry and on exit) or null
  float* partials;              // [B][blocks][12]
  unsigned* counters;           // [B]
};

struct SsimParams {
  int B, C, h, w;
  float c1, c2;
  int mode;                     // 0 both factors, 1 C1 = inf, 2 C2 = inf
  const float* x;
  const float* y;
  const float* weight;          // [B,1,h,w] or null (plain SSIM)
  float* out;                   // [B,C,h,w]
  float* avg_w;                 // [B,1,h,w] or null
  const float* g_out;
  float* g_x;                   // or null
  float* g_y;                   // or null
  float* coef;                  // [6][B*C][h*w]
};

struct SmoothParams {
  int B, C, h, w;
  const float* depth;           // [B,1,h,w]
  const float* image;           // [B,C,h,w]
  float* loss;                  // [1]
  float* stats;                 // [B][2] mean inverse depth, per-image loss
  float* partials;              // [B][blocks][4]
  double* fin;                  // [B]
  unsigned* counters;           // [1 + B]
  const float* g_loss;          // [1]
  float* g_depth;
};


}  // namespace sde

namespace sde {

struct McParams {   // motion_consistency_loss, translation term
  int B, h, w;
  const float* coords;      // [B,h,w,2]
  const float* mask;        // [B,1,h,w]
  const float* R;           // [B,3,3] A->B
  const float* t_ab;        // [B,3,h,w] translation field A->B, or its residual part when pose_ab is given (may then be nullptr)
  const float* t_ba;        // [B,3,h,w] likewise B->A
  const float* pose_ab;     // optional [B,4,4]: its translation column is added to t_ab (the field is never materialised)
  const float* pose_ba;     // optional [B,4,4]
  float* loss;              // [1]
  const float* g_loss;
  float* g_t_ab;            // [B,3,h,w] (nullptr when t_ab is)
  long long* g_t_ba_fix;    // [B,3,h,w] fixed-point accumulator (zero on entry and exit)
  float* g_R;               // [B,3,3]
  float* g_pose_t_ab;       // optional [B,3]: sum over pixels of d / d t_ab
  float* g_pose_t_ba;       // optional [B,3]: sum over pixels and in-range taps of d / d t_hat
  int scatter;              // backward: scatter d / d t_hat into g_t_ba_fix (t_ba given)
  float* slots;             // forward: [B*blocks]; backward: [B*blocks][16]
  unsigned* counters;       // [1 + B]
};

struct MfieldParams {   // fused regularisers of the residual translation field (motion_field.cu)
  int B, h, w;
  const float* pose;        // [B,4,4] or nullptr (zero translation)
  const float* field;       // [B,3,h,w]
  float* losses;            // [2] smoothness, sparsity
  float* stats;             // [B][12]
  const float* g_losses;    // [2]
  float* g_field;           // [B,3,h,w]
  float* g_pose_t;          // [B,3] or nullptr
  float* slots;             // [B][blocks][8]
  double* fin;              // [B][2]
  unsigned* counters;       // [1 + B]
};

struct MregParams {   // motion_smoothness_loss_fn / motion_sparsity_loss_fn
  int B, C, h, w;
  const float* field;       // [B,C,h,w]
  float* loss;              // [1]
  float* stats;             // sparsity: [B*C] mean |m|
  const float* g_loss;
  float* g_field;
  float* slots;             // [B*C*blocks]
  unsigned* counters;       // [1 + B*C]
};

struct VarParams {   // variance_loss
  long long n;              // elements of the depth tensor
  const float* depth;
  float* loss;              // [1]
  float* stats;             // [2] mean, centred second moment
  const float* g_loss;
  float* g_depth;
  float* slots;             // [blocks]
  unsigned* counters;       // [1]
};

constexpr int kPyrFrames = SDE_MAX_SOURCES + 1;   // target + sources

struct PyramidParams {   // resize_img of several frames to several sizes, one launch
  int n_frames, n_levels, planes, sh, sw;
  int dh[SDE_MAX_SCALES], dw[SDE_MAX_SCALES];
  int blk_start[SDE_MAX_SCALES + 1];   // first pixel block of level l
  float rh[SDE_MAX_SCALES], rw[SDE_MAX_SCALES];
  const void* src[kPyrFrames];          // fp32 planes, or uint8 planes (src_u8: value = byte / 255, torchvision ToTensor)
  float* dst[kPyrFrames][SDE_MAX_SCALES];
  int src_u8;
};

struct SilogParams {   // silog_loss
  long long n;
  const float* est;
  const float* gt;
  float vf;                 // variance_focus
  float* loss;              // [1]
  float* stats;             // [3] mean(d), sqrt(mean(d^2) - vf mean(d)^2), number of masked elements
  const float* g_loss;
  float* g_est;
  float* slots;             // [blocks][3]
  unsigned* counter;        // [1]
};

struct DispParams {    // disp_to_depth
  long long n;
  const float* disp;
  float min_disp, range;    // 1 / max_depth, 1 / min_depth - 1 / max_depth
  float* scaled;            // optional
  float* depth;
  const float* g_scaled;    // optional
  const float* g_depth;     // optional
  float* g_disp;
};

struct PoseVecParams { // pose_vec2mat
  int B;
  const float* vec;         // [B,6]
  float* mat;               // [B,4,4]
  const float* g_mat;       // [B,4,4]
  float* g_vec;             // [B,6]
};

}  // namespace sde
