// Shared device helpers for the sm_100a view-synthesis loss kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sde_loss.h"

namespace sde {

// ---------------------------------------------------------------------------------------------
// Packed fp32 pairs.  Blackwell (sm_100) issues add/mul/fma on two fp32 lanes per instruction
// (PTX *.f32x2 -> SASS FADD2/FMUL2/FFMA2, with free LO/HI swizzles); every kernel here maps two
// horizontally adjacent pixels onto one f2 so the SSIM arithmetic costs half the issue slots.
// ---------------------------------------------------------------------------------------------
struct f2 {
  float x, y;
};

__device__ __forceinline__ unsigned long long f2_pack(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long r) {
  f2 a;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r));
  return a;
}
__device__ __forceinline__ f2 mk2(float x, float y) {
  f2 r;
  r.x = x;
  r.y = y;
  return r;
}
__device__ __forceinline__ f2 bc2(float v) { return mk2(v, v); }
__device__ __forceinline__ f2 swp(f2 a) { return mk2(a.y, a.x); }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(r);
}
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return fma2(b, bc2(-1.0f), a); }
__device__ __forceinline__ f2 neg2(f2 a) { return mk2(-a.x, -a.y); }
__device__ __forceinline__ f2 abs2(f2 a) { return mk2(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ f2 sat2(f2 a) { return mk2(__saturatef(a.x), __saturatef(a.y)); }
// a / b with the fast reciprocal (2 ulp); operands here are O(1e-8 .. 10), far from the
// 2^126 range where __fdividef degrades.
__device__ __forceinline__ f2 fdiv2(f2 a, f2 b) { return mk2(__fdividef(a.x, b.x), __fdividef(a.y, b.y)); }
// a / b: hardware reciprocal plus one Newton correction on the quotient (two packed FMAs).
// Within 1 ulp of the IEEE quotient and exactly 1 when a == b.
__device__ __forceinline__ f2 div2(f2 a, f2 b) {
  const f2 r = mk2(__frcp_rn(b.x), __frcp_rn(b.y));
  const f2 q = a * r;
  return fma2(r, fma2(neg2(q), b, a), q);
}
__device__ __forceinline__ f2 ld2(const float* p) {
  float2 v = *reinterpret_cast<const float2*>(p);
  return mk2(v.x, v.y);
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reflect index as nn.ReflectionPad2d(1) does (-1 -> 1, n -> n-2), then clamp for tiles that
// overhang the image (those positions are never used by a valid output)
__device__ __forceinline__ int reflect_clamp(int i, int n) {
  i = i < 0 ? -i : i;
  i = i >= n ? 2 * (n - 1) - i : i;
  return min(max(i, 0), n - 1);
}

// Per-(sample, source) camera terms, computed once per CTA (camera.py:25-37,172-178).
struct Cam {
  float ki[9];   // K^-1 as the reference builds it (clone of K with 4 entries replaced)
  float fx, fy, sk, cx, cy;
};
struct Proj {
  float m[9];    // M = K R
  float tau[3];  // tau = K t
  float r[9];    // R (backward needs R^T)
};

__device__ __forceinline__ void load_cam(Cam& c, float k[9], const float* __restrict__ K, int b, float sx, float sy) {
#pragma unroll
  for (int i = 0; i < 9; ++i) k[i] = K[b * 9 + i];
  // scale_intrinsics, camera.py:14-22
  k[0] *= sx; k[2] *= sx; k[4] *= sy; k[5] *= sy;
#pragma unroll
  for (int i = 0; i < 9; ++i) c.ki[i] = k[i];
  c.ki[0] = 1.0f / k[0];
  c.ki[4] = 1.0f / k[4];
  c.ki[2] = -1.0f * k[2] / k[0];
  c.ki[5] = -1.0f * k[5] / k[4];
  c.fx = k[0]; c.sk = k[1]; c.cx = k[2]; c.fy = k[4]; c.cy = k[5];
}

__device__ __forceinline__ void load_proj(Proj& q, const float k[9], const float* __restrict__ pose, int b) {
  const float* T = pose + b * 16;
  float t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) q.r[i * 3 + j] = T[i * 4 + j];
    t[i] = T[i * 4 + 3];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
      q.m[i * 3 + j] = k[i * 3] * q.r[j] + k[i * 3 + 1] * q.r[3 + j] + k[i * 3 + 2] * q.r[6 + j];
    q.tau[i] = k[i * 3] * t[0] + k[i * 3 + 1] * t[1] + k[i * 3 + 2] * t[2];
  }
}

}  // namespace sde
