// Shared device helpers for the sm_100a view-synthesis loss kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sde_loss.h"

// Checked build (-DSDE_CHECKED; tools/build_variant.sh checked -DSDE_CHECKED): index assertions on the shared-memory
// plane positions, the argmin / list arrays and the global pixel offsets of the tile kernels.  compute-sanitizer is not
// available on the GPU pool, so the odd-size parity tests are run once against this build instead; a violated
// assertion prints its location and traps (the context dies, every later test fails).
#ifdef SDE_CHECKED
#include <stdio.h>
#define SDE_CHECK(cond)                                                                                          \
  do {                                                                                                           \
    if (!(cond)) {                                                                                               \
      printf("SDE_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                                  \
    }                                                                                                            \
  } while (0)
#else
#define SDE_CHECK(cond) ((void)0)
#endif

namespace sde {

// ---------------------------------------------------------------------------------------------
// Packed fp32 pairs.  Blackwell (sm_100) issues add/mul/fma on two fp32 lanes per instruction
// (PTX *.f32x2 -> SASS FADD2/FMUL2/FFMA2, with free LO/HI swizzles); every kernel here maps two
// horizontally adjacent pixels onto one f2 so the SSIM arithmetic costs half the issue slots.
// ---------------------------------------------------------------------------------------------
// An f2 lives in one 64-bit register pair from load to store, so no repacking MOVs are needed.
struct f2 {
  unsigned long long v;
};

__device__ __forceinline__ f2 mk2(float x, float y) {
  f2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r.v) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float lo(f2 a) {
  float x, y;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return x;
}
__device__ __forceinline__ float hi(f2 a) {
  float x, y;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return y;
}
__device__ __forceinline__ f2 bc2(float v) { return mk2(v, v); }
__device__ __forceinline__ f2 swp(f2 a) { return mk2(hi(a), lo(a)); }   // folds into a LO_HI operand swizzle
__device__ __forceinline__ f2 operator+(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// a - b as fma(b, -1, a): one FFMA2 with an immediate
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return fma2(b, bc2(-1.0f), a); }
// a / b: hardware reciprocal plus one Newton correction on the quotient (two packed FMAs).
// Within 1 ulp of the IEEE quotient and exactly 1 when a == b.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ f2 div2(f2 a, f2 b) {
  const f2 r = mk2(rcp_approx(lo(b)), rcp_approx(hi(b)));
  const f2 q = a * r;
  return fma2(r, fma2(q * bc2(-1.0f), b, a), q);
}
// -(a / b), bit for bit the negated div2(a, b), in one packed instruction less: the sign rides on the reciprocal
// (the negation of b folds into the MUFU operand, rcp(-b) == -rcp(b) exactly), so the residual a - q b is
// fma(-q, b, a) without a separate negation.  Callers fold the sign into their next constant.
__device__ __forceinline__ f2 ndiv2(f2 a, f2 b) {
  const f2 r = mk2(rcp_approx(-lo(b)), rcp_approx(-hi(b)));   // -1/b
  const f2 q = a * r;                                           // -q
  return fma2(r, fma2(q, b, a), q);                             // -(q + (a - q b)/b)
}
__device__ __forceinline__ f2 ldg2(const float* p) {   // 8-byte aligned, global, read-only path
  f2 r;
  r.v = __ldg(reinterpret_cast<const unsigned long long*>(p));
  return r;
}
__device__ __forceinline__ f2 ld2(const float* p) {   // 8-byte aligned
  f2 r;
  r.v = *reinterpret_cast<const unsigned long long*>(p);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): the kernels of a step are chained on one stream, and each one lets the
// next be scheduled while it drains (launch_dependents, once the heavy part of a CTA is done) and itself touches no tensor
// before the kernels ahead of it in the stream have completed and flushed (pdl_wait).  What runs before pdl_wait
// is index arithmetic, shared-memory initialisation and barrier set-up only, so the effect is that the launch
// latency and the CTA prologue of a kernel overlap the tail of its predecessor.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Flags between kernels that overlap under programmatic dependent launch (mono_params.cuh: flow).  The producer's CTA
// stores its data, synchronises, and ONE thread fences and sets the flag; a consumer thread spins with acquire loads.
// Every CTA of the producer grid is resident (or has exited) before the consumer grid is scheduled -- that is what
// the launch attribute guarantees once all of them have executed launch_dependents -- so the spin cannot starve the
// producer.  A flag that stays clear for seconds means a broken protocol: trap instead of hanging the device.
__device__ __forceinline__ void flag_publish(unsigned* f) {
  __threadfence();
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
}
__device__ __forceinline__ void flag_wait(const unsigned* f) {
  unsigned v, spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v != 0u) return;
    __nanosleep(100);
    if (++spins > (1u << 24)) __trap();
  }
}
// What the forward kernel of the same step produced (argmin bytes, smoothness gradient, per-image statistics) is read
// with ld.global.cg: under flow the producer grid may still be running, and L1 must not keep a line that straddles a
// finished and an unfinished image (these tensors are [B, h, w]: one image follows the other).
template <class T>
__device__ __forceinline__ T ld_prod(const T* p) { return __ldcg(p); }
// The derivative planes of the warp kernel keep the L1-allocating read-only path (ld.global.nc).  Measured at cfg2: the
// backward kernel takes 130.8 us with it, 138.3 us with ld.global.nc.L1::no_allocate (SASS LDG.E.NA) and 138-140 us with
// ld.global.cg -- the 60-wide gradient tiles share 128-byte lines with their neighbours, and the compiler issues these
// loads a phase early.  Why no stale line can be hit under flow: a step's warp kernel is an ordinary launch behind the
// previous step's backward kernel, so L1 is invalidated before the planes are rewritten; within the step nobody reads a
// line of planes 3..8 of an image before the image's flag, and such a line never holds another image's L1-read data
// (planes 0..2 on either side go through the copy engine).  `launder` hides a pointer's value from the compiler at a
// point behind the flag wait, so that no such load can be hoisted above it.
template <class T>
__device__ __forceinline__ T ld_plane(const T* p) { return __ldg(p); }
__device__ __forceinline__ unsigned launder_zero() { unsigned z = 0; asm volatile("" : "+r"(z)); return z; }
#ifdef SDE_TRACE
// developer build (tools/trace_step.py): per-CTA %globaltimer stamps [kernel][block][start, end, SM id, after the wait]
constexpr int kTraceBlocks = 8192;
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void trace_begin(unsigned long long* tr, int k) {
  if (tr && threadIdx.x == 0 && blockIdx.x < kTraceBlocks) {
    unsigned sm; asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    tr[(k * kTraceBlocks + blockIdx.x) * 4] = gtimer(); tr[(k * kTraceBlocks + blockIdx.x) * 4 + 2] = sm;
  }
}
__device__ __forceinline__ void trace_mark(unsigned long long* tr, int k, int slot) {
  if (tr && threadIdx.x == 0 && blockIdx.x < kTraceBlocks) tr[(k * kTraceBlocks + blockIdx.x) * 4 + slot] = gtimer();
}
#define SDE_TRACE_BEGIN(p, k) trace_begin((p).trace, k)
#define SDE_TRACE_MARK(p, k, slot) trace_mark((p).trace, k, slot)
#else
#define SDE_TRACE_BEGIN(p, k)
#define SDE_TRACE_MARK(p, k, slot)
#endif
// data another SM wrote with ordinary stores is read by the copy engine (async proxy) next
__device__ __forceinline__ void flag_proxy_fence() { asm volatile("fence.proxy.async;" ::: "memory"); }

#ifndef __CUDACC_RTC__
// Host side: launch with the programmatic-stream-serialization attribute (SDE_DISABLE_PDL=1: plain launch).
bool pdl_enabled(int which);
template <class... KArgs, class... Args>
inline cudaError_t launch_chained(int which, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(which) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#ifndef __CUDACC_RTC__
// CTAs of `kernel` that are resident at once on the current device (occupancy x SMs): the grid of a persistent launch
template <class K>
inline unsigned resident_ctas(K kernel, int threads, size_t smem) {
  int dev = 0, sms = 0, occ = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess)
    return 0;
  return (unsigned)(occ * sms);
}
#endif

// Fence between publishing a partial-result slot and taking the ticket (and, on the reading side, between the ticket and
// the slots).  SDE_FENCE_ACQREL: the release / acquire form instead of the sequentially consistent __threadfence().
__device__ __forceinline__ void publish_fence() {
#ifdef SDE_FENCE_ACQREL
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
  __threadfence();
#endif
}

// ---------------------------------------------------------------------------------------------
// Deterministic scatter accumulation: 64-bit fixed point (scale 2^44, resolution 5.7e-14) with INTEGER atomics -- the
// sum does not depend on the order of the additions, unlike a float atomicAdd.  Range: a contribution must be below
// 2^15 in magnitude and a sum below 2^16; what is not -- including NaN / inf, which ATen's float scatter would
// propagate -- poisons the element (atomicMax to 2^62, idempotent), and fix_to_float returns NaN for any accumulator at
// or beyond +-2^60.
// ---------------------------------------------------------------------------------------------
constexpr double kFixScale = 17592186044416.0;        // 2^44
constexpr double kFixInv = 1.0 / 17592186044416.0;
__device__ __forceinline__ void fix_add(long long* dst, float v) {
  // |v| < 2^15: v * 2^44 is exact in fp32 (a power-of-two scale, no overflow) and below 2^59, so the conversion is exact
  // too -- the same integer as the fp64 product, without the fp64 multiply and the two fp64 conversions per contribution
  if (fabsf(v) < 32768.0f)
    atomicAdd(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)__float2ll_rn(v * 17592186044416.0f));
  else
    atomicMax(dst, 1LL << 62);   // NaN, inf or out of range
}
__device__ __forceinline__ float fix_to_float(long long a) {
  return (a >= (1LL << 60) || a <= -(1LL << 60)) ? __int_as_float(0x7fc00000) : (float)((double)a * kFixInv);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums 16 per-lane values over the warp with 16 shuffles instead of 80: at every step a lane keeps one half of its
// values and hands the other half to its partner (offsets 16, 8, 4, 2), then the last exchange (offset 1) completes
// the sums.  Returns the total of value warp_slot(lane) -- lanes 2k and 2k+1 hold the same total.  Fixed order.
__device__ __forceinline__ int warp_slot(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = up ? v[i + half] : v[i];
      const float send = up ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
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
  // (pointer form: K + b * 9 may also be the k[] of a CamRaw with b == 0)
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

// The same in two steps, so that the global loads can be in flight while the thread does something else:
// raw loads (no dependent arithmetic) now, camera terms later.
struct CamRaw {
  float k[9], T[16];
};
__device__ __forceinline__ void load_cam_raw(CamRaw& r, const float* __restrict__ K, const float* __restrict__ pose, int b) {
#pragma unroll
  for (int i = 0; i < 9; ++i) r.k[i] = __ldg(K + b * 9 + i);
#pragma unroll
  for (int i = 0; i < 16; ++i) r.T[i] = __ldg(pose + b * 16 + i);
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
