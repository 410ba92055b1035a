// TMA (cp.async.bulk.tensor) + mbarrier wrappers for staging halo'd image tiles into shared memory.
//
// A tile plane (18 rows x 68 floats, see mono_device.cuh) is one 3-D box {68, 18, 1} of a tensor map over
// a [planes, h, w] fp32 tensor: the copy engine writes it densely (row pitch 68 floats = kPitch) while the
// CTA's warps do arithmetic, no thread issues a load and no register holds data in flight.  Coordinates may
// start outside the image (halo): out-of-bounds elements arrive as zeros and the one-pixel reflection ring
// is patched in shared memory afterwards (reflect_fixup).
#pragma once
#include <cuda.h>          // CUtensorMap (type only; the encoder is resolved through the runtime, no -lcuda)
#include <cuda_runtime.h>
#include <stdint.h>

namespace sde {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barrier visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// generic-proxy accesses (LDS/STS) before, async-proxy accesses (TMA writes) after
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one {68, 18, 1} box starting at (x, y, z) -> smem_dst (128-byte aligned); completes on `bar`
__device__ __forceinline__ void tma_load_plane(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace sde
