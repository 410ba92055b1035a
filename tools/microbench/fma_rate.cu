// microbenchmark: issue rate of FFMA vs FFMA2 (packed f32x2) on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[8]; unsigned long long p[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  unsigned long long cc = ((unsigned long long)__float_as_uint(0.5f) << 32) | __float_as_uint(0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = fmaf(a[i], s, 0.5f + a[(i + 1) & 7] * 0.f);
        if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(ss), "l"(cc));
        if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s), "f"(a[(i+3)&7]));
        if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
        if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ss));
      }
    }
  }
  float r = 0; for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out) {
  int iters = 4096; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(out, 16, 1.0001f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winst = 148.0 * 8 * 8 * iters * 32.0;  // warp instructions
  double per_smsp_clk = winst / (148 * 4) / (ms * 1e-3 * 1.965e9);
  printf("%s: %.3f ms, %.3f warp-instr/clk/SMSP (at 1965 MHz)\n", name, ms, per_smsp_clk);
}
int main() { float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<2>("FFMA  3-reg", out); run<1>("FFMA2 packed", out); run<3>("FADD2 packed", out); run<4>("FMUL2 packed", out); run<2>("FFMA  3-reg", out); return 0; }
