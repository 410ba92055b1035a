// microbenchmark: scattered 4-tap x 3-channel gather through LDG vs texture point fetches vs tex2Dgather (pitch2D linear memory)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int B = 12, C = 3, H = 192, W = 640;
__device__ __forceinline__ void coords(const float* __restrict__ jit, int b, int pix, int& x0, int& y0, float& ax, float& ay) {
  const int y = pix / W, x = pix - y * W;
  const float X = fminf(fmaxf((float)x + 20.0f + jit[(size_t)b * H * W + pix], 0.0f), (float)(W - 1));
  const float Y = fminf(fmaxf((float)y + 0.3f, 0.0f), (float)(H - 1));
  x0 = min((int)floorf(X), W - 2); y0 = min((int)floorf(Y), H - 2); ax = X - x0; ay = Y - y0;
}
template <int MODE>
__global__ void __launch_bounds__(256, 4) gather(const float* __restrict__ src, cudaTextureObject_t tex, const float* __restrict__ jit, float* __restrict__ out) {
  const int chunks = H * W / 1024;
  const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int pix = chunk * 1024 + it * 256 + threadIdx.x;
    int x0, y0; float ax, ay;
    coords(jit, b, pix, x0, y0, ax, ay);
    const float bx = 1.f - ax, by = 1.f - ay;
    float t[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (MODE == 0) {
        const float* q = src + ((size_t)(b * C + c) * H + y0) * W + x0;
        t[c][0] = __ldg(q); t[c][1] = __ldg(q + 1); t[c][2] = __ldg(q + W); t[c][3] = __ldg(q + W + 1);
      } else if (MODE == 1) {
        const float fy = (float)((b * C + c) * H + y0) + 0.5f, fx = (float)x0 + 0.5f;
        t[c][0] = tex2D<float>(tex, fx, fy); t[c][1] = tex2D<float>(tex, fx + 1.f, fy);
        t[c][2] = tex2D<float>(tex, fx, fy + 1.f); t[c][3] = tex2D<float>(tex, fx + 1.f, fy + 1.f);
      } else {
        const float fy = (float)((b * C + c) * H + y0) + 1.0f, fx = (float)x0 + 1.0f;
        const float4 g = tex2Dgather<float4>(tex, fx, fy, 0);   // w: (x0,y0) z: (x0+1,y0) x: (x0,y0+1) y: (x0+1,y0+1)
        t[c][0] = g.w; t[c][1] = g.z; t[c][2] = g.x; t[c][3] = g.y;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[(size_t)(b * C + c) * H * W + pix] = t[c][0] * bx * by + t[c][1] * ax * by + t[c][2] * bx * ay + t[c][3] * ax * ay;
  }
}
int main() {
  const size_t n = (size_t)B * C * H * W, np = (size_t)B * H * W;
  std::vector<float> hs(n), hj(np);
  srand(1);
  for (auto& v : hs) v = rand() / (float)RAND_MAX;
  for (auto& v : hj) { float u = 0; for (int k = 0; k < 12; ++k) u += rand() / (float)RAND_MAX; v = (u - 6.0f) * 4.6f; }
  float *src, *jit, *out[3];
  CK(cudaMalloc(&src, n * 4)); CK(cudaMalloc(&jit, np * 4));
  for (int k = 0; k < 3; ++k) CK(cudaMalloc(&out[k], n * 4));
  CK(cudaMemcpy(src, hs.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(jit, hj.data(), np * 4, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = src; rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  rd.res.pitch2D.width = W; rd.res.pitch2D.height = (size_t)B * C * H; rd.res.pitch2D.pitchInBytes = W * 4;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = B * H * W / 1024;
  auto run = [&](int mode, float* o) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      for (int k = 0; k < 20; ++k) {
        if (mode == 0) gather<0><<<grid, 256>>>(src, tex, jit, o);
        if (mode == 1) gather<1><<<grid, 256>>>(src, tex, jit, o);
        if (mode == 2) gather<2><<<grid, 256>>>(src, tex, jit, o);
      }
      cudaEventRecord(e1); cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 2) printf("mode %d: %.2f us per launch (%d px x 3 ch)\n", mode, ms * 1000 / 20, B * H * W);
    }
  };
  run(0, out[0]); run(1, out[1]); run(2, out[2]);
  std::vector<float> h0(n), h1(n), h2(n);
  cudaMemcpy(h0.data(), out[0], n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h1.data(), out[1], n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), out[2], n * 4, cudaMemcpyDeviceToHost);
  size_t d1 = 0, d2 = 0; for (size_t i = 0; i < n; ++i) { d1 += h0[i] != h1[i]; d2 += h0[i] != h2[i]; }
  printf("mismatches: tex point %zu, tex gather %zu of %zu\n", d1, d2, n);
  return 0;
}
