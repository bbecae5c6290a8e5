// Instruction-fetch probe: cycles per instruction of cold straight-line code on one SM, for
//   A: 1 warp / SM            B: 4 warps / SM, same code        C: 4 warps / SM, four distinct copies of the code
//   D: the same instruction count as a loop whose body fits the instruction caches
// Used to decide how the model-specialised step kernel should be laid out.
#include <cstdio>
#include <cuda_runtime.h>

#define R8(x) x x x x x x x x
#define R64(x) R8(R8(x))
#define R512(x) R8(R64(x))
#define BODY(c) a0 = fmaf(a0, c, b0); a1 = fmaf(a1, c, b1); a2 = fmaf(a2, c, b2); a3 = fmaf(a3, c, b3); \
                a4 = fmaf(a4, c, b0); a5 = fmaf(a5, c, b1); a6 = fmaf(a6, c, b2); a7 = fmaf(a7, c, b3);

template <int ID> __device__ __noinline__ float straight(float x) {
  float a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  const float b0 = 0.5f + ID, b1 = 0.25f, b2 = 0.125f, b3 = 0.0625f;
  const float c = 0.999f - 0.001f * ID;
  R512(BODY(c)) R512(BODY(c))   // 2 * 512 * 8 = 8192 FFMA = 128 KB of code
  return a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__device__ __noinline__ float looped(float x, int iters) {
  float a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  const float b0 = 0.5f, b1 = 0.25f, b2 = 0.125f, b3 = 0.0625f, c = 0.999f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) { R64(BODY(c)) }   // 512 FFMA = 8 KB body
  return a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void probe(int variant, float* out, long long* clk) {
  const int warp = threadIdx.x >> 5;
  float r = 0;
  const long long t0 = clock64();
  if (variant == 3) r = looped(threadIdx.x, 16);
  else if (variant == 2) {
    switch (warp) { case 0: r = straight<0>(threadIdx.x); break; case 1: r = straight<1>(threadIdx.x); break;
                    case 2: r = straight<2>(threadIdx.x); break; default: r = straight<3>(threadIdx.x); }
  } else r = straight<0>(threadIdx.x);
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 4 + warp] = t1 - t0;
}
int main() {
  float* out; long long* clk; long long h[148 * 4];
  cudaMalloc(&out, 148 * 128 * 4); cudaMalloc(&clk, sizeof(h));
  const char* names[4] = {"A 1 warp/SM straight", "B 4 warps/SM same straight code", "C 4 warps/SM distinct straight code", "D 4 warps/SM loop (8 KB body x16)"};
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 4; ++v) {
      const int threads = v == 0 ? 32 : 128;
      cudaMemset(clk, 0, sizeof(h));
      probe<<<128, threads>>>(v, out, clk);
      cudaDeviceSynchronize();
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0, mx = 0; int n = 0;
      for (int b = 0; b < 128; ++b) for (int w = 0; w < threads / 32; ++w) { s += h[b * 4 + w]; if (h[b * 4 + w] > mx) mx = h[b * 4 + w]; ++n; }
      printf("rep %d  %-40s mean %.0f cycles, max %.0f  -> %.2f cycles/instr (8192 FFMA)\n", rep, names[v], s / n, mx, s / n / 8192.0);
    }
  return 0;
}
