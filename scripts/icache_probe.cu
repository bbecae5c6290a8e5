// Instruction-cache probe, second generation: for straight-line bodies of S KB executed R times inside one launch,
// cycles per instruction of the FIRST pass (cold within the launch) and of the later passes (warm if the body fits an
// instruction cache level).  Answers: (1) up to which code size does re-executed code run at the issue rate, (2) does
// the instruction cache keep a kernel's code from one launch to the next, (3) what two / four different streams on one
// SM cost each other.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/icache_probe.bin scripts/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define R4(x) x x x x
#define R8(x) x x x x x x x x
#define R32(x) R4(R8(x))
#define BODY(c) a0 = fmaf(a0, c, b0); a1 = fmaf(a1, c, b1); a2 = fmaf(a2, c, b2); a3 = fmaf(a3, c, b3); \
                a4 = fmaf(a4, c, b0); a5 = fmaf(a5, c, b1); a6 = fmaf(a6, c, b2); a7 = fmaf(a7, c, b3);
// one chunk = 32 * 8 FFMA = 256 instructions = 4 KB
#define CHUNK R32(BODY(c))

template <int KB4, int ID> __device__ __forceinline__ void body(float& a0, float& a1, float& a2, float& a3, float& a4, float& a5, float& a6,
                                                                 float& a7, const float b0, const float b1, const float b2, const float b3, const float) {
  constexpr float c = 0.999f - 0.001f * ID;   // an immediate: the copies differ in their encoding
  if constexpr (KB4 > 0) { CHUNK body<KB4 - 1, ID>(a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3, c); }
}
// KB4 chunks of 4 KB, R passes; clk[pass] per warp
template <int KB4, int ID> __global__ void __launch_bounds__(128) probe(int passes, float* out, long long* clk, int nstream) {
  const int warp = threadIdx.x >> 5;
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b0 = 0.5f + ID, b1 = 0.25f, b2 = 0.125f, b3 = 0.0625f, c = 0.999f - 0.001f * ID;
  long long t = clock64();
#pragma unroll 1
  for (int p = 0; p < passes; ++p) {
    body<KB4, ID>(a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3, c);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clk[((size_t)blockIdx.x * 4 + warp) * 8 + p] = t1 - t;
    t = t1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// nstream distinct copies of the same-size body on the warps of one CTA (warp w runs copy w % nstream)
template <int KB4, int ID> __device__ __noinline__ float copy(float x) {
  float a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  const float b0 = 0.5f + ID, b1 = 0.25f, b2 = 0.125f, b3 = 0.0625f;
  body<KB4, ID>(a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3, 0.f);
  return a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int KB4> __global__ void __launch_bounds__(128) probe_streams(int passes, float* out, long long* clk, int nstream) {
  const int warp = threadIdx.x >> 5;
  float r = threadIdx.x;
  long long t = clock64();
#pragma unroll 1
  for (int p = 0; p < passes; ++p) {
    switch (warp % nstream) {
      case 0: r = copy<KB4, 0>(r); break;
      case 1: r = copy<KB4, 1>(r); break;
      case 2: r = copy<KB4, 2>(r); break;
      default: r = copy<KB4, 3>(r); break;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clk[((size_t)blockIdx.x * 4 + warp) * 8 + p] = t1 - t;
    t = t1;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

static float* g_out; static long long* g_clk; static long long h[592 * 4 * 8];
template <class K> static void run(const char* name, K kern, int kb, int grid, int threads, int passes, int nstream, int launches) {
  for (int l = 0; l < launches; ++l) {
    cudaMemset(g_clk, 0, sizeof(h));
    kern<<<grid, threads>>>(passes, g_out, g_clk, nstream);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, g_clk, sizeof(h), cudaMemcpyDeviceToHost);
    const double ninstr = kb * 64.0;   // 16-byte instructions
    printf("%-26s %4d KB grid %3d warps %d launch %d:", name, kb, grid, threads / 32, l);
    for (int p = 0; p < passes; ++p) {
      double s = 0; int n = 0;
      for (int b = 0; b < grid; ++b) for (int w = 0; w < threads / 32; ++w) { s += h[((size_t)b * 4 + w) * 8 + p]; ++n; }
      printf(" %.2f", s / n / ninstr);
    }
    printf("  cycles/instr per pass\n");
  }
}
#define SWEEP(KB4) run("same stream", probe<KB4, 0>, KB4 * 4, 148, 128, 5, 1, 2); run("same stream", probe<KB4, 0>, KB4 * 4, 148, 32, 5, 1, 1);
int main() {
  cudaMalloc(&g_out, 592 * 128 * 4); cudaMalloc(&g_clk, sizeof(h));
  SWEEP(1) SWEEP(2) SWEEP(4) SWEEP(6) SWEEP(8) SWEEP(10) SWEEP(12) SWEEP(16) SWEEP(24) SWEEP(32)
  // distinct streams on one SM: 2 and 4 copies, 16 KB and 64 KB each
  run("2 streams (4 warps)", probe_streams<4>, 16, 148, 128, 5, 2, 1);
  run("4 streams (4 warps)", probe_streams<4>, 16, 148, 128, 5, 4, 1);
  run("2 streams (4 warps)", probe_streams<16>, 64, 148, 128, 5, 2, 1);
  run("4 streams (4 warps)", probe_streams<16>, 64, 148, 128, 5, 4, 1);
  run("2 streams (2 warps)", probe_streams<16>, 64, 148, 64, 5, 2, 1);
  // two CTAs per SM on the same code (296 CTAs): does a second tile ride on the first one's fetches?
  run("same stream 2 CTA/SM", probe<16, 0>, 64, 296, 128, 5, 1, 1);
  run("same stream 4 CTA/SM", probe<16, 0>, 64, 592, 128, 5, 1, 1);
  return 0;
}
