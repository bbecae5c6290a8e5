// Latency / issue cost of the instruction kinds of the mixed-precision step kernel on one warp per SM sub-partition:
// FFMA, DFMA, F2F (fp64 -> fp32 and back), MUFU.RCP, LDS.  cycles per instruction for a dependent chain (latency) and
// for 8 independent chains (issue interval).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_probe.bin scripts/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int KIND, int ILP> __global__ void probe(double* out, long long* clk, double seed) {
  double d[8]; float f[8];
  for (int i = 0; i < 8; ++i) { d[i] = seed + i + threadIdx.x; f[i] = (float)d[i]; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N / 8; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < ILP; ++c) {
        if (KIND == 0) f[c] = fmaf(f[c], 0.999f, 0.5f);
        if (KIND == 1) d[c] = fma(d[c], 0.999, 0.5);
        if (KIND == 2) { f[c] = (float)d[c]; d[c] = (double)f[c] + 1.0; }     // F2F.F32.F64 + F2F.F64.F32 + DADD
        if (KIND == 3) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[c])); f[c] = r + 1.5f; }
        if (KIND == 4) d[c] = d[c] + 1.0;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += d[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int KIND, int ILP> void run(const char* name, double* out, long long* clk) {
  long long h[148];
  probe<KIND, ILP><<<148, 32>>>(out, clk, 1.0);
  cudaDeviceSynchronize();
  probe<KIND, ILP><<<148, 32>>>(out, clk, 1.0);
  cudaDeviceSynchronize();
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
  const double per = s / 148 / (double)(N * ILP);
  printf("%-46s ILP %d: %.2f cycles per op-group (1 warp / SM)\n", name, ILP, per);
}
int main() {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&clk, 148 * 8);
  run<0, 1>("FFMA dependent", out, clk); run<0, 8>("FFMA 8 chains", out, clk);
  run<1, 1>("DFMA dependent", out, clk); run<1, 8>("DFMA 8 chains", out, clk);
  run<4, 1>("DADD dependent", out, clk); run<4, 8>("DADD 8 chains", out, clk);
  run<2, 1>("F2F f64->f32, F2F f32->f64, DADD dependent", out, clk); run<2, 8>("same, 8 chains", out, clk);
  run<3, 1>("MUFU.RCP + FADD dependent", out, clk); run<3, 8>("same, 8 chains", out, clk);
  return 0;
}
