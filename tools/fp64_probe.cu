// Microbenchmark: FP64 throughput of DFMA (vector pipe) vs DMMA.8x8x4 (tensor sub-pipe) on
// sm_100a, alone and interleaved, plus exp/log1p/div.  Output feeds DESIGN.md's roofline notes.
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>   // per iteration: NF DFMAs (per thread) and NM DMMAs (per warp)
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b) {
  double f[16], m[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { f[i] = threadIdx.x * 1e-3 + i; m[i][0] = i; m[i][1] = -i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NF / 4; ++i) f[(r * (NF / 4) + i) & 15] = fma(f[(r * (NF / 4) + i) & 15], a, b);
#pragma unroll
      for (int i = 0; i < NM / 4; ++i) dmma884(m[(r * (NM / 4) + i) & 15][0], m[(r * (NM / 4) + i) & 15][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i] + m[i][0] + m[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe_fn(double* out, int iters, double a) {
  double x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = -1e-3 * threadIdx.x - i * 0.1 - a;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (MODE == 0) x[i] = -exp(x[i]);                 // x stays in (-1, 0)
      if (MODE == 1) x[i] = -log1p(-x[i] * 0.5);        // in (-0.41, 0)
      if (MODE == 2) x[i] = -1.0 / (1.0 - x[i]);        // in (-1, 0)
      if (MODE == 3) x[i] = -log(1.0 - x[i]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}

template <typename F>
float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const int grid = 148 * 8, block = 256, iters = 4000;
  double* out; cudaMalloc(&out, sizeof(double) * grid * block);
  const double warps = (double)grid * block / 32;
#define RUN(NF, NM)                                                                          \
  {                                                                                          \
    float ms = time_ms([&] { probe<NF, NM><<<grid, block>>>(out, iters, 0.999, 1e-3); });    \
    double dfma = warps * 32.0 * NF * iters * 2, dmma = warps * NM * iters * 512.0;          \
    printf("NF=%2d NM=%2d: %.3f ms  DFMA %.2f TF  DMMA %.2f TF  total %.2f TF\n", NF, NM, ms, \
           dfma / ms / 1e9, dmma / ms / 1e9, (dfma + dmma) / ms / 1e9);                      \
  }
  RUN(16, 0) RUN(0, 16) RUN(0, 8) RUN(16, 4) RUN(16, 8) RUN(8, 8) RUN(16, 16) RUN(4, 16) RUN(8, 16)
  const char* names[4] = {"exp", "log1p", "div", "log"};
  for (int mode = 0; mode < 4; ++mode) {
    float ms = 0;
    if (mode == 0) ms = time_ms([&] { probe_fn<0><<<grid, block>>>(out, iters / 4, 0.1); });
    if (mode == 1) ms = time_ms([&] { probe_fn<1><<<grid, block>>>(out, iters / 4, 0.1); });
    if (mode == 2) ms = time_ms([&] { probe_fn<2><<<grid, block>>>(out, iters / 4, 0.1); });
    if (mode == 3) ms = time_ms([&] { probe_fn<3><<<grid, block>>>(out, iters / 4, 0.1); });
    double calls = (double)grid * block * 4 * (iters / 4);
    printf("%s: %.3f ms  %.1f Gcall/s  (= %.1f DFMA-slots per call at 18.6 T DFMA/s)\n", names[mode], ms,
           calls / ms / 1e6, 18.6e12 / (calls / ms * 1e3));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
