// DMMA probe 3: 4x4 register tile per warp (16 DMMAs per k-step), operands (a) loop-invariant,
// (b) loaded from shared memory every k-step, (c) as (b) with one DMUL per B fragment,
// (d) as (c) but software-pipelined (fragments of k-step n+1 loaded before the DMMAs of k-step n).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(double* out, long long* cyc, int ksteps, int ZS) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, lr = lane & 3, lc = lane >> 2;
  for (int i = threadIdx.x; i < 64 * ZS + 256; i += blockDim.x) sm[i] = 1.0 + 1e-6 * i;
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const double* sw = sm + 64 * ZS;
  int offA[4], offB[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) { offA[f] = 8 * (f + (warp & 3)) + lc; offB[f] = 8 * (f + 4 + (warp >> 2)) + lc; }
  double fa[4], fb[4];
#pragma unroll
  for (int f = 0; f < 4; ++f) { fa[f] = sm[lr * ZS + offA[f]]; fb[f] = sm[lr * ZS + offB[f]]; }
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int ks = 0; ks < ksteps; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
  } else if (MODE == 1 || MODE == 2) {
    for (int ks = 0; ks < ksteps; ++ks) {
      const int n = 4 * (ks & 15) + lr;
      const double* zr = sm + n * ZS;
      const double w = sw[n];
#pragma unroll
      for (int f = 0; f < 4; ++f) { fa[f] = zr[offA[f]]; fb[f] = zr[offB[f]]; if (MODE == 2) fb[f] *= w; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
    }
  } else {
    // software pipelined: next fragments in flight during the DMMAs
    double na[4], nb[4];
    for (int ks = 0; ks < ksteps; ++ks) {
      const int n = 4 * ((ks + 1) & 15) + lr;
      const double* zr = sm + n * ZS;
      const double w = sw[n];
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(na[f]) : "r"((unsigned)__cvta_generic_to_shared(zr + offA[f])));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(nb[f]) : "r"((unsigned)__cvta_generic_to_shared(zr + offB[f])));
      }
      if (MODE == 4) {
#pragma unroll
        for (int f = 0; f < 4; ++f) fb[f] *= w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
#pragma unroll
      for (int f = 0; f < 4; ++f) { fa[f] = na[f]; fb[f] = nb[f]; }
    }
  }
  const long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) r += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(int warps, double* out, long long* cyc) {
  const int ksteps = 4000, ZS = 108;
  const size_t smem = sizeof(double) * (64 * ZS + 256);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int r = 0; r < 2; ++r) probe<MODE><<<148, 32 * warps, smem>>>(out, cyc, ksteps, ZS);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("MODE=%d warps/SM=%2d: cycles per DMMA per SMSP %.2f (%.1f%% of the 16-cycle peak)  %s\n", MODE, warps,
         (double)hc / ksteps / 16 / (warps / 4.0), 100.0 * 16 / ((double)hc / ksteps / 16 / (warps / 4.0)),
         cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 512); cudaMalloc(&cyc, 8 * 148);
  for (int w : {4, 8, 16}) run<0>(w, out, cyc);
  for (int w : {4, 8, 16}) run<1>(w, out, cyc);
  for (int w : {4, 8, 16}) run<2>(w, out, cyc);
  for (int w : {4, 8, 16}) run<3>(w, out, cyc);
  for (int w : {4, 8, 16}) run<4>(w, out, cyc);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
