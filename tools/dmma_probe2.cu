// DMMA / DMUL interleave probe: does switching between DMUL and DMMA in the FP64 pipe cost bubbles?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void vdmul(double& m, double s) {
  asm volatile("mul.f64 %0, %0, %1;\n" : "+d"(m) : "d"(s));
}
// MODE 0: 12 DMUL then 16 DMMA.  MODE 1: interleaved D M M? (12 groups).  MODE 2: 24 DMUL then 32 DMMA.
// MODE 3: 16 DMMA only. MODE 4: DMUL results feed DMMA operands (grouped). MODE 5: feed, interleaved
template <int MODE>
__global__ void probe(double* out, long long* cyc, int iters, double s) {
  double c[16][2], a[16], b[16], m[12];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = i; c[i][1] = -i; a[i] = 1e-3 * threadIdx.x + i; b[i] = 1.0 - 1e-3 * i + s; }
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 12; ++i) vdmul(m[i], s);
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { if (i < 12) vdmul(m[i], s); dmma(c[i][0], c[i][1], a[i], b[i]); }
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 12; ++i) vdmul(m[i], s);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
    } else if (MODE == 4) {
      double t[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) { t[i] = m[i]; vdmul(t[i], s); }
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a[i], t[i % 12]);
    } else if (MODE == 5) {
      double t[12];
#pragma unroll
      for (int i = 0; i < 16; ++i) { if (i < 12) { t[i] = m[i]; vdmul(t[i], s); } dmma(c[i][0], c[i][1], a[i], t[i % 12]); }
    }
  }
  const long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 12; ++i) r += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(int wps, double* out, long long* cyc) {
  const int iters = 2000;
  const int ctas = wps > 8 ? 2 : 1;
  const int block = 32 * wps / ctas;
  for (int r = 0; r < 2; ++r) probe<MODE><<<148 * ctas, block>>>(out, cyc, iters, 1.0000001);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (MODE == 2) ? 2.0 : 1.0;
  const double warps_per_smsp = wps / 4.0;
  printf("MODE=%d warps/SM=%2d: cycles per (12 DMUL + 16 DMMA) per SMSP %.1f  (%s)\n", MODE, wps,
         (double)hc / iters / per / warps_per_smsp, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 2 * 1024);
  cudaMalloc(&cyc, 8 * 148 * 2);
  for (int w : {4, 8, 16}) run<3>(w, out, cyc);
  for (int w : {4, 8, 16}) run<0>(w, out, cyc);
  for (int w : {4, 8, 16}) run<1>(w, out, cyc);
  for (int w : {4, 8, 16}) run<2>(w, out, cyc);
  for (int w : {4, 8, 16}) run<4>(w, out, cyc);
  for (int w : {4, 8, 16}) run<5>(w, out, cyc);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
