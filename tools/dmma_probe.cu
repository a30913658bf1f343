// DMMA.8x8x4 issue / latency probe on sm_100a: throughput vs warps per SM sub-partition and
// independent accumulators per warp, with distinct operand registers; and DMUL interleave.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/dmma_probe tools/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP, int NMUL>
__global__ void probe(double* out, long long* cyc, int iters, double s) {
  double c[ILP][2], a[ILP], b[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    c[i][0] = i; c[i][1] = -i;
    a[i] = 1e-3 * threadIdx.x + i; b[i] = 1.0 - 1e-3 * i + s;
  }
  double m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NMUL; ++i) m[i & 7] *= s;
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a[i], b[i]);
  }
  const long long t1 = clock64();
  double r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int NMUL>
void run(int warps_per_sm, double* out, long long* cyc) {
  const int iters = 2000;
  const int ctas = warps_per_sm > 16 ? 2 : 1;       // > 16 warps: two CTAs per SM
  const int block = 32 * warps_per_sm / ctas;
  probe<ILP, NMUL><<<148 * ctas, block>>>(out, cyc, iters, 1.0000001);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<ILP, NMUL><<<148 * ctas, block>>>(out, cyc, iters, 1.0000001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  const double n_dmma = 148.0 * warps_per_sm * ILP * iters;
  const double per_smsp = (double)hc / (warps_per_sm / 4.0 > 1 ? warps_per_sm / 4.0 : 1) / ILP / iters;
  printf("ILP=%2d NMUL=%d warps/SM=%2d: %.3f ms  %.2f TF  cycles/iter/warp %.1f  cycles per DMMA per SMSP %.2f\n",
         ILP, NMUL, warps_per_sm, ms, n_dmma * 512 / ms / 1e9, (double)hc / iters, per_smsp);
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 2 * 1024);
  cudaMalloc(&cyc, 8 * 148 * 2);
  for (int w : {4, 8, 16, 32}) run<1, 0>(w, out, cyc);
  for (int w : {4, 8, 16, 32}) run<2, 0>(w, out, cyc);
  for (int w : {4, 8, 16, 32}) run<4, 0>(w, out, cyc);
  for (int w : {4, 8, 16, 32}) run<8, 0>(w, out, cyc);
  for (int w : {4, 8, 16, 32}) run<16, 0>(w, out, cyc);
  for (int w : {4, 8, 16}) run<16, 4>(w, out, cyc);
  for (int w : {4, 8, 16}) run<16, 12>(w, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
