// DMMA probe 4: heterogeneous warp jobs (ni x nj tiles per warp, from argv) + a block barrier every
// `period` k-steps -- the compute phase of k_gram_big without its data movement.
//   dmma_probe4 <period> <ni,nj> x 16
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NI, int NJ>
__device__ __forceinline__ void ksteps(double (&acc)[4][4][2], const double* sm, const double* sw, const int (&offA)[4],
                                       const int (&offB)[4], int ZS, int k0, int k1, int lr) {
  for (int ks = k0; ks < k1; ++ks) {
    const int n = 4 * (ks & 15) + lr;
    const double* zr = sm + n * ZS;
    const double w = sw[n];
    double fa[4], fb[4];
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f < NI) fa[f] = zr[offA[f]];
      if (f < NJ) fb[f] = zr[offB[f]] * w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i < NI && j < NJ) dmma(acc[i][j][0], acc[i][j][1], fa[i], fb[j]);
  }
}
struct Jobs { int ni[16], nj[16]; };
__global__ void __launch_bounds__(512, 1) probe(double* out, long long* cyc, int total, int period, int ZS, Jobs jobs) {
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
  const int ni = jobs.ni[warp], nj = jobs.nj[warp];
  const long long t0 = clock64();
  for (int k0 = 0; k0 < total; k0 += period) {
    const int k1 = k0 + period;
    switch (ni * 4 + nj) {
#define C(a, b) case a * 4 + b: ksteps<a, b>(acc, sm, sw, offA, offB, ZS, k0, k1, lr); break;
      C(1, 1) C(1, 2) C(1, 3) C(1, 4) C(2, 1) C(2, 2) C(2, 3) C(2, 4) C(3, 1) C(3, 2) C(3, 3) C(3, 4) C(4, 1) C(4, 2) C(4, 3) C(4, 4)
      default: break;
    }
    __syncthreads();
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
int main(int argc, char** argv) {
  const int period = atoi(argv[1]);
  Jobs jobs; int load[4] = {0, 0, 0, 0}, tot = 0;
  for (int w = 0; w < 16; ++w) {
    int a = 0, b = 0; if (2 + w < argc) sscanf(argv[2 + w], "%d,%d", &a, &b);
    jobs.ni[w] = a; jobs.nj[w] = b; load[w % 4] += a * b; tot += a * b;
  }
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 512); cudaMalloc(&cyc, 8 * 148);
  const int total = 4096, ZS = 108;
  const size_t smem = sizeof(double) * (64 * ZS + 256);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int r = 0; r < 2; ++r) probe<<<148, 512, smem>>>(out, cyc, total, period, ZS, jobs);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  int mx = 0; for (int q = 0; q < 4; ++q) mx = load[q] > mx ? load[q] : mx;
  printf("period %4d  tiles %3d  SMSP loads %d %d %d %d : %.1f cycles per k-step, pipe efficiency %.1f%% (vs max-loaded SMSP %.1f%%)  %s\n",
         period, tot, load[0], load[1], load[2], load[3], (double)hc / total, 100.0 * tot * 4.0 / ((double)hc / total),
         100.0 * mx * 16.0 / ((double)hc / total), cudaGetErrorString(cudaGetLastError()));
  return 0;
}
