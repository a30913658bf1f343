// Stand-alone check + timing of the small-K packed Gram kernel (csrc/gram_small.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/build/gram_bench tools/gram_bench.cu
//   tools/build/gram_bench [N] [K] [reps] [ctas_per_sm]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../linearresponsevariationalbayes.py_b200/csrc/gram_small.cuh"

namespace lrvb {
void set_error(const char*, ...) {}
long long g_launches = 0;
}
using namespace lrvb;

__global__ void k_fill(double* p, size_t n, unsigned seed, double lo, double hi) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  p[i] = lo + (hi - lo) * ((z >> 11) * (1.0 / 9007199254740992.0));
}

// naive reference: one CTA per (family, p, q)
__global__ void k_ref(const double* X, const double* Wabc, double* out, int64_t N, int64_t ldw, int K) {
  __shared__ double red[32];
  const int fam = blockIdx.x / (K * K), p = (blockIdx.x / K) % K, q = blockIdx.x % K;
  const double* w = Wabc + (int64_t)fam * ldw;
  double s = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += blockDim.x) {
    double xp = X[n * K + p], xq = X[n * K + q];
    if (fam == 2) xp *= xp;
    if (fam >= 1) xq *= xq;
    s += w[n] * xp * xq;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// sum the per-CTA partials and unpack the packed upper triangle into the three families
__global__ void k_unpack(const double* part, double* out, int K, int T2, int NT, int ncta) {
  const int t = blockIdx.x, e = threadIdx.x;   // tile slot, element
  int j = 0;
  while ((j + 1) * (j + 2) / 2 <= t) ++j;
  const int i = t - j * (j + 1) / 2;
  double s = 0.0;
  for (int c = 0; c < ncta; ++c) s += part[((size_t)c * NT + t) * 64 + e];
  const int p = 8 * i + (e >> 3), q = 8 * j + (e & 7);
  if (p > q || q >= 2 * K) return;
  if (q < K) { out[p * K + q] = s; out[q * K + p] = s; }
  else if (p < K) out[K * K + p * K + (q - K)] = s;
  else { out[2 * K * K + (p - K) * K + (q - K)] = s; out[2 * K * K + (q - K) * K + (p - K)] = s; }
}

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 1000000;
  const int K = argc > 2 ? atoi(argv[2]) : 20;
  const int reps = argc > 3 ? atoi(argv[3]) : 20;
  const int cps = argc > 4 ? atoi(argv[4]) : 2;
  const int do_flush = argc > 5 ? atoi(argv[5]) : 1;
  const GramSmallShape sh = gram_small_shape(K);
  const int grid = 148 * cps;
  double *X, *W, *part, *out, *ref, *flush;
  cudaMalloc(&X, sizeof(double) * N * K);
  const int64_t ldw = (N + 7) / 8 * 8;
  cudaMalloc(&W, sizeof(double) * 3 * ldw);
  cudaMalloc(&part, sizeof(double) * (size_t)grid * sh.NT * 64);
  cudaMalloc(&out, sizeof(double) * 3 * K * K);
  cudaMalloc(&ref, sizeof(double) * 3 * K * K);
  const size_t nflush = 32u << 20;
  cudaMalloc(&flush, sizeof(double) * nflush);
  k_fill<<<(unsigned)((N * K + 255) / 256), 256>>>(X, (size_t)N * K, 1, -1.5, 1.5);
  k_fill<<<(unsigned)((3 * ldw + 255) / 256), 256>>>(W, (size_t)3 * ldw, 2, -1.0, 0.5);
  cudaMemset(out, 0, sizeof(double) * 3 * K * K);

  if (!launch_gram_small(X, W, part, N, ldw, K, grid, 0)) { printf("K=%d unsupported\n", K); return 2; }
  k_unpack<<<sh.NT, 64>>>(part, out, K, sh.T2, sh.NT, grid);
  k_ref<<<3 * K * K, 256>>>(X, W, ref, N, ldw, K);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<double> ho(3 * K * K), hr(3 * K * K);
  cudaMemcpy(ho.data(), out, sizeof(double) * 3 * K * K, cudaMemcpyDeviceToHost);
  cudaMemcpy(hr.data(), ref, sizeof(double) * 3 * K * K, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < 3 * K * K; ++i) {
    maxerr = std::max(maxerr, fabs(ho[i] - hr[i]));
    maxref = std::max(maxref, fabs(hr[i]));
  }
  printf("N=%lld K=%d T2=%d T0=%d M=%d: max abs err %.3e (max |ref| %.3e) rel %.3e %s\n", (long long)N, K,
         sh.T2, sh.T0, sh.has_m, maxerr, maxref, maxerr / maxref, maxerr / maxref < 1e-11 ? "OK" : "FAIL");

  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> ms(reps);
  for (int r = 0; r < reps + 3; ++r) {
    if (do_flush) k_fill<<<(unsigned)((nflush + 255) / 256), 256>>>(flush, nflush, r, 0, 1);   // L2 flush
    cudaEventRecord(e0);
    launch_gram_small(X, W, part, N, ldw, K, grid, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    if (r >= 3) cudaEventElapsedTime(&ms[r - 3], e0, e1);
  }
  std::sort(ms.begin(), ms.end());
  const double flops = (double)N * (4.0 * K * K + 2.0 * K);
  const double exec = (double)N / 4 * (sh.NT + (sh.has_m ? 1 : 0)) * 512.0;
  printf("  time min %.1f us  median %.1f us : algorithmic %.2f TFLOP/s (%.3f of 37.1), executed %.2f TFLOP/s, X+W %.0f GB/s\n",
         ms[0] * 1e3, ms[reps / 2] * 1e3, flops / ms[reps / 2] / 1e9, flops / ms[reps / 2] / 1e9 / 37.1,
         exec / ms[reps / 2] / 1e9, (double)N * (8.0 * K + 24) / ms[reps / 2] / 1e6);
  e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return (e != cudaSuccess) || !(maxerr / maxref < 1e-11);
}
