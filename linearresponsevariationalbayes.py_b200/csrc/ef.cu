// Batched exponential-family entropies / expectations and the elementwise Gauss-Hermite
// logistic term.  fp64, one thread per factor, coalesced along the factor axis.
//
// Restates ExponentialFamilies.py:5-120 and Modeling.py:35-52 (aggregate_all=False) for M
// independent factors per launch (the reference evaluates them with numpy ufuncs + scipy
// digamma / gammaln).
#include "common.cuh"

namespace lrvb {

#define LRVB_GRID_STRIDE(i, n)                                                       \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);          \
       i += (int64_t)gridDim.x * blockDim.x)

constexpr double kLog2Pi = 1.8378770664093454836;
constexpr double kLogPi = 1.1447298858494001741;
constexpr double kLog2 = 0.69314718055994530942;

// :33-35 gamma_entropy and :111-112 get_e_log_gamma in ONE pass over (shape, rate): they share digamma(a) and
// log b.  Either output may be NULL.  16 B in, 8 - 16 B out per factor.
__global__ void __launch_bounds__(256)
k_gamma_terms(const double* __restrict__ shape, const double* __restrict__ rate, int64_t M,
              double* __restrict__ entropy, double* __restrict__ e_log) {
  LRVB_GRID_STRIDE(i, M) {
    const double a = shape[i], b = rate[i];
    const PsiLg f = digamma_lgamma_pos(a, entropy != nullptr);
    const double lb = log(b);
    if (entropy) entropy[i] = a - lb + f.lg + (1.0 - a) * f.psi;
    if (e_log) e_log[i] = f.psi - lb;
  }
}
// :23-25 (per factor)
__global__ void k_uvn_entropy(const double* __restrict__ info, int64_t M, double* __restrict__ out) {
  LRVB_GRID_STRIDE(i, M) out[i] = 0.5 * (-log(info[i]) + 1.0 + kLog2Pi);
}
// :43-52 dirichlet_entropy and :118-120 get_e_log_dirichlet in ONE pass; alpha (d, M): simplex dimension is
// axis 0.  entropy (M,) and / or e_log (d, M); either may be NULL.
__global__ void __launch_bounds__(256)
k_dirichlet_terms(const double* __restrict__ alpha, int d, int64_t M, double* __restrict__ entropy,
                  double* __restrict__ e_log) {
  LRVB_GRID_STRIDE(i, M) {
    double sum_alpha = 0.0, sum_lg = 0.0, sum_ad = 0.0;
    for (int j = 0; j < d; ++j) {
      const double a = alpha[(int64_t)j * M + i];
      const PsiLg f = digamma_lgamma_pos(a, entropy != nullptr);
      sum_alpha += a;
      sum_lg += f.lg;
      sum_ad += (a - 1.0) * f.psi;
      if (e_log) e_log[(int64_t)j * M + i] = f.psi;       // minus digamma(sum) below
    }
    const PsiLg fs = digamma_lgamma_pos(sum_alpha, entropy != nullptr);
    if (entropy) entropy[i] = (sum_lg - fs.lg) - ((double)d - sum_alpha) * fs.psi - sum_ad;
    if (e_log)
      for (int j = 0; j < d; ++j) e_log[(int64_t)j * M + i] -= fs.psi;
  }
}
// :54-69 per row of tau (M,2)
__global__ void k_beta_entropy(const double* __restrict__ tau, int64_t M, double* __restrict__ out) {
  LRVB_GRID_STRIDE(i, M) {
    const double2 t = reinterpret_cast<const double2*>(tau)[i];
    const double s = t.x + t.y;
    const PsiLg fx = digamma_lgamma_pos(t.x, true), fy = digamma_lgamma_pos(t.y, true), fs = digamma_lgamma_pos(s, true);
    out[i] = (fx.lg + fy.lg - fs.lg) - (t.x - 1.0) * fx.psi - (t.y - 1.0) * fy.psi + (s - 2.0) * fs.psi;
  }
}
// :20-21, p (M, d)
__global__ void k_multinoulli_entropy(const double* __restrict__ p, int d, int64_t M,
                                      double min_prob, double* __restrict__ out) {
  LRVB_GRID_STRIDE(i, M) {
    double s = 0.0;
    for (int j = 0; j < d; ++j) {
      const double v = p[i * d + j];
      s += v * log(v + min_prob);
    }
    out[i] = -s;
  }
}

// :5-13
__device__ __forceinline__ double mv_digamma(double x, int k) {
  double s = 0.0;
  for (int j = 0; j < k; ++j) s += digamma_lgamma_pos(x - 0.5 * j, false).psi;
  return s;
}
// multivariate digamma and log-gamma of the same argument, sharing every factor's evaluation
__device__ __forceinline__ void mv_digamma_gammaln(double x, int k, double& dg, double& lg) {
  dg = 0.0;
  lg = 0.0;
  for (int j = 0; j < k; ++j) {
    const PsiLg f = digamma_lgamma_pos(x - 0.5 * j, true);
    dg += f.psi;
    lg += f.lg;
  }
  lg += 0.25 * kLogPi * k * (k - 1.0);
}

// :72-82 wishart_entropy, :88-94 e_log_det_wishart, :97-102 e_log_inv_wishart_diag, batched:
// per factor an in-register Cholesky of v (k <= 8) gives log det and diag(v^-1).
// status: set to 1 if any v is not positive definite (the reference asserts sign > 0).
__global__ void k_wishart(const double* __restrict__ df, const double* __restrict__ v, int k,
                          int64_t M, double* __restrict__ entropy, double* __restrict__ e_log_det,
                          double* __restrict__ e_log_inv_diag, int* __restrict__ status) {
  LRVB_GRID_STRIDE(i, M) {
    double Lc[8][8];
    const double* vi = v + i * k * k;
    bool ok = true;
    double logdet = 0.0;
    for (int c = 0; c < k; ++c) {
      double d = vi[c * k + c];
      for (int j = 0; j < c; ++j) d -= Lc[c][j] * Lc[c][j];
      if (!(d > 0.0)) { ok = false; d = 1.0; }
      const double l = sqrt(d);
      Lc[c][c] = l;
      logdet += 2.0 * log(l);
      for (int r = c + 1; r < k; ++r) {
        double s = vi[r * k + c];
        for (int j = 0; j < c; ++j) s -= Lc[r][j] * Lc[c][j];
        Lc[r][c] = s / l;
      }
    }
    if (!ok) atomicExch(status, 1);
    const double n = df[i], kk = (double)k;
    double mvd, mvl;
    mv_digamma_gammaln(0.5 * n, k, mvd, mvl);
    if (entropy) {
      entropy[i] = 0.5 * (kk + 1.0) * logdet + 0.5 * kk * (kk + 1.0) * kLog2 +
                   mvl - 0.5 * (n - kk - 1.0) * mvd + 0.5 * n * kk;
    }
    if (e_log_det) e_log_det[i] = mvd + kk * kLog2 + logdet;
    if (e_log_inv_diag) {
      // diag(v^-1)[c] = sum_r (L^-1[r][c])^2 ; column c of L^-1 by forward substitution
      const double dg = digamma_lgamma_pos(0.5 * (n - kk + 1.0), false).psi;
      for (int c = 0; c < k; ++c) {
        double col[8];
        double acc = 0.0;
        for (int r = c; r < k; ++r) {
          double s = (r == c) ? 1.0 : 0.0;
          for (int j = c; j < r; ++j) s -= Lc[r][j] * col[j];
          col[r] = s / Lc[r][r];
          acc += col[r] * col[r];
        }
        e_log_inv_diag[i * k + c] = log(acc) - dg - kLog2;
      }
    }
  }
}

// Modeling.py:35-52 with aggregate_all=False
__global__ void k_gh_logistic(const double* __restrict__ z_mean, const double* __restrict__ z_sd,
                              int64_t M, const double* __restrict__ gh, int Q,
                              double* __restrict__ out) {
  __shared__ double c[kMaxQ], w[kMaxQ];
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    c[q] = gh[q];
    w[q] = gh[Q + q];
  }
  __syncthreads();
  LRVB_GRID_STRIDE(i, M) {
    const double zm = z_mean[i], zs = z_sd[i];
    double s = 0.0;
    for (int q = 0; q < Q; ++q) {
      const double t = fma(zs, c[q], zm);
      s = fma(w[q], fmax(t, 0.0) + log1p(exp(-fabs(t))), s);
    }
    out[i] = s;
  }
}

// deterministic two-stage sum
__global__ void __launch_bounds__(256)
k_sum_stage(const double* __restrict__ x, int64_t n, double* __restrict__ part) {
  __shared__ double red[32];
  double s = 0.0;
  LRVB_GRID_STRIDE(i, n) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

static int ew_grid(int64_t M) {
  int64_t g = (M + 255) / 256;
  if (g > 16 * kNumSMs) g = 16 * kNumSMs;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace lrvb

using namespace lrvb;

#define EF_ARGS_OK(cond, name) LRVB_REQUIRE(cond, name ": bad argument")

extern "C" {

int lrvb_ef_gamma_entropy(const double* shape, const double* rate, int64_t M, double* out,
                          void* stream) {
  EF_ARGS_OK(M >= 0 && (M == 0 || (shape && rate && out)), "lrvb_ef_gamma_entropy");
  if (M == 0) return LRVB_OK;
  k_gamma_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(shape, rate, M, out, nullptr);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_e_log_gamma(const double* shape, const double* rate, int64_t M, double* out,
                        void* stream) {
  EF_ARGS_OK(M >= 0 && (M == 0 || (shape && rate && out)), "lrvb_ef_e_log_gamma");
  if (M == 0) return LRVB_OK;
  k_gamma_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(shape, rate, M, nullptr, out);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_gamma_terms(const double* shape, const double* rate, int64_t M, double* entropy, double* e_log,
                        void* stream) {
  EF_ARGS_OK(M >= 0 && (M == 0 || (shape && rate && (entropy || e_log))), "lrvb_ef_gamma_terms");
  if (M == 0) return LRVB_OK;
  k_gamma_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(shape, rate, M, entropy, e_log);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_dirichlet_terms(const double* alpha, int32_t d, int64_t M, double* entropy, double* e_log,
                            void* stream) {
  EF_ARGS_OK(d >= 1 && M >= 0 && (M == 0 || (alpha && (entropy || e_log))), "lrvb_ef_dirichlet_terms");
  if (M == 0) return LRVB_OK;
  k_dirichlet_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(alpha, d, M, entropy, e_log);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_uvn_entropy(const double* info, int64_t M, double* out, void* stream) {
  EF_ARGS_OK(M >= 0 && (M == 0 || (info && out)), "lrvb_ef_uvn_entropy");
  if (M == 0) return LRVB_OK;
  k_uvn_entropy<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(info, M, out);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_dirichlet_entropy(const double* alpha, int32_t d, int64_t M, double* out,
                              void* stream) {
  EF_ARGS_OK(d >= 1 && M >= 0 && (M == 0 || (alpha && out)), "lrvb_ef_dirichlet_entropy");
  if (M == 0) return LRVB_OK;
  k_dirichlet_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(alpha, d, M, out, nullptr);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_e_log_dirichlet(const double* alpha, int32_t d, int64_t M, double* out,
                            void* stream) {
  EF_ARGS_OK(d >= 1 && M >= 0 && (M == 0 || (alpha && out)), "lrvb_ef_e_log_dirichlet");
  if (M == 0) return LRVB_OK;
  k_dirichlet_terms<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(alpha, d, M, nullptr, out);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_beta_entropy(const double* tau, int64_t M, double* out, void* stream) {
  EF_ARGS_OK(M >= 0 && (M == 0 || (tau && out)), "lrvb_ef_beta_entropy");
  LRVB_REQUIRE((((uintptr_t)tau) & 15) == 0, "lrvb_ef_beta_entropy: tau not 16-byte aligned");
  if (M == 0) return LRVB_OK;
  k_beta_entropy<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(tau, M, out);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_multinoulli_entropy(const double* p, int32_t d, int64_t M, double min_prob,
                                double* out, void* stream) {
  EF_ARGS_OK(d >= 1 && M >= 0 && (M == 0 || (p && out)), "lrvb_ef_multinoulli_entropy");
  if (M == 0) return LRVB_OK;
  k_multinoulli_entropy<<<ew_grid(M), 256, 0, (cudaStream_t)stream>>>(p, d, M, min_prob, out);
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_ef_wishart(const double* df, const double* v, int32_t k, int64_t M, double* entropy,
                    double* e_log_det, double* e_log_inv_diag, void* stream) {
  EF_ARGS_OK(k >= 1 && k <= 8 && M >= 0 && (M == 0 || (df && v)), "lrvb_ef_wishart");
  if (M == 0) return LRVB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int* dstat = nullptr;
  LRVB_CUDA(cudaMalloc((void**)&dstat, sizeof(int)));
  cudaMemsetAsync(dstat, 0, sizeof(int), st);
  k_wishart<<<ew_grid(M), 128, 0, st>>>(df, v, k, M, entropy, e_log_det, e_log_inv_diag, dstat);
  int hstat = 0;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hstat, dstat, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(dstat);
  if (e != cudaSuccess) {
    set_error("lrvb_ef_wishart failed: %s", cudaGetErrorString(e));
    return LRVB_ECUDA;
  }
  LRVB_REQUIRE(hstat == 0, "lrvb_ef_wishart: a scale matrix v is not positive definite");
  return LRVB_OK;
}

int lrvb_gh_logistic_term(const double* z_mean, const double* z_sd, int64_t M,
                          const double* gh_x_host, const double* gh_w_host, int32_t Q,
                          double* out, void* stream) {
  EF_ARGS_OK(M >= 0 && Q >= 1 && Q <= kMaxQ && gh_x_host && gh_w_host &&
                 (M == 0 || (z_mean && z_sd && out)),
             "lrvb_gh_logistic_term");
  if (M == 0) return LRVB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  double hb[2 * kMaxQ];
  for (int q = 0; q < Q; ++q) {
    hb[q] = sqrt(2.0) * gh_x_host[q];
    hb[Q + q] = gh_w_host[q] / sqrt(M_PI);
  }
  double* dgh = nullptr;
  LRVB_CUDA(cudaMalloc((void**)&dgh, sizeof(double) * 2 * Q));
  cudaError_t e = cudaMemcpyAsync(dgh, hb, sizeof(double) * 2 * Q, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    k_gh_logistic<<<ew_grid(M), 256, 0, st>>>(z_mean, z_sd, M, dgh, Q, out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // hb / dgh lifetime
  cudaFree(dgh);
  if (e != cudaSuccess) {
    set_error("lrvb_gh_logistic_term failed: %s", cudaGetErrorString(e));
    return LRVB_ECUDA;
  }
  return LRVB_OK;
}

int lrvb_sum(const double* x_dev, int64_t M, double* out_dev, void* stream) {
  EF_ARGS_OK(M >= 0 && out_dev && (M == 0 || x_dev), "lrvb_sum");
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) {
    LRVB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double), st));
    return LRVB_OK;
  }
  int grid = (int)((M + 2047) / 2048);
  if (grid > 1024) grid = 1024;
  double* part = nullptr;
  LRVB_CUDA(cudaMallocAsync((void**)&part, sizeof(double) * grid, st));
  k_sum_stage<<<grid, 256, 0, st>>>(x_dev, M, part);
  k_sum_stage<<<1, 256, 0, st>>>(part, grid, out_dev);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(part, st);
  if (e != cudaSuccess) {
    set_error("lrvb_sum failed: %s", cudaGetErrorString(e));
    return LRVB_ECUDA;
  }
  return LRVB_OK;
}

}  // extern "C"
