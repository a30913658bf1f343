// Batched constrained <-> free maps of the matrix- and simplex-valued parameter types, with their
// Jacobians and Hessians.
//
// Replaces, for M parameters at a time,
//   MatrixParameters.py:101-129  pack_posdef_matrix / unpack_posdef_matrix / pos_def_matrix_free_to_vector
//                                (log-Cholesky: free = vec_ld(log_diag(chol(A - lb I))))
//   MatrixParameters.py:131-134  autograd.jacobian / autograd.hessian of that map
//   SimplexParams.py:11-27       constrain_simplex_matrix / unconstrain_simplex_matrix
//   SimplexParams.py:33-63       constrain_grad_from_moment / constrain_hess_from_moment
// The reference loops over the M parameters in Python (PosDefMatrixParamVector.set_free,
// MatrixParameters.py:236-243; SimplexParam.free_to_vector_jac, SimplexParams.py:105-127) and
// differentiates each k x k matrix with autograd; here every map is closed form:
//   A = L L^T + lb I,  L_ij = f_ij (i > j),  L_ii = exp(f_ii),  packed row-major over the lower
//   triangle (numpy.tril_indices order: index(i, j) = i (i + 1) / 2 + j);  with D_ij = dL_ij/df_ij
//   (= L_ii on the diagonal, 1 below it)
//     dA_ab / df_ij          = D_ij ([a = i] L_bj + [b = i] L_aj)
//     d2A_ab / df_ij df_pq   = D_ij D_pq [j = q] ([a = i][b = p] + [b = i][a = p])
//                              + [ij = pq, i = j] L_ii ([a = i] L_bj + [b = i] L_aj)
//   z = softmax([0, f]):  dz_k/df_a = z_k ([k = a+1] - z_{a+1})
//     d2z_k/df_a df_b = z_k (([k = a+1] - z_{a+1}) ([k = b+1] - z_{b+1}) - z_{a+1} ([a = b] - z_{b+1}))
// All kernels are HBM-bound streaming maps: the value maps run one thread per parameter with the
// k x k factor in registers; the derivative maps of small parameters (k <= 4 / 3, d <= 8 / 6) run one
// thread per parameter with compile-time indices and stage their outputs in shared memory for
// coalesced stores; larger ones run one warp per parameter with lanes over the output elements.
#include "common.cuh"
#include "../../include/lrvb_b200.h"

namespace lrvb {

constexpr int kPdMaxK = 8;        // matrix size limit (registers); v = k (k + 1) / 2 <= 36
constexpr int kPdMaxV = kPdMaxK * (kPdMaxK + 1) / 2;
constexpr int kSxMaxD = 64;       // simplex size limit

// ---- shared-memory tiles: a thread owns one parameter, the CTA moves the data -------------------
// Consecutive parameters are contiguous in memory, so a CTA of T parameters reads T * IN and writes
// T * OUT contiguous doubles.  Threads compute from / into their own row of a tile; the CTA streams
// the tiles in and out with fully coalesced accesses.  Row strides IN | 1 and OUT | 1 doubles keep the
// row-per-thread accesses bank-conflict free.
template <int IN, int OUT>
struct TileGeom {
  static constexpr int kInStride = IN > 0 ? (IN | 1) : 0;
  static constexpr int kStride = OUT | 1;
  static constexpr int kMaxT = (96 * 1024) / (8 * (kInStride + kStride));
  static constexpr int kT = kMaxT >= 128 ? 128 : (kMaxT / 32) * 32;     // threads (= parameters) per CTA
  static constexpr size_t kSmem = sizeof(double) * kT * (kInStride + kStride);
  // tiny parameters (a few doubles in and out) are bound by their exp / log / sqrt, not by the
  // access pattern: they skip the staging (measured: profiles/r01_packing_throughput_v3.log)
  static constexpr bool kTiled = IN == 0 || IN + OUT >= 12 || (IN > OUT && IN + OUT >= 8);
};

template <int IN, int T>
__device__ __forceinline__ void tile_load(double* __restrict__ tile, const double* __restrict__ in, int64_t m0,
                                          int64_t M) {
  const int64_t left = M - m0;
  const int n = (int)((left < T ? left : T) * IN);
  const double* src = in + m0 * IN;
  for (int e = threadIdx.x; e < n; e += T) tile[(e / IN) * (IN | 1) + (e % IN)] = src[e];
  __syncthreads();
}

template <int OUT, int T>
__device__ __forceinline__ void tile_flush(const double* __restrict__ tile, double* __restrict__ out,
                                           int64_t m0, int64_t M) {
  __syncthreads();
  const int64_t left = M - m0;
  const int n = (int)((left < T ? left : T) * OUT);
  double* o = out + m0 * OUT;
  for (int e = threadIdx.x; e < n; e += T) o[e] = tile[(e / OUT) * (OUT | 1) + (e % OUT)];
}

// ---- positive-definite matrices: value maps, one thread per matrix -----------------------------
template <int K>
__device__ __forceinline__ void pd_load_factor(const double* __restrict__ f, double (&L)[K][K]) {
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      if (j > i) L[i][j] = 0.0;
      else {
        const double v = f[i * (i + 1) / 2 + j];
        L[i][j] = (i == j) ? exp(v) : v;
      }
    }
}

// MODE 0: full (k, k) matrix; MODE 1: its lower triangle in packed order (free_to_vector)
template <int K, int MODE>
__global__ void __launch_bounds__(TileGeom<K * (K + 1) / 2, MODE == 0 ? K * K : K * (K + 1) / 2>::kT)
k_pd_unpack(const double* __restrict__ free_v, double* __restrict__ out, int64_t M, double lb) {
  pdl_sync();
  constexpr int V = K * (K + 1) / 2;
  constexpr int OUT = MODE == 0 ? K * K : V;
  using G = TileGeom<V, OUT>;
  extern __shared__ __align__(16) double tile_sm[];
  double* tin = tile_sm;
  double* tout = tile_sm + G::kT * G::kInStride;
  const int64_t m0 = (int64_t)blockIdx.x * G::kT, m = m0 + threadIdx.x;
  if (G::kTiled) tile_load<V, G::kT>(tin, free_v, m0, M);
  if (m < M) {
    double L[K][K];
    pd_load_factor<K>(G::kTiled ? tin + threadIdx.x * G::kInStride : free_v + m * V, L);
    double* row = G::kTiled ? tout + threadIdx.x * G::kStride : out + m * OUT;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
      for (int b = 0; b <= a; ++b) {
        double s = (a == b) ? lb : 0.0;
#pragma unroll
        for (int c = 0; c <= b; ++c) s = fma(L[a][c], L[b][c], s);
        if (MODE == 0) {
          row[a * K + b] = s;
          row[b * K + a] = s;
        } else {
          row[a * (a + 1) / 2 + b] = s;
        }
      }
  }
  if (G::kTiled) tile_flush<OUT, G::kT>(tout, out, m0, M);
}

// log-Cholesky of (A - lb I); reads the lower triangle.  A matrix that is not positive definite
// gives NaN rows and is counted in *bad (numpy.linalg.cholesky raises LinAlgError there).
template <int K>
__global__ void __launch_bounds__(TileGeom<K * K, K * (K + 1) / 2>::kT)
k_pd_pack(const double* __restrict__ mat, double* __restrict__ free_v, int64_t M, double lb,
          int* __restrict__ bad) {
  pdl_sync();
  constexpr int V = K * (K + 1) / 2;
  using G = TileGeom<K * K, V>;
  extern __shared__ __align__(16) double tile_sm[];
  double* tin = tile_sm;
  double* tout = tile_sm + G::kT * G::kInStride;
  const int64_t m0 = (int64_t)blockIdx.x * G::kT, m = m0 + threadIdx.x;
  if (G::kTiled) tile_load<K * K, G::kT>(tin, mat, m0, M);
  if (m < M) {
    double L[K][K];
    const double* A = G::kTiled ? tin + threadIdx.x * G::kInStride : mat + m * K * K;
    bool ok = true;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        double s = A[i * K + j] - ((i == j) ? lb : 0.0);
#pragma unroll
        for (int c = 0; c < j; ++c) s -= L[i][c] * L[j][c];
        if (i == j) {
          if (!(s > 0.0)) ok = false;
          L[i][i] = sqrt(s);
        } else {
          L[i][j] = s / L[j][j];
        }
      }
    if (!ok && bad) atomicAdd(bad, 1);
    double* row = G::kTiled ? tout + threadIdx.x * G::kStride : free_v + m * V;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const double v = (i == j) ? log(L[i][i]) : L[i][j];
        row[i * (i + 1) / 2 + j] = ok ? v : __longlong_as_double(0x7ff8000000000000LL);
      }
  }
  if (G::kTiled) tile_flush<V, G::kT>(tout, free_v, m0, M);
}

// ---- derivative maps: one warp per matrix, lanes over the output elements ---------------------
__device__ __forceinline__ void pd_index(int c, int& i, int& j) {   // packed index -> (row, col)
  i = 0;
  while ((i + 1) * (i + 2) / 2 <= c) ++i;
  j = c - i * (i + 1) / 2;
}

// HESS = false: jac (M, v, v), jac[r][c] = dA_r / df_c;  HESS = true: hess (M, v, v, v)
template <bool HESS>
__global__ void __launch_bounds__(256)
k_pd_derivs(const double* __restrict__ free_v, double* __restrict__ out, int64_t M, int k) {
  pdl_sync();
  __shared__ double Ls[8][kPdMaxK * kPdMaxK];
  __shared__ unsigned char ri[kPdMaxV], ci[kPdMaxV];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int v = k * (k + 1) / 2;
  if (threadIdx.x < v) {
    int i, j;
    pd_index(threadIdx.x, i, j);
    ri[threadIdx.x] = (unsigned char)i;
    ci[threadIdx.x] = (unsigned char)j;
  }
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * 8 + w;
  if (m >= M) return;
  double* L = Ls[w];
  for (int e = lane; e < k * k; e += 32) L[e] = 0.0;
  __syncwarp();
  for (int c = lane; c < v; c += 32) {
    const double f = free_v[m * v + c];
    L[ri[c] * k + ci[c]] = (ri[c] == ci[c]) ? exp(f) : f;
  }
  __syncwarp();
  if (!HESS) {
    double* o = out + m * v * v;
    for (int e = lane; e < v * v; e += 32) {
      const int r = e / v, c = e - r * v;
      const int a = ri[r], b = ci[r], i = ri[c], j = ci[c];
      const double D = (i == j) ? L[i * k + i] : 1.0;
      double s = 0.0;
      if (a == i) s += L[b * k + j];
      if (b == i) s += L[a * k + j];
      o[e] = D * s;
    }
  } else {
    double* o = out + m * v * v * v;
    const int vv = v * v;
    for (int e = lane; e < v * vv; e += 32) {
      const int r = e / vv, rem = e - r * vv, c1 = rem / v, c2 = rem - c1 * v;
      const int a = ri[r], b = ci[r], i = ri[c1], j = ci[c1], p = ri[c2], q = ci[c2];
      double s = 0.0;
      if (j == q) {
        const double D1 = (i == j) ? L[i * k + i] : 1.0, D2 = (p == q) ? L[p * k + p] : 1.0;
        double t = 0.0;
        if (a == i && b == p) t += 1.0;
        if (b == i && a == p) t += 1.0;
        s = D1 * D2 * t;
      }
      if (c1 == c2 && i == j) {
        double t = 0.0;
        if (a == i) t += L[b * k + j];
        if (b == i) t += L[a * k + j];
        s = fma(L[i * k + i], t, s);
      }
      o[e] = s;
    }
  }
}

// ---- simplexes ---------------------------------------------------------------------------------
// z (M, d) = softmax of [0, free (M, d-1)] along the row (SimplexParams.py:11-18), one thread per row
__global__ void __launch_bounds__(128)
k_simplex_constrain(const double* __restrict__ free_v, double* __restrict__ z, int64_t M, int d) {
  pdl_sync();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const double* f = free_v + m * (d - 1);
  double mx = 0.0;
  for (int a = 0; a < d - 1; ++a) mx = fmax(mx, f[a]);
  double s = exp(-mx);
  for (int a = 0; a < d - 1; ++a) s += exp(f[a] - mx);
  const double ln = mx + log(s);       // logsumexp
  z[m * d] = exp(-ln);
  for (int a = 0; a < d - 1; ++a) z[m * d + a + 1] = exp(f[a] - ln);
}

// free (M, d-1) = log z[:, 1:] - log z[:, 0] (SimplexParams.py:21-23)
__global__ void __launch_bounds__(128)
k_simplex_unconstrain(const double* __restrict__ z, double* __restrict__ free_v, int64_t M, int d) {
  pdl_sync();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const double l0 = log(z[m * d]);
  for (int a = 0; a < d - 1; ++a) free_v[m * (d - 1) + a] = log(z[m * d + a + 1]) - l0;
}

// HESS = false: jac (M, d, d-1); HESS = true: hess (M, d, d-1, d-1).  One warp per simplex.
template <bool HESS>
__global__ void __launch_bounds__(256)
k_simplex_derivs(const double* __restrict__ free_v, double* __restrict__ out, int64_t M, int d) {
  pdl_sync();
  __shared__ double zs[8][kSxMaxD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * 8 + w;
  if (m >= M) return;
  const int df = d - 1;
  const double* f = free_v + m * df;
  double mx = 0.0;
  for (int a = lane; a < df; a += 32) mx = fmax(mx, f[a]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = (lane == 0) ? exp(-mx) : 0.0;
  for (int a = lane; a < df; a += 32) s += exp(f[a] - mx);
  s = warp_sum(s);
  const double ln = mx + log(s);
  double* z = zs[w];
  if (lane == 0) z[0] = exp(-ln);
  for (int a = lane; a < df; a += 32) z[a + 1] = exp(f[a] - ln);
  __syncwarp();
  if (!HESS) {
    double* o = out + m * d * df;
    for (int e = lane; e < d * df; e += 32) {
      const int k = e / df, a = e - k * df;
      o[e] = z[k] * (((k == a + 1) ? 1.0 : 0.0) - z[a + 1]);
    }
  } else {
    double* o = out + m * d * df * df;
    for (int e = lane; e < d * df * df; e += 32) {
      const int k = e / (df * df), rem = e - k * df * df, a = rem / df, b = rem - a * df;
      const double ta = ((k == a + 1) ? 1.0 : 0.0) - z[a + 1];
      const double tb = ((k == b + 1) ? 1.0 : 0.0) - z[b + 1];
      o[e] = z[k] * (ta * tb - z[a + 1] * (((a == b) ? 1.0 : 0.0) - z[b + 1]));
    }
  }
}

// ---- small parameters: one thread per parameter, outputs staged in shared memory ---------------
// For small k / d a warp per parameter leaves most lanes idle and spends its time decoding indices.
// Here a thread computes all outputs of ITS parameter with compile-time indices (the structural
// zeros of the log-Cholesky derivatives cost nothing) into a row of a shared-memory tile; the CTA then
// streams the tile out -- consecutive parameters are contiguous in memory, so the stores are
// fully coalesced.  Row stride OUT | 1 doubles keeps the row-per-thread writes bank-conflict free.
constexpr __host__ __device__ int pd_row(int c) {
  int i = 0;
  while ((i + 1) * (i + 2) / 2 <= c) ++i;
  return i;
}
constexpr __host__ __device__ int pd_col(int c) { return c - pd_row(c) * (pd_row(c) + 1) / 2; }

template <int K, bool HESS>
__global__ void __launch_bounds__(TileGeom<0, HESS ? (K * (K + 1) / 2) * (K * (K + 1) / 2) * (K * (K + 1) / 2)
                                                   : (K * (K + 1) / 2) * (K * (K + 1) / 2)>::kT)
k_pd_derivs_small(const double* __restrict__ free_v, double* __restrict__ out, int64_t M) {
  pdl_sync();
  constexpr int V = K * (K + 1) / 2;
  constexpr int OUT = HESS ? V * V * V : V * V;
  using G = TileGeom<0, OUT>;
  extern __shared__ __align__(16) double tile_sm[];
  const int64_t m0 = (int64_t)blockIdx.x * G::kT, m = m0 + threadIdx.x;
  double* row = tile_sm + threadIdx.x * G::kStride;
  if (m < M) {
    double L[K][K];
    pd_load_factor<K>(free_v + m * V, L);
#pragma unroll
    for (int r = 0; r < V; ++r) {
      const int a = pd_row(r), b = pd_col(r);
#pragma unroll
      for (int c1 = 0; c1 < V; ++c1) {
        const int i = pd_row(c1), j = pd_col(c1);
        if (!HESS) {
          double s = 0.0;
          if (a == i) s += L[b][j];
          if (b == i) s += L[a][j];
          row[r * V + c1] = ((i == j) ? L[i][i] : 1.0) * s;
        } else {
#pragma unroll
          for (int c2 = 0; c2 < V; ++c2) {
            const int p = pd_row(c2), q = pd_col(c2);
            double s = 0.0;
            if (j == q) {
              const double t = ((a == i && b == p) ? 1.0 : 0.0) + ((b == i && a == p) ? 1.0 : 0.0);
              if (t != 0.0) s = ((i == j) ? L[i][i] : 1.0) * ((p == q) ? L[p][p] : 1.0) * t;
            }
            if (c1 == c2 && i == j) {
              double t = 0.0;
              if (a == i) t += L[b][j];
              if (b == i) t += L[a][j];
              s = fma(L[i][i], t, s);
            }
            row[(r * V + c1) * V + c2] = s;
          }
        }
      }
    }
  }
  tile_flush<OUT, G::kT>(tile_sm, out, m0, M);
}

template <int D, bool HESS>
__global__ void __launch_bounds__(TileGeom<0, HESS ? D * (D - 1) * (D - 1) : D * (D - 1)>::kT)
k_simplex_derivs_small(const double* __restrict__ free_v, double* __restrict__ out, int64_t M) {
  pdl_sync();
  constexpr int DF = D - 1;
  constexpr int OUT = HESS ? D * DF * DF : D * DF;
  using G = TileGeom<0, OUT>;
  extern __shared__ __align__(16) double tile_sm[];
  const int64_t m0 = (int64_t)blockIdx.x * G::kT, m = m0 + threadIdx.x;
  double* row = tile_sm + threadIdx.x * G::kStride;
  if (m < M) {
    double z[D];
    double mx = 0.0;
#pragma unroll
    for (int a = 0; a < DF; ++a) {
      z[a + 1] = free_v[m * DF + a];
      mx = fmax(mx, z[a + 1]);
    }
    double s = exp(-mx);
#pragma unroll
    for (int a = 0; a < DF; ++a) s += exp(z[a + 1] - mx);
    const double ln = mx + log(s);
    z[0] = exp(-ln);
#pragma unroll
    for (int a = 0; a < DF; ++a) z[a + 1] = exp(z[a + 1] - ln);
#pragma unroll
    for (int k = 0; k < D; ++k)
#pragma unroll
      for (int a = 0; a < DF; ++a) {
        const double ta = ((k == a + 1) ? 1.0 : 0.0) - z[a + 1];
        if (!HESS) {
          row[k * DF + a] = z[k] * ta;
        } else {
#pragma unroll
          for (int b = 0; b < DF; ++b) {
            const double tb = ((k == b + 1) ? 1.0 : 0.0) - z[b + 1];
            row[(k * DF + a) * DF + b] = z[k] * (ta * tb - z[a + 1] * (((a == b) ? 1.0 : 0.0) - z[b + 1]));
          }
        }
      }
  }
  tile_flush<OUT, G::kT>(tile_sm, out, m0, M);
}

// simplex value maps for d <= 8: DIR 0 constrain (d-1 -> d), DIR 1 unconstrain (d -> d-1)
template <int D, int DIR>
__global__ void __launch_bounds__(TileGeom<DIR == 0 ? D - 1 : D, DIR == 0 ? D : D - 1>::kT)
k_simplex_value_small(const double* __restrict__ in, double* __restrict__ out, int64_t M) {
  pdl_sync();
  constexpr int IN = DIR == 0 ? D - 1 : D, OUT = DIR == 0 ? D : D - 1;
  using G = TileGeom<IN, OUT>;
  extern __shared__ __align__(16) double tile_sm[];
  double* tin = tile_sm;
  double* tout = tile_sm + G::kT * G::kInStride;
  const int64_t m0 = (int64_t)blockIdx.x * G::kT, m = m0 + threadIdx.x;
  if (G::kTiled) tile_load<IN, G::kT>(tin, in, m0, M);
  if (m < M) {
    const double* f = G::kTiled ? tin + threadIdx.x * G::kInStride : in + m * IN;
    double* row = G::kTiled ? tout + threadIdx.x * G::kStride : out + m * OUT;
    if (DIR == 0) {
      double mx = 0.0;
#pragma unroll
      for (int a = 0; a < D - 1; ++a) mx = fmax(mx, f[a]);
      double s = exp(-mx);
#pragma unroll
      for (int a = 0; a < D - 1; ++a) s += exp(f[a] - mx);
      const double ln = mx + log(s);
      row[0] = exp(-ln);
#pragma unroll
      for (int a = 0; a < D - 1; ++a) row[a + 1] = exp(f[a] - ln);
    } else {
      const double l0 = log(f[0]);
#pragma unroll
      for (int a = 0; a < D - 1; ++a) row[a] = log(f[a + 1]) - l0;
    }
  }
  if (G::kTiled) tile_flush<OUT, G::kT>(tout, out, m0, M);
}

template <int D, int DIR>
static cudaError_t launch_sx_value_small(const double* in, double* out, int64_t M, cudaStream_t st) {
  using G = TileGeom<DIR == 0 ? D - 1 : D, DIR == 0 ? D : D - 1>;
  return launch_pdl(k_simplex_value_small<D, DIR>, dim3((unsigned)((M + G::kT - 1) / G::kT)), dim3(G::kT),
                    G::kSmem, st, in, out, M);
}

template <int K, bool HESS>
static cudaError_t launch_pd_small(const double* f, double* out, int64_t M, cudaStream_t st) {
  constexpr int V = K * (K + 1) / 2;
  using G = TileGeom<0, HESS ? V * V * V : V * V>;
  const size_t smem = G::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_pd_derivs_small<K, HESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  return launch_pdl(k_pd_derivs_small<K, HESS>, dim3((unsigned)((M + G::kT - 1) / G::kT)), dim3(G::kT), smem, st,
                    f, out, M);
}

template <int D, bool HESS>
static cudaError_t launch_sx_small(const double* f, double* out, int64_t M, cudaStream_t st) {
  using G = TileGeom<0, HESS ? D * (D - 1) * (D - 1) : D * (D - 1)>;
  const size_t smem = G::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_simplex_derivs_small<D, HESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  return launch_pdl(k_simplex_derivs_small<D, HESS>, dim3((unsigned)((M + G::kT - 1) / G::kT)), dim3(G::kT), smem,
                    st, f, out, M);
}

template <int K, int MODE>
static cudaError_t launch_unpack_k(const double* f, double* out, int64_t M, double lb, cudaStream_t st) {
  using G = TileGeom<K * (K + 1) / 2, MODE == 0 ? K * K : K * (K + 1) / 2>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_pd_unpack<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
    configured = true;
  }
  return launch_pdl(k_pd_unpack<K, MODE>, dim3((unsigned)((M + G::kT - 1) / G::kT)), dim3(G::kT), G::kSmem, st,
                    f, out, M, lb);
}

template <int MODE>
static int launch_unpack(const double* f, double* out, int k, int64_t M, double lb, cudaStream_t st) {
#define LRVB_PD(KK)                                                      \
  case KK:                                                               \
    LRVB_CUDA((launch_unpack_k<KK, MODE>(f, out, M, lb, st)));           \
    break;
  switch (k) {
    LRVB_PD(1) LRVB_PD(2) LRVB_PD(3) LRVB_PD(4) LRVB_PD(5) LRVB_PD(6) LRVB_PD(7) LRVB_PD(8)
  }
#undef LRVB_PD
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

template <int K>
static cudaError_t launch_pack_k(const double* mat, double* f, int64_t M, double lb, int* bad, cudaStream_t st) {
  using G = TileGeom<K * K, K * (K + 1) / 2>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(k_pd_pack<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
    configured = true;
  }
  return launch_pdl(k_pd_pack<K>, dim3((unsigned)((M + G::kT - 1) / G::kT)), dim3(G::kT), G::kSmem, st, mat, f, M,
                    lb, bad);
}

}  // namespace lrvb

using namespace lrvb;

#define PD_ARGS(fn)                                                                             \
  LRVB_REQUIRE(k >= 1 && k <= kPdMaxK, fn ": matrix size k = %d outside [1, %d]", k, kPdMaxK);   \
  LRVB_REQUIRE(M >= 0, fn ": M must be non-negative");                                          \
  LRVB_REQUIRE(diag_lb >= 0.0, fn ": diag_lb must be non-negative");                            \
  if (M == 0) return LRVB_OK

extern "C" {

int lrvb_posdef_unpack(const double* free_dev, int32_t k, int64_t M, double diag_lb, double* mat_dev,
                       void* stream) {
  PD_ARGS("lrvb_posdef_unpack");
  LRVB_REQUIRE(free_dev && mat_dev, "lrvb_posdef_unpack: NULL pointer");
  return launch_unpack<0>(free_dev, mat_dev, k, M, diag_lb, (cudaStream_t)stream);
}

int lrvb_posdef_free_to_vector(const double* free_dev, int32_t k, int64_t M, double diag_lb,
                               double* vec_dev, void* stream) {
  PD_ARGS("lrvb_posdef_free_to_vector");
  LRVB_REQUIRE(free_dev && vec_dev, "lrvb_posdef_free_to_vector: NULL pointer");
  return launch_unpack<1>(free_dev, vec_dev, k, M, diag_lb, (cudaStream_t)stream);
}

int lrvb_posdef_pack(const double* mat_dev, int32_t k, int64_t M, double diag_lb, double* free_dev,
                     int32_t* not_posdef_dev, void* stream) {
  PD_ARGS("lrvb_posdef_pack");
  LRVB_REQUIRE(mat_dev && free_dev, "lrvb_posdef_pack: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
#define LRVB_PD(KK)                                                                              \
  case KK:                                                                                       \
    LRVB_CUDA((launch_pack_k<KK>(mat_dev, free_dev, M, diag_lb, (int*)not_posdef_dev, st)));      \
    break;
  switch (k) {
    LRVB_PD(1) LRVB_PD(2) LRVB_PD(3) LRVB_PD(4) LRVB_PD(5) LRVB_PD(6) LRVB_PD(7) LRVB_PD(8)
  }
#undef LRVB_PD
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_posdef_free_to_vector_jac(const double* free_dev, int32_t k, int64_t M, double diag_lb,
                                   double* jac_dev, void* stream) {
  PD_ARGS("lrvb_posdef_free_to_vector_jac");
  LRVB_REQUIRE(free_dev && jac_dev, "lrvb_posdef_free_to_vector_jac: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: LRVB_CUDA((launch_pd_small<1, false>(free_dev, jac_dev, M, st))); break;
    case 2: LRVB_CUDA((launch_pd_small<2, false>(free_dev, jac_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_pd_small<3, false>(free_dev, jac_dev, M, st))); break;
    case 4: LRVB_CUDA((launch_pd_small<4, false>(free_dev, jac_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_pd_derivs<false>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, free_dev,
                           jac_dev, M, (int)k));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_posdef_free_to_vector_hess(const double* free_dev, int32_t k, int64_t M, double diag_lb,
                                    double* hess_dev, void* stream) {
  PD_ARGS("lrvb_posdef_free_to_vector_hess");
  LRVB_REQUIRE(free_dev && hess_dev, "lrvb_posdef_free_to_vector_hess: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (k) {
    case 1: LRVB_CUDA((launch_pd_small<1, true>(free_dev, hess_dev, M, st))); break;
    case 2: LRVB_CUDA((launch_pd_small<2, true>(free_dev, hess_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_pd_small<3, true>(free_dev, hess_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_pd_derivs<true>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, free_dev,
                           hess_dev, M, (int)k));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

#define SX_ARGS(fn)                                                                            \
  LRVB_REQUIRE(d >= 2 && d <= kSxMaxD, fn ": simplex size d = %d outside [2, %d]", d, kSxMaxD); \
  LRVB_REQUIRE(M >= 0, fn ": M must be non-negative");                                         \
  if (M == 0) return LRVB_OK

int lrvb_simplex_constrain(const double* free_dev, int64_t M, int32_t d, double* z_dev, void* stream) {
  SX_ARGS("lrvb_simplex_constrain");
  LRVB_REQUIRE(free_dev && z_dev, "lrvb_simplex_constrain: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: LRVB_CUDA((launch_sx_value_small<2, 0>(free_dev, z_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_sx_value_small<3, 0>(free_dev, z_dev, M, st))); break;
    case 4: LRVB_CUDA((launch_sx_value_small<4, 0>(free_dev, z_dev, M, st))); break;
    case 5: LRVB_CUDA((launch_sx_value_small<5, 0>(free_dev, z_dev, M, st))); break;
    case 6: LRVB_CUDA((launch_sx_value_small<6, 0>(free_dev, z_dev, M, st))); break;
    case 7: LRVB_CUDA((launch_sx_value_small<7, 0>(free_dev, z_dev, M, st))); break;
    case 8: LRVB_CUDA((launch_sx_value_small<8, 0>(free_dev, z_dev, M, st))); break;
    case 9: LRVB_CUDA((launch_sx_value_small<9, 0>(free_dev, z_dev, M, st))); break;
    case 10: LRVB_CUDA((launch_sx_value_small<10, 0>(free_dev, z_dev, M, st))); break;
    case 11: LRVB_CUDA((launch_sx_value_small<11, 0>(free_dev, z_dev, M, st))); break;
    case 12: LRVB_CUDA((launch_sx_value_small<12, 0>(free_dev, z_dev, M, st))); break;
    case 13: LRVB_CUDA((launch_sx_value_small<13, 0>(free_dev, z_dev, M, st))); break;
    case 14: LRVB_CUDA((launch_sx_value_small<14, 0>(free_dev, z_dev, M, st))); break;
    case 15: LRVB_CUDA((launch_sx_value_small<15, 0>(free_dev, z_dev, M, st))); break;
    case 16: LRVB_CUDA((launch_sx_value_small<16, 0>(free_dev, z_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_simplex_constrain, dim3((unsigned)((M + 127) / 128)), dim3(128), 0, st, free_dev,
                           z_dev, M, (int)d));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_simplex_unconstrain(const double* z_dev, int64_t M, int32_t d, double* free_dev, void* stream) {
  SX_ARGS("lrvb_simplex_unconstrain");
  LRVB_REQUIRE(free_dev && z_dev, "lrvb_simplex_unconstrain: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: LRVB_CUDA((launch_sx_value_small<2, 1>(z_dev, free_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_sx_value_small<3, 1>(z_dev, free_dev, M, st))); break;
    case 4: LRVB_CUDA((launch_sx_value_small<4, 1>(z_dev, free_dev, M, st))); break;
    case 5: LRVB_CUDA((launch_sx_value_small<5, 1>(z_dev, free_dev, M, st))); break;
    case 6: LRVB_CUDA((launch_sx_value_small<6, 1>(z_dev, free_dev, M, st))); break;
    case 7: LRVB_CUDA((launch_sx_value_small<7, 1>(z_dev, free_dev, M, st))); break;
    case 8: LRVB_CUDA((launch_sx_value_small<8, 1>(z_dev, free_dev, M, st))); break;
    case 9: LRVB_CUDA((launch_sx_value_small<9, 1>(z_dev, free_dev, M, st))); break;
    case 10: LRVB_CUDA((launch_sx_value_small<10, 1>(z_dev, free_dev, M, st))); break;
    case 11: LRVB_CUDA((launch_sx_value_small<11, 1>(z_dev, free_dev, M, st))); break;
    case 12: LRVB_CUDA((launch_sx_value_small<12, 1>(z_dev, free_dev, M, st))); break;
    case 13: LRVB_CUDA((launch_sx_value_small<13, 1>(z_dev, free_dev, M, st))); break;
    case 14: LRVB_CUDA((launch_sx_value_small<14, 1>(z_dev, free_dev, M, st))); break;
    case 15: LRVB_CUDA((launch_sx_value_small<15, 1>(z_dev, free_dev, M, st))); break;
    case 16: LRVB_CUDA((launch_sx_value_small<16, 1>(z_dev, free_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_simplex_unconstrain, dim3((unsigned)((M + 127) / 128)), dim3(128), 0, st, z_dev,
                           free_dev, M, (int)d));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_simplex_jac(const double* free_dev, int64_t M, int32_t d, double* jac_dev, void* stream) {
  SX_ARGS("lrvb_simplex_jac");
  LRVB_REQUIRE(free_dev && jac_dev, "lrvb_simplex_jac: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: LRVB_CUDA((launch_sx_small<2, false>(free_dev, jac_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_sx_small<3, false>(free_dev, jac_dev, M, st))); break;
    case 4: LRVB_CUDA((launch_sx_small<4, false>(free_dev, jac_dev, M, st))); break;
    case 5: LRVB_CUDA((launch_sx_small<5, false>(free_dev, jac_dev, M, st))); break;
    case 6: LRVB_CUDA((launch_sx_small<6, false>(free_dev, jac_dev, M, st))); break;
    case 7: LRVB_CUDA((launch_sx_small<7, false>(free_dev, jac_dev, M, st))); break;
    case 8: LRVB_CUDA((launch_sx_small<8, false>(free_dev, jac_dev, M, st))); break;
    case 9: LRVB_CUDA((launch_sx_small<9, false>(free_dev, jac_dev, M, st))); break;
    case 10: LRVB_CUDA((launch_sx_small<10, false>(free_dev, jac_dev, M, st))); break;
    case 11: LRVB_CUDA((launch_sx_small<11, false>(free_dev, jac_dev, M, st))); break;
    case 12: LRVB_CUDA((launch_sx_small<12, false>(free_dev, jac_dev, M, st))); break;
    case 13: LRVB_CUDA((launch_sx_small<13, false>(free_dev, jac_dev, M, st))); break;
    case 14: LRVB_CUDA((launch_sx_small<14, false>(free_dev, jac_dev, M, st))); break;
    case 15: LRVB_CUDA((launch_sx_small<15, false>(free_dev, jac_dev, M, st))); break;
    case 16: LRVB_CUDA((launch_sx_small<16, false>(free_dev, jac_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_simplex_derivs<false>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, free_dev,
                           jac_dev, M, (int)d));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_simplex_hess(const double* free_dev, int64_t M, int32_t d, double* hess_dev, void* stream) {
  SX_ARGS("lrvb_simplex_hess");
  LRVB_REQUIRE(free_dev && hess_dev, "lrvb_simplex_hess: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 2: LRVB_CUDA((launch_sx_small<2, true>(free_dev, hess_dev, M, st))); break;
    case 3: LRVB_CUDA((launch_sx_small<3, true>(free_dev, hess_dev, M, st))); break;
    case 4: LRVB_CUDA((launch_sx_small<4, true>(free_dev, hess_dev, M, st))); break;
    case 5: LRVB_CUDA((launch_sx_small<5, true>(free_dev, hess_dev, M, st))); break;
    case 6: LRVB_CUDA((launch_sx_small<6, true>(free_dev, hess_dev, M, st))); break;
    default:
      LRVB_CUDA(launch_pdl(k_simplex_derivs<true>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, free_dev,
                           hess_dev, M, (int)d));
  }
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

}  // extern "C"
