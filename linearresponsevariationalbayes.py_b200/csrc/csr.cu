// Device-side CSR export of the arrowhead Hessian.
//
// Replaces get_sparse_sub_hessian / get_sparse_sub_matrix (SparseObjectives.py:591-619) followed
// by scipy's COO->CSR canonicalisation: entries that are EXACTLY 0.0 are dropped, each
// coordinate appears once, column indices are sorted, indices / indptr are int32.
// Row r of H in the reference layout [globals | u.mean (G) | u.info (G)]:
//   r <  Dg      : A[r,:] , B[g,0,r] for g = 0..G-1 , B[g,1,r] for g = 0..G-1
//   r = Dg+g     : B[g,0,:] , L[g].mm , L[g].mi
//   r = Dg+G+g   : B[g,1,:] , L[g].mi , L[g].ii
// so the column order is already sorted and every kernel below only has to compact in order.
#include "common.cuh"

namespace lrvb {

__device__ __forceinline__ double global_row_elem(const double* __restrict__ A,
                                                  const double* __restrict__ B, int r, int64_t j,
                                                  int Dg, int G) {
  if (j < Dg) return A[(size_t)r * Dg + j];
  j -= Dg;
  if (j < G) return B[(size_t)j * 2 * Dg + r];
  j -= G;
  return B[(size_t)j * 2 * Dg + Dg + r];
}

// block-wide exclusive scan of a 0/1 flag, in thread order; returns offset, total in *total
__device__ __forceinline__ int block_excl_scan_flag(bool f, int* wsum, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, f);
  const int inw = __popc(m & ((1u << lane) - 1));
  __syncthreads();
  if (lane == 0) wsum[wid] = __popc(m);
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < nw; ++w) {
    const int c = wsum[w];
    if (w < wid) off += c;
    tot += c;
  }
  *total = tot;
  return off + inw;
}

// FILL = false: count -> rowcnt[r].  FILL = true: write indices/data at indptr[r].
template <bool FILL>
__global__ void __launch_bounds__(256)
k_csr_global_rows(const double* __restrict__ A, const double* __restrict__ B, int Dg, int G,
                  int32_t* __restrict__ rowcnt, const int32_t* __restrict__ indptr,
                  int32_t* __restrict__ indices, double* __restrict__ data) {
  __shared__ int wsum[8];
  const int r = blockIdx.x;
  const int64_t len = Dg + 2 * (int64_t)G;
  int64_t base = FILL ? indptr[r] : 0;
  int count = 0;
  for (int64_t j0 = 0; j0 < len; j0 += blockDim.x) {
    const int64_t j = j0 + threadIdx.x;
    const double v = (j < len) ? global_row_elem(A, B, r, j, Dg, G) : 0.0;
    const bool nz = (v != 0.0);
    int tot;
    const int off = block_excl_scan_flag(nz, wsum, &tot);
    if (FILL && nz) {
      indices[base + off] = (int32_t)j;
      data[base + off] = v;
    }
    base += tot;
    count += tot;
  }
  if (!FILL && threadIdx.x == 0) rowcnt[r] = count;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_csr_local_rows(const double* __restrict__ B, const double* __restrict__ L, int Dg, int G,
                 int32_t* __restrict__ rowcnt, const int32_t* __restrict__ indptr,
                 int32_t* __restrict__ indices, double* __restrict__ data) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wg >= 2 * (int64_t)G) return;
  const int which = (wg >= G) ? 1 : 0;
  const int gi = (int)(wg - (which ? G : 0));
  const int64_t row = Dg + wg;
  const double* b = B + (size_t)gi * 2 * Dg + (which ? Dg : 0);
  int64_t base = FILL ? indptr[row] : 0;
  int count = 0;
  for (int c0 = 0; c0 < Dg; c0 += 32) {
    const int c = c0 + lane;
    const double v = (c < Dg) ? b[c] : 0.0;
    const bool nz = (v != 0.0);
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if (FILL && nz) {
      const int off = __popc(m & ((1u << lane) - 1));
      indices[base + off] = c;
      data[base + off] = v;
    }
    base += __popc(m);
    count += __popc(m);
  }
  if (lane == 0) {
    const double l0 = L[(size_t)gi * 3], l1 = L[(size_t)gi * 3 + 1], l2 = L[(size_t)gi * 3 + 2];
    const double va = which ? l1 : l0;   // column u.mean_g
    const double vb = which ? l2 : l1;   // column u.info_g
    if (va != 0.0) {
      if (FILL) { indices[base] = Dg + gi; data[base] = va; }
      ++base; ++count;
    }
    if (vb != 0.0) {
      if (FILL) { indices[base] = Dg + G + gi; data[base] = vb; }
      ++base; ++count;
    }
    if (!FILL) rowcnt[row] = count;
  }
}

// ---- exclusive scan of rowcnt (n entries) into indptr (n+1 entries), 3 phases --------------------
constexpr int kScanChunk = 2048;  // per CTA (256 threads x 8)

__global__ void __launch_bounds__(256)
k_scan_sum(const int32_t* __restrict__ cnt, int64_t n, int64_t* __restrict__ blk) {
  __shared__ double red[32];
  const int64_t b0 = (int64_t)blockIdx.x * kScanChunk;
  long long s = 0;
  for (int i = threadIdx.x; i < kScanChunk; i += blockDim.x)
    if (b0 + i < n) s += cnt[b0 + i];
  // counts < 2^31 each and at most 2048 per chunk: exact in double
  const double t = block_sum((double)s, red);
  if (threadIdx.x == 0) blk[blockIdx.x] = (int64_t)t;
}

__global__ void k_scan_top(int64_t* __restrict__ blk, int nblk, int64_t* __restrict__ total) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t run = 0;
  for (int i = 0; i < nblk; ++i) {
    const int64_t c = blk[i];
    blk[i] = run;
    run += c;
  }
  *total = run;
}

__global__ void __launch_bounds__(256)
k_scan_apply(const int32_t* __restrict__ cnt, int64_t n, const int64_t* __restrict__ blk,
             const int64_t* __restrict__ total, int32_t* __restrict__ indptr) {
  __shared__ int wsum[8];
  __shared__ int carry_s;
  const int64_t b0 = (int64_t)blockIdx.x * kScanChunk;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t carry = blk[blockIdx.x];
  for (int s0 = 0; s0 < kScanChunk; s0 += 256) {
    const int64_t i = b0 + s0 + threadIdx.x;
    const int c = (i < n) ? cnt[i] : 0;
    int v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    __syncthreads();
    if (lane == 31) wsum[wid] = v;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (w < wid) off += wsum[w];
      tot += wsum[w];
    }
    if (i < n) indptr[i] = (int32_t)(carry + off + v - c);
    carry += tot;
  }
  (void)carry_s;
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) indptr[n] = (int32_t)(*total);
}

}  // namespace lrvb

using namespace lrvb;

extern "C" {

int lrvb_glmm_hessian_csr_nnz(lrvb_glmm* h, int64_t* nnz, void* stream) {
  LRVB_REQUIRE(h != nullptr && nnz != nullptr, "lrvb_glmm_hessian_csr_nnz: NULL argument");
  if (!h->hess_valid) {
    set_error("lrvb_glmm_hessian_csr_nnz: no Hessian cached (call lrvb_glmm_eval with order 2)");
    return LRVB_ESTATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Dg = h->Dg, G = h->G;
  const int64_t D = h->D;
  const int nblk = cdiv(D, kScanChunk);
  if (!h->rowcnt) {
    LRVB_CUDA(cudaMalloc((void**)&h->rowcnt, sizeof(int32_t) * (size_t)(D + 1)));
    LRVB_CUDA(cudaMalloc((void**)&h->scanblk, sizeof(int64_t) * ((size_t)nblk + 2)));
  }
  k_csr_global_rows<false><<<Dg, 256, 0, st>>>(h->A, h->B, Dg, G, h->rowcnt, nullptr, nullptr, nullptr);
  LRVB_CHECK_LAUNCH();
  if (G > 0) {
    k_csr_local_rows<false><<<cdiv(2 * (int64_t)G, 8), 256, 0, st>>>(h->B, h->L, Dg, G, h->rowcnt,
                                                                     nullptr, nullptr, nullptr);
    LRVB_CHECK_LAUNCH();
  }
  int64_t* blk = (int64_t*)h->scanblk;
  k_scan_sum<<<nblk, 256, 0, st>>>(h->rowcnt, D, blk);
  LRVB_CHECK_LAUNCH();
  k_scan_top<<<1, 32, 0, st>>>(blk, nblk, blk + nblk);
  LRVB_CHECK_LAUNCH();
  int64_t total = 0;
  LRVB_CUDA(cudaMemcpyAsync(&total, blk + nblk, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  LRVB_CUDA(cudaStreamSynchronize(st));
  LRVB_REQUIRE(total < (int64_t)2147483647, "Hessian has %lld nonzeros: exceeds int32 CSR indices",
               (long long)total);
  h->csr_nnz = total;
  *nnz = total;
  return LRVB_OK;
}

int lrvb_glmm_hessian_csr_fill(lrvb_glmm* h, int32_t* indptr_dev, int32_t* indices_dev,
                               double* data_dev, void* stream) {
  LRVB_REQUIRE(h != nullptr && indptr_dev && indices_dev && data_dev,
               "lrvb_glmm_hessian_csr_fill: NULL argument");
  if (!h->hess_valid || h->csr_nnz < 0) {
    set_error("lrvb_glmm_hessian_csr_fill: call lrvb_glmm_hessian_csr_nnz first");
    return LRVB_ESTATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Dg = h->Dg, G = h->G;
  const int64_t D = h->D;
  const int nblk = cdiv(D, kScanChunk);
  int64_t* blk = (int64_t*)h->scanblk;
  k_scan_apply<<<nblk, 256, 0, st>>>(h->rowcnt, D, blk, blk + nblk, indptr_dev);
  LRVB_CHECK_LAUNCH();
  k_csr_global_rows<true><<<Dg, 256, 0, st>>>(h->A, h->B, Dg, G, nullptr, indptr_dev, indices_dev,
                                              data_dev);
  LRVB_CHECK_LAUNCH();
  if (G > 0) {
    k_csr_local_rows<true><<<cdiv(2 * (int64_t)G, 8), 256, 0, st>>>(h->B, h->L, Dg, G, nullptr,
                                                                    indptr_dev, indices_dev, data_dev);
    LRVB_CHECK_LAUNCH();
  }
  return LRVB_OK;
}

}  // extern "C"
