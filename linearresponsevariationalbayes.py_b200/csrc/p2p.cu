// One-shot all-reduce (sum, fp64) of the small replicated blocks of a sharded GLMM over NVLink
// peer memory -- the collective that follows every sharded evaluation / HVP / Schur build
// ([KL, grad_g (Dg), H_gg (Dg*Dg)], 16 KB at K = 20, 87 KB at K = 50).
//
// The messages are far below the size at which link bandwidth matters; what a sharded step pays is
// latency.  So instead of a ring: every rank owns a WINDOW of device memory that its peers map
// (CUDA IPC, one process per GPU), and one kernel per rank, one thread per element,
//   1. PUSHES its value into slot [parity][my rank][element] of every peer's window.  A slot is a
//      16-byte line {lo, epoch, hi, epoch}: each 8-byte half carries its own copy of the 32-bit
//      epoch, so the data needs no fence and no separate flag -- a half is valid exactly when its
//      epoch matches (8-byte stores are single-copy atomic; this is the "LL" protocol of NCCL);
//   2. POLLS the `world` lines of its element in its OWN window until both halves carry the epoch;
//   3. SUMS them in RANK ORDER into the caller's buffer.
// One NVLink traversal end to end, no intermediate hop, no round trip; and because every rank adds
// the same numbers in the same order the result is bitwise identical on all ranks and from run to
// run.  Slots are double-buffered by epoch parity: a rank can only reach epoch e+2 (and overwrite
// parity e) after every peer has pushed e+1, which a peer does only once its kernel of epoch e --
// hence its reads -- has completed (stream order).  The kernel runs on the caller's stream behind
// the producer (programmatic dependent launch like every other kernel of the library).  A thread
// waits only for the same element of the peers, which run on OTHER devices, so nothing needs to
// be co-resident.  A bounded spin (~5 s) turns a missing peer into a launch failure instead of a hung device.
#include <new>
#include "common.cuh"
#include "../../include/lrvb_b200.h"

namespace lrvb {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pThreads = 256;
constexpr long long kP2pSpinCycles = 10000000000LL;  // ~5 s at 1.9 GHz

struct P2pPeers {
  uint4* win[kP2pMaxWorld];                   // window of rank r as mapped in this process
};

__device__ __forceinline__ void st_line(uint4* p, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(flag), "r"(hi),
               "r"(flag)
               : "memory");
}
__device__ __forceinline__ bool ld_line(const uint4* p, unsigned flag, double& v) {
  unsigned lo, f0, hi, f1;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1)
               : "l"(p)
               : "memory");
  v = __hiloint2double((int)hi, (int)lo);
  return f0 == flag && f1 == flag;
}

// window layout: lines [2 parities][world][max_elems]
__global__ void __launch_bounds__(kP2pThreads)
k_p2p_allreduce(double* __restrict__ buf, int64_t n, P2pPeers peers, int rank, int world,
                unsigned long long epoch, int64_t max_elems, int* __restrict__ status) {
  pdl_sync();
  const int64_t e = (int64_t)blockIdx.x * kP2pThreads + threadIdx.x;
  if (e >= n) return;
  const unsigned flag = (unsigned)(epoch & 0xffffffffull);
  const size_t slot0 = (size_t)(epoch & 1ull) * world * (size_t)max_elems + (size_t)e;
  const double mine = buf[e];
  for (int p = 0; p < world; ++p)
    if (p != rank) st_line(peers.win[p] + slot0 + (size_t)rank * max_elems, mine, flag);
  double s = 0.0;
  const uint4* own = peers.win[rank] + slot0;
  for (int r = 0; r < world; ++r) {
    double v = mine;
    if (r != rank) {
      const uint4* line = own + (size_t)r * max_elems;
      long long t0 = 0;
      unsigned polls = 0;
      while (!ld_line(line, flag, v)) {
        if ((++polls & 1023u) == 0) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > kP2pSpinCycles) {
            // rank r never arrived: record it and abort the launch -- every later CUDA call of this
            // process then fails loudly instead of continuing with a partial sum
            atomicExch(status, 1 + r);
            __threadfence_system();
            __trap();
          }
        }
      }
    }
    s += v;
  }
  buf[e] = s;
}

}  // namespace lrvb

using namespace lrvb;

struct lrvb_p2p {
  int rank = 0, world = 1, device = 0;
  int64_t max_elems = 0;
  size_t bytes = 0;
  uint4* window = nullptr;         // ours (cudaMalloc)
  P2pPeers peers;
  bool opened[kP2pMaxWorld];
  bool connected = false;
  unsigned long long epoch = 0;
  int* status = nullptr;           // device word: 0 ok, 1 + r = rank r did not arrive
};

extern "C" {

int lrvb_p2p_create(lrvb_p2p** out, int32_t rank, int32_t world, int64_t max_elems) {
  LRVB_REQUIRE(out != nullptr, "lrvb_p2p_create: out is NULL");
  LRVB_REQUIRE(world >= 1 && world <= kP2pMaxWorld, "lrvb_p2p_create: world = %d outside [1, %d]", world,
               kP2pMaxWorld);
  LRVB_REQUIRE(rank >= 0 && rank < world, "lrvb_p2p_create: rank %d outside [0, %d)", rank, world);
  LRVB_REQUIRE(max_elems >= 1, "lrvb_p2p_create: max_elems must be positive");
  lrvb_p2p* h = new (std::nothrow) lrvb_p2p();
  LRVB_REQUIRE(h != nullptr, "lrvb_p2p_create: out of host memory");
  h->rank = rank;
  h->world = world;
  h->max_elems = max_elems;
  h->bytes = sizeof(uint4) * 2 * (size_t)world * (size_t)h->max_elems;
  for (int r = 0; r < kP2pMaxWorld; ++r) { h->peers.win[r] = nullptr; h->opened[r] = false; }
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->window, h->bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->status, sizeof(int));
  if (e == cudaSuccess) e = cudaMemset(h->window, 0, h->bytes);
  if (e == cudaSuccess) e = cudaMemset(h->status, 0, sizeof(int));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("lrvb_p2p_create: %s", cudaGetErrorString(e));
    if (h->window) cudaFree(h->window);
    if (h->status) cudaFree(h->status);
    delete h;
    return LRVB_ECUDA;
  }
  h->peers.win[rank] = h->window;
  if (world == 1) h->connected = true;
  *out = h;
  return LRVB_OK;
}

int lrvb_p2p_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int lrvb_p2p_export(lrvb_p2p* h, void* handle_out) {
  LRVB_REQUIRE(h != nullptr && handle_out != nullptr, "lrvb_p2p_export: NULL argument");
  cudaIpcMemHandle_t mh;
  LRVB_CUDA(cudaIpcGetMemHandle(&mh, h->window));
  memcpy(handle_out, &mh, sizeof(mh));
  return LRVB_OK;
}

int lrvb_p2p_connect(lrvb_p2p* h, const void* handles) {
  LRVB_REQUIRE(h != nullptr && handles != nullptr, "lrvb_p2p_connect: NULL argument");
  LRVB_REQUIRE(!h->connected, "lrvb_p2p_connect: already connected");
  const char* src = static_cast<const char*>(handles);
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, src + (size_t)r * sizeof(mh), sizeof(mh));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("lrvb_p2p_connect: cannot map the window of rank %d: %s", r, cudaGetErrorString(e));
      return LRVB_ECUDA;
    }
    h->peers.win[r] = static_cast<uint4*>(p);
    h->opened[r] = true;
  }
  h->connected = true;
  return LRVB_OK;
}

int lrvb_p2p_allreduce_sum(lrvb_p2p* h, double* buf_dev, int64_t n, void* stream) {
  LRVB_REQUIRE(h != nullptr, "lrvb_p2p_allreduce_sum: handle is NULL");
  LRVB_REQUIRE(h->connected, "lrvb_p2p_allreduce_sum: peers are not connected");
  LRVB_REQUIRE(n >= 0 && n <= h->max_elems, "lrvb_p2p_allreduce_sum: n = %lld exceeds the window (%lld)",
               (long long)n, (long long)h->max_elems);
  if (n == 0) return LRVB_OK;
  LRVB_REQUIRE(buf_dev != nullptr, "lrvb_p2p_allreduce_sum: buffer is NULL");
  h->epoch += 1;
  if ((h->epoch & 0xffffffffull) == 0) h->epoch += 2;   // epoch 0 is the empty window; keeps the parity sequence
  const int grid = (int)((n + kP2pThreads - 1) / kP2pThreads);
  LRVB_CUDA(launch_pdl(k_p2p_allreduce, dim3(grid), dim3(kP2pThreads), 0, (cudaStream_t)stream, buf_dev, n,
                       h->peers, h->rank, h->world, h->epoch, h->max_elems, h->status));
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_p2p_status(lrvb_p2p* h, int32_t* status_out, void* stream) {
  LRVB_REQUIRE(h != nullptr && status_out != nullptr, "lrvb_p2p_status: NULL argument");
  int s = 0;
  LRVB_CUDA(cudaMemcpyAsync(&s, h->status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  LRVB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *status_out = s;
  return LRVB_OK;
}

int lrvb_p2p_destroy(lrvb_p2p* h) {
  if (!h) return LRVB_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < h->world; ++r)
    if (h->opened[r]) cudaIpcCloseMemHandle(h->peers.win[r]);
  if (h->window) cudaFree(h->window);
  if (h->status) cudaFree(h->status);
  cudaGetLastError();
  delete h;
  return LRVB_OK;
}

}  // extern "C"
