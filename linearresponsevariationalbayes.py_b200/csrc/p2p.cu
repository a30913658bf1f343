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
// be co-resident.  A bounded spin (default 120 s, LRVB_P2P_TIMEOUT_S or lrvb_p2p_set_timeout) turns a
// missing peer into a STATUS: the waiting thread records 1 + r in a word of mapped pinned HOST memory,
// poisons its output with NaN and returns -- the context survives (no trap), the host reads the word
// without touching the device and raises at the next call.
// Wait accounting (lrvb_p2p_set_stats): every thread takes %globaltimer before and after its poll
// loops; the per-call maximum over all elements goes (one atomicMax per CTA, the last CTA of the call
// folds it) into running totals that bench.py reports as collective_wait_us -- the time a rank spent
// waiting for its SLOWEST peer, i.e. inter-rank skew plus one NVLink traversal.
#include <new>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "../../include/lrvb_b200.h"

namespace lrvb {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pThreads = 256;
constexpr double kP2pDefaultTimeoutS = 120.0;

// stats block (device memory): [0] max wait (ns) of the call in flight, [1] CTAs of the call that are done,
// [2] calls, [3] sum over calls of the per-call max wait (ns), [4] largest per-call max wait (ns)
constexpr int kP2pStatWords = 8;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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
                unsigned long long epoch, int64_t max_elems, volatile int* __restrict__ status,
                unsigned long long timeout_ns, unsigned long long* __restrict__ stats,
                volatile int* __restrict__ dead_dev) {
  pdl_sync();
  __shared__ unsigned long long wmax[kP2pThreads / 32];
  const int64_t e = (int64_t)blockIdx.x * kP2pThreads + threadIdx.x;
  unsigned long long waited = 0;
  if (e < n) {
    const unsigned flag = (unsigned)(epoch & 0xffffffffull);
    const size_t slot0 = (size_t)(epoch & 1ull) * world * (size_t)max_elems + (size_t)e;
    const double mine = buf[e];
    for (int p = 0; p < world; ++p)
      if (p != rank) st_line(peers.win[p] + slot0 + (size_t)rank * max_elems, mine, flag);
    double s = 0.0;
    const uint4* own = peers.win[rank] + slot0;
    const unsigned long long t_begin = stats ? globaltimer_ns() : 0ull;
    bool dead = (*dead_dev != 0);        // a peer already timed out in an earlier call: do not wait again
    for (int r = 0; r < world; ++r) {
      double v = mine;
      if (r != rank) {
        const uint4* line = own + (size_t)r * max_elems;
        unsigned long long t0 = 0;
        unsigned polls = 0;
        while (!dead && !ld_line(line, flag, v)) {
          if ((++polls & 1023u) == 0) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > timeout_ns) {
              // rank r never arrived: record it where the host can read it without the device
              // (mapped pinned memory), poison the result and leave -- no trap, the context survives
              *status = 1 + r;
              *dead_dev = 1 + r;
              __threadfence_system();
              dead = true;
            }
          }
        }
        if (dead) v = __longlong_as_double(0x7ff8000000000000LL);
      }
      s += v;
    }
    if (stats) waited = globaltimer_ns() - t_begin;
    buf[e] = s;
  }
  if (stats) {
    // per-call maximum of the wait over all elements: warp max -> CTA max -> one atomicMax per CTA; the
    // last CTA of the call folds the call's maximum into the running totals
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, waited, o);
      waited = other > waited ? other : waited;
    }
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = waited;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long m = 0;
      for (int i = 0; i < kP2pThreads / 32; ++i) m = wmax[i] > m ? wmax[i] : m;
      atomicMax(&stats[0], m);
      __threadfence();
      if (atomicAdd(&stats[1], 1ull) == (unsigned long long)gridDim.x - 1) {
        __threadfence();
        const unsigned long long call_max = atomicExch(&stats[0], 0ull);
        stats[1] = 0;
        stats[2] += 1;
        stats[3] += call_max;
        if (call_max > stats[4]) stats[4] = call_max;
      }
    }
  }
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
  int* status_host = nullptr;      // mapped pinned host word: 0 ok, 1 + r = rank r did not arrive (sticky)
  int* status = nullptr;           // its device alias
  unsigned long long timeout_ns = 0;
  unsigned long long* stats = nullptr;   // device, kP2pStatWords (+ one word: device copy of the status)
  int stats_on = 0;
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
  // the status word lives in mapped pinned HOST memory: it stays readable whatever happens to the
  // stream, and the host polls it without a device round trip
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&h->status_host, sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *h->status_host = 0;
    e = cudaHostGetDevicePointer((void**)&h->status, h->status_host, 0);
  }
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->stats, sizeof(unsigned long long) * (kP2pStatWords + 1));
  if (e == cudaSuccess) e = cudaMemset(h->window, 0, h->bytes);
  if (e == cudaSuccess) e = cudaMemset(h->stats, 0, sizeof(unsigned long long) * (kP2pStatWords + 1));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("lrvb_p2p_create: %s", cudaGetErrorString(e));
    if (h->window) cudaFree(h->window);
    if (h->status_host) cudaFreeHost(h->status_host);
    if (h->stats) cudaFree(h->stats);
    delete h;
    return LRVB_ECUDA;
  }
  double tmo = kP2pDefaultTimeoutS;
  if (const char* env = getenv("LRVB_P2P_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0.0) tmo = v;
  }
  h->timeout_ns = (unsigned long long)(tmo * 1e9);
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
  {
    const int st = *(volatile int*)h->status_host;
    if (st != 0) {
      set_error("lrvb_p2p_allreduce_sum: rank %d did not arrive within %.1f s in an earlier call; results since "
                "then are NaN and this communicator is unusable", st - 1, (double)h->timeout_ns * 1e-9);
      return LRVB_ESTATE;
    }
  }
  h->epoch += 1;
  if ((h->epoch & 0xffffffffull) == 0) h->epoch += 2;   // epoch 0 is the empty window; keeps the parity sequence
  const int grid = (int)((n + kP2pThreads - 1) / kP2pThreads);
  LRVB_CUDA(launch_pdl(k_p2p_allreduce, dim3(grid), dim3(kP2pThreads), 0, (cudaStream_t)stream, buf_dev, n,
                       h->peers, h->rank, h->world, h->epoch, h->max_elems, (volatile int*)h->status, h->timeout_ns,
                       h->stats_on ? h->stats : (unsigned long long*)nullptr,
                       (volatile int*)(h->stats + kP2pStatWords)));
  LRVB_CHECK_LAUNCH();
  return LRVB_OK;
}

int lrvb_p2p_status(lrvb_p2p* h, int32_t* status_out, void* stream) {
  LRVB_REQUIRE(h != nullptr && status_out != nullptr, "lrvb_p2p_status: NULL argument");
  // wait for the calls in flight; a failed stream does not matter, the word is host memory
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) cudaGetLastError();
  *status_out = *(volatile int*)h->status_host;
  return LRVB_OK;
}

int lrvb_p2p_status_nowait(lrvb_p2p* h, int32_t* status_out) {
  LRVB_REQUIRE(h != nullptr && status_out != nullptr, "lrvb_p2p_status_nowait: NULL argument");
  *status_out = *(volatile int*)h->status_host;
  return LRVB_OK;
}

int lrvb_p2p_set_timeout(lrvb_p2p* h, double seconds) {
  LRVB_REQUIRE(h != nullptr, "lrvb_p2p_set_timeout: handle is NULL");
  LRVB_REQUIRE(seconds > 0.0 && seconds < 1e7, "lrvb_p2p_set_timeout: seconds must be in (0, 1e7)");
  h->timeout_ns = (unsigned long long)(seconds * 1e9);
  return LRVB_OK;
}

int lrvb_p2p_set_stats(lrvb_p2p* h, int32_t enable, void* stream) {
  LRVB_REQUIRE(h != nullptr, "lrvb_p2p_set_stats: handle is NULL");
  h->stats_on = enable ? 1 : 0;
  LRVB_CUDA(cudaMemsetAsync(h->stats, 0, sizeof(unsigned long long) * kP2pStatWords, (cudaStream_t)stream));
  return LRVB_OK;
}

int lrvb_p2p_get_stats(lrvb_p2p* h, double* out3_host, void* stream) {
  LRVB_REQUIRE(h != nullptr && out3_host != nullptr, "lrvb_p2p_get_stats: NULL argument");
  unsigned long long w[kP2pStatWords];
  LRVB_CUDA(cudaMemcpyAsync(w, h->stats, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  LRVB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  out3_host[0] = (double)w[2];
  out3_host[1] = (double)w[3] * 1e-3;
  out3_host[2] = (double)w[4] * 1e-3;
  return LRVB_OK;
}

int lrvb_p2p_destroy(lrvb_p2p* h) {
  if (!h) return LRVB_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < h->world; ++r)
    if (h->opened[r]) cudaIpcCloseMemHandle(h->peers.win[r]);
  if (h->window) cudaFree(h->window);
  if (h->status_host) cudaFreeHost(h->status_host);
  if (h->stats) cudaFree(h->stats);
  cudaGetLastError();
  delete h;
  return LRVB_OK;
}

}  // extern "C"
