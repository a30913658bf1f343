// Order-2 evaluation in ONE pass over X (K <= 62): quadrature, per-group sums and the packed weighted Gram
// on the same 32-row stage of X in shared memory, by HOMOGENEOUS teams of P warps.  sm_100a.
//
// Run back to back, the observation pass (obs_fused.cuh) keeps the FP64 pipe ~50 % busy -- its exp / log1p
// chains are latency-bound and its 161 registers allow 12 warps per SM -- and the Gram kernel (gram_mid.cuh)
// ~80 %; both read X from HBM.  A first one-pass kernel with SPECIALISED warps (one quadrature warp feeding
// P DMMA warps through mbarriers; profiles/r02_onepass_specialised.md) was exactly as slow as the two kernels
// together: the quadrature warps were the serial bottleneck, the DMMA warps waited.  Here every warp does
// both: a team of P warps owns a contiguous multiple-of-32 range of rows and walks it stage by stage
// (32 rows, TMA into a 2-slot ring, issued by the team's warp 0):
//   Q  lane = observation, every warp: z_mean / z_sd of its row (replicated), then ITS share of the
//      Gauss-Hermite nodes (node q belongs to warp q mod P); the P partial node sums go through shared
//      memory (double-buffered by stage parity), ONE team barrier, and every warp adds them in warp
//      order and forms the row's weights l_m, l_v, a, b, c in its own copy `wq` (warp 0 also stores W
//      and accumulates the KL partial);
//   C  lane = column: warp c of the team keeps the running per-group sums of column chunk c (32 columns;
//      P = 1: the one warp keeps all chunks) -- groups inside the team's range are written directly,
//      head / tail pieces of straddling groups go to bval and k_obs_fixup adds them in row order;
//   D  warp r multiplies ITS range of tiles of the packed [x|s] upper triangle (gram_mid's column-major
//      ranges of equal DMMA count) over the 8 k-steps of the stage, accumulators in registers.
// The slot of stage s is refilled (stage s + 2) once every warp of the team has passed the barrier of
// stage s + 1.  Warps of different teams are in different phases, so every SM sub-partition always holds
// DMMA work for the pipe while some warp waits on a quadrature chain.  X is read from HBM once; the
// weights never leave the SM on their way to the tensor pipe.  No atomics, fixed summation orders.
#pragma once
#include "common.cuh"
#include "gram_small.cuh"
#include "gram_mid.cuh"
#include "obs_fused.cuh"

namespace lrvb {

constexpr int kTeRows = 32;
constexpr int kTeSlots = 2;
#ifndef LRVB_TEAM_MAXWARPS
#define LRVB_TEAM_MAXWARPS 12
#endif
constexpr int kTeMaxWarps = LRVB_TEAM_MAXWARPS;      // 12 warps x 168 registers: no spills (see TeamQC)
#ifndef LRVB_TEAM_UNROLL
#define LRVB_TEAM_UNROLL 4
#endif

// warps per team by tile-grid size: the accumulator tiles of a warp (2 registers each) must fit a
// 128-register budget next to the quadrature temporaries
__host__ __device__ constexpr int team_P(int T2) { return T2 <= 6 ? 1 : T2 <= 8 ? 2 : T2 <= 13 ? 4 : 8; }
__host__ __device__ inline int team_slot_elems(int K) { return kTeRows * K + 2 * kTeRows + kTeRows / 2; }
// shared memory for `warps` warps (a multiple of P)
inline size_t team_smem(int K, int Q, int T2, int warps) {
  const int P = team_P(T2), teams = warps / P;
  const size_t ring = sizeof(double) * (size_t)teams * kTeSlots * team_slot_elems(K);
  const size_t perwarp = sizeof(double) * (size_t)warps * (6 * kTeRows + (P > 1 ? 2 * 6 * kTeRows : 0));
  const size_t red = sizeof(double) * (size_t)(T2 * (T2 + 1) / 2) * 64;
  const size_t body = ring + perwarp > red ? ring + perwarp : red;
  return body + sizeof(double) * (2 * (size_t)K + 2 * Q + (size_t)teams * 2 * K + teams) +
         sizeof(unsigned long long) * (size_t)teams * kTeSlots;
}
// the largest warp count (multiple of P, <= kTeMaxWarps) whose shared memory fits
inline int team_max_warps(int K, int Q, int T2) {
  const int P = team_P(T2);
  int w = kTeMaxWarps / P * P;
  while (w > P && team_smem(K, Q, T2, w) > 225 * 1024) w -= P;
  return w;
}

struct FusedArgs {
  const double* X; const double* y; const int32_t* g; const double* w; const double* vec; const double* gh;
  const int32_t* gptr; double* W; int64_t ldw; double* klpart; double* gradpart; double* gsc; double* BR;
  double* bval; double* grampart; int64_t N; int K, G, Q; int64_t rows_per_team;
};

__device__ __forceinline__ void team_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Per-warp state of the quadrature / group-sum phases (scalar-replaced into registers: everything is
// inlined).  Nothing here may live in local memory: with ~215 KB of the SM's 256 KB SRAM carved out as
// shared memory the L1 that would back spills or an ABI call's register saves is a few KB for 384 threads,
// and a first version that called the Q / C phases as a non-inlined function (accumulators saved around
// the call once per stage) ran 1.5 - 2.4x SLOWER than the two separate kernels for exactly that reason
// (profiles/r02_onepass_attempts.md).  Hence 12 warps x 168 registers rather than 16 x 128.
template <int NCHW>
struct TeamQC {
  const double *bm, *bv, *ghc, *ghw;
  double *ring, *wq, *part, *gred_row, *kred_slot;
  unsigned ring_u, full_u;
  int64_t gw, rs, re;
  int nst, R, team, skew, chunk0, cur_g;
  bool c_warp;
  double klacc;
  double gm[NCHW], gv[NCHW], q0[NCHW], q1[NCHW], q2[NCHW], q3[NCHW], q4[NCHW], q5[NCHW];
};

template <int P, int NCHW>
__device__ __forceinline__ void team_issue(const FusedArgs& a, const TeamQC<NCHW>& c, int st, int slot) {
  const int lane = threadIdx.x & 31;
  const int K = a.K;
  const int64_t n0 = c.rs + (int64_t)st * kTeRows;
  const unsigned nops = a.w ? 4u : 3u;
  if (st < c.nst && n0 + kTeRows <= a.N && lane < (int)nops) {
    const unsigned xbytes = (unsigned)(kTeRows * K * sizeof(double));
    const unsigned vbytes = (unsigned)(kTeRows * sizeof(double));
    const unsigned gbytes = (unsigned)(kTeRows * sizeof(int32_t));
    const unsigned bar = c.full_u + 8 * slot;
    const unsigned dst = c.ring_u + (unsigned)(slot * team_slot_elems(K) * sizeof(double));
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, xbytes + (a.w ? 2 : 1) * vbytes + gbytes);
      bulk_g2s(dst, a.X + n0 * K, xbytes, bar);
    } else if (lane == 1) {
      bulk_g2s(dst + xbytes, a.y + n0, vbytes, bar);
    } else if (lane == 2) {
      bulk_g2s(dst + xbytes + 2 * vbytes, a.g + n0, gbytes, bar);
    } else {
      bulk_g2s(dst + xbytes + vbytes, a.w + n0, vbytes, bar);
    }
  }
}

template <int NCHW>
__device__ __forceinline__ void team_flush(const FusedArgs& a, TeamQC<NCHW>& c) {
  if (c.cur_g < 0) return;
  const int lane = threadIdx.x & 31;
  const int K = a.K;
  const int nb = 5 + 4 * K;
  const int64_t gb = a.gptr[c.cur_g], ge = a.gptr[c.cur_g + 1];
  double* dbr;
  double* dsc;
  if (gb >= c.rs && ge <= c.re) {          // the whole group lies inside the team's range
    dbr = a.BR + (size_t)c.cur_g * 4 * K;
    dsc = a.gsc + (size_t)c.cur_g * 5;
  } else {                                 // head (starts before the range) or tail piece
    double* rec = a.bval + ((size_t)c.gw * 2 + (gb < c.rs ? 0 : 1)) * nb;
    dsc = rec;
    dbr = rec + 5;
  }
#pragma unroll
  for (int i = 0; i < NCHW; ++i) {
    const int k = lane + 32 * (c.chunk0 + i);
    if (k < K) {
      dbr[k] = c.q2[i];
      dbr[K + k] = c.q3[i];
      dbr[2 * K + k] = c.q4[i];
      dbr[3 * K + k] = c.q5[i];
      c.gm[i] += c.q0[i];
      c.gv[i] += c.q1[i];
    } else if (k == K) {                   // ones column: sum l_m, l_v, a, b, c
      dsc[0] = c.q0[i];
      dsc[1] = c.q1[i];
      dsc[2] = c.q2[i];
      dsc[3] = c.q3[i];
      dsc[4] = c.q5[i];
    }
    c.q0[i] = c.q1[i] = c.q2[i] = c.q3[i] = c.q4[i] = c.q5[i] = 0.0;
  }
}

// Phases Q and C of stage `st` (slot `slot`, full-barrier parity `phase`) for one warp.
template <int P, int NCHW>
__device__ __forceinline__ void team_qc_stage(const FusedArgs& a, TeamQC<NCHW>& c, int st, int slot, unsigned phase) {
  constexpr int UNR = LRVB_TEAM_UNROLL;
  const int lane = threadIdx.x & 31;
  const int K = a.K, G = a.G, Q = a.Q;
  const int64_t N = a.N, ldw = a.ldw;
  const double* __restrict__ vec = a.vec;
  const double* __restrict__ w = a.w;
  const double* __restrict__ bm = c.bm;
  const double* __restrict__ bv = c.bv;
  const double* __restrict__ ghc = c.ghc;
  const double* __restrict__ ghw = c.ghw;
  double* __restrict__ W = a.W;
  const int R = c.R;
  const int slot_elems = team_slot_elems(K);
  const int64_t um0 = 4 + 2 * (int64_t)K, ui0 = um0 + G;
  const int64_t n0 = c.rs + (int64_t)st * kTeRows;
  double* xs = c.ring + (size_t)slot * slot_elems;
  double* ys = xs + kTeRows * K;
  const int32_t* gs = reinterpret_cast<const int32_t*>(ys + 2 * kTeRows);
  double* wq = c.wq;
  const int rows = (int)((c.re - n0 < kTeRows) ? (c.re - n0) : kTeRows);
  if (R == 0 && n0 + kTeRows > N) {
    // ragged last stage of the data set: filled by warp 0 itself (zero rows beyond N); every warp of
    // the team has passed the barrier of stage st - 1, so the slot's previous use (st - 2) is over
    const int vr = (int)(N - n0);
    for (int e = lane; e < kTeRows * K; e += 32) xs[e] = (e < vr * K) ? a.X[n0 * K + e] : 0.0;
    ys[lane] = (lane < vr) ? a.y[n0 + lane] : 0.0;
    ys[kTeRows + lane] = (lane < vr && w) ? w[n0 + lane] : 0.0;
    reinterpret_cast<int32_t*>(ys + 2 * kTeRows)[lane] = (lane < vr) ? a.g[n0 + lane] : -1;
    __syncwarp();
    if (lane == 0) mbar_arrive(c.full_u + 8 * slot);
  }
  mbar_wait(c.full_u + 8 * slot, phase);

  // ---- Q: lane = observation ----
  unsigned segmask = 0;
  {
    const int64_t n = n0 + lane;
    const bool valid = lane < rows;
    const int gi = valid ? gs[lane] : 0;
    if (c.c_warp) {
      const int gprev = __shfl_up_sync(0xffffffffu, gi, 1);
      segmask = __ballot_sync(0xffffffffu, valid && (lane == 0 ? gi != c.cur_g : gi != gprev));
    }
    double zm = vec[um0 + gi];
    double zv = 1.0 / vec[ui0 + gi];
    const double* xr = xs + (size_t)lane * K;
    const int skew = c.skew;
    for (int k = skew; k < K; ++k) {
      const double x = xr[k];
      zm = fma(x, bm[k], zm);
      zv = fma(x * x, bv[k], zv);
    }
    for (int k = 0; k < skew; ++k) {
      const double x = xr[k];
      zm = fma(x, bm[k], zm);
      zv = fma(x * x, bv[k], zv);
    }
    const double zs = sqrt(zv);
    GHSumsF s = {0, 0, 0, 0, 0, 0};
    GHSumsF s2 = {0, 0, 0, 0, 0, 0};
    int q = R;
    for (; q + (UNR - 1) * P < Q; q += UNR * P) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const double cq = ghc[q + u * P];
        gh_node_f<2>(fma(zs, cq, zm), cq, ghw[q + u * P], (u & 1) ? s2 : s);
      }
    }
    for (; q + P < Q; q += 2 * P) {      // two nodes in flight for the remainder (and when Q / P < UNR)
      const double c0 = ghc[q], c1 = ghc[q + P];
      gh_node_f<2>(fma(zs, c0, zm), c0, ghw[q], s);
      gh_node_f<2>(fma(zs, c1, zm), c1, ghw[q + P], s2);
    }
    for (; q < Q; q += P) {
      const double c0 = ghc[q];
      gh_node_f<2>(fma(zs, c0, zm), c0, ghw[q], s);
    }
    s.A += s2.A; s.Am += s2.Am; s.As += s2.As; s.Amm += s2.Amm; s.Ams += s2.Ams; s.Ass += s2.Ass;
    if (P > 1) {
      // exchange the partial node sums: per warp [parity][6][32], summed in warp order
      constexpr int PW = 3 * 6 * kTeRows;          // per-warp block: wq | partials (parity 0) | partials (parity 1)
      double* mp = c.part + (size_t)(st & 1) * 6 * kTeRows;
      mp[lane] = s.A; mp[32 + lane] = s.Am; mp[64 + lane] = s.As;
      mp[96 + lane] = s.Amm; mp[128 + lane] = s.Ams; mp[160 + lane] = s.Ass;
      team_bar(1 + c.team, 32 * P);
      // every warp of the team is past stage st - 1: its slot is free for stage st + 1
      if (R == 0 && st >= 1) team_issue<P, NCHW>(a, c, st + 1, slot ^ 1);
      const double* tp = mp - (size_t)R * PW;      // warp 0's block of this parity
      s.A = s.Am = s.As = s.Amm = s.Ams = s.Ass = 0.0;
#pragma unroll
      for (int r2 = 0; r2 < P; ++r2) {
        const double* op = tp + (size_t)r2 * PW;
        s.A += op[lane]; s.Am += op[32 + lane]; s.As += op[64 + lane];
        s.Amm += op[96 + lane]; s.Ams += op[128 + lane]; s.Ass += op[160 + lane];
      }
    }
    const double wn = valid ? (w ? ys[kTeRows + lane] : 1.0) : 0.0;
    const double yn = ys[lane];
    const double h = 0.5 / zs;
    const double lm = wn * (yn - s.Am);
    const double lv = -wn * s.As * h;
    const double wa = -wn * s.Amm;
    const double wb = -wn * s.Ams * h;
    const double wc = -wn * (s.Ass - s.As / zs) / (4.0 * zv);     // l_vv = -(A_ss / (4 z_v) - A_s / (4 z_s^3))
    double2* wrow = reinterpret_cast<double2*>(wq + 6 * lane);
    wrow[0] = make_double2(lm, lv);
    wrow[1] = make_double2(wa, wb);
    wrow[2] = make_double2(wc, 0.0);
    if (R == 0) {
      c.klacc += wn * (yn * zm - s.A);
      if (valid && W) {
        W[n] = lm;
        W[ldw + n] = lv;
        W[2 * ldw + n] = wa;
        W[3 * ldw + n] = wb;
        W[4 * ldw + n] = wc;
      }
    }
  }
  __syncwarp();

  // ---- C: lane = column; per-group running sums of this warp's chunk(s) ----
  if (c.c_warp) {
    const double2* w2 = reinterpret_cast<const double2*>(wq);
    int koff[NCHW];
    bool isone[NCHW];
    double q0[NCHW], q1[NCHW], q2[NCHW], q3[NCHW], q4[NCHW], q5[NCHW];
#pragma unroll
    for (int i = 0; i < NCHW; ++i) {
      const int k = lane + 32 * (c.chunk0 + i);
      koff[i] = (k < K) ? k : 0;
      isone[i] = (k == K);
      q0[i] = c.q0[i]; q1[i] = c.q1[i]; q2[i] = c.q2[i]; q3[i] = c.q3[i]; q4[i] = c.q4[i]; q5[i] = c.q5[i];
    }
    auto rows_acc = [&](int r, auto nrow) {
      constexpr int NR = decltype(nrow)::value;
      double2 wl[NR], wab[NR], wc2[NR];
      double x[NR][NCHW];
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        wl[u] = w2[3 * (r + u)];
        wab[u] = w2[3 * (r + u) + 1];
        wc2[u] = w2[3 * (r + u) + 2];
#pragma unroll
        for (int i = 0; i < NCHW; ++i) {
          const double v = xs[(size_t)(r + u) * K + koff[i]];
          x[u][i] = isone[i] ? 1.0 : v;
        }
      }
#pragma unroll
      for (int i = 0; i < NCHW; ++i) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0;
#pragma unroll
        for (int u = 0; u < NR; ++u) {
          const double xv = x[u][i], xx = xv * xv;
          t0 = fma(wl[u].x, xv, t0);
          t1 = fma(wl[u].y, xx, t1);
          t2 = fma(wab[u].x, xv, t2);
          t3 = fma(wab[u].y, xv, t3);
          t4 = fma(wab[u].y, xx, t4);
          t5 = fma(wc2[u].x, xx, t5);
        }
        q0[i] += t0; q1[i] += t1; q2[i] += t2; q3[i] += t3; q4[i] += t4; q5[i] += t5;
      }
    };
    int r = 0;
    const unsigned m = segmask;
    while (r < rows) {
      if ((m >> r) & 1u) {            // row r opens a new group
#pragma unroll
        for (int i = 0; i < NCHW; ++i) {
          c.q0[i] = q0[i]; c.q1[i] = q1[i]; c.q2[i] = q2[i]; c.q3[i] = q3[i]; c.q4[i] = q4[i]; c.q5[i] = q5[i];
        }
        team_flush<NCHW>(a, c);
#pragma unroll
        for (int i = 0; i < NCHW; ++i) q0[i] = q1[i] = q2[i] = q3[i] = q4[i] = q5[i] = 0.0;
        c.cur_g = gs[r];
      }
      const unsigned rest = (r + 1 < 32) ? (m >> (r + 1)) : 0u;
      const int nxt = rest ? (r + 1 + __ffs((int)rest) - 1) : rows;
      const int r1 = nxt < rows ? nxt : rows;
      for (; r + 4 <= r1; r += 4) rows_acc(r, std::integral_constant<int, 4>());
      for (; r < r1; ++r) rows_acc(r, std::integral_constant<int, 1>());
    }
#pragma unroll
    for (int i = 0; i < NCHW; ++i) {
      c.q0[i] = q0[i]; c.q1[i] = q1[i]; c.q2[i] = q2[i]; c.q3[i] = q3[i]; c.q4[i] = q4[i]; c.q5[i] = q5[i];
    }
  }
}

// last flush + the warp's KL / global-gradient partials
template <int NCHW>
__device__ __forceinline__ void team_qc_finish(const FusedArgs& a, TeamQC<NCHW>& c) {
  const int lane = threadIdx.x & 31;
  const int K = a.K;
  if (c.c_warp) team_flush<NCHW>(a, c);
  if (c.R == 0) {
    const double kl = warp_sum(c.klacc);
    if (lane == 0) *c.kred_slot = kl;
  }
  if (c.c_warp) {
#pragma unroll
    for (int i = 0; i < NCHW; ++i) {
      const int k = lane + 32 * (c.chunk0 + i);
      if (k < K) {
        c.gred_row[k] = c.gm[i];
        c.gred_row[K + k] = c.gv[i];
      }
    }
  }
}

// One warp of a team: role R of P, tile range [TLO, THI).  NCHW = column chunks this warp may own.
template <int T2, int T0, bool HAS_M, int P, int R, int TLO, int THI, int NCHW>
__device__ __forceinline__ void team_run(const FusedArgs& a, double* ring, unsigned ring_u, unsigned full_u,
                                         double* wq, double* part, const double* bm, const double* bv,
                                         const double* ghc, const double* ghw, double* gred_row, double* kred_slot,
                                         int64_t gw, int team, double* red, int warp, int nwarps) {
  constexpr int TS = HAS_M ? T0 : -1;
  constexpr int TB = HAS_M ? T0 + 1 : T0;
  constexpr int JLO = gram_mid_col(TLO), JHI = gram_mid_col(THI - 1) + 1;
  constexpr int NTL = THI - TLO;
  constexpr int KSTEPS = kTeRows / 4;
  auto mine = [](int i, int j) constexpr { return j * (j + 1) / 2 + i >= TLO && j * (j + 1) / 2 + i < THI; };
  const int lane = threadIdx.x & 31;
  const int lr = lane & 3, lc = lane >> 2;
  const int K = a.K;
  const int slot_elems = team_slot_elems(K);
  const unsigned wq_u = smem_u32(wq);

  TeamQC<NCHW> c;
  c.bm = bm; c.bv = bv; c.ghc = ghc; c.ghw = ghw;
  c.ring = ring; c.wq = wq; c.part = part; c.gred_row = gred_row; c.kred_slot = kred_slot;
  c.ring_u = ring_u; c.full_u = full_u;
  c.gw = gw;
  c.rs = gw * a.rows_per_team;
  c.re = (c.rs + a.rows_per_team < a.N) ? c.rs + a.rows_per_team : a.N;
  c.nst = (c.rs < c.re) ? (int)((c.re - c.rs + kTeRows - 1) / kTeRows) : 0;
  c.R = R; c.team = team;
  {
    int gcd16 = 1;
    while (gcd16 < 16 && (K % (gcd16 * 2)) == 0) gcd16 *= 2;
    int skew = ((lane & 15) * gcd16) >> 4;      // bank-conflict skew of the row-per-lane reads
    c.skew = skew >= K ? 0 : skew;
  }
  const int nch_total = (K + 1 + 31) / 32;
  c.c_warp = (P == 1) || (R < nch_total);
  c.chunk0 = (P == 1) ? 0 : R;
  c.cur_g = -1;
  c.klacc = 0.0;
#pragma unroll
  for (int i = 0; i < NCHW; ++i) c.gm[i] = c.gv[i] = c.q0[i] = c.q1[i] = c.q2[i] = c.q3[i] = c.q4[i] = c.q5[i] = 0.0;
  const int nst = c.nst;

  // ---- D operands: per-lane offsets of the packed columns inside a staged row ----
  const int base = lr * K + lc;
  bool cls1 = false, valid_m = true, valid_last = true;
  int off_m = 0, off_last = 0;
  if (HAS_M) {
    const int col = 8 * TS + lc;
    cls1 = col >= K;
    valid_m = col < 2 * K;
    off_m = lr * K + (valid_m ? (cls1 ? col - K : col) : 0);
  }
  {
    const int col = 8 * (T2 - 1) + lc;
    valid_last = col < 2 * K;
    off_last = lr * K + (valid_last ? col - K : 0);
  }
  double acc[NTL][2];
#pragma unroll
  for (int t = 0; t < NTL; ++t) acc[t][0] = acc[t][1] = 0.0;

  if (R == 0) {
#pragma unroll
    for (int p = 0; p < kTeSlots; ++p) team_issue<P, NCHW>(a, c, p, p);
  }

  int slot = 0;
  unsigned phase = 0;
#pragma unroll 1
  for (int st = 0; st < nst; ++st) {
    team_qc_stage<P, NCHW>(a, c, st, slot, phase);

    // ---- D: this warp's tiles of the packed triangle over the 8 k-steps of the stage ----
    {
      const unsigned xs_u = ring_u + (unsigned)(slot * slot_elems * sizeof(double));
      const unsigned wl_u = wq_u + 8u * (unsigned)(6 * lr);
#pragma unroll 1
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const unsigned row_u = xs_u + 8u * (unsigned)(4 * ks * K);
        double z[JHI];
#pragma unroll
        for (int t = 0; t < JHI; ++t) {
          int o;
          if (t == TS) o = off_m;
          else if (t == T2 - 1 && t >= TB) o = off_last;
          else o = base + ((t < T0) ? 8 * t : 8 * t - K);
          z[t] = lds_f64(row_u + 8u * (unsigned)o);
        }
        const unsigned wrow_u = wl_u + 8u * (unsigned)(24 * ks);     // row 4 ks + lr, 6 doubles per row
        const double wa = lds_f64(wrow_u + 16u);
        const double wb = lds_f64(wrow_u + 24u);
        const double wc = lds_f64(wrow_u + 32u);
#pragma unroll
        for (int t = 0; t < JHI; ++t) {
          if (t == TS) {
            const double xx = vmul(z[t], z[t]);
            z[t] = cls1 ? xx : z[t];
            if (!valid_m) z[t] = 0.0;
          } else if (t >= TB) {
            z[t] = vmul(z[t], z[t]);
            if (t == T2 - 1 && !valid_last) z[t] = 0.0;
          }
        }
        double aw2 = 0.0;
        if (HAS_M && JHI > TS) aw2 = vmul(z[(HAS_M && JHI > TS) ? TS : 0], cls1 ? wc : wb);
#pragma unroll
        for (int j = JLO; j < JHI; ++j) {
          const int cb = j * (j + 1) / 2 - TLO;
          bool need0 = (HAS_M && j == TS && mine(j, j)), needc = false;
#pragma unroll
          for (int i = 0; i < T0; ++i)
            if (i <= j && mine(i, j)) need0 = true;
#pragma unroll
          for (int i = TB; i < T2; ++i)
            if (i <= j && mine(i, j)) needc = true;
          double bw0 = 0.0, bc = 0.0;
          if (need0) bw0 = vmul(z[j], (j < T0) ? wa : ((j == TS) ? (cls1 ? wb : wa) : wb));
          if (needc) bc = vmul(z[j], wc);
#pragma unroll
          for (int i = 0; i < T0; ++i)
            if (i <= j && mine(i, j)) dmma884(acc[cb + i][0], acc[cb + i][1], z[i], bw0);
          if (HAS_M && j == TS && mine(j, j)) {
            dmma884(acc[cb + j][0], acc[cb + j][1], cls1 ? 0.0 : z[j], bw0);
            dmma884(acc[cb + j][0], acc[cb + j][1], cls1 ? z[j] : 0.0, aw2);
          }
          if (HAS_M && j > TS && mine(HAS_M ? TS : 0, j))
            dmma884(acc[cb + (HAS_M ? TS : 0)][0], acc[cb + (HAS_M ? TS : 0)][1], aw2, z[j]);
#pragma unroll
          for (int i = TB; i < T2; ++i)
            if (i <= j && mine(i, j)) dmma884(acc[cb + i][0], acc[cb + i][1], z[i], bc);
        }
      }
    }
    __syncwarp();     // every lane is done with the slot and with wq
    if (P == 1) team_issue<P, NCHW>(a, c, st + kTeSlots, slot);
    slot ^= 1;
    if (slot == 0) phase ^= 1u;
  }
  team_qc_finish<NCHW>(a, c);

  // ---- the rings become the (NT, 64) tile buffer; warps add their accumulators in warp order ----
  __syncthreads();
  for (int e = threadIdx.x; e < (T2 * (T2 + 1) / 2) * 64; e += blockDim.x) red[e] = 0.0;
  __syncthreads();
  const int e0 = (lane >> 2) * 8 + 2 * (lane & 3);
#pragma unroll 1
  for (int w2 = 0; w2 < nwarps; ++w2) {
    if (warp == w2) {
#pragma unroll
      for (int t = 0; t < NTL; ++t) {
        double* d = red + (size_t)(TLO + t) * 64 + e0;
        d[0] += acc[t][0];
        d[1] += acc[t][1];
      }
    }
    __syncthreads();
  }
}

template <int T2, int T0, bool HAS_M, int P, int R, int NCHW>
__device__ __forceinline__ void team_dispatch(int role, const FusedArgs& a, double* ring, unsigned ring_u,
                                              unsigned full_u, double* wq, double* part, const double* bm,
                                              const double* bv, const double* ghc, const double* ghw, double* gred_row,
                                              double* kred_slot, int64_t gw, int team, double* red, int warp, int nwarps) {
  if constexpr (R < P) {
    if (role == R) {
      constexpr int LO = gram_mid_bound(T2, T0, HAS_M, P, R), HI = gram_mid_bound(T2, T0, HAS_M, P, R + 1);
      team_run<T2, T0, HAS_M, P, R, LO, HI, NCHW>(a, ring, ring_u, full_u, wq, part, bm, bv, ghc, ghw, gred_row,
                                                  kred_slot, gw, team, red, warp, nwarps);
    } else {
      team_dispatch<T2, T0, HAS_M, P, R + 1, NCHW>(role, a, ring, ring_u, full_u, wq, part, bm, bv, ghc, ghw,
                                                   gred_row, kred_slot, gw, team, red, warp, nwarps);
    }
  }
}

// blockDim.x = 32 * warps, warps a multiple of P; warps [P t, P t + P) are team t
template <int T2, int T0, bool HAS_M, int P, int NCHW>
__global__ void __launch_bounds__(32 * kTeMaxWarps, 1)
k_team_eval(const FusedArgs a) {
  pdl_sync();
  constexpr int NT = T2 * (T2 + 1) / 2;
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int teams = nwarps / P;
  const int team = warp / P, role = warp % P;
  const int K = a.K, Q = a.Q;
  const int slot_elems = team_slot_elems(K);
  const size_t ring_elems = (size_t)teams * kTeSlots * slot_elems;
  const size_t pw = 6 * kTeRows + (P > 1 ? 2 * 6 * kTeRows : 0);          // per-warp: wq [+ 2 partial blocks]
  const size_t body = ring_elems + (size_t)nwarps * pw;
  const size_t red_elems = (size_t)NT * 64;
  double* tail = sm + (body > red_elems ? body : red_elems);
  double* bm = tail;
  double* bv = bm + K;
  double* ghc = bv + K;
  double* ghw = ghc + Q;
  double* gred = ghw + Q;                           // teams x 2K
  double* kred = gred + (size_t)teams * 2 * K;      // teams
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(kred + teams);
  double* ring = sm + (size_t)team * kTeSlots * slot_elems;
  double* wq = sm + ring_elems + (size_t)warp * pw;
  double* part = wq + 6 * kTeRows;
  const unsigned ring_u = smem_u32(ring);
  const unsigned full_u = smem_u32(bars + (size_t)team * kTeSlots);

  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    bm[k] = a.vec[4 + k];
    bv[k] = 1.0 / a.vec[4 + K + k];
  }
  for (int q = threadIdx.x; q < Q; q += blockDim.x) {
    ghc[q] = a.gh[q];
    ghw[q] = a.gh[Q + q];
  }
  for (int e = threadIdx.x; e < teams * 2 * K + teams; e += blockDim.x) gred[e] = 0.0;
  if (role == 0 && (threadIdx.x & 31) == 0) {
#pragma unroll
    for (int p = 0; p < kTeSlots; ++p) mbar_init(full_u + 8 * p, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int64_t gw = (int64_t)blockIdx.x * teams + team;
  team_dispatch<T2, T0, HAS_M, P, 0, NCHW>(role, a, ring, ring_u, full_u, wq, part, bm, bv, ghc, ghw,
                                           gred + (size_t)team * 2 * K, kred + team, gw, team, sm, warp, nwarps);
  // ---- CTA outputs: Gram tiles, KL partial, global-gradient partials (fixed order) ----
  double* out = a.grampart + (size_t)blockIdx.x * NT * 64;
  for (int e = threadIdx.x; e < NT * 64; e += blockDim.x) out[e] = sm[e];
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < teams; ++i) s += kred[i];
    a.klpart[blockIdx.x] = s;
  }
  for (int k = threadIdx.x; k < 2 * K; k += blockDim.x) {
    double s = 0.0;
    for (int i = 0; i < teams; ++i) s += gred[(size_t)i * 2 * K + k];
    a.gradpart[(size_t)k * gridDim.x + blockIdx.x] = s;
  }
}

// launch the instantiation for K with `warps` warps per CTA; false when K / alignment is outside its range
inline bool launch_team_eval(const FusedArgs& a, int grid, int warps, int Q, cudaStream_t st) {
  const int K = a.K;
  if (K < 1 || K > kOfMaxK || (a.ldw & 1) || (((uintptr_t)a.X) & 15)) return false;
  const int T2 = (2 * K + 7) / 8, T0 = K / 8;
  const bool M = (K % 8) != 0;
  const int nch = (K + 1 + 31) / 32;
  const size_t smem = team_smem(K, Q, T2, warps);
#define LRVB_TE(T2_, T0_, M_)                                                                        \
  if (T2 == T2_ && T0 == T0_ && M == M_) {                                                           \
    constexpr int P_ = team_P(T2_);                                                                  \
    constexpr int NCHW_ = (P_ == 1 && T2_ * 4 + 1 > 32) ? 2 : 1;                                      \
    if (warps % P_ != 0 || (P_ > 1 && nch > P_)) return false;                                       \
    static size_t configured = 48 * 1024;                                                            \
    if (smem > configured) {                                                                         \
      cudaFuncSetAttribute(k_team_eval<T2_, T0_, M_, P_, NCHW_>,                                     \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
      configured = smem;                                                                             \
    }                                                                                                \
    return launch_pdl(k_team_eval<T2_, T0_, M_, P_, NCHW_>, dim3(grid), dim3(32 * warps), smem, st, a) == \
           cudaSuccess;                                                                              \
  }
  LRVB_TE(1, 0, true)
  LRVB_TE(2, 0, true)
  LRVB_TE(2, 1, false)
  LRVB_TE(3, 1, true)
  LRVB_TE(4, 1, true)
  LRVB_TE(4, 2, false)
  LRVB_TE(5, 2, true)
  LRVB_TE(6, 2, true)
  LRVB_TE(6, 3, false)
  LRVB_TE(7, 3, true)
  LRVB_TE(8, 3, true)
  LRVB_TE(8, 4, false)
  LRVB_TE(9, 4, true)
  LRVB_TE(10, 4, true)
  LRVB_TE(10, 5, false)
  LRVB_TE(11, 5, true)
  LRVB_TE(12, 5, true)
  LRVB_TE(12, 6, false)
  LRVB_TE(13, 6, true)
  LRVB_TE(14, 6, true)
  LRVB_TE(14, 7, false)
  LRVB_TE(15, 7, true)
  LRVB_TE(16, 7, true)
#undef LRVB_TE
  return false;
}

}  // namespace lrvb
