// agp_tail3.cuh -- third-generation m x m tail (fp64): ONE persistent launch factorises P_v = R R^T and accumulates X = R^-1
// for EVERY latent GP this rank owns (global_update! of inference/inference.jl:25-28: Sigma = inv(-2 eta2), kept as X with
// Sigma_v = X^T X).
//
// Same block algorithm and tile arithmetic as agp_tail2.cuh (64-wide block steps, DMMA tile products, panel potf2 with the
// inverse factor accumulated in the same pass), but the nblk dependent launches per latent are replaced by a dataflow inside one
// grid of nlat x G co-resident CTAs (one per SM, grid <= 148):
//   role 0 of each team ("chain CTA")  : D(0); then for k = 0 .. nblk-2:  U(k+1,k+1,k) (look-ahead SYRK, tile kept in shared
//                                        memory) and D(k+1) = Cholesky + inverse of the diagonal tile.  It never leaves the
//                                        critical path and never waits for a launch.
//   roles 1 .. G-1 ("helpers")         : the other tile tasks of block step k, dealt round-robin (task t -> helper t mod H):
//       U(i,j,k), k<j<=i : A_ij -= L_ik L_jk^T        with L_ik = A_ik X_kk^T formed inside the CTA
//       W(i,c,k), c<=k<i : W_ic  = [c<k] W_ic - L_ik Wn_kc,  Wn_kc = X_kk W_kc (c<k) or X_kk (c=k)
//       F(k,c),   c<k    : Xout_kc = X_kk W_kc         (final rows of X)
//     in priority order: the two tiles the chain CTA needs for its NEXT look-ahead first, then the rest of block column k+1.
// Dependencies are per-tile version words in global memory, written with st.release.gpu after a CTA barrier and polled with
// ld.acquire.gpu; every word carries the launch epoch in its high bits, so nothing is cleared between launches (the last CTA of
// a team bumps the epoch).  Tiles written by other CTAs are read with ld.global.cg (L2), never through L1.
// Measured (profiles/r2/tail3_*.txt, B200): per block step the chain CTA spends 10.6 us in the diagonal tile (4 panels x [16 pivots
// ~1.5 us + strip / trailing-head phases]), 3.9 us in the look-ahead L / SYRK products (DMMA: 64 FMA/clk/SM), 1.5 us loading its two
// tiles from L2, ~1 us in flag traffic; it never waits for a helper.  One latent at m = 512: 138 us (the multi-launch tail2 chain:
// 126 us, programmatic dependent launches cost less than release / acquire round trips); eight latents in one launch: 146 us
// (8 x 126 us sequentially).  A variant of the diagonal-tile routine in which warp 0 alone carries the critical path (chain -> 16 x 16
// strip -> next diagonal block, no CTA barrier) was bit-identical and no faster (the pivots themselves slow down to ~2 us per panel
// when the other warps' DMMA work overlaps them), so tile2_potf2_inv stays.
// Every wait is bounded (~1 s): on expiry the CTA raises ST_TAIL_TIMEOUT, stops waiting and runs to completion on whatever
// data it finds, so a scheduling problem surfaces as AGP_ERR_STATE instead of a hung GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agp_tail2.cuh"

namespace agp {

typedef unsigned long long u64;

// timeline instrumentation for profiles/microbench/tail3_test.cu (compiled in only with -DAGP_T3_TRACE): per CTA a list of
// (tag, clock64, globaltimer) records; tag = k * 1000 + task type * 100 + event
#ifdef AGP_T3_TRACE
constexpr int T3_TRACE_SLOTS = 512;
__device__ unsigned long long agp_t3_trace[160 * T3_TRACE_SLOTS * 3];
__device__ int agp_t3_trace_n[160];
__device__ __forceinline__ void t3_trace(int tag) {
  if (threadIdx.x == 0) {
    int n = atomicAdd(&agp_t3_trace_n[blockIdx.x], 1);
    if (n < T3_TRACE_SLOTS) {
      unsigned long long g;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
      unsigned long long* r = agp_t3_trace + ((size_t)blockIdx.x * T3_TRACE_SLOTS + n) * 3;
      r[0] = (unsigned long long)tag; r[1] = (unsigned long long)clock64(); r[2] = g;
    }
  }
}
#define T3T(tag) t3_trace(tag)
#else
#define T3T(tag) do { } while (0)
#endif

struct Tail3Lat {
  double* P; double* W; double* Xout; double* Dinv;   // [mp][ld] x3, [nblk][64][64]
  double* logdet;
  u64* flags;                                         // [2 + nblk + 2 nblk^2]: epoch, done, D[k], verA[i][j], verW[i][c]
};
struct Tail3Params {
  const Tail3Lat* lat;   // device array, one entry per latent of this launch
  int nlat, G, nblk;
  int64_t ld;
  int* status;
};
static inline size_t tail3_flag_words(int nblk) { return 2 + (size_t)nblk + 2 * (size_t)nblk * nblk; }
// number of helper tasks of block step k (the look-ahead tile (k+1,k+1) belongs to the chain CTA)
__host__ __device__ __forceinline__ int tail3_nU(int nblk, int k) { const int r = nblk - 1 - k; return r >= 2 ? r * (r + 1) / 2 - 1 : 0; }
__host__ __device__ __forceinline__ int tail3_nW(int nblk, int k) { return (nblk - 1 - k) * (k + 1); }
static inline int tail3_max_tasks(int nblk) {
  int mx = 0;
  for (int k = 0; k < nblk; ++k) { int t = tail3_nU(nblk, k) + tail3_nW(nblk, k) + k; mx = t > mx ? t : mx; }
  return mx;
}

__device__ __forceinline__ u64 t3_ld_acquire(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void t3_st_release(u64* p, u64 v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

// one thread: wait until *p >= want (words only grow: epoch << 16 | version)
__device__ __forceinline__ void t3_wait(const u64* p, u64 want, int* __restrict__ status, volatile int* s_dead) {
  if (*s_dead) return;
  if (t3_ld_acquire(p) >= want) return;
  const long long t0 = clock64();
  while (t3_ld_acquire(p) < want) {
    if (clock64() - t0 > 2000000000LL) {
      atomicOr(status, ST_TAIL_TIMEOUT);
      *s_dead = 1;
      return;
    }
  }
}
// whole CTA: everything this CTA stored so far becomes visible to whoever acquires the flag.  The release is cumulative over
// the CTA barrier (PTX memory model: bar.sync orders the other threads' stores before it), and it is issued by the LAST warp so
// that the ~1 us it takes does not delay the waits / loads that threads 0..127 start right after.
__device__ __forceinline__ void t3_signal(u64* p, u64 v) {
  __syncthreads();
  if (threadIdx.x == TAIL_THREADS - 1) t3_st_release(p, v);
}

// 64 x 64 tile, global (written by another CTA of this launch: L2 loads) -> shared [64][T2LD]
__device__ __forceinline__ void tile3_load(double* s, const double* g, int64_t ld) {
  double2 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    v[u] = __ldcg(reinterpret_cast<const double2*>(g + (int64_t)(e >> 5) * ld + (e & 31) * 2));
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    *reinterpret_cast<double2*>(s + (e >> 5) * T2LD + (e & 31) * 2) = v[u];
  }
}


constexpr int TAIL3_SMEM = TAIL2_SMEM;

struct T3Ctx {
  const Tail3Lat* L;
  int nblk; int64_t ld; int* status;
  u64 ep;                  // epoch << 16
  u64 *fD, *fA, *fW;
  volatile int* s_dead;
  double *s0, *s1, *s2, *s3, *s4, *sl, *vec;
  __device__ __forceinline__ double* Pt(int i, int j) const { return L->P + (int64_t)i * TNB * ld + (int64_t)j * TNB; }
  __device__ __forceinline__ double* Wt(int i, int c) const { return L->W + (int64_t)i * TNB * ld + (int64_t)c * TNB; }
  __device__ __forceinline__ double* Xt(int i, int c) const { return L->Xout + (int64_t)i * TNB * ld + (int64_t)c * TNB; }
  __device__ __forceinline__ double* Dt(int k) const { return L->Dinv + (int64_t)k * TNB * TNB; }
  __device__ __forceinline__ void waitD(int k) const { t3_wait(fD + k, ep | 1ull, status, s_dead); }
  __device__ __forceinline__ void waitA(int i, int j, int ver) const { if (ver > 0) t3_wait(fA + i * nblk + j, ep | (u64)ver, status, s_dead); }
  __device__ __forceinline__ void waitW(int i, int c, int cnt) const { if (cnt > 0) t3_wait(fW + i * nblk + c, ep | (u64)cnt, status, s_dead); }
};

// ---- chain CTA -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tail3_chain(const T3Ctx& c) {
  double* xb[2] = {c.s0, c.s2};       // X_kk of the current / next diagonal tile (alternating)
  double acc[8][2];
  T3T(0);
  tile3_load(c.s1, c.Pt(0, 0), c.ld);
  __syncthreads();
  T3T(1);
  tile2_potf2_inv(c.s1, xb[0], c.sl, c.vec, c.Xt(0, 0), c.ld, c.Dt(0), c.L->logdet, c.status);
  T3T(2);
  t3_signal(c.fD + 0, c.ep | 1ull);
  T3T(3);
  for (int k = 0; k + 1 < c.nblk; ++k) {
    const double* sX = xb[k & 1];
    double* sXn = xb[(k + 1) & 1];
    // look-ahead U(k+1,k+1,k): A_{k+1,k} must have received block steps 0..k-1, and so must the diagonal tile
    if (threadIdx.x == 0) c.waitA(k + 1, k, k);
    if (threadIdx.x == 32) c.waitA(k + 1, k + 1, k);
    __syncthreads();
    T3T((k + 1) * 1000 + 0);
    const double* At = c.Pt(k + 1, k + 1);
    const SyrkTiles tl = syrk_tiles();
    const int c2 = 2 * (threadIdx.x & 3);
    double2 old5[5];
    double acc5[5][2];
#pragma unroll
    for (int s_ = 0; s_ < 5; ++s_)
      old5[s_] = tl.live[s_] ? __ldcg(reinterpret_cast<const double2*>(At + (int64_t)tl.row[s_] * c.ld + tl.col[s_] + c2)) : make_double2(0.0, 0.0);
    tile3_load(c.s1, c.Pt(k + 1, k), c.ld);
    __syncthreads();
    T3T((k + 1) * 1000 + 1);
    tile2_prod<0, true, 1, false>(c.s1, sX, acc); acc2_to_smem<0>(c.s3, acc);         // L_{k+1,k} = A_{k+1,k} X_kk^T
    __syncthreads();
    tile2_syrk_lower(c.s3, tl, acc5);
#pragma unroll
    for (int s_ = 0; s_ < 5; ++s_)
      if (tl.live[s_]) *reinterpret_cast<double2*>(c.s1 + tl.row[s_] * T2LD + tl.col[s_] + c2) = make_double2(old5[s_].x - acc5[s_][0], old5[s_].y - acc5[s_][1]);
    __syncthreads();
    T3T((k + 1) * 1000 + 2);
    tile2_potf2_inv(c.s1, sXn, c.sl, c.vec, c.Xt(k + 1, k + 1), c.ld, c.Dt(k + 1), c.L->logdet, c.status);
    T3T((k + 1) * 1000 + 3);
    t3_signal(c.fD + k + 1, c.ep | 1ull);
    T3T((k + 1) * 1000 + 4);
  }
}

// ---- helper tasks ----------------------------------------------------------------------------------------------------------
// U(i,j,k): A_ij -= L_ik L_jk^T.  sX (= c.s0) holds X_kk when have_x; otherwise it is loaded after D(k) has been signalled.
__device__ __forceinline__ void tail3_task_U(const T3Ctx& c, int i, int j, int k, bool& have_x) {
  double acc[8][2];
  T3T(k * 1000 + 100 + 0);
  if (threadIdx.x == 32) c.waitA(i, k, k);
  if (threadIdx.x == 64 && j != i) c.waitA(j, k, k);
  if (threadIdx.x == 96) c.waitA(i, j, k);
  __syncthreads();
  T3T(k * 1000 + 100 + 1);
  double* At = c.Pt(i, j);
  tile3_load(c.s1, c.Pt(i, k), c.ld);
  if (j != i) tile3_load(c.s2, c.Pt(j, k), c.ld);
  T3T(k * 1000 + 100 + 2);
  if (!have_x) {
    if (threadIdx.x == 0) c.waitD(k);
    __syncthreads();
    T3T(k * 1000 + 100 + 3);
    tile3_load(c.s0, c.Dt(k), TNB);
    have_x = true;
  }
  T3T(k * 1000 + 100 + 4);
  if (j != i) {
    double2 old[8];   // the tile being updated, prefetched in the accumulator layout (MAP 0)
#pragma unroll
    for (int t = 0; t < 8; ++t) old[t] = __ldcg(reinterpret_cast<const double2*>(At + (int64_t)acc_row<0>(t) * c.ld + acc_col<0>(t)));
    __syncthreads();
    tile2_prod<0, true, 1, false>(c.s1, c.s0, acc); acc2_to_smem<0>(c.s3, acc);
    tile2_prod<0, true, 1, false>(c.s2, c.s0, acc); acc2_to_smem<0>(c.s4, acc);
    __syncthreads();
    tile2_prod<0, true, 0, false>(c.s3, c.s4, acc);
#pragma unroll
    for (int t = 0; t < 8; ++t)
      *reinterpret_cast<double2*>(At + (int64_t)acc_row<0>(t) * c.ld + acc_col<0>(t)) = make_double2(old[t].x - acc[t][0], old[t].y - acc[t][1]);
  } else {
    const SyrkTiles tl = syrk_tiles();
    const int c2 = 2 * (threadIdx.x & 3);
    double2 old5[5];
    double acc5[5][2];
#pragma unroll
    for (int s_ = 0; s_ < 5; ++s_)
      old5[s_] = tl.live[s_] ? __ldcg(reinterpret_cast<const double2*>(At + (int64_t)tl.row[s_] * c.ld + tl.col[s_] + c2)) : make_double2(0.0, 0.0);
    __syncthreads();
    tile2_prod<0, true, 1, false>(c.s1, c.s0, acc); acc2_to_smem<0>(c.s3, acc);
    __syncthreads();
    tile2_syrk_lower(c.s3, tl, acc5);
#pragma unroll
    for (int s_ = 0; s_ < 5; ++s_)
      if (tl.live[s_]) *reinterpret_cast<double2*>(At + (int64_t)tl.row[s_] * c.ld + tl.col[s_] + c2) = make_double2(old5[s_].x - acc5[s_][0], old5[s_].y - acc5[s_][1]);
  }
  T3T(k * 1000 + 100 + 5);
  t3_signal(c.fA + i * c.nblk + j, c.ep | (u64)(k + 1));
  T3T(k * 1000 + 100 + 6);
}
// W(i,c,k): W_ic = [c<k] W_ic - L_ik Wn_kc
__device__ __forceinline__ void tail3_task_W(const T3Ctx& c, int i, int cc, int k, bool& have_x) {
  double acc[8][2];
  if (threadIdx.x == 32) c.waitA(i, k, k);
  if (threadIdx.x == 64 && cc < k) c.waitW(k, cc, k - cc);      // W_kc has received block steps cc .. k-1: final
  if (threadIdx.x == 96 && cc < k) c.waitW(i, cc, k - cc);
  __syncthreads();
  double* Wt = c.Wt(i, cc);
  tile3_load(c.s1, c.Pt(i, k), c.ld);
  if (cc < k) tile3_load(c.s2, c.Wt(k, cc), c.ld);
  if (!have_x) {
    if (threadIdx.x == 0) c.waitD(k);
    __syncthreads();
    tile3_load(c.s0, c.Dt(k), TNB);
    have_x = true;
  }
  double2 old[8];
#pragma unroll
  for (int t = 0; t < 8; ++t)
    old[t] = (cc < k) ? __ldcg(reinterpret_cast<const double2*>(Wt + (int64_t)acc_row<0>(t) * c.ld + acc_col<0>(t))) : make_double2(0.0, 0.0);
  __syncthreads();
  tile2_prod<0, true, 1, false>(c.s1, c.s0, acc); acc2_to_smem<0>(c.s3, acc);                        // L_ik
  if (cc < k) { tile2_prod<1, false, 2, false>(c.s0, c.s2, acc); acc2_to_smem<1>(c.s4, acc); }      // Wn_kc = X_kk W_kc
  __syncthreads();
  if (cc < k) tile2_prod<0, false, 0, false>(c.s3, c.s4, acc);
  else tile2_prod<0, false, 3, false>(c.s3, c.s0, acc);                                              // Wn = X_kk (lower triangular)
#pragma unroll
  for (int t = 0; t < 8; ++t)
    *reinterpret_cast<double2*>(Wt + (int64_t)acc_row<0>(t) * c.ld + acc_col<0>(t)) = make_double2(old[t].x - acc[t][0], old[t].y - acc[t][1]);
  t3_signal(c.fW + i * c.nblk + cc, c.ep | (u64)(k - cc + 1));
}
// F(k,c): Xout_kc = X_kk W_kc
__device__ __forceinline__ void tail3_task_F(const T3Ctx& c, int cc, int k, bool& have_x) {
  double acc[8][2];
  if (threadIdx.x == 64) c.waitW(k, cc, k - cc);
  __syncthreads();
  tile3_load(c.s2, c.Wt(k, cc), c.ld);
  if (!have_x) {
    if (threadIdx.x == 0) c.waitD(k);
    __syncthreads();
    tile3_load(c.s0, c.Dt(k), TNB);
    have_x = true;
  }
  __syncthreads();
  tile2_prod<1, false, 2, false>(c.s0, c.s2, acc);
  double* Xt = c.Xt(k, cc);
#pragma unroll
  for (int t = 0; t < 8; ++t) *reinterpret_cast<double2*>(Xt + (int64_t)acc_row<1>(t) * c.ld + acc_col<1>(t)) = make_double2(acc[t][0], acc[t][1]);
  __syncthreads();     // s0 / s2 are reused by the next task
}

__device__ __forceinline__ void tail3_helper(const T3Ctx& c, int h, int H) {
  const int nblk = c.nblk;
  for (int k = 0; k < nblk; ++k) {
    const int r = nblk - 1 - k;
    const int nU = tail3_nU(nblk, k), nW = tail3_nW(nblk, k), nT = nU + nW + k;
    bool have_x = false;
    for (int t = h; t < nT; t += H) {
      if (t < nU) {
        // priority order: (k+2,k+1), (k+2,k+2) -- the chain CTA's next look-ahead operands -- then the rest of column k+1,
        // then the remaining tiles of the trailing matrix row by row
        int i, j;
        if (t == 0) { i = k + 2; j = k + 1; }
        else if (t == 1) { i = k + 2; j = k + 2; }
        else if (t < r) { i = k + 1 + t; j = k + 1; }
        else {
          const int u = t - r + 1;
          int ii = 0;
          while ((ii + 1) * (ii + 2) / 2 <= u) ++ii;
          i = k + 2 + ii; j = k + 2 + (u - ii * (ii + 1) / 2);
        }
        tail3_task_U(c, i, j, k, have_x);
      } else if (t < nU + nW) {
        const int b = t - nU;
        tail3_task_W(c, k + 1 + b / (k + 1), b % (k + 1), k, have_x);     // row k+1 first: the next block step needs it final
      } else {
        tail3_task_F(c, t - nU - nW, k, have_x);
      }
    }
  }
}

__global__ void __launch_bounds__(TAIL_THREADS, 1) tail3_kernel(const Tail3Params p) {
  extern __shared__ double sm[];
  __shared__ u64 s_ep;
  __shared__ int s_dead;
  pdl_launch_dependents();   // (dependents only start once every CTA of this grid is resident: no slot is taken from a late CTA)
  pdl_wait();
  const int team = blockIdx.x / p.G, role = blockIdx.x % p.G;
  if (team >= p.nlat) return;
  const Tail3Lat* L = p.lat + team;
  if (threadIdx.x == 0) { s_ep = __ldcg(L->flags); s_dead = 0; }
  __syncthreads();
  T3Ctx c;
  c.L = L; c.nblk = p.nblk; c.ld = p.ld; c.status = p.status;
  c.ep = s_ep << 16;
  c.fD = L->flags + 2; c.fA = c.fD + p.nblk; c.fW = c.fA + p.nblk * p.nblk;
  c.s_dead = &s_dead;
  c.s0 = sm; c.s1 = sm + 1 * T2_TILE; c.s2 = sm + 2 * T2_TILE; c.s3 = sm + 3 * T2_TILE; c.s4 = sm + 4 * T2_TILE;
  c.sl = sm + 5 * T2_TILE; c.vec = c.sl + TNB * T2SL;
  if (role == 0) tail3_chain(c);
  else tail3_helper(c, role - 1, p.G - 1);
  // the last CTA of the team to finish opens the next epoch (every CTA has read the current one by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const u64 done = atomicAdd(L->flags + 1, 1ull);
    if (done == (u64)(p.G - 1)) {
      L->flags[1] = 0ull;
      L->flags[0] = s_ep + 1ull;
      __threadfence();
    }
  }
}

}  // namespace agp
