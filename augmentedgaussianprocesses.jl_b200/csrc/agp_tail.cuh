// agp_tail.cuh -- the per-iteration m x m tail in fp64:  P_v = R R^T (Cholesky) and X = R^-1, fused.
//
// global_update! (inference/inference.jl:25-28) needs Sigma = inv(-2 eta2).  In the whitened basis the engine never
// forms Sigma_v = X^T X in the hot loop: var_f = |X v_b|^2 + Ktilde and mu_v = X^T (X eta1_v) only need the inverse
// Cholesky factor X, which is accumulated DURING the right-looking factorisation (the row operations that eliminate
// block column k are applied to [P | W], W starting as I), so there is no separate triangular-inverse phase and the
// factor R itself is never stored.
//
// One launch per 64-wide block step k (nblk launches per iteration, captured in the step's CUDA graph):
//   A tiles (i,j), k<j<=i : A_ij -= L_ik L_jk^T          with L_ik = A_ik X_kk^T formed inside the CTA
//   W tiles (i,c), c<=k<i : W_ic  = [c<k] W_ic - L_ik Wn_kc,  Wn_kc = X_kk W_kc (c<k) or X_kk (c=k)
//   F tiles (k,c), c<k    : Xout_kc = X_kk W_kc            (final rows of X = R^-1)
// and the CTA that owns the next diagonal tile (k+1,k+1) factorises it right after updating it (look-ahead), which
// produces X_{k+1,k+1} for the next launch.  All tile products are 64x64x64 fp64 from shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agp_kernels.cuh"

namespace agp {

#ifdef AGP_TAIL_DEBUG
__device__ long long agp_dbg[64 * 4];
__device__ double agp_dbgf[64];
__device__ long long agp_own[64 * 4];
#endif

constexpr int TNB = 64;         // tile size
constexpr int TLD = TNB + 1;    // padded leading dimension of a shared-memory tile (conflict-free column reads)
constexpr int TAIL_THREADS = 256;
constexpr int POTF2_THREADS = TAIL_THREADS;
constexpr int TAIL_SMEM = (5 * TNB * TLD + 2 * 72 + 3 * TNB) * (int)sizeof(double);

struct TailStepParams {
  double* P; double* W; double* Xout; double* Dinv;  // [mp][ld] x3, [nblk][64][64]
  int64_t ld;
  int nblk, k;
  double* logdet; int* status;
  // tail2_potf2_first_kernel only: non-null = the in-stream predecessor is combine_kernel with TailParams::tile0_flag set; the kernel
  // starts as soon as that counter says tile (0, 0) of P_v is written instead of waiting for the whole grid
  int* early_flag;
};

// ---- shared-memory tile helpers (256 threads) -------------------------------------------------------
// 64 x 64 fp64 tile, global (ld, 16-byte aligned rows) -> shared [64][TLD]: 8 independent 16-byte loads per thread
__device__ __forceinline__ void tile_load(double* s, const double* __restrict__ g, int64_t ld) {
  double2 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;  // double2 index: 32 per row
    v[u] = *reinterpret_cast<const double2*>(g + (int64_t)(e >> 5) * ld + (e & 31) * 2);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    double* d = s + (e >> 5) * TLD + (e & 31) * 2;
    d[0] = v[u].x; d[1] = v[u].y;
  }
}
__device__ __forceinline__ void tile_load_dense(double* s, const double* __restrict__ g) { tile_load(s, g, TNB); }
// acc[i][j] (+)= sum_k A[r_i][k] * B'[k][c_j],  r_i = ty + 16 i, c_j = tx + 16 j
//   BT == false: B' = B^T, i.e. B is stored [c][k]  ("NT");   BT == true: B stored [k][c]  ("NN")
template <bool BT>
__device__ __forceinline__ void tile_mma(const double* __restrict__ A, const double* __restrict__ B, double (&acc)[4][4]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 8
  for (int k = 0; k < TNB; ++k) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = A[(ty + 16 * i) * TLD + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = BT ? B[k * TLD + tx + 16 * j] : B[(tx + 16 * j) * TLD + k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}
__device__ __forceinline__ void acc_zero(double (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
}
__device__ __forceinline__ void acc_to_smem(double* s, const double (&acc)[4][4], double scale) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s[(ty + 16 * i) * TLD + tx + 16 * j] = scale * acc[i][j];
}

// ---- Cholesky + inverse of one 64 x 64 tile held in shared memory ------------------------------------
// sa: [64][TLD] SPD tile (lower triangle read).  Result: X = chol(sa)^-1 (lower triangular, zeros above) written to
// Xg (ld ldx) and densely to Dg; logdet += sum log d_j.  Right-looking elimination on [A | W] (W starts as I), state
// in registers, 256 threads (lo = t & 63, hi = t >> 6):
//   A piece, ROW layout   : row i = lo, columns 16 hi .. 16 hi + 15
//   W piece, COLUMN layout: column c = lo, rows 16 hi .. 16 hi + 15
// With these two layouts every vector a pivot has to broadcast -- column j of A, row j of W -- is held ONE ELEMENT PER
// THREAD by the 64 threads of group hi = j / 16, so publishing costs one shared store per thread (BAR.SYNC drains
// pending shared stores; a single thread issuing 16+ of them per pivot dominated the pivot chain in the first
// version).  Pivots are unrolled by 16 so the register index of column/row j (jj) and the "below the pivot"
// predicates are compile-time.  1/sqrt(d): fp32 seed + Newton in fp64, by the diagonal's owner, before the barrier.
__device__ __forceinline__ void tile_potf2_inv(const double* sa, double* lcol /*[2][LCS]*/, double* prow /*[2][64]*/, double* dvals /*[64]*/,
                                               double* __restrict__ Xg, int64_t ldx, double* __restrict__ Dg,
                                               double* __restrict__ logdet, int* __restrict__ status) {
  constexpr int LCS = 72;  // pivot column (64) + published 1/sqrt(d), 1/d
  const int t = threadIdx.x, lo = t & 63, hi = t >> 6, g0 = hi * 16;
  double xa[16], xw[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    xa[q] = sa[lo * TLD + g0 + q];          // a[lo][g0+q]  (entries above the diagonal are never used)
    xw[q] = (g0 + q == lo) ? 1.0 : 0.0;     // w[g0+q][lo]
  }
#pragma unroll 1
  for (int J = 0; J < 4; ++J) {
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = J * 16 + jj;
      double* lc = lcol + (jj & 1) * LCS;
      double* pr = prow + (jj & 1) * TNB;
      if (hi == J) {                         // warp-uniform: this group holds column j of A and row j of W
        lc[lo] = xa[jj];
        pr[lo] = xw[jj];
        if (lo == j) {
          double d = xa[jj];
          if (!(d > 0.0)) { atomicOr(status, ST_NOT_POSDEF); d = 1.0; }
          // only 1/d sits on the pivot chain (the sqrt scaling of W's rows is applied once at the end):
          // MUFU.RCP64H seed + two Newton steps
          double y;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
          y = y * (2.0 - d * y);
          y = y * (2.0 - d * y);
          lc[64] = y;
          dvals[j] = d;
        }
      }
      __syncthreads();
#ifdef AGP_ABL_NO_UPDATE
      if (false) {
#else
      if (hi >= J) {
#endif
        const double inv_d = lc[64];
        double l16[16];                      // column j of A at rows/cols g0 .. g0+15 (shared by both updates)
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          double2 v = *reinterpret_cast<const double2*>(lc + g0 + q);
          l16[q] = v.x; l16[q + 1] = v.y;
        }
        // A part: a_ic -= a_ij a_cj / d   for rows i = lo > j and my columns c = g0+q > j
        const double ga = (lo > j) ? -lc[lo] * inv_d : 0.0;
        // W part: w_ic -= a_ij w_jc / d   for my rows i = g0+q > j and column c = lo <= j
        const double pw = pr[lo];
        const double gw = (lo <= j) ? -pw * inv_d : 0.0;
#ifdef AGP_ABL_NO_FMA
        if (false) {
#else
        if (hi > J) {
#endif
#pragma unroll
          for (int q = 0; q < 16; ++q) { xa[q] = fma(l16[q], ga, xa[q]); xw[q] = fma(l16[q], gw, xw[q]); }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) if (q > jj) { xa[q] = fma(l16[q], ga, xa[q]); xw[q] = fma(l16[q], gw, xw[q]); }
          // row j of W is final up to its scaling by 1/sqrt(d_j), applied after the loop
        }
      }
    }
  }
  __syncthreads();                         // dvals complete
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    int i = g0 + q;
    double v = (lo <= i) ? xw[q] * rsqrt(dvals[i]) : 0.0;   // X = diag(d)^-1/2 W
    Xg[(int64_t)i * ldx + lo] = v;
    Dg[i * TNB + lo] = v;
  }
  if (t < TNB) {  // logdet += sum_j log d_j  (= 2 sum log R_jj), off the pivot chain
    double l = warp_sum(log(dvals[t]));
    if ((t & 31) == 0) atomicAdd(logdet, l);
  }
}

// reciprocal on the pivot chain: MUFU.RCP64H seed + NEWTON Newton steps (branch-free)
template <int NEWTON>
__device__ __forceinline__ double rcp_chain(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
  for (int it = 0; it < NEWTON; ++it) y = fma(y, fma(-d, y, 1.0), y);
  return y;
}
// generation-1 tail (kept selectable with AGP_TAIL_VARIANT=0 as the A/B reference of agp_tail2.cuh); VAR is 0
template <int VAR>
__device__ __forceinline__ void tile_potf2_dispatch(const double* sa, double* vec, double* __restrict__ Xg, int64_t ldx,
                                                    double* __restrict__ Dg, double* __restrict__ logdet, int* __restrict__ status) {
  tile_potf2_inv(sa, vec, vec + 2 * 72, vec + 2 * 72 + 2 * TNB, Xg, ldx, Dg, logdet, status);
}

// first diagonal block (no update precedes it)
template <int VAR>
__global__ void __launch_bounds__(POTF2_THREADS, 1) tail_potf2_first_kernel(const TailStepParams p) {
  extern __shared__ double sm[];
  double* sa = sm;
  double* vec = sm + 5 * TNB * TLD;
  tile_load(sa, p.P, p.ld);
  __syncthreads();
  tile_potf2_dispatch<VAR>(sa, vec, p.Xout, p.ld, p.Dinv, p.logdet, p.status);
}

template <int VAR>
__global__ void __launch_bounds__(TAIL_THREADS, 1) tail_step_kernel(const TailStepParams p) {
  extern __shared__ double sm[];
  double* sX = sm;                   // X_kk
  double* s1 = sm + 1 * TNB * TLD;   // A_ik
  double* s2 = sm + 2 * TNB * TLD;   // A_jk / W_kc
  double* s3 = sm + 3 * TNB * TLD;   // L_ik
  double* s4 = sm + 4 * TNB * TLD;   // L_jk / Wn_kc / updated tile
  double* vec = sm + 5 * TNB * TLD;
  const int k = p.k, nblk = p.nblk, r = nblk - 1 - k;
  const int nA = r * (r + 1) / 2, nW = r * (k + 1);
  int b = blockIdx.x;
  const int64_t ld = p.ld;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];

  tile_load_dense(sX, p.Dinv + (int64_t)k * TNB * TNB);

  if (b < nA) {
    // ---- A tile (i, j), k < j <= i ----
    int ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= b) ++ii;
    int jj = b - ii * (ii + 1) / 2;
    const int i = k + 1 + ii, j = k + 1 + jj;
    tile_load(s1, p.P + (int64_t)i * TNB * ld + (int64_t)k * TNB, ld);
    if (j != i) tile_load(s2, p.P + (int64_t)j * TNB * ld + (int64_t)k * TNB, ld);
    __syncthreads();
    acc_zero(acc); tile_mma<false>(s1, sX, acc); acc_to_smem(s3, acc, 1.0);        // L_ik = A_ik X_kk^T
    const double* Lj = s3;
    if (j != i) { acc_zero(acc); tile_mma<false>(s2, sX, acc); acc_to_smem(s4, acc, 1.0); Lj = s4; }
    __syncthreads();
    acc_zero(acc); tile_mma<false>(s3, Lj, acc);                                   // L_ik L_jk^T
    double* At = p.P + (int64_t)i * TNB * ld + (int64_t)j * TNB;
    const bool lookahead = (i == k + 1 && j == k + 1);
    __syncthreads();  // s1/s2 are free again
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int rr = ty + 16 * a, cc = tx + 16 * c;
        double v = At[(int64_t)rr * ld + cc] - acc[a][c];
        if (lookahead) s1[rr * TLD + cc] = v; else At[(int64_t)rr * ld + cc] = v;
      }
    if (lookahead) {
      __syncthreads();
      tile_potf2_dispatch<VAR>(s1, vec, p.Xout + (int64_t)(k + 1) * TNB * (ld + 1), ld, p.Dinv + (int64_t)(k + 1) * TNB * TNB, p.logdet, p.status);
    }
  } else if (b < nA + nW) {
    // ---- W tile (i, c), c <= k < i ----
    b -= nA;
    const int i = k + 1 + b / (k + 1), c = b % (k + 1);
    tile_load(s1, p.P + (int64_t)i * TNB * ld + (int64_t)k * TNB, ld);
    if (c < k) tile_load(s2, p.W + (int64_t)k * TNB * ld + (int64_t)c * TNB, ld);
    __syncthreads();
    acc_zero(acc); tile_mma<false>(s1, sX, acc); acc_to_smem(s3, acc, 1.0);        // L_ik
    const double* Wn = sX;
    if (c < k) { acc_zero(acc); tile_mma<true>(sX, s2, acc); acc_to_smem(s4, acc, 1.0); Wn = s4; }  // Wn_kc = X_kk W_kc
    __syncthreads();
    acc_zero(acc); tile_mma<true>(s3, Wn, acc);                                    // L_ik Wn_kc
    double* Wt = p.W + (int64_t)i * TNB * ld + (int64_t)c * TNB;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int cc4 = 0; cc4 < 4; ++cc4) {
        int rr = ty + 16 * a, cc = tx + 16 * cc4;
        double old = (c < k) ? Wt[(int64_t)rr * ld + cc] : 0.0;
        Wt[(int64_t)rr * ld + cc] = old - acc[a][cc4];
      }
  } else {
    // ---- F tile (k, c), c < k : final row block of X ----
    const int c = b - nA - nW;
    tile_load(s2, p.W + (int64_t)k * TNB * ld + (int64_t)c * TNB, ld);
    __syncthreads();
    acc_zero(acc); tile_mma<true>(sX, s2, acc);
    double* Xt = p.Xout + (int64_t)k * TNB * ld + (int64_t)c * TNB;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int cc4 = 0; cc4 < 4; ++cc4) Xt[(int64_t)(ty + 16 * a) * ld + tx + 16 * cc4] = acc[a][cc4];
  }
}

// One block per row i of the lower-triangular X = chol(P_v)^-1:  T shadow (zeros above the diagonal) for the B x m
// contraction, optional TF32 hi/lo split of that shadow, and t_i = (X eta1_v)_i  (mu_v = X^T t is never needed in
// the hot loop: mean_f = V mu_v = (V X^T) t).
template <typename T>
__device__ __forceinline__ void x_finalize_row(const int i, const double* __restrict__ X, int64_t ld, int m, const double* __restrict__ eta1v,
                                               T* __restrict__ shadow, int64_t lds, float* __restrict__ hi, float* __restrict__ lo,
                                               double* __restrict__ tvec) {
  double s = 0.0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    double v = (j <= i) ? X[(int64_t)i * ld + j] : 0.0;
    s += v * eta1v[j];
    T tv = (T)v;
    shadow[(int64_t)i * lds + j] = tv;
    if (hi) {
      float fv = (float)tv;
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(fv));
      float h = __uint_as_float(u);
      hi[(int64_t)i * lds + j] = h;
      float lv = fv - h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(lv));
      lo[(int64_t)i * lds + j] = __uint_as_float(u);
    }
  }
  __shared__ double sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
    tvec[i] = a;
  }
}
template <typename T>
__global__ void x_finalize_kernel(const double* __restrict__ X, int64_t ld, int m, const double* __restrict__ eta1v,
                                  T* __restrict__ shadow, int64_t lds, float* __restrict__ hi, float* __restrict__ lo,
                                  double* __restrict__ tvec, double* __restrict__ lr_next, int64_t* __restrict__ counters,
                                  double rm_kappa, double rm_tau, int bump, int row0 = 0) {
  pdl_prologue();
  const int i = blockIdx.x + row0;    // launches may cover a block of rows (rows leave the tail block by block)
  // Robbins-Monro step size of the NEXT iteration (inference/optimisers.jl:14-19; the counter is bumped after this kernel):
  // one thread of one block, hidden behind the rest of the grid instead of sitting on the next step's chain
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (lr_next) *lr_next = pow(rm_tau + (double)(counters[0] + 1), -rm_kappa);
    if (bump) { counters[0] += 1; counters[1] += 1; }   // end of the step: Robbins-Monro counter and minibatch-list cursor
  }
  x_finalize_row<T>(i, X, ld, m, eta1v, shadow, lds, hi, lo, tvec);
}
// the latents of a batch in one launch (blockIdx.y = latent; pointers by value, see SMALL_NB in agp_kernels.cuh); lr_next / bump are
// passed by the launch that holds the step's LAST latent only
constexpr int FINALIZE_NB = 16;
template <typename T>
struct FinalizeBatch { const double* X[FINALIZE_NB]; const double* eta1v[FINALIZE_NB]; T* shadow[FINALIZE_NB]; float* hi[FINALIZE_NB]; float* lo[FINALIZE_NB]; double* tvec[FINALIZE_NB]; };
template <typename T>
__global__ void x_finalize_batched_kernel(const FinalizeBatch<T> bt, int64_t ld, int m, int64_t lds, double* __restrict__ lr_next,
                                          int64_t* __restrict__ counters, double rm_kappa, double rm_tau, int bump) {
  pdl_prologue();
  const int z = blockIdx.y;
  if (blockIdx.x == 0 && z == 0 && threadIdx.x == 0) {
    if (lr_next) *lr_next = pow(rm_tau + (double)(counters[0] + 1), -rm_kappa);
    if (bump) { counters[0] += 1; counters[1] += 1; }
  }
  x_finalize_row<T>(blockIdx.x, bt.X[z], ld, m, bt.eta1v[z], bt.shadow[z], lds, bt.hi[z], bt.lo[z], bt.tvec[z]);
}

// End of a step whose tail was the experimental Newton-Schulz refinement (no factor X, so no x_finalize_kernel): the step-size /
// counter duties of x_finalize_kernel, plus the convergence check: resid2[i] = |I - Y P|_F^2 at the START of iteration i, the
// error squares per iteration, so the result is accepted when the last recorded residual is below sqrt-tolerance.
__global__ void ns_finalize_kernel(const double* __restrict__ resid2, int iters, double thr2, int* __restrict__ status,
                                   double* __restrict__ lr_next, int64_t* __restrict__ counters, double rm_kappa, double rm_tau, int bump) {
  pdl_prologue();
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  if (lr_next) *lr_next = pow(rm_tau + (double)(counters[0] + 1), -rm_kappa);
  if (bump) { counters[0] += 1; counters[1] += 1; }
  if (iters > 0 && !(resid2[iters - 1] < thr2)) atomicOr(status, ST_NS_NOCONV);
}

// out[0] += |X|_F^2 (= tr(Sigma_v)),  out[1] += |mu_v - mu0_v|^2 ; one warp per row of X, 8 rows per block
__global__ void gauss_kl_x_kernel(const double* __restrict__ X, int64_t ld, int m, const double* __restrict__ muv,
                                  const double* __restrict__ mu0v, double* __restrict__ out) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= m) return;
  double tr = 0.0;
  for (int j = lane; j <= i; j += 32) { double v = X[(int64_t)i * ld + j]; tr += v * v; }
  tr = warp_sum(tr);
  if (lane == 0) {
    double d = muv[i] - mu0v[i];
    atomicAdd(out + 0, tr);
    atomicAdd(out + 1, d * d);
  }
}

}  // namespace agp
