// agp_tail2.cuh -- second-generation m x m tail (fp64): P_v = R R^T and X = R^-1, fused, on the fp64 tensor path.
//
// Same block algorithm and launch structure as agp_tail.cuh (one launch per 64-wide block step, look-ahead
// factorisation of the next diagonal tile inside the step kernel; global_update! of inference/inference.jl:25-28),
// but every tile product runs on DMMA (mma.sync.m8n8k4.f64) from shared memory tiles with leading dimension 68
// (68 = 4 mod 16: the 8x4 / 4x8 fragment loads are bank-conflict free), triangular operands skip their zero
// k-range, and the 64 x 64 diagonal tile is factorised by PANELS of 16 columns:
//   (a) one warp eliminates the 16 x 16 diagonal block (pivot chain; rank-1 updates on 16 x 16 only, lanes 0-15 hold
//       the rows of the block, lanes 16-31 the columns of its inverse factor, one 16-double broadcast per pivot),
//   (b) all warps: strip L_IJ = A_IJ X_JJ^T and the finished row block X_J,: = X_JJ T_J,:           (DMMA, K = 16)
//   (c) all warps: trailing A_trail -= L L^T and T_I,: -= L_IJ X_J,: for the rows below            (DMMA, K = 16)
// so the per-pivot work drops from a 64 x 128 register-resident rank-1 update (measured 568 cycles per pivot, bound by
// shared-memory operand traffic and fp64 issue) to a 16 x 32 one (~110 cycles: STS -> LDS -> rcp -> FMA).
// Measured on B200 (profiles/r1): DFMA latency 8.4 cycles, 62 DFMA/clk/SM; DMMA m8n8k4 27.5 cycles latency, 64 FMA/clk/SM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agp_tail.cuh"

namespace agp {

constexpr int T2LD = 68;                 // leading dimension of a 64 x 64 fp64 tile in shared memory
constexpr int T2SL = 20;                 // leading dimension of the 64 x 16 panel strip (20 = 4 mod 16)
constexpr int T2_TILE = TNB * T2LD;      // doubles per tile buffer
constexpr int TAIL2_SMEM = (5 * T2_TILE + TNB * T2SL + 2 * 16 + 16 + TNB + 8) * (int)sizeof(double);

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 64 x 64 tile, global (ld, 16-byte aligned rows) -> shared [64][T2LD]
__device__ __forceinline__ void tile2_load(double* s, const double* __restrict__ g, int64_t ld) {
  double2 v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    v[u] = *reinterpret_cast<const double2*>(g + (int64_t)(e >> 5) * ld + (e & 31) * 2);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    *reinterpret_cast<double2*>(s + (e >> 5) * T2LD + (e & 31) * 2) = v[u];
  }
}
__device__ __forceinline__ void tile2_store(double* __restrict__ g, int64_t ld, const double* s) {
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int e = threadIdx.x + u * TAIL_THREADS;
    *reinterpret_cast<double2*>(g + (int64_t)(e >> 5) * ld + (e & 31) * 2) = *reinterpret_cast<const double2*>(s + (e >> 5) * T2LD + (e & 31) * 2);
  }
}

// ---- 64 x 64 x 64 tile product on DMMA -----------------------------------------------------------------------
// acc[t] = the 8 x 8 output tile t of the 8 tiles this warp owns.
//   MAP 0: warp w owns output row block w  (tiles t = column block 0..7);  MAP 1: warp w owns column block w (t = row block)
//   BT : B is stored [n][k] ("A B^T" form) / otherwise [k][n]
//   KLIM 0: full k range; 1: k < 8 (nb + 1)  (B^T with B lower triangular); 2: k < 8 (mb + 1)  (A lower triangular);
//        3: k >= 8 nb  (B stored [k][n], lower triangular)
//   LOWER: only tiles with nb <= mb
template <int MAP, bool BT, int KLIM, bool LOWER>
__device__ __forceinline__ void tile2_prod(const double* __restrict__ A, const double* __restrict__ B, double (&acc)[8][2]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, r = lane >> 2, kk = lane & 3;
#pragma unroll
  for (int t = 0; t < 8; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
#pragma unroll 2
  for (int k4 = 0; k4 < 16; ++k4) {
    const int k = 4 * k4;
    if (MAP == 0) {
      const int mb = w;
      if (KLIM == 2 && k >= 8 * (mb + 1)) break;
      const double a = A[(8 * mb + r) * T2LD + k + kk];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        if (LOWER && nb > mb) continue;
        if (KLIM == 1 && k >= 8 * (nb + 1)) continue;
        if (KLIM == 3 && k + 4 <= 8 * nb) continue;
        const double b = BT ? B[(8 * nb + r) * T2LD + k + kk] : B[(k + kk) * T2LD + 8 * nb + r];
        dmma884(acc[nb][0], acc[nb][1], a, b);
      }
    } else {
      const int nb = w;
      if (KLIM == 1 && k >= 8 * (nb + 1)) break;
      if (KLIM == 3 && k + 4 <= 8 * nb) continue;
      const double b = BT ? B[(8 * nb + r) * T2LD + k + kk] : B[(k + kk) * T2LD + 8 * nb + r];
#pragma unroll
      for (int mb = 0; mb < 8; ++mb) {
        if (LOWER && nb > mb) continue;
        if (KLIM == 2 && k >= 8 * (mb + 1)) continue;
        const double a = A[(8 * mb + r) * T2LD + k + kk];
        dmma884(acc[mb][0], acc[mb][1], a, b);
      }
    }
  }
}
// Lower triangle of L L^T (36 of the 64 8 x 8 tiles), balanced: tile id = warp + 8 s (row-major enumeration of the lower
// tiles), 4 or 5 tiles per warp instead of 1..8 with the row-owner mapping, so the diagonal-tile update on the critical path
// costs 72-80 DMMAs per warp instead of 128.
struct SyrkTiles { int row[5], col[5]; bool live[5]; };
__device__ __forceinline__ SyrkTiles syrk_tiles() {
  SyrkTiles tl;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int s_ = 0; s_ < 5; ++s_) {
    const int id = w + 8 * s_;
    int mb = 0;
    while ((mb + 1) * (mb + 2) / 2 <= id) ++mb;
    const int nb = id - mb * (mb + 1) / 2;
    tl.live[s_] = id < 36;
    tl.row[s_] = 8 * mb + (lane >> 2);     // accumulator row of this lane; also the A-fragment row
    tl.col[s_] = 8 * nb;                   // first column of the tile (B-fragment row block)
  }
  return tl;
}
__device__ __forceinline__ void tile2_syrk_lower(const double* __restrict__ Lm, const SyrkTiles& tl, double (&acc)[5][2]) {
  const int lane = threadIdx.x & 31, r = lane >> 2, kk = lane & 3;
#pragma unroll
  for (int s_ = 0; s_ < 5; ++s_) { acc[s_][0] = 0.0; acc[s_][1] = 0.0; }
#pragma unroll 2
  for (int k = 0; k < TNB; k += 4) {
#pragma unroll
    for (int s_ = 0; s_ < 5; ++s_) {
      if (!tl.live[s_]) continue;
      const double a = Lm[tl.row[s_] * T2LD + k + kk];
      const double b = Lm[(tl.col[s_] + r) * T2LD + k + kk];
      dmma884(acc[s_][0], acc[s_][1], a, b);
    }
  }
}

// global / shared address of this lane's two accumulator elements of tile t
template <int MAP>
__device__ __forceinline__ int acc_row(int t) { return 8 * (MAP == 0 ? (int)(threadIdx.x >> 5) : t) + ((threadIdx.x & 31) >> 2); }
template <int MAP>
__device__ __forceinline__ int acc_col(int t) { return 8 * (MAP == 0 ? t : (int)(threadIdx.x >> 5)) + 2 * (threadIdx.x & 3); }
template <int MAP>
__device__ __forceinline__ void acc2_to_smem(double* s, const double (&acc)[8][2]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) *reinterpret_cast<double2*>(s + acc_row<MAP>(t) * T2LD + acc_col<MAP>(t)) = make_double2(acc[t][0], acc[t][1]);
}

// ---- Cholesky + inverse of one 64 x 64 tile, by 16-column panels -------------------------------------------------
// sa: [64][T2LD] SPD tile (lower triangle read, destroyed).  sx: [64][T2LD] work tile -> X = chol(sa)^-1 (lower, zeros above).
// sl: [64][T2SL] panel strip.  vec: col[2][16], rs[16], dvals[64].  Result also written to Xg (ld ldx) and densely to Dg.
//
// Warp-specialised panel pipeline: warp 0 runs the pivot chain of panel J while warps 1-7 finish the trailing update (c)
// of panel J-1 (the three 8 x 8 tiles of the next diagonal block first, handed over through named barrier 1) and stream
// the finished rows of X to global memory, so only the chain and the small (b) phase are on the critical path.
// ABL (measurement only, results wrong when != 0): bit 0 skip the pivot chain, 1 skip (b), 2 skip (c), 3 skip the result stores
__device__ __forceinline__ void bar_arrive1() { asm volatile("bar.arrive 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_sync1() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// rows [r0, r0+16) of X (lower triangular) from sx to the two global destinations; nthr threads, thread index tt
__device__ __forceinline__ void x_rows_store(const double* sx, int r0, int tt, int nthr, double* __restrict__ Xg, int64_t ldx, double* __restrict__ Dg) {
  for (int e = tt; e < 16 * 32; e += nthr) {
    const int i = r0 + (e >> 5), c = (e & 31) * 2;
    double2 v = *reinterpret_cast<const double2*>(sx + i * T2LD + c);
    if (c > i) v.x = 0.0;
    if (c + 1 > i) v.y = 0.0;
    *reinterpret_cast<double2*>(Xg + (int64_t)i * ldx + c) = v;
    *reinterpret_cast<double2*>(Dg + i * TNB + c) = v;
  }
}

// ---- the 16-pivot chain of one panel: two implementations (CHAIN 0: scalar chain with the reciprocal started one pivot ahead, the
// round-1 version; CHAIN 1: next column kept locally + seeded reciprocals) ----
__device__ __forceinline__ void chain16_scalar(double* sa, double* sx, double* vec, const int c0, const int lane, int* __restrict__ status) {
  double* col = vec;            // [2][20]  (16 column entries + the next diagonal element)
  double* dvals = vec + 48;     // [64]
      const int rr = lane & 15;
      const bool arow = lane < 16;
      bool bad = false;
      double x[16], rsv[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int hi_ = rr > q ? rr : q, lo_ = rr > q ? q : rr;
        x[q] = arow ? sa[(c0 + hi_) * T2LD + c0 + lo_] : (q == rr ? 1.0 : 0.0);  // row rr of the block (symmetric) | column rr of W
      }
      // Two interleaved reciprocal chains: 1/d_{j+1} = d_j / (a_{j+1,j+1} d_j - a_{j+1,j}^2) only needs the state BEFORE pivot j,
      // so the reciprocal for pivot j+1 is started one pivot ahead and a pivot costs max(exchange, rcp / 2) instead of rcp.
      double dprev = 0.0, Rprev = 0.0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double* cj = col + (j & 1) * 20;
        if (arow) cj[rr] = x[j];                                  // column j of the block: a_ij from the lane that holds row i
        if (j < 15 && arow && rr == j + 1) cj[16] = x[j + 1];     // a_{j+1,j+1} before pivot j
        __syncwarp();
        double d = cj[j];
        if (!(d > 0.0)) { bad = true; d = 1.0; }                  // PosDefException is raised after the chain (no branch on the chain)
        const double inv = (j == 0) ? rcp_chain<1>(d) : dprev * Rprev;   // MUFU.RCP64H + one Newton step: relative error ~1e-14
        if (j < 15) {
          const double q1 = cj[j + 1], p1 = cj[16];
          Rprev = rcp_chain<1>(fma(p1, d, -q1 * q1));
          dprev = d;
        }
        double g = -x[j] * inv;                  // rows: -a_rj / d_j ; W columns: -w_jr / d_j
        if (arow && rr <= j) g = 0.0;            // rows at or above the pivot are not touched
#pragma unroll
        for (int q = j + 1; q < 16; ++q) x[q] = fma(cj[q], g, x[q]);   // a_rq -= a_rj a_qj / d | w_qr -= a_qj w_jr / d
        rsv[j] = d;
        if (lane == j) dvals[c0 + j] = d;
      }
      if (bad && lane == 0) atomicOr(status, ST_NOT_POSDEF);
      // d^-1/2 for the row scaling, branch-free (library rsqrt() carries a slow-path branch per call, and basic-block
      // boundaries inside the pivot loop stop the scheduler from interleaving the chains): MUFU.RSQ64H + two Newton steps
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(rsv[q]));
        double e = fma(-rsv[q] * y, y, 1.0);
        y = fma(0.5 * y, e, y);
        e = fma(-rsv[q] * y, y, 1.0);
        rsv[q] = fma(0.5 * y, e, y);
      }
      // X_JJ = diag(d)^-1/2 W  (rows of W scaled)
      if (!arow) {
#pragma unroll
        for (int q = 0; q < 16; ++q) sx[(c0 + q) * T2LD + c0 + rr] = (q >= rr) ? x[q] * rsv[q] : 0.0;
      }
}

__device__ __forceinline__ void chain16_lookahead(double* sa, double* sx, double* vec, const int c0, const int lane, int* __restrict__ status) {
  double* col = vec;            // [2][24]
  double* dvals = vec + 48;     // [64]
      const int rr = lane & 15;
      const bool arow = lane < 16;
      bool bad = false;
      double x[16], rsv[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int hi_ = rr > q ? rr : q, lo_ = rr > q ? q : rr;
        x[q] = arow ? sa[(c0 + hi_) * T2LD + c0 + lo_] : (q == rr ? 1.0 : 0.0);  // row rr of the block (symmetric) | column rr of W
      }
      // The pivot recurrence, as short as the arithmetic allows.  Two latencies used to sit on it -- the shared-memory round trip that
      // broadcasts column j+1 after pivot j has updated it (~100 cycles) and the reciprocal of the pivot (MUFU.RCP64H ~80) -- 126
      // cycles per pivot in total.  Now:
      //  * every lane keeps its own copy of the NEXT column and applies pivot j's rank-1 update to it itself (cA: column j, cN: column
      //    j+1, both in registers; 15 - j extra DFMAs per lane), so column j+1 is known without waiting for anybody; what travels
      //    through shared memory is column j+2 (published after pivot j, needed during pivot j+1): one whole pivot of slack;
      //  * 1 / d_{j+2} is SEEDED at pivot j from the leading 3 x 3 minor of the block as it stands then (det2 / det3; the raw
      //    MUFU.RCP64H result is refined by the next pivot, so its latency never stalls this in-order warp), and pivot j+2 spends one
      //    Newton step y <- y (2 - d y) on the true pivot: the seed's error (cancellation in det3) is squared, 1 / d is good to ~1e-15.
      // Recurrence per pivot: d = cA[j] -> Newton (2 DFMA) -> f = cA[j+1] / d (DMUL) -> cN[q] -= cA[q] f (DFMA).
      double cA[16], cN[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        cA[q] = sa[(c0 + q) * T2LD + c0];                                                   // column 0 (lower triangle of the tile)
        cN[q] = (q >= 1) ? sa[(c0 + q) * T2LD + c0 + 1] : sa[(c0 + 1) * T2LD + c0];          // column 1 (symmetric)
      }
      double a22n = sa[(c0 + 2) * T2LD + c0 + 2];    // diagonal entry j+2 of the current state (for the seed)
      double seed_cur = 0.0, seed_first = 0.0;       // seed of 1 / d_j; 1 / d_1 from the first pair
      double p_det2 = 1.0, p_det3 = 1.0, p_raw = 1.0;   // determinants / raw reciprocal issued by the previous pivot
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        double* pub = col + (j & 1) * 24;          // column j+2 (+ the diagonal entry j+3) of the state AFTER this pivot
        double d = cA[j];
        if (!(d > 0.0)) { bad = true; d = 1.0; }                  // PosDefException is raised after the chain (no branch on the chain)
        double inv;
        if (j == 0) inv = rcp_chain<2>(d);
        else inv = fma(seed_cur, fma(-d, seed_cur, 1.0), seed_cur);
        // seed of the pivot after next, from the 3 x 3 minor (rows / columns j, j+1, j+2) of the state before this pivot
        double seed2 = 0.0;
        if (j >= 1 && j < 15) seed2 = p_det2 * fma(p_raw, fma(-p_det3, p_raw, 1.0), p_raw);      // 1 / d_{j+1}, issued by pivot j-1
        if (j < 15) {
          const double a01 = cA[j + 1], a11 = cN[j + 1];
          const double det2 = fma(a11, d, -a01 * a01);            // d_j d_{j+1}
          if (j == 0) seed_first = d * rcp_chain<1>(det2);
          if (j < 14) {
            const double a02 = cA[j + 2], a12 = cN[j + 2];
            const double t = fma(d * a12, a12, fma(-2.0 * a01 * a02, a12, a11 * a02 * a02));
            p_det2 = det2; p_det3 = fma(det2, a22n, -t);
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(p_raw) : "d"(p_det3));   // consumed by the NEXT pivot
          }
        }
        // column j+1 after this pivot, locally
        if (j < 15) {
          const double f = cA[j + 1] * inv;
#pragma unroll
          for (int q = j + 1; q < 16; ++q) cN[q] = fma(-cA[q], f, cN[q]);
        }
        // own row (A lanes) / own column of W (W lanes)
        double g = -x[j] * inv;                  // rows: -a_rj / d_j ; W columns: -w_jr / d_j
        if (arow && rr <= j) g = 0.0;            // rows at or above the pivot are not touched
#pragma unroll
        for (int q = j + 1; q < 16; ++q) x[q] = fma(cA[q], g, x[q]);   // a_rq -= a_rj a_qj / d | w_qr -= a_qj w_jr / d
        // publish column j+2 of the new state (and the diagonal entry j+3) for the pivot after next
        if (j < 14) {
          if (arow) pub[rr] = x[j + 2];
          if (j < 13 && arow && rr == j + 3) pub[16] = x[j + 3];
        }
        rsv[j] = d;
        if (lane == j) dvals[c0 + j] = d;
        __syncwarp();
        // rotate: column j+1 becomes the pivot column; fetch column j+2 (published just now; its consumers are the NEXT pivot's updates)
        seed_cur = (j == 0) ? seed_first : seed2;
#pragma unroll
        for (int q = 0; q < 16; ++q) cA[q] = cN[q];
        if (j < 14) {
#pragma unroll
          for (int q = j + 2; q < 16; ++q) cN[q] = pub[q];
          a22n = (j < 13) ? pub[16] : 0.0;
        }
      }
      if (bad && lane == 0) atomicOr(status, ST_NOT_POSDEF);
      // d^-1/2 for the row scaling, branch-free (library rsqrt() carries a slow-path branch per call, and basic-block
      // boundaries inside the pivot loop stop the scheduler from interleaving the chains): MUFU.RSQ64H + two Newton steps
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(rsv[q]));
        double e = fma(-rsv[q] * y, y, 1.0);
        y = fma(0.5 * y, e, y);
        e = fma(-rsv[q] * y, y, 1.0);
        rsv[q] = fma(0.5 * y, e, y);
      }
      // X_JJ = diag(d)^-1/2 W  (rows of W scaled)
      if (!arow) {
#pragma unroll
        for (int q = 0; q < 16; ++q) sx[(c0 + q) * T2LD + c0 + rr] = (q >= rr) ? x[q] * rsv[q] : 0.0;
      }
}

template <int ABL = 0, int CHAIN = 1>
__device__ __forceinline__ void tile2_potf2_inv(double* sa, double* sx, double* sl, double* vec, double* __restrict__ Xg, int64_t ldx,
                                                double* __restrict__ Dg, double* __restrict__ logdet, int* __restrict__ status) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int r = lane >> 2, kk = lane & 3;
  double* col = vec;            // [2][24]  (16 column entries + three entries of the next two rows, see the seeds below)
  double* dvals = vec + 48;     // [64]
#pragma unroll 1
  for (int J = 0; J < 4; ++J) {
    const int c0 = 16 * J, nbelow = TNB - c0 - 16;
    if (w == 0) {
      // ---- (a) pivot chain on the 16 x 16 diagonal block ---------------------------------------------------------
      if (J > 0) bar_sync1();                    // the diagonal block has received the trailing update of panel J-1
      if (!(ABL & 1)) {
      if (CHAIN == 0) chain16_scalar(sa, sx, vec, c0, lane, status);
      else chain16_lookahead(sa, sx, vec, c0, lane, status);
      }
    } else {
      const int tt = t - 32;
      if (J == 0) {
        // T starts as the identity; the 16 x 16 diagonal blocks are written whole by the chains and never read before
        for (int e = tt; e < TNB * TNB; e += TAIL_THREADS - 32) {
          const int i = e >> 6, c = e & 63;
          if ((i >> 4) != (c >> 4)) sx[i * T2LD + c] = 0.0;
        }
      } else {
        // ---- (c) of panel J-1: trailing A -= L L^T (lower 8 x 8 tiles) and T_I,0:c0 -= L_I,J-1 X_J-1,0:c0 for the rows below ----
        const int cp = c0 - 16, nbel = TNB - c0;          // previous panel start, rows below it
        const int nb8 = nbel / 8, n1 = nb8 * (nb8 + 1) / 2, ncb = c0 / 8, n2 = nb8 * ncb;
        bool arrived = false;
        if (w > 3) { bar_arrive1(); arrived = true; }     // only warps 1-3 touch the next diagonal block (tiles 0..2)
        if (!(ABL & 4)) {
#pragma unroll 1
        for (int id = w - 1; id < n1 + n2; id += 7) {
          double a0 = 0.0, a1 = 0.0;
          if (id < n1) {
            int mb = 0;
            while ((mb + 1) * (mb + 2) / 2 <= id) ++mb;
            const int nb = id - mb * (mb + 1) / 2;
            const double* Ap = sl + (8 * mb + r) * T2SL;
            const double* Bp = sl + (8 * nb + r) * T2SL;
#pragma unroll
            for (int k = 0; k < 16; k += 4) dmma884(a0, a1, Ap[k + kk], Bp[k + kk]);
            double2* dst = reinterpret_cast<double2*>(sa + (c0 + 8 * mb + r) * T2LD + c0 + 8 * nb + 2 * kk);
            double2 o = *dst; o.x -= a0; o.y -= a1; *dst = o;
          } else {
            const int id2 = id - n1, mb = id2 / ncb, nb = id2 % ncb;
            const double* Ap = sl + (8 * mb + r) * T2SL;
            const double* Bp = sx + cp * T2LD + 8 * nb + r;          // X_J-1[k][c]
#pragma unroll
            for (int k = 0; k < 16; k += 4) dmma884(a0, a1, Ap[k + kk], Bp[(k + kk) * T2LD]);
            double2* dst = reinterpret_cast<double2*>(sx + (c0 + 8 * mb + r) * T2LD + 8 * nb + 2 * kk);
            double2 o = *dst; o.x -= a0; o.y -= a1; *dst = o;
          }
          if (!arrived) { bar_arrive1(); arrived = true; }   // tiles 0..2 (first round of warps 1-3) are the next diagonal block
        }
        }
        if (!arrived) bar_arrive1();
        // rows of panel J-1 of X are final: stream them out while the chain runs
        if (!(ABL & 8)) x_rows_store(sx, cp, tt, TAIL_THREADS - 32, Xg, ldx, Dg);
      }
    }
    __syncthreads();
    // ---- (b) strip L_IJ = A_IJ X_JJ^T (rows below, -> sl) and X_J,0:c0 = X_JJ T_J,0:c0 (in registers until the barrier) ----
    // 8 x 8 output tiles, K = 16: strip tiles (nbelow/8) x 2, row-block tiles 2 x (c0/8); at most two tiles per warp
    {
      const int n1 = (nbelow / 8) * 2, n2 = 2 * (c0 / 8);
      double keep[2][2]; int keep_off[2];
#pragma unroll
      for (int s_ = 0; s_ < 2; ++s_) {
        const int id = w + 8 * s_;
        keep_off[s_] = -1;
        if (id >= n1 + n2 || (ABL & 2)) continue;
        double a0 = 0.0, a1 = 0.0;
        if (id < n1) {
          const int mb = id >> 1, nb = id & 1;           // rows c0+16+8mb.., strip columns 8nb..
          const double* Ap = sa + (c0 + 16 + 8 * mb + r) * T2LD + c0;
          const double* Bp = sx + (c0 + 8 * nb + r) * T2LD + c0;   // X_JJ[c][k], "B^T" form, k <= c
#pragma unroll
          for (int k = 0; k < 16; k += 4) if (k < 8 * (nb + 1)) dmma884(a0, a1, Ap[k + kk], Bp[k + kk]);
          *reinterpret_cast<double2*>(sl + (8 * mb + r) * T2SL + 8 * nb + 2 * kk) = make_double2(a0, a1);
        } else {
          const int id2 = id - n1, mb = id2 / (c0 / 8), nb = id2 % (c0 / 8);   // rows c0+8mb.., columns 8nb..
          const double* Ap = sx + (c0 + 8 * mb + r) * T2LD + c0;  // X_JJ[r][k], k <= r
          const double* Bp = sx + c0 * T2LD + 8 * nb + r;          // T_J[k][c]
#pragma unroll
          for (int k = 0; k < 16; k += 4) if (k < 8 * (mb + 1)) dmma884(a0, a1, Ap[k + kk], Bp[(k + kk) * T2LD]);
          keep[s_][0] = a0; keep[s_][1] = a1; keep_off[s_] = (c0 + 8 * mb + r) * T2LD + 8 * nb + 2 * kk;
        }
      }
      if (c0 > 0) {
        __syncthreads();   // every reader of T_J is done: overwrite it with X_J,0:c0
#pragma unroll
        for (int s_ = 0; s_ < 2; ++s_)
          if (keep_off[s_] >= 0) *reinterpret_cast<double2*>(sx + keep_off[s_]) = make_double2(keep[s_][0], keep[s_][1]);
      }
    }
    __syncthreads();
  }
  // ---- last row block of X, logdet ----
  if (!(ABL & 8)) x_rows_store(sx, 48, t, TAIL_THREADS, Xg, ldx, Dg);
  if (t < TNB) {
    double l = warp_sum(log(dvals[t]));
    if ((t & 31) == 0) atomicAdd(logdet, l);
  }
}

template <int ABL = 0, int CHAIN = 1>
__global__ void __launch_bounds__(TAIL_THREADS, 1) tail2_potf2_first_kernel(const TailStepParams p) {
  extern __shared__ double sm[];
  pdl_launch_dependents();   // let the next block step get resident while this one runs; it waits in pdl_wait()
  if (p.early_flag) {
    // Start on tile (0, 0) while the natural-parameter update is still writing the other tiles of P_v: its 64 blocks of rows 0..63 /
    // column block 0 release-count themselves in early_flag (combine_kernel).  The spin is bounded (~0.5 ms): on a timeout the full
    // programmatic dependency below makes the tile valid anyway, so the counter only ever shortens the wait.
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      int v;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p.early_flag) : "memory");
      } while (v < 64 && clock64() - t0 < 1000000LL);
      if (v < 64) pdl_wait();
      *p.early_flag = 0;     // every producer of this step has arrived (or finished): ready for the next step
    }
    __syncthreads();
  } else {
    pdl_wait();
  }
  tile2_load(sm, p.P, p.ld);
  __syncthreads();
  tile2_potf2_inv<ABL, CHAIN>(sm, sm + T2_TILE, sm + 5 * T2_TILE, sm + 5 * T2_TILE + TNB * T2SL, p.Xout, p.ld, p.Dinv, p.logdet, p.status);
  // the block steps that follow read every tile of P_v: this grid must not complete before the producer grid has (their programmatic
  // dependency is on THIS kernel only)
  if (p.early_flag) pdl_wait();
}

template <int CHAIN = 1>
__global__ void __launch_bounds__(TAIL_THREADS, 1) tail2_step_kernel(const TailStepParams p) {
  extern __shared__ double sm[];
  double* sX = sm;                  // X_kk
  double* s1 = sm + 1 * T2_TILE;    // A_ik
  double* s2 = sm + 2 * T2_TILE;    // A_jk / W_kc
  double* s3 = sm + 3 * T2_TILE;    // L_ik
  double* s4 = sm + 4 * T2_TILE;    // L_jk / Wn_kc
  double* sl = sm + 5 * T2_TILE;
  double* vec = sl + TNB * T2SL;
  const int k = p.k, nblk = p.nblk, r = nblk - 1 - k;
  const int nA = r * (r + 1) / 2, nW = r * (k + 1);
  int b = blockIdx.x;
  const int64_t ld = p.ld;
  double acc[8][2];

  pdl_launch_dependents();
  pdl_wait();                // everything below reads what the previous block step wrote
  tile2_load(sX, p.Dinv + (int64_t)k * TNB * TNB, TNB);

  if (b < nA) {
    // ---- A tile (i, j), k < j <= i :  A_ij -= L_ik L_jk^T,  L_ik = A_ik X_kk^T ----
    int ii = 0;
    while ((ii + 1) * (ii + 2) / 2 <= b) ++ii;
    int jj = b - ii * (ii + 1) / 2;
    const int i = k + 1 + ii, j = k + 1 + jj;
    double* At = p.P + (int64_t)i * TNB * ld + (int64_t)j * TNB;
    tile2_load(s1, p.P + (int64_t)i * TNB * ld + (int64_t)k * TNB, ld);
    if (j != i) tile2_load(s2, p.P + (int64_t)j * TNB * ld + (int64_t)k * TNB, ld);
    const bool lookahead = (i == k + 1 && j == k + 1);
    if (j != i) {
      double2 old[8];   // the tile being updated, prefetched in the accumulator layout (MAP 0)
#pragma unroll
      for (int t = 0; t < 8; ++t) old[t] = *reinterpret_cast<const double2*>(At + (int64_t)acc_row<0>(t) * ld + acc_col<0>(t));
      __syncthreads();
      tile2_prod<0, true, 1, false>(s1, sX, acc); acc2_to_smem<0>(s3, acc);
      tile2_prod<0, true, 1, false>(s2, sX, acc); acc2_to_smem<0>(s4, acc);
      __syncthreads();
      tile2_prod<0, true, 0, false>(s3, s4, acc);
#pragma unroll
      for (int t = 0; t < 8; ++t)
        *reinterpret_cast<double2*>(At + (int64_t)acc_row<0>(t) * ld + acc_col<0>(t)) = make_double2(old[t].x - acc[t][0], old[t].y - acc[t][1]);
    } else {
      // diagonal tile: lower 8 x 8 tiles only, balanced over the warps (this is the critical path of the block step)
      const SyrkTiles tl = syrk_tiles();
      const int c2 = 2 * (threadIdx.x & 3);
      double2 old5[5];
      double acc5[5][2];
#pragma unroll
      for (int s_ = 0; s_ < 5; ++s_)
        old5[s_] = tl.live[s_] ? *reinterpret_cast<const double2*>(At + (int64_t)tl.row[s_] * ld + tl.col[s_] + c2) : make_double2(0.0, 0.0);
      __syncthreads();
      tile2_prod<0, true, 1, false>(s1, sX, acc); acc2_to_smem<0>(s3, acc);
      __syncthreads();
      tile2_syrk_lower(s3, tl, acc5);
      if (!lookahead) {
#pragma unroll
        for (int s_ = 0; s_ < 5; ++s_)
          if (tl.live[s_]) *reinterpret_cast<double2*>(At + (int64_t)tl.row[s_] * ld + tl.col[s_] + c2) = make_double2(old5[s_].x - acc5[s_][0], old5[s_].y - acc5[s_][1]);
      } else {
#pragma unroll
        for (int s_ = 0; s_ < 5; ++s_)
          if (tl.live[s_]) *reinterpret_cast<double2*>(s1 + tl.row[s_] * T2LD + tl.col[s_] + c2) = make_double2(old5[s_].x - acc5[s_][0], old5[s_].y - acc5[s_][1]);
        __syncthreads();
        tile2_potf2_inv<0, CHAIN>(s1, s2, sl, vec, p.Xout + (int64_t)(k + 1) * TNB * (ld + 1), ld, p.Dinv + (int64_t)(k + 1) * TNB * TNB, p.logdet, p.status);
      }
    }
  } else if (b < nA + nW) {
    // ---- W tile (i, c), c <= k < i :  W_ic = [c<k] W_ic - L_ik Wn_kc,  Wn_kc = X_kk W_kc (c<k) or X_kk (c=k) ----
    b -= nA;
    const int i = k + 1 + b / (k + 1), c = b % (k + 1);
    double* Wt = p.W + (int64_t)i * TNB * ld + (int64_t)c * TNB;
    tile2_load(s1, p.P + (int64_t)i * TNB * ld + (int64_t)k * TNB, ld);
    if (c < k) tile2_load(s2, p.W + (int64_t)k * TNB * ld + (int64_t)c * TNB, ld);
    double2 old[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) old[t] = (c < k) ? *reinterpret_cast<const double2*>(Wt + (int64_t)acc_row<0>(t) * ld + acc_col<0>(t)) : make_double2(0.0, 0.0);
    __syncthreads();
    tile2_prod<0, true, 1, false>(s1, sX, acc); acc2_to_smem<0>(s3, acc);                        // L_ik
    if (c < k) { tile2_prod<1, false, 2, false>(sX, s2, acc); acc2_to_smem<1>(s4, acc); }      // Wn_kc = X_kk W_kc
    __syncthreads();
    if (c < k) tile2_prod<0, false, 0, false>(s3, s4, acc);
    else tile2_prod<0, false, 3, false>(s3, sX, acc);                                            // Wn = X_kk (lower triangular)
#pragma unroll
    for (int t = 0; t < 8; ++t)
      *reinterpret_cast<double2*>(Wt + (int64_t)acc_row<0>(t) * ld + acc_col<0>(t)) = make_double2(old[t].x - acc[t][0], old[t].y - acc[t][1]);
  } else {
    // ---- F tile (k, c), c < k : final row block of X ----
    const int c = b - nA - nW;
    tile2_load(s2, p.W + (int64_t)k * TNB * ld + (int64_t)c * TNB, ld);
    __syncthreads();
    tile2_prod<1, false, 2, false>(sX, s2, acc);
    double* Xt = p.Xout + (int64_t)k * TNB * ld + (int64_t)c * TNB;
#pragma unroll
    for (int t = 0; t < 8; ++t) *reinterpret_cast<double2*>(Xt + (int64_t)acc_row<1>(t) * ld + acc_col<1>(t)) = make_double2(acc[t][0], acc[t][1]);
  }
}

}  // namespace agp
