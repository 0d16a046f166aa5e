// agp_umma.cu -- tcgen05 (5th-gen tensor core) path of the B x m contractions, written for sm_100a.
//
//   C[M x N] = A[M x K] * B[N x K]^T  in "3xTF32": each fp32 operand tile is split inside the CTA into
//   hi = rna_tf32(a) and lo = rna_tf32(a - hi), and the tensor cores accumulate lo*hi + hi*lo + hi*hi in fp32 TMEM
//   (error ~2^-22 per product, fp32 class).
//
// One CTA computes one 128 x 128 output tile (UMMA 128x128x8, kind::tf32, cta_group::1), 192 threads:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle): the RAW fp32 A and B tiles of one 32-wide
//                 k-block (2 x 16 KB) per stage -- operands cross L2 -> SM once, not as separate hi and lo copies
//                 (the first version streamed pre-split operands and was L2-bandwidth bound: 168 MB / launch)
//   warps 2..9  : two converter groups (alternating k-blocks): raw tile -> hi / lo tiles at the same (swizzled) offsets, fence.proxy.async, then
//                 hand the stage to the MMA warp; after the main loop the same warps run the epilogue
//                 (tcgen05.ld 32x32b -> registers -> fp32 row-major global stores)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane, 12 tcgen05.mma per k-block, tcgen05.commit
//                 releases the hi/lo buffers / signals the epilogue)
// 2-stage ring, 96 KB per stage (raw A, raw B, A_hi, A_lo, B_hi, B_lo).
// Used for V = Knm L^-T and V X^T (k-blocks above the diagonal of the lower-triangular right operand are skipped) and
// the split-K Gram product U^T U with U = diag(sqrt(rho w)) V (upper-triangular tiles only).
#include "agp_umma.h"
#include "agp_tc.cuh"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace agp {

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
#ifndef AGP_UMMA_RS
#define AGP_UMMA_RS 4
#endif
#ifndef AGP_UMMA_CS
#define AGP_UMMA_CS 2
#endif
constexpr int RS = AGP_UMMA_RS;                // raw (TMA) ring depth: prefetch distance of 4 k-blocks
constexpr int CS = AGP_UMMA_CS;                // converted-operand ring depth
constexpr int TILE_BYTES = BM * BK * 4;        // 16 KB : 128 rows x 128 B (one 128B-swizzle atom wide)
constexpr int RAW_BYTES = 2 * TILE_BYTES;      // raw A, raw B
constexpr int CONV_BYTES = 2 * TILE_BYTES;     // B_hi, B_lo   (A_hi / A_lo live in TMEM)
constexpr int RING_BYTES = RS * RAW_BYTES + CS * CONV_BYTES;
constexpr int NUM_CONV_THREADS = 128;
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;     // two fp32 accumulators [0,128) [128,256), then per converted slot 32 columns A_hi + 32 columns A_lo
constexpr int TMEM_A0 = 256;
constexpr int NUM_CONV_GROUPS = 2;   // worker groups (4 warps each): group g converts AND drains the CTA's tiles lt = g (mod 2)
constexpr int NUM_THREADS = 64 + NUM_CONV_GROUPS * 128;

using namespace tc;
constexpr uint32_t kIdesc = make_idesc_tf32(BM, BN);

// Work decomposition of one launch.  A work unit = one 128 x 128 output tile (x one split-K slice).
//   tri_mode 0: all tiles | 1: B operand lower-triangular (B[n][k] = 0 for k > n): the k-extent stops at the tile's last
//   column, units are ordered longest first | 2: symmetric output, only tiles with tile_n >= tile_m (x splits)
struct GemmWork {
  int ntm, ntn, nsplit, total;     // tiles along M, N, split-K slices, number of units
  int total_kb, kb_per_split, tri_mode;
  int u0;                          // first unit of this launch in the unit enumeration (launches restricted to a range of N tiles)
};
struct WorkUnit { int tile_m, tile_n, split, kb0, nkb; };

__device__ __forceinline__ WorkUnit get_unit(const GemmWork& w, int u) {
  WorkUnit r;
  u += w.u0;
  if (w.tri_mode == 2) {
    r.split = u % w.nsplit;
    int p = u / w.nsplit;                      // p-th upper tile, row-major over (tile_m <= tile_n)
    int tm = 0;
    while (p >= w.ntn - tm) { p -= w.ntn - tm; ++tm; }
    r.tile_m = tm; r.tile_n = tm + p;
  } else if (w.tri_mode == 1) {
    r.split = 0;
    r.tile_n = w.ntn - 1 - u / w.ntm;          // longest k-extent first
    r.tile_m = u % w.ntm;
  } else {
    r.split = 0;
    r.tile_m = u / w.ntn; r.tile_n = u % w.ntn;
  }
  r.kb0 = r.split * w.kb_per_split;
  int kb1 = min(w.total_kb, r.kb0 + w.kb_per_split);
  if (w.tri_mode == 1) kb1 = min(kb1, (r.tile_n * BN + BN) / BK);
  r.nkb = max(kb1 - r.kb0, 0);
  return r;
}
// persistent CTAs take units in "snake" order (round r: CTA c takes r*G + c, or r*G + G-1-c on odd rounds), which pairs
// long units with short ones when the list is sorted by length
__device__ __forceinline__ int unit_index(int round, int cta, int G) { return round * G + ((round & 1) ? (G - 1 - cta) : cta); }

// per-role timeline instrumentation for profiles/microbench/gemm_trace.cu (compiled in only with -DAGP_UMMA_TRACE):
// role 0 = TMA thread, 1 = MMA thread, 2 / 3 = first thread of worker group 0 / 1; records (tag, clock64)
#ifdef AGP_UMMA_TRACE
constexpr int UT_SLOTS = 400;
__device__ unsigned long long agp_ut_trace[160 * 4 * UT_SLOTS * 2];
__device__ int agp_ut_n[160 * 4];
__device__ __forceinline__ void ut_trace(int role, int tag) {
  const int n = agp_ut_n[blockIdx.x * 4 + role]++;
  if (n < UT_SLOTS) {
    unsigned long long* r = agp_ut_trace + (((size_t)blockIdx.x * 4 + role) * UT_SLOTS + n) * 2;
    r[0] = (unsigned long long)tag; r[1] = (unsigned long long)clock64();
  }
}
#define UTT(role, tag) ut_trace(role, tag)
#else
#define UTT(role, tag) do { } while (0)
#endif

__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
                    int64_t ldc, int64_t c_split_stride, const GemmWork work, const UmmaEpilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RS + s); };
  auto conv_full = [&](int s) { return bars + 8u * (2 * RS + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RS + CS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + 2 + b); };
  auto unit_conv_done = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + 4 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + RING_BYTES + 8 * (2 * RS + 2 * CS + 6));
  const uint32_t conv_base = smem_base + RS * RAW_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    for (int s = 0; s < RS; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), NUM_CONV_THREADS); }
    for (int s = 0; s < CS; ++s) { mbar_init(conv_full(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); mbar_init(unit_conv_done(b), NUM_CONV_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel of the chain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) UTT(0, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: raw fp32 tiles, one ring across all units of this CTA =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= work.total) break;
        const WorkUnit wu = get_unit(work, u);
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int s = g % RS;
          mbar_wait(raw_empty(s), ((g / RS) & 1) ^ 1);
          UTT(0, 1000 + g);
          const uint32_t dst = smem_base + s * RAW_BYTES;
          mbar_expect_tx(raw_full(s), 2 * TILE_BYTES);
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, &tmA, raw_full(s), k, wu.tile_m * BM);
          tma_load_2d(dst + 1 * TILE_BYTES, &tmB, raw_full(s), k, wu.tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator buffer lt & 1, so the epilogue of unit lt overlaps the main loop of unit lt + 1 =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      const int ab = lt & 1;
      mbar_wait(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);      // the epilogue two units ago has drained this accumulator
      tc_fence_after();
      if (lane == 0) UTT(1, 100000 + lt);
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS;
        mbar_wait(conv_full(s), (g / CS) & 1);
        tc_fence_after();
        if (lane == 0) UTT(1, 2000 + g);
        if (elect_one()) {
          const uint32_t b_hi = conv_base + s * CONV_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;  // 8 tf32 = 32 bytes along K inside the 128B swizzle atom (B); 8 TMEM columns (A)
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));                          // frees B_hi/B_lo and the TMEM A columns of this slot
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));   // accumulator complete
        }
        __syncwarp();
        if (lane == 0) UTT(1, 3000 + g);
      }
    }
  } else {
    // ===== worker groups: raw -> hi / lo conversion of the group's units, then their epilogue =====
    const int grp = (warp - 2) >> 2;
    const int ct = (threadIdx.x - 64) & 127;         // 0..127 within the group
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int arow = q * 32 + lane;                  // A-tile row = TMEM lane handled by this thread
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      if ((lt & 1) != grp) { g += wu.nkb; continue; }
      // A group only sees the mbarrier phases of its own units.  Parity waits are unambiguous only within one phase, so
      // do not start before the other group has issued every conversion of the previous unit (the TMA ring and the MMA
      // warp are then at most one phase behind on every barrier this group is about to wait on).
      if (lt > 0) mbar_wait(unit_conv_done(grp ^ 1), ((lt - 1) >> 1) & 1);
      if (ct == 0) UTT(2 + grp, 200000 + lt);
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int rs = g % RS, s = g % CS;
        mbar_wait(raw_full(rs), (g / RS) & 1);               // TMA landed the raw tiles
        if (ct == 0) UTT(2 + grp, 4000 + g);
        mbar_wait(mma_done(s), ((g / CS) & 1) ^ 1);          // previous MMAs on this slot's converted operands retired
        tc_fence_after();
        if (ct == 0) UTT(2 + grp, 5000 + g);
        uint8_t* base = smem_gen + rs * RAW_BYTES;
        uint8_t* cbase_s = smem_gen + RS * RAW_BYTES + s * CONV_BYTES;
        {
          // A: row `arow` of the raw tile (128 B, 16-byte chunks XOR-swizzled with row & 7) -> hi / lo -> TMEM
          const float4* rowp = reinterpret_cast<const float4*>(base + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        {
          // B: raw -> hi / lo tiles at the same (swizzled) offsets in shared memory
          const float4* raw = reinterpret_cast<const float4*>(base + 1 * TILE_BYTES);
          float4* hi = reinterpret_cast<float4*>(cbase_s);
          float4* lo = reinterpret_cast<float4*>(cbase_s + TILE_BYTES);
#pragma unroll
          for (int uu = 0; uu < TILE_BYTES / 16 / NUM_CONV_THREADS; ++uu) {
            const int e = ct + uu * NUM_CONV_THREADS;
            const float4 v = raw[e];
            float4 h, l;
            h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
            h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
            h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
            h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
            hi[e] = h; lo[e] = l;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(raw_empty(rs));   // the raw tiles may be overwritten by the next TMA
        mbar_arrive(conv_full(s));    // operands ready for the MMA warp
        if (ct == 0) UTT(2 + grp, 6000 + g);
      }
      mbar_arrive(unit_conv_done(grp));
      // ----- epilogue of this unit (the other group is already converting the next one) -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      float* cbase = C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq4[4] = {0.0, 0.0, 0.0, 0.0}, acc_dot4[4] = {0.0, 0.0, 0.0, 0.0};   // fused row statistics: fp64 sums, four interleaved chains each
      if (wu.nkb > 0) {
        mbar_wait(tmem_full(ab), (lt >> 1) & 1);
        tc_fence_after();
        if (ct == 0) UTT(2 + grp, 300000 + lt);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t rr[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
          TMEM_LD32(taddr, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c == BN / 32 - 1) {                            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
          }
          if (ep.mode != UMMA_EPI_STATS_ONLY) {
            float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          }
          if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const double v = (double)__uint_as_float(rr[j]);
              acc_sq4[j & 3] = fma(v, v, acc_sq4[j & 3]);
              if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot4[j & 3] = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + j], acc_dot4[j & 3]);
            }
          }
          if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
            // symmetric product: also write the transposed tile (lanes = consecutive addresses)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
          }
        }
        const double acc_sq = (acc_sq4[0] + acc_sq4[1]) + (acc_sq4[2] + acc_sq4[3]), acc_dot = (acc_dot4[0] + acc_dot4[1]) + (acc_dot4[2] + acc_dot4[3]);
        if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, acc_sq);
        if (ep.mode == UMMA_EPI_STATS_ONLY) { atomicAdd(ep.acc0 + row, acc_sq); atomicAdd(ep.acc1 + row, acc_dot); }
        if (ct == 0) UTT(2 + grp, 400000 + lt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) UTT(0, 2);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// Pre-split right operand (default for the two products whose B operand is an m x m matrix that changes at most once per step:
// V = K_nm L^-T and V X^T; AGP_UMMA_PS=0 selects the kernel above).  L^-1 is split into TF32 hi / lo planes once per agp_refresh_K, X by
// x_finalize_kernel; the planes arrive by TMA next to the raw A tile (48 KB per stage, 4 stages), so the worker groups only split A
// (32 instead of 64 values per thread and k-block, no shared-memory stores) and the MMA warp's commit releases the ring stage.
// Everything else (unit order, group alternation by unit, epilogues) is the kernel above.
namespace ps {
constexpr int RSP = 4;
constexpr int STAGE_BYTES = 3 * TILE_BYTES;                      // A raw | B hi | B lo
constexpr int RING_BYTES = RSP * STAGE_BYTES;                    // 192 KB
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
}
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm_ps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, float* __restrict__ C,
                    int64_t ldc, int64_t c_split_stride, const GemmWork work, const UmmaEpilogue ep, const UmmaRowFinish fin) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  using namespace ps;
  const uint32_t bars = smem_base + ps::RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };                 // TMA landed stage s (A raw, B hi, B lo)
  auto raw_empty = [&](int s) { return bars + 8u * (RSP + s); };        // the MMAs that read stage s have retired (tcgen05.commit)
  auto conv_full = [&](int s) { return bars + 8u * (2 * RSP + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RSP + CS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + 2 + b); };
  auto unit_conv_done = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + 4 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + ps::RING_BYTES + 8 * (2 * RSP + 2 * CS + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmBhi); tma_prefetch_desc(&tmBlo);
    for (int s = 0; s < RSP; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 1); }
    for (int s = 0; s < CS; ++s) { mbar_init(conv_full(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); mbar_init(unit_conv_done(b), NUM_CONV_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel of the chain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) UTT(0, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: raw fp32 tiles, one ring across all units of this CTA =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= work.total) break;
        const WorkUnit wu = get_unit(work, u);
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int s = g % RSP;
          mbar_wait(raw_empty(s), ((g / RSP) & 1) ^ 1);
          UTT(0, 1000 + g);
          const uint32_t dst = smem_base + s * STAGE_BYTES;
          mbar_expect_tx(raw_full(s), 3 * TILE_BYTES);
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, &tmA, raw_full(s), k, wu.tile_m * BM);
          tma_load_2d(dst + 1 * TILE_BYTES, &tmBhi, raw_full(s), k, wu.tile_n * BN);
          tma_load_2d(dst + 2 * TILE_BYTES, &tmBlo, raw_full(s), k, wu.tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator buffer lt & 1, so the epilogue of unit lt overlaps the main loop of unit lt + 1 =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      const int ab = lt & 1;
      mbar_wait(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);      // the epilogue two units ago has drained this accumulator
      tc_fence_after();
      if (lane == 0) UTT(1, 100000 + lt);
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS;
        mbar_wait(conv_full(s), (g / CS) & 1);
        tc_fence_after();
        if (lane == 0) UTT(1, 2000 + g);
        if (elect_one()) {
          const uint32_t b_hi = smem_base + (g % RSP) * STAGE_BYTES + TILE_BYTES, b_lo = b_hi + TILE_BYTES;   // pre-split B, straight from TMA
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;  // 8 tf32 = 32 bytes along K inside the 128B swizzle atom (B); 8 TMEM columns (A)
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));                          // frees the TMEM A columns of this slot
          tc_commit(raw_empty(g % RSP));                   // ... and the ring stage (A raw was consumed earlier, B hi / lo just now)
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));   // accumulator complete
        }
        __syncwarp();
        if (lane == 0) UTT(1, 3000 + g);
      }
    }
  } else {
    // ===== worker groups: raw -> hi / lo conversion of the group's units, then their epilogue =====
    const int grp = (warp - 2) >> 2;
    const int ct = (threadIdx.x - 64) & 127;         // 0..127 within the group
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int arow = q * 32 + lane;                  // A-tile row = TMEM lane handled by this thread
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      if ((lt & 1) != grp) { g += wu.nkb; continue; }
      // A group only sees the mbarrier phases of its own units.  Parity waits are unambiguous only within one phase, so
      // do not start before the other group has issued every conversion of the previous unit (the TMA ring and the MMA
      // warp are then at most one phase behind on every barrier this group is about to wait on).
      if (lt > 0) mbar_wait(unit_conv_done(grp ^ 1), ((lt - 1) >> 1) & 1);
      if (ct == 0) UTT(2 + grp, 200000 + lt);
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int rs = g % RSP, s = g % CS;
        mbar_wait(raw_full(rs), (g / RSP) & 1);              // TMA landed the stage
        if (ct == 0) UTT(2 + grp, 4000 + g);
        mbar_wait(mma_done(s), ((g / CS) & 1) ^ 1);          // previous MMAs on this slot's converted operands retired
        tc_fence_after();
        if (ct == 0) UTT(2 + grp, 5000 + g);
        uint8_t* base = smem_gen + rs * STAGE_BYTES;
        {
          // A: row `arow` of the raw tile (128 B, 16-byte chunks XOR-swizzled with row & 7) -> hi / lo -> TMEM
          const float4* rowp = reinterpret_cast<const float4*>(base + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(conv_full(s));    // A operand ready for the MMA warp (the ring stage is released by the MMA warp's commit)
        if (ct == 0) UTT(2 + grp, 6000 + g);
      }
      mbar_arrive(unit_conv_done(grp));
      // ----- epilogue of this unit (the other group is already converting the next one) -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      float* cbase = C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq4[4] = {0.0, 0.0, 0.0, 0.0}, acc_dot4[4] = {0.0, 0.0, 0.0, 0.0};   // fused row statistics: fp64 sums, four interleaved chains each
      if (wu.nkb > 0) {
        mbar_wait(tmem_full(ab), (lt >> 1) & 1);
        tc_fence_after();
        if (ct == 0) UTT(2 + grp, 300000 + lt);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t rr[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
          TMEM_LD32(taddr, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c == BN / 32 - 1) {                            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
          }
          if (ep.mode != UMMA_EPI_STATS_ONLY) {
            float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          }
          if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const double v = (double)__uint_as_float(rr[j]);
              acc_sq4[j & 3] = fma(v, v, acc_sq4[j & 3]);
              if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot4[j & 3] = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + j], acc_dot4[j & 3]);
            }
          }
          if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
            // symmetric product: also write the transposed tile (lanes = consecutive addresses)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
          }
        }
        const double acc_sq = (acc_sq4[0] + acc_sq4[1]) + (acc_sq4[2] + acc_sq4[3]), acc_dot = (acc_dot4[0] + acc_dot4[1]) + (acc_dot4[2] + acc_dot4[3]);
        if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, acc_sq);
        if (ep.mode == UMMA_EPI_STATS_ONLY) { atomicAdd(ep.acc0 + row, acc_sq); atomicAdd(ep.acc1 + row, acc_dot); }
        if (ep.mode == UMMA_EPI_STATS_ONLY && fin.cnt) {
          // fused row finish: the last of this row's N-tile units (all have nkb > 0 with a triangular right operand) completes the sample
          __threadfence();
          const unsigned prev = atomicAdd(fin.cnt + row, 1u);
          if (prev == (unsigned)work.ntn - 1u) {
            __threadfence();
            fin.cnt[row] = 0u;
            if (row < fin.B) {
              const double ssq = __ldcg(ep.acc0 + row), dvs = __ldcg(ep.acc1 + row);
              const double kt = fin.kdiag_jit - fin.sumsq_v[row];
              fin.Ktilde[row] = kt;
              if (!(kt > 0.0)) atomicOr(fin.status, ST_KTILDE);  // latentgp.jl:213
              const_cast<double*>(fin.lp.mean_f)[row] = dvs;
              const_cast<double*>(fin.lp.var_f)[row] = ssq + kt;
              double r0 = 0.0, r1 = 0.0;
              lik_update_sample(fin.lp, row, r0, r1);
            }
          }
        }
        if (ct == 0) UTT(2 + grp, 400000 + lt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) UTT(0, 2);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}



// One entry per latent GP of a grouped launch (device array): the operands' tensor maps and the epilogue targets.
struct alignas(64) GemmGroup {
  CUtensorMap tmA, tmB;
  float* C; double* acc0; double* acc1; const double* tvec;
  CUtensorMap tmBhi, tmBlo;    // pre-split planes of the right operand (umma_gemm_grouped_ps_kernel)
};
// work unit u of a grouped launch: `work` describes ONE group; units are ordered so that the longest k-extents of every group
// come first (tri_mode 1) and the groups interleave, which keeps the persistent CTAs balanced under the snake order
__device__ __forceinline__ WorkUnit get_unit_grouped(const GemmWork& w, int ngroups, int u, int& grp) {
  int ul;
  if (w.tri_mode == 1) {
    const int per = w.ntm * ngroups;
    const int p = u / per, rem = u - p * per;
    grp = rem / w.ntm;
    ul = p * w.ntm + (rem - grp * w.ntm);          // get_unit: tile_n = ntn - 1 - ul / ntm, tile_m = ul % ntm
  } else if (w.tri_mode == 2) {
    const int per = w.nsplit * ngroups;            // all splits of all groups' p-th upper tile are neighbours
    const int p = u / per, rem = u - p * per;
    grp = rem / w.nsplit;
    ul = p * w.nsplit + (rem - grp * w.nsplit);    // get_unit: split = ul % nsplit, tile = ul / nsplit
  } else {
    grp = u / w.total;
    ul = u - grp * w.total;
  }
  return get_unit(w, ul);
}

// Grouped variant of umma_gemm_nt_kernel: the same pipeline, one launch for the same-shaped products of several latent GPs
// (tensor maps and epilogue targets come from a device array instead of kernel parameters).  Launch-bound products (C2 / C4
// sized: ~8 us of fixed cost per launch against 0.65 us per k-block) become one long persistent loop.
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm_grouped_kernel(const GemmGroup* __restrict__ groups, const int ngroups, int64_t ldc, int64_t c_split_stride, const GemmWork work,
                         const int epi_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RS + s); };
  auto conv_full = [&](int s) { return bars + 8u * (2 * RS + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RS + CS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + 2 + b); };
  auto unit_conv_done = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + 4 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + RING_BYTES + 8 * (2 * RS + 2 * CS + 6));
  const uint32_t conv_base = smem_base + RS * RAW_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&groups[0].tmA); tma_prefetch_desc(&groups[0].tmB);
    for (int s = 0; s < RS; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), NUM_CONV_THREADS); }
    for (int s = 0; s < CS; ++s) { mbar_init(conv_full(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); mbar_init(unit_conv_done(b), NUM_CONV_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel of the chain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: raw fp32 tiles, one ring across all units of this CTA =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= work.total * ngroups) break;
        int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int s = g % RS;
          mbar_wait(raw_empty(s), ((g / RS) & 1) ^ 1);
          const uint32_t dst = smem_base + s * RAW_BYTES;
          mbar_expect_tx(raw_full(s), 2 * TILE_BYTES);
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, &groups[gq].tmA, raw_full(s), k, wu.tile_m * BM);
          tma_load_2d(dst + 1 * TILE_BYTES, &groups[gq].tmB, raw_full(s), k, wu.tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator buffer lt & 1, so the epilogue of unit lt overlaps the main loop of unit lt + 1 =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total * ngroups) break;
      int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
      const int ab = lt & 1;
      mbar_wait(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);      // the epilogue two units ago has drained this accumulator
      tc_fence_after();
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS;
        mbar_wait(conv_full(s), (g / CS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = conv_base + s * CONV_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;  // 8 tf32 = 32 bytes along K inside the 128B swizzle atom (B); 8 TMEM columns (A)
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));                          // frees B_hi/B_lo and the TMEM A columns of this slot
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===== worker groups: raw -> hi / lo conversion of the group's units, then their epilogue =====
    const int grp = (warp - 2) >> 2;
    const int ct = (threadIdx.x - 64) & 127;         // 0..127 within the group
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int arow = q * 32 + lane;                  // A-tile row = TMEM lane handled by this thread
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total * ngroups) break;
      int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
      if ((lt & 1) != grp) { g += wu.nkb; continue; }
      // A group only sees the mbarrier phases of its own units.  Parity waits are unambiguous only within one phase, so
      // do not start before the other group has issued every conversion of the previous unit (the TMA ring and the MMA
      // warp are then at most one phase behind on every barrier this group is about to wait on).
      if (lt > 0) mbar_wait(unit_conv_done(grp ^ 1), ((lt - 1) >> 1) & 1);
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int rs = g % RS, s = g % CS;
        mbar_wait(raw_full(rs), (g / RS) & 1);               // TMA landed the raw tiles
        mbar_wait(mma_done(s), ((g / CS) & 1) ^ 1);          // previous MMAs on this slot's converted operands retired
        tc_fence_after();
        uint8_t* base = smem_gen + rs * RAW_BYTES;
        uint8_t* cbase_s = smem_gen + RS * RAW_BYTES + s * CONV_BYTES;
        {
          // A: row `arow` of the raw tile (128 B, 16-byte chunks XOR-swizzled with row & 7) -> hi / lo -> TMEM
          const float4* rowp = reinterpret_cast<const float4*>(base + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        {
          // B: raw -> hi / lo tiles at the same (swizzled) offsets in shared memory
          const float4* raw = reinterpret_cast<const float4*>(base + 1 * TILE_BYTES);
          float4* hi = reinterpret_cast<float4*>(cbase_s);
          float4* lo = reinterpret_cast<float4*>(cbase_s + TILE_BYTES);
#pragma unroll
          for (int uu = 0; uu < TILE_BYTES / 16 / NUM_CONV_THREADS; ++uu) {
            const int e = ct + uu * NUM_CONV_THREADS;
            const float4 v = raw[e];
            float4 h, l;
            h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
            h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
            h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
            h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
            hi[e] = h; lo[e] = l;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(raw_empty(rs));   // the raw tiles may be overwritten by the next TMA
        mbar_arrive(conv_full(s));    // operands ready for the MMA warp
      }
      mbar_arrive(unit_conv_done(grp));
      // ----- epilogue of this unit (the other group is already converting the next one) -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      UmmaEpilogue ep;
      ep.mode = epi_mode; ep.acc0 = groups[gq].acc0; ep.acc1 = groups[gq].acc1; ep.tvec = groups[gq].tvec; ep.cin = nullptr;
      float* cbase = groups[gq].C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq4[4] = {0.0, 0.0, 0.0, 0.0}, acc_dot4[4] = {0.0, 0.0, 0.0, 0.0};   // fused row statistics: fp64 sums, four interleaved chains each
      if (wu.nkb > 0) {
        mbar_wait(tmem_full(ab), (lt >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t rr[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
          TMEM_LD32(taddr, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c == BN / 32 - 1) {                            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
          }
          if (ep.mode != UMMA_EPI_STATS_ONLY) {
            float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          }
          if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const double v = (double)__uint_as_float(rr[j]);
              acc_sq4[j & 3] = fma(v, v, acc_sq4[j & 3]);
              if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot4[j & 3] = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + j], acc_dot4[j & 3]);
            }
          }
          if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
            // symmetric product: also write the transposed tile (lanes = consecutive addresses)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
          }
        }
        const double acc_sq = (acc_sq4[0] + acc_sq4[1]) + (acc_sq4[2] + acc_sq4[3]), acc_dot = (acc_dot4[0] + acc_dot4[1]) + (acc_dot4[2] + acc_dot4[3]);
        if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, acc_sq);
        if (ep.mode == UMMA_EPI_STATS_ONLY) { atomicAdd(ep.acc0 + row, acc_sq); atomicAdd(ep.acc1 + row, acc_dot); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// grouped launch form of umma_gemm_ps_kernel (pre-split right operand)
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm_grouped_ps_kernel(const GemmGroup* __restrict__ groups, const int ngroups, int64_t ldc, int64_t c_split_stride, const GemmWork work,
                         const int epi_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  using namespace ps;
  const uint32_t bars = smem_base + ps::RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RSP + s); };
  auto conv_full = [&](int s) { return bars + 8u * (2 * RSP + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RSP + CS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + 2 + b); };
  auto unit_conv_done = [&](int b) { return bars + 8u * (2 * RSP + 2 * CS + 4 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + ps::RING_BYTES + 8 * (2 * RSP + 2 * CS + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&groups[0].tmA); tma_prefetch_desc(&groups[0].tmBhi); tma_prefetch_desc(&groups[0].tmBlo);
    for (int s = 0; s < RSP; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 1); }
    for (int s = 0; s < CS; ++s) { mbar_init(conv_full(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); mbar_init(unit_conv_done(b), NUM_CONV_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: barrier init / TMEM allocation above overlap the previous kernel of the chain
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: raw fp32 tiles, one ring across all units of this CTA =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= work.total * ngroups) break;
        int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int s = g % RSP;
          mbar_wait(raw_empty(s), ((g / RSP) & 1) ^ 1);
          const uint32_t dst = smem_base + s * STAGE_BYTES;
          mbar_expect_tx(raw_full(s), 3 * TILE_BYTES);
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, &groups[gq].tmA, raw_full(s), k, wu.tile_m * BM);
          tma_load_2d(dst + 1 * TILE_BYTES, &groups[gq].tmBhi, raw_full(s), k, wu.tile_n * BN);
          tma_load_2d(dst + 2 * TILE_BYTES, &groups[gq].tmBlo, raw_full(s), k, wu.tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator buffer lt & 1, so the epilogue of unit lt overlaps the main loop of unit lt + 1 =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total * ngroups) break;
      int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
      const int ab = lt & 1;
      mbar_wait(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);      // the epilogue two units ago has drained this accumulator
      tc_fence_after();
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS;
        mbar_wait(conv_full(s), (g / CS) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = smem_base + (g % RSP) * STAGE_BYTES + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;  // 8 tf32 = 32 bytes along K inside the 128B swizzle atom (B); 8 TMEM columns (A)
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));                          // frees the TMEM A columns of this slot
          tc_commit(raw_empty(g % RSP));                   // ... and the ring stage
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===== worker groups: raw -> hi / lo conversion of the group's units, then their epilogue =====
    const int grp = (warp - 2) >> 2;
    const int ct = (threadIdx.x - 64) & 127;         // 0..127 within the group
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int arow = q * 32 + lane;                  // A-tile row = TMEM lane handled by this thread
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total * ngroups) break;
      int gq; const WorkUnit wu = get_unit_grouped(work, ngroups, u, gq);
      if ((lt & 1) != grp) { g += wu.nkb; continue; }
      // A group only sees the mbarrier phases of its own units.  Parity waits are unambiguous only within one phase, so
      // do not start before the other group has issued every conversion of the previous unit (the TMA ring and the MMA
      // warp are then at most one phase behind on every barrier this group is about to wait on).
      if (lt > 0) mbar_wait(unit_conv_done(grp ^ 1), ((lt - 1) >> 1) & 1);
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int rs = g % RSP, s = g % CS;
        mbar_wait(raw_full(rs), (g / RSP) & 1);              // TMA landed the stage
        mbar_wait(mma_done(s), ((g / CS) & 1) ^ 1);          // previous MMAs on this slot's converted operands retired
        tc_fence_after();
        uint8_t* base = smem_gen + rs * STAGE_BYTES;
        {
          // A: row `arow` of the raw tile (128 B, 16-byte chunks XOR-swizzled with row & 7) -> hi / lo -> TMEM
          const float4* rowp = reinterpret_cast<const float4*>(base + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(conv_full(s));    // A operand ready for the MMA warp
      }
      mbar_arrive(unit_conv_done(grp));
      // ----- epilogue of this unit (the other group is already converting the next one) -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      UmmaEpilogue ep;
      ep.mode = epi_mode; ep.acc0 = groups[gq].acc0; ep.acc1 = groups[gq].acc1; ep.tvec = groups[gq].tvec; ep.cin = nullptr;
      float* cbase = groups[gq].C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq4[4] = {0.0, 0.0, 0.0, 0.0}, acc_dot4[4] = {0.0, 0.0, 0.0, 0.0};   // fused row statistics: fp64 sums, four interleaved chains each
      if (wu.nkb > 0) {
        mbar_wait(tmem_full(ab), (lt >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t rr[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
          TMEM_LD32(taddr, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c == BN / 32 - 1) {                            // accumulator drained: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
          }
          if (ep.mode != UMMA_EPI_STATS_ONLY) {
            float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          }
          if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const double v = (double)__uint_as_float(rr[j]);
              acc_sq4[j & 3] = fma(v, v, acc_sq4[j & 3]);
              if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot4[j & 3] = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + j], acc_dot4[j & 3]);
            }
          }
          if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
            // symmetric product: also write the transposed tile (lanes = consecutive addresses)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
          }
        }
        const double acc_sq = (acc_sq4[0] + acc_sq4[1]) + (acc_sq4[2] + acc_sq4[3]), acc_dot = (acc_dot4[0] + acc_dot4[1]) + (acc_dot4[2] + acc_dot4[3]);
        if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, acc_sq);
        if (ep.mode == UMMA_EPI_STATS_ONLY) { atomicAdd(ep.acc0 + row, acc_sq); atomicAdd(ep.acc1 + row, acc_dot); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// =====================================================================================================
// Third-generation main loop (opt-in: AGP_UMMA_V3=1).  MEASURED on a B200 (profiles/r2/gemm_bench_v1_v3.txt, gemm_trace_v3.txt):
// bit-identical results, but 4 % SLOWER than the kernel above (25.0 / 22.7 / 22.7 us against 24.1 / 20.8 / 20.6 us at C2; 110 / 97 /
// 100 us against 106 / 90 / 90 us at C3).  With two groups writing their A operands into tensor memory (tcgen05.st) while the MMA
// warp's tcgen05.mma read theirs from it, single 12-MMA issue bursts stall for ~3000 cycles and the conversions stall with them:
// tensor-memory traffic of the TS operand form, not converter throughput, is what paces the 3xTF32 main loop.  Kept as the
// record of that experiment.  Motivation was: the per-role timeline of the
// kernel above (profiles/microbench/gemm_trace.cu, profiles/r2/gemm_trace_v1.txt) shows a rigid 1500-cycle cadence per 32-deep
// k-block at every size -- ONE worker group converts at a time (985 cycles of split arithmetic + ~520 cycles of mbarrier / fence /
// tcgen05.wait::st round trips, nothing overlapped), while the MMA warp issues the block's 12 tcgen05.mma in ~900 cycles and then
// idles.  Here BOTH worker groups convert concurrently: a group that becomes free takes the next k-block of the CTA from a
// shared counter (so two conversions are always in flight, whichever unit they belong to), and the group that converts the LAST
// k-block of a unit drains that unit's accumulator while the other group keeps the MMA warp fed.  Same operand layouts, same MMA
// order (lo.hi, hi.lo, hi.hi), same epilogues as the kernel above; every mbarrier wait is bounded (trap instead of a hung GPU).
// GROUPED: tensor maps / epilogue targets per latent from a device array (umma_gemm_grouped_kernel's launch form).
namespace v3 {
constexpr int RS3 = 3, CS3 = 3;
constexpr int RING_BYTES = RS3 * RAW_BYTES + CS3 * CONV_BYTES;           // 192 KB
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
constexpr int NBAR = 2 * RS3 + 2 * CS3 + 4;                              // raw_full/empty, conv_full, mma_done, tmem_full/empty
}

__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("umma v3: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

template <bool GROUPED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gemm3_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, float* __restrict__ C1, const UmmaEpilogue ep1,
                  const GemmGroup* __restrict__ groups, const int ngroups, int64_t ldc, int64_t c_split_stride, const GemmWork work, const int epi_mode) {
  using namespace v3;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + v3::RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RS3 + s); };
  auto conv_full = [&](int s) { return bars + 8u * (2 * RS3 + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RS3 + CS3 + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RS3 + 2 * CS3 + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RS3 + 2 * CS3 + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + v3::RING_BYTES + 8 * NBAR);
  int* grab = reinterpret_cast<int*>(smem_gen + v3::RING_BYTES + 8 * NBAR + 8);           // next k-block of this CTA to convert
  volatile int* gsel = reinterpret_cast<volatile int*>(smem_gen + v3::RING_BYTES + 8 * NBAR + 16);   // [2 groups][2 parities]
  const uint32_t conv_base = smem_base + RS3 * RAW_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int total_units = GROUPED ? work.total * ngroups : work.total;
  auto unit_of = [&](int u, int& gq) -> WorkUnit {
    if (GROUPED) return get_unit_grouped(work, ngroups, u, gq);
    gq = 0;
    return get_unit(work, u);
  };

  if (warp == 0 && lane == 0) {
    if (GROUPED) { tma_prefetch_desc(&groups[0].tmA); tma_prefetch_desc(&groups[0].tmB); }
    else { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB1); }
    for (int s = 0; s < RS3; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), NUM_CONV_THREADS); }
    for (int s = 0; s < CS3; ++s) { mbar_init(conv_full(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
    *grab = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) UTT(0, 1);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= total_units) break;
        int gq; const WorkUnit wu = unit_of(u, gq);
        const CUtensorMap* ma = GROUPED ? &groups[gq].tmA : &tmA1;
        const CUtensorMap* mb = GROUPED ? &groups[gq].tmB : &tmB1;
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int s = g % RS3;
          mbar_wait_bounded(raw_empty(s), ((g / RS3) & 1) ^ 1);
          UTT(0, 1000 + g);
          const uint32_t dst = smem_base + s * RAW_BYTES;
          mbar_expect_tx(raw_full(s), 2 * TILE_BYTES);
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, ma, raw_full(s), k, wu.tile_m * BM);
          tma_load_2d(dst + 1 * TILE_BYTES, mb, raw_full(s), k, wu.tile_n * BN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (k-blocks in order; accumulator lt & 1) =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r) {
      const int u = unit_index(r, cta, G);
      if (u >= total_units) break;
      int gq; const WorkUnit wu = unit_of(u, gq);
      if (wu.nkb <= 0) continue;
      const int ab = lt & 1;
      mbar_wait_bounded(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) UTT(1, 100000 + lt);
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS3;
        mbar_wait_bounded(conv_full(s), (g / CS3) & 1);
        tc_fence_after();
        if (lane == 0) UTT(1, 2000 + g);
        if (elect_one()) {
          const uint32_t b_hi = conv_base + s * CONV_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);  // small terms first
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));
        }
        __syncwarp();
        if (lane == 0) UTT(1, 3000 + g);
      }
      ++lt;
    }
  } else {
    // ===== worker groups: whichever group is free converts the CTA's next k-block; the converter of a unit's last k-block
    // drains that unit =====
    const int grp = (warp - 2) >> 2;
    const int ct = (threadIdx.x - 64) & 127;
    const int q = warp & 3;
    const int arow = q * 32 + lane;
    int cur_r = 0, cur_g0 = 0, cur_lt = 0;      // cursor: round of the current unit, its first k-block index, its accumulator count
    int it = 0;
    for (;; ++it) {
      if (ct == 0) gsel[grp * 2 + (it & 1)] = atomicAdd(grab, 1);
      group_bar(2 + grp);
      const int g = gsel[grp * 2 + (it & 1)];
      // locate the unit that contains k-block g (units without k-blocks own no accumulator)
      WorkUnit wu; int gq = 0; bool done = false;
      for (;;) {
        const int u = unit_index(cur_r, cta, G);
        if (u >= total_units) { done = true; break; }
        wu = unit_of(u, gq);
        if (wu.nkb > 0 && g < cur_g0 + wu.nkb) break;
        if (wu.nkb > 0) { cur_g0 += wu.nkb; ++cur_lt; }
        ++cur_r;
      }
      if (done) break;
      const int i = g - cur_g0, lt = cur_lt;
      {
        const int rs = g % RS3, s = g % CS3;
        if (ct == 0) UTT(2 + grp, 7000 + g);
        mbar_wait_bounded(raw_full(rs), (g / RS3) & 1);
        if (ct == 0) UTT(2 + grp, 4000 + g);
        mbar_wait_bounded(mma_done(s), ((g / CS3) & 1) ^ 1);
        tc_fence_after();
        if (ct == 0) UTT(2 + grp, 5000 + g);
        uint8_t* base = smem_gen + rs * RAW_BYTES;
        uint8_t* cbase_s = smem_gen + RS3 * RAW_BYTES + s * CONV_BYTES;
        {
          const float4* rowp = reinterpret_cast<const float4*>(base + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        {
          const float4* raw = reinterpret_cast<const float4*>(base + 1 * TILE_BYTES);
          float4* hi = reinterpret_cast<float4*>(cbase_s);
          float4* lo = reinterpret_cast<float4*>(cbase_s + TILE_BYTES);
#pragma unroll
          for (int uu = 0; uu < TILE_BYTES / 16 / NUM_CONV_THREADS; ++uu) {
            const int e = ct + uu * NUM_CONV_THREADS;
            const float4 v = raw[e];
            float4 h, l;
            h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
            h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
            h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
            h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
            hi[e] = h; lo[e] = l;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(raw_empty(rs));
        mbar_arrive(conv_full(s));
        if (ct == 0) UTT(2 + grp, 6000 + g);
      }
      if (i != wu.nkb - 1) continue;
      // ----- epilogue of unit lt (the other group goes on converting) -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      UmmaEpilogue ep;
      if (GROUPED) { ep.mode = epi_mode; ep.acc0 = groups[gq].acc0; ep.acc1 = groups[gq].acc1; ep.tvec = groups[gq].tvec; ep.cin = nullptr; }
      else ep = ep1;
      float* cbase = (GROUPED ? groups[gq].C : C1) + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq[4] = {0.0, 0.0, 0.0, 0.0}, acc_dot[4] = {0.0, 0.0, 0.0, 0.0};   // four interleaved fp64 chains per statistic
      mbar_wait_bounded(tmem_full(ab), (lt >> 1) & 1);
      tc_fence_after();
      if (ct == 0) UTT(2 + grp, 300000 + lt);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t rr[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
        TMEM_LD32(taddr, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(ab));
        }
        if (ep.mode != UMMA_EPI_STATS_ONLY) {
          float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
        if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
          const double* tv = ep.tvec + wu.tile_n * BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const double v = (double)__uint_as_float(rr[j]);
            acc_sq[j & 3] = fma(v, v, acc_sq[j & 3]);
            if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot[j & 3] = fma(v, tv[j], acc_dot[j & 3]);
          }
        }
        if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
        }
      }
      if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, (acc_sq[0] + acc_sq[1]) + (acc_sq[2] + acc_sq[3]));
      if (ep.mode == UMMA_EPI_STATS_ONLY) {
        atomicAdd(ep.acc0 + row, (acc_sq[0] + acc_sq[1]) + (acc_sq[2] + acc_sq[3]));
        atomicAdd(ep.acc1 + row, (acc_dot[0] + acc_dot[1]) + (acc_dot[2] + acc_dot[3]));
      }
      if (ct == 0) UTT(2 + grp, 400000 + lt);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) UTT(0, 2);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// =====================================================================================================
// Second-generation kernel (opt-in: AGP_UMMA_V2=1; NOT yet measured or parity-checked on a B200 -- written
// after the round's GPU budget was spent, from the line-level stall profile profiles/r1/umma_stalls_by_line.txt).
// What the profile showed for the kernel above: the two worker groups alternate by UNIT, so one group converts
// while the other drains an accumulator and then idles -- a single 4-warp group cannot hide the shared-memory /
// tcgen05.st latency of a k-block (~2000 cycles per k-block against 768 cycles of tcgen05.mma), and half of the
// conversion work is the m x m operand that every CTA splits again.  Changes:
//   * roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 a dedicated epilogue group, warps 6..13 two
//     converter groups that alternate by K-BLOCK (two conversions in flight at any time)
//   * the B operand may arrive pre-split (L^-1 once per refresh_K, X by x_finalize_kernel): its hi / lo tiles go
//     from TMA straight to the tensor core and the converters only split the A rows into TMEM.  Without a
//     pre-split copy (Gram product) the raw B tile is converted in place (hi) with lo beside it.
//   * rings: raw A 4 x 16 KB (freed by the converters), B 4 x 32 KB and 4 TMEM A slots (both freed by
//     tcgen05.commit), so A prefetch is not gated by the MMA
//   * every mbarrier wait is bounded (about a second) and traps, so a protocol error is a launch failure, not a hang
// The barrier protocol is model-checked on the CPU by tools/umma_v2_protocol_sim.py.
namespace v2 {
constexpr int RA = 4, RB = 4, TS = 4;             // raw-A stages, B stages, TMEM A slots (RB == TS: one release barrier)
constexpr int A_BYTES = TILE_BYTES, B_BYTES = 2 * TILE_BYTES;
constexpr int RING_BYTES = RA * A_BYTES + RB * B_BYTES;   // 192 KB
constexpr int NCG = 2;                            // converter groups; must divide RA and TS
constexpr int CONV_WARP0 = 6;                     // warps 2 .. CONV_WARP0-1: epilogue group
constexpr int NUM_THREADS = 64 + 128 + NCG * 128; // 448
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
static_assert(RB == TS && RA % NCG == 0 && TS % NCG == 0, "ring / group shape");
static_assert(TMEM_A0 + TS * 64 <= TMEM_COLS, "TMEM budget");

__device__ __forceinline__ void wait_bounded(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++n & 1023u) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}
}  // namespace v2

__global__ void __launch_bounds__(v2::NUM_THREADS, 1)
umma_gemm_nt_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                       const __grid_constant__ CUtensorMap tmB1, float* __restrict__ C, int64_t ldc, int64_t c_split_stride,
                       const GemmWork work, const UmmaEpilogue ep, const int b_presplit) {
  using namespace v2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + v2::RING_BYTES;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (RA + s); };
  auto b_full = [&](int s) { return bars + 8u * (2 * RA + s); };
  auto conv_done = [&](int s) { return bars + 8u * (2 * RA + RB + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RA + RB + TS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RA + RB + 2 * TS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RA + RB + 2 * TS + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + v2::RING_BYTES + 8 * (2 * RA + RB + 2 * TS + 4));
  const uint32_t b_base = smem_base + RA * A_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB0);
    if (b_presplit) tma_prefetch_desc(&tmB1);
    for (int s = 0; s < RA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), NUM_CONV_THREADS); }
    for (int s = 0; s < RB; ++s) mbar_init(b_full(s), 1);
    for (int s = 0; s < TS; ++s) { mbar_init(conv_done(s), NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: raw A tiles (ring freed by the converters) and B tiles (ring freed by the MMA) =====
      int g = 0;
      for (int r = 0;; ++r) {
        const int u = unit_index(r, cta, G);
        if (u >= work.total) break;
        const WorkUnit wu = get_unit(work, u);
        for (int i = 0; i < wu.nkb; ++i, ++g) {
          const int k = (wu.kb0 + i) * BK;
          const int sa = g % RA, sb = g % RB;
          wait_bounded(a_empty(sa), ((g / RA) & 1) ^ 1);
          mbar_expect_tx(a_full(sa), A_BYTES);
          tma_load_2d(smem_base + sa * A_BYTES, &tmA, a_full(sa), k, wu.tile_m * BM);
          wait_bounded(mma_done(sb), ((g / RB) & 1) ^ 1);         // the MMAs of k-block g - RB have read this stage
          const uint32_t dst = b_base + sb * B_BYTES;
          mbar_expect_tx(b_full(sb), b_presplit ? 2 * TILE_BYTES : TILE_BYTES);
          tma_load_2d(dst, &tmB0, b_full(sb), k, wu.tile_n * BN);                                   // hi (or raw)
          if (b_presplit) tma_load_2d(dst + TILE_BYTES, &tmB1, b_full(sb), k, wu.tile_n * BN);      // lo
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      const int ab = lt & 1;
      wait_bounded(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % TS;
        wait_bounded(conv_done(s), (g / TS) & 1);                 // A hi / lo in TMEM (and B converted when not pre-split)
        if (b_presplit) wait_bounded(b_full(s), (g / RB) & 1);    // B hi / lo landed by TMA
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = b_base + s * B_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));                                 // frees the B stage and the TMEM A slot
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));
        }
        __syncwarp();
      }
    }
  } else if (warp < CONV_WARP0) {
    // ===== epilogue group: drains accumulator lt & 1 while the main loop of unit lt + 1 runs =====
    const int q = warp & 3;
    int lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      if (wu.nkb <= 0) continue;
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      float* cbase = C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      double acc_sq = 0.0, acc_dot = 0.0;
      float fsq = 0.f;
      wait_bounded(tmem_full(ab), (lt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t rr[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
        TMEM_LD32(taddr, rr);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(ab));
        }
        if (ep.mode == UMMA_EPI_EYE_MINUS) {
          const int col0 = wu.tile_n * BN + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = -__uint_as_float(rr[j]);
            if (col0 + j == row) v += 1.f;
            fsq = fmaf(v, v, fsq);
            rr[j] = __float_as_uint(v);
          }
        } else if (ep.mode == UMMA_EPI_ADD) {
          const float4* src = reinterpret_cast<const float4*>(ep.cin + (int64_t)row * ldc + (int64_t)wu.tile_n * BN + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = src[j];
            rr[4 * j + 0] = __float_as_uint(__uint_as_float(rr[4 * j + 0]) + a.x);
            rr[4 * j + 1] = __float_as_uint(__uint_as_float(rr[4 * j + 1]) + a.y);
            rr[4 * j + 2] = __float_as_uint(__uint_as_float(rr[4 * j + 2]) + a.z);
            rr[4 * j + 3] = __float_as_uint(__uint_as_float(rr[4 * j + 3]) + a.w);
          }
        }
        if (ep.mode != UMMA_EPI_STATS_ONLY && ep.mode != UMMA_EPI_STATS_SIGMA) {
          float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
        }
        if (ep.mode == UMMA_EPI_STORE_SUMSQ || ep.mode == UMMA_EPI_STATS_ONLY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const double v = (double)__uint_as_float(rr[j]);
            acc_sq = fma(v, v, acc_sq);
            if (ep.mode == UMMA_EPI_STATS_ONLY) acc_dot = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + j], acc_dot);
          }
        }
        if (ep.mode == UMMA_EPI_STATS_SIGMA) {
          const float4* vrow = reinterpret_cast<const float4*>(ep.cin + (int64_t)row * ldc + (int64_t)wu.tile_n * BN + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = vrow[j];
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const double v = (double)__uint_as_float(rr[4 * j + e]);
              acc_sq = fma(v, (double)av[e], acc_sq);
              acc_dot = fma(v, ep.tvec[wu.tile_n * BN + c * 32 + 4 * j + e], acc_dot);
            }
          }
        }
        if (ep.mode == UMMA_EPI_STORE_MIRROR && wu.tile_n != wu.tile_m) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
        }
      }
      if (ep.mode == UMMA_EPI_STORE_SUMSQ) atomicAdd(ep.acc0 + row, acc_sq);
      if (ep.mode == UMMA_EPI_STATS_ONLY || ep.mode == UMMA_EPI_STATS_SIGMA) { atomicAdd(ep.acc0 + row, acc_sq); atomicAdd(ep.acc1 + row, acc_dot); }
      if (ep.mode == UMMA_EPI_EYE_MINUS) {
        double sq = (double)fsq;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) atomicAdd(ep.acc0, sq);
      }
    }
  } else {
    // ===== converter groups: group c splits the k-blocks g = c (mod NCG) =====
    const int grp = (warp - CONV_WARP0) >> 2;
    const int ct = (threadIdx.x - CONV_WARP0 * 32) & 127;
    const int q = warp & 3;
    const int arow = q * 32 + lane;
    int g = 0;
    for (int r = 0;; ++r) {
      const int u = unit_index(r, cta, G);
      if (u >= work.total) break;
      const WorkUnit wu = get_unit(work, u);
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        if (g % NCG != grp) continue;
        const int sa = g % RA, s = g % TS;
        wait_bounded(mma_done(s), ((g / TS) & 1) ^ 1);            // TMEM A slot (and B stage) of k-block g - TS released
        wait_bounded(a_full(sa), (g / RA) & 1);
        tc_fence_after();
        {
          const float4* rowp = reinterpret_cast<const float4*>(smem_gen + sa * A_BYTES + arow * 128);
          uint32_t h[32], l[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = rowp[c ^ (arow & 7)];
            float t;
            t = tf32_rna(v.x); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v.x - t));
            t = tf32_rna(v.y); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v.y - t));
            t = tf32_rna(v.z); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v.z - t));
            t = tf32_rna(v.w); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v.w - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
        }
        if (!b_presplit) {
          wait_bounded(b_full(s), (g / RB) & 1);
          float4* hi = reinterpret_cast<float4*>(smem_gen + RA * A_BYTES + s * B_BYTES);   // raw tile, split in place
          float4* lo = reinterpret_cast<float4*>(smem_gen + RA * A_BYTES + s * B_BYTES + TILE_BYTES);
#pragma unroll
          for (int uu = 0; uu < TILE_BYTES / 16 / NUM_CONV_THREADS; ++uu) {
            const int e = ct + uu * NUM_CONV_THREADS;
            const float4 v = hi[e];
            float4 hh, ll;
            hh.x = tf32_rna(v.x); ll.x = tf32_rna(v.x - hh.x);
            hh.y = tf32_rna(v.y); ll.y = tf32_rna(v.y - hh.y);
            hh.z = tf32_rna(v.z); ll.z = tf32_rna(v.z - hh.z);
            hh.w = tf32_rna(v.w); ll.w = tf32_rna(v.w - hh.w);
            hi[e] = hh; lo[e] = ll;
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        if (!b_presplit) fence_proxy_async();
        mbar_arrive(a_empty(sa));
        mbar_arrive(conv_done(s));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Gram product straight from V (default for single-latent steps; AGP_GRAM_TN=0 goes back to scale_transpose_kernel + the kernel at the
// top): split-K partials of U^T U, U = diag(sqrt(rho w)) V, with V read as stored ([B][ldm], samples = the K dimension along the ROWS).
// The transposition and the scaling happen in the worker threads, where the operands are rewritten anyway for the hi / lo split:
//   * TMA lands un-swizzled [32 samples][128 columns] boxes of V (one per operand; one box for a diagonal tile);
//   * the TMA warp's 32 lanes put sf[k] = (float)sqrt(max(rho w_k, 0)) and gf[k] = (float)g_k of the k-block next to the stage;
//   * worker thread r reads COLUMN r of the box (lanes = consecutive words: conflict-free), multiplies by sf[k], splits, and writes
//     row r of the A operand to tensor memory and row r of the K-major 128B-swizzled B_hi / B_lo tiles (a quarter warp writes eight
//     different 16-byte chunk positions: conflict-free);
//   * on diagonal tiles the same thread accumulates (V^T g)_r (transpose(kappa) * grad_mu, analyticVI.jl:168, whitened): fp32 over the
//     32 samples of a k-block, fp64 across k-blocks, one atomicAdd per unit.
// This removes scale_transpose_kernel (9.3 us: 16.8 MB read + 16.8 MB written) from the critical chain of the step.  MMA warp,
// accumulators, unit enumeration (tri_mode 2) and the mirror epilogue are those of umma_gemm_nt_kernel.
namespace tn {
constexpr int SC_BYTES = 256;                                    // sf[32] | gf[32]
constexpr int SC_OFF = RS * RAW_BYTES + CS * CONV_BYTES;         // after the operand rings (which need 1024-byte alignment)
constexpr int RING_BYTES = SC_OFF + RS * SC_BYTES;
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;
}
// GROUPED: one launch for the Gram products of several latent GPs (multi-latent steps): the tensor map of V, the weights, the gradient
// vector and the outputs of unit u's latent come from a device array, the units of all latents interleave (get_unit_grouped).
struct alignas(64) GramTnGroup {
  CUtensorMap tmV;
  float* C; const double* w; const double* g; double* v1;
};
template <bool GROUPED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
umma_gram_tn_kernel(const __grid_constant__ CUtensorMap tmV1, float* __restrict__ C1, int64_t ldc, int64_t c_split_stride, const GemmWork work,
                    const double* __restrict__ wvec1, const double rho, const double* __restrict__ gvec1, double* __restrict__ v1_1,
                    const GramTnGroup* __restrict__ groups, const int ngroups) {
  const int total_units = GROUPED ? work.total * ngroups : work.total;
  auto unit_of = [&](int u, int& gq) -> WorkUnit {
    if (GROUPED) return get_unit_grouped(work, ngroups, u, gq);
    gq = 0;
    return get_unit(work, u);
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + tn::RING_BYTES;
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RS + s); };
  auto conv_full = [&](int s) { return bars + 8u * (2 * RS + s); };
  auto mma_done = [&](int s) { return bars + 8u * (2 * RS + CS + s); };
  auto tmem_full = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + b); };
  auto tmem_empty = [&](int b) { return bars + 8u * (2 * RS + 2 * CS + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + tn::RING_BYTES + 8 * (2 * RS + 2 * CS + 6));
  const uint32_t conv_base = smem_base + RS * RAW_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(GROUPED ? &groups[0].tmV : &tmV1);
    for (int s = 0; s < RS; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 2 * NUM_CONV_THREADS); }
    for (int s = 0; s < CS; ++s) { mbar_init(conv_full(s), 2 * NUM_CONV_THREADS); mbar_init(mma_done(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===== producer warp: all lanes stage the sample scales of the k-block, lane 0 issues the TMA boxes =====
    int g = 0;
    for (int r = 0;; ++r) {
      const int u = unit_index(r, cta, G);
      if (u >= total_units) break;
      int gq; const WorkUnit wu = unit_of(u, gq);
      const bool diag = wu.tile_m == wu.tile_n;
      const CUtensorMap* tmV = GROUPED ? &groups[gq].tmV : &tmV1;
      const double* __restrict__ wvec = GROUPED ? groups[gq].w : wvec1;
      const double* __restrict__ gvec = GROUPED ? groups[gq].g : gvec1;
      double wk = wu.nkb > 0 ? wvec[wu.kb0 * BK + lane] : 0.0, gk = wu.nkb > 0 ? gvec[wu.kb0 * BK + lane] : 0.0;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % RS;
        const float sfv = (float)sqrt(fmax(rho * wk, 0.0)), gfv = (float)gk;
        if (i + 1 < wu.nkb) { wk = wvec[(wu.kb0 + i + 1) * BK + lane]; gk = gvec[(wu.kb0 + i + 1) * BK + lane]; }   // next k-block, in flight during the wait
        mbar_wait(raw_empty(s), ((g / RS) & 1) ^ 1);
        if (lane == 0) UTT(0, 1000 + g);
        float* sc = reinterpret_cast<float*>(smem_gen + tn::SC_OFF + s * tn::SC_BYTES);
        sc[lane] = sfv; sc[32 + lane] = gfv;
        __syncwarp();
        if (lane == 0) {
          const uint32_t dst = smem_base + s * RAW_BYTES;
          mbar_expect_tx(raw_full(s), diag ? TILE_BYTES : 2 * TILE_BYTES);     // release: the scale words above are visible to the waiters
          const int k = (wu.kb0 + i) * BK;
          tma_load_2d(dst + 0 * TILE_BYTES, tmV, raw_full(s), wu.tile_m * BM, k);
          if (!diag) tma_load_2d(dst + 1 * TILE_BYTES, tmV, raw_full(s), wu.tile_n * BN, k);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (as umma_gemm_nt_kernel) =====
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= total_units) break;
      int gq; const WorkUnit wu = unit_of(u, gq);
      const int ab = lt & 1;
      mbar_wait(tmem_empty(ab), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + ab * BN;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int s = g % CS;
        mbar_wait(conv_full(s), (g / CS) & 1);
        tc_fence_after();
        if (lane == 0) UTT(1, 2000 + g);
        if (elect_one()) {
          const uint32_t b_hi = conv_base + s * CONV_BYTES, b_lo = b_hi + TILE_BYTES;
          const uint32_t a_hi = tmem_base + TMEM_A0 + s * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint32_t off = kk * 32;
            const uint64_t dbh = make_desc(b_hi + off), dbl = make_desc(b_lo + off);
            tc_mma_tf32_ts(acc, a_lo + kk * 8, dbh, kIdesc, (i > 0 || kk > 0) ? 1u : 0u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbl, kIdesc, 1u);
            tc_mma_tf32_ts(acc, a_hi + kk * 8, dbh, kIdesc, 1u);
          }
          tc_commit(mma_done(s));
          if (i == wu.nkb - 1) tc_commit(tmem_full(ab));
        }
        __syncwarp();
        if (lane == 0) UTT(1, 3000 + g);
      }
    }
  } else {
    // ===== worker groups: group 0 builds the A operand (tensor memory) and V^T g, group 1 the B operand (shared memory) of EVERY
    // k-block -- a Gram launch has one unit per CTA, so alternating the groups by unit would leave one of them idle, and one group
    // alone converts at 2400 cycles per k-block against 1490 for the MMAs; then both drain half of the accumulator columns each =====
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int arow = q * 32 + lane;                  // operand row = TMEM lane = column of the V box
    int g = 0, lt = 0;
    for (int r = 0;; ++r, ++lt) {
      const int u = unit_index(r, cta, G);
      if (u >= total_units) break;
      int gq; const WorkUnit wu = unit_of(u, gq);
      const bool diag = wu.tile_m == wu.tile_n;
      float* __restrict__ C = GROUPED ? groups[gq].C : C1;
      double* __restrict__ v1 = GROUPED ? groups[gq].v1 : v1_1;
      double vdot = 0.0;
      for (int i = 0; i < wu.nkb; ++i, ++g) {
        const int rs = g % RS, s = g % CS;
        mbar_wait(raw_full(rs), (g / RS) & 1);
        if ((threadIdx.x & 127) == 64) UTT(2 + grp, 4000 + g);
        mbar_wait(mma_done(s), ((g / CS) & 1) ^ 1);
        tc_fence_after();
        if ((threadIdx.x & 127) == 64) UTT(2 + grp, 5000 + g);
        const float4* sc4 = reinterpret_cast<const float4*>(smem_gen + tn::SC_OFF + rs * tn::SC_BYTES);
        if (grp == 0) {
          const float* rawA = reinterpret_cast<const float*>(smem_gen + rs * RAW_BYTES) + arow;
          uint32_t h[32], l[32];
          float dot32 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 sf = sc4[c];
            const float a0 = rawA[(4 * c + 0) * BM], a1 = rawA[(4 * c + 1) * BM], a2 = rawA[(4 * c + 2) * BM], a3 = rawA[(4 * c + 3) * BM];
            if (diag) {
              const float4 gf = sc4[8 + c];
              dot32 = fmaf(a0, gf.x, dot32); dot32 = fmaf(a1, gf.y, dot32); dot32 = fmaf(a2, gf.z, dot32); dot32 = fmaf(a3, gf.w, dot32);
            }
            float v, t;
            v = a0 * sf.x; t = tf32_rna(v); h[4 * c + 0] = __float_as_uint(t); l[4 * c + 0] = __float_as_uint(tf32_rna(v - t));
            v = a1 * sf.y; t = tf32_rna(v); h[4 * c + 1] = __float_as_uint(t); l[4 * c + 1] = __float_as_uint(tf32_rna(v - t));
            v = a2 * sf.z; t = tf32_rna(v); h[4 * c + 2] = __float_as_uint(t); l[4 * c + 2] = __float_as_uint(tf32_rna(v - t));
            v = a3 * sf.w; t = tf32_rna(v); h[4 * c + 3] = __float_as_uint(t); l[4 * c + 3] = __float_as_uint(tf32_rna(v - t));
          }
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + s * 64;
          TMEM_ST32(ta, h);
          TMEM_ST32(ta + 32, l);
          if (diag) vdot += (double)dot32;
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
        } else {
          const float* rawB = reinterpret_cast<const float*>(smem_gen + rs * RAW_BYTES + (diag ? 0 : TILE_BYTES)) + arow;
          uint8_t* cbase_s = smem_gen + RS * RAW_BYTES + s * CONV_BYTES;
          float4* bhi = reinterpret_cast<float4*>(cbase_s + arow * 128);
          float4* blo = reinterpret_cast<float4*>(cbase_s + TILE_BYTES + arow * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 sf = sc4[c];
            float4 hv, lv;
            float v;
            v = rawB[(4 * c + 0) * BM] * sf.x; hv.x = tf32_rna(v); lv.x = tf32_rna(v - hv.x);
            v = rawB[(4 * c + 1) * BM] * sf.y; hv.y = tf32_rna(v); lv.y = tf32_rna(v - hv.y);
            v = rawB[(4 * c + 2) * BM] * sf.z; hv.z = tf32_rna(v); lv.z = tf32_rna(v - hv.z);
            v = rawB[(4 * c + 3) * BM] * sf.w; hv.w = tf32_rna(v); lv.w = tf32_rna(v - hv.w);
            bhi[c ^ (arow & 7)] = hv;     // row arow of the K-major tile, 16-byte chunk c at its 128B-swizzle position
            blo[c ^ (arow & 7)] = lv;
          }
          fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
        }
        mbar_arrive(raw_empty(rs));
        mbar_arrive(conv_full(s));
        if ((threadIdx.x & 127) == 64) UTT(2 + grp, 6000 + g);
      }
      if (grp == 0 && diag && wu.nkb > 0) atomicAdd(v1 + wu.tile_m * BM + arow, vdot);
      // ----- epilogue: C tile and, off the diagonal, its transpose (UMMA_EPI_STORE_MIRROR of umma_gemm_nt_kernel); group g drains the
      // column chunks 2g and 2g + 1 -----
      const int ab = lt & 1;
      const int row = wu.tile_m * BM + q * 32 + lane;
      float* cbase = C + (int64_t)wu.split * c_split_stride;
      float* crow = cbase + (int64_t)row * ldc + (int64_t)wu.tile_n * BN;
      if (wu.nkb > 0) {
        mbar_wait(tmem_full(ab), (lt >> 1) & 1);
        tc_fence_after();
        if ((threadIdx.x & 127) == 64) UTT(2 + grp, 300000 + lt);
#pragma unroll 1
        for (int c = 2 * grp; c < 2 * grp + 2; ++c) {
          uint32_t rr[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * BN + c * 32);
          TMEM_LD32(taddr, rr);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (c == 2 * grp + 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(ab));
          }
          float4* dst = reinterpret_cast<float4*>(crow + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3]));
          if (!diag) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              cbase[(int64_t)(wu.tile_n * BN + c * 32 + j) * ldc + row] = __uint_as_float(rr[j]);
          }
        }
        if ((threadIdx.x & 127) == 64) UTT(2 + grp, 400000 + lt);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// hi = rna_tf32(src), lo = rna_tf32(src - hi) for an m x m operand that every CTA of the v2 GEMM would otherwise split again
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo,
                                                         int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    float4 h, l;
    h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
    h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
    h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
    h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
}

// U^T = (diag(sqrt(rho w)) V)^T:  V is [B][ldv], output is [m][ldt] fp32 (split to hi/lo inside the GEMM); fused with
// v1[j] += sum_b V[b][j] g[b]  (transpose(kappa) * grad_mu, analyticVI.jl:168, whitened).  Block = 32 columns x 128 rows.
__global__ void __launch_bounds__(256) scale_transpose_kernel(const float* __restrict__ V, int64_t ldv, const double* __restrict__ w, double rho,
                                                              float* __restrict__ UT, int64_t ldt, const double* __restrict__ g,
                                                              double* __restrict__ v1) {
  __shared__ float tile[128][33];
  __shared__ float sw[128];
  __shared__ double sg[128];
  __shared__ double red[8][33];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int j0 = blockIdx.x * 32, b0 = blockIdx.y * 128;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;  // 32 x 8
  if (tid < 128) {
    sw[tid] = (float)sqrt(fmax(rho * w[b0 + tid], 0.0));
    sg[tid] = g[b0 + tid];
  }
  float v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) v[u] = V[(int64_t)(b0 + ty + 8 * u) * ldv + j0 + tx];   // 16 independent coalesced loads
  __syncthreads();
  double dot = 0.0;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int r = ty + 8 * u;
    dot = fma((double)v[u], sg[r], dot);
    tile[r][tx] = v[u] * sw[r];
  }
  red[ty][tx] = dot;
  __syncthreads();
  // U^T rows j0 .. j0+31, columns b0 .. b0+127: each warp writes 128 contiguous floats of one row per iteration
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int jr = ty + 8 * u;
#pragma unroll
    for (int sblk = 0; sblk < 4; ++sblk) UT[(int64_t)(j0 + jr) * ldt + b0 + tx + 32 * sblk] = tile[tx + 32 * sblk][jr];
  }
  if (ty == 0) {
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) a += red[r][tx];
    atomicAdd(v1 + j0 + tx, a);
  }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// K-major fp32 matrix [rows][cols] with leading dimension ld -> tensor map with a (32 x 128) 128B-swizzled box
bool make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// MN-major view of the same kind of matrix for umma_gram_tn_kernel: (128 columns x 32 rows) boxes, no swizzle (the worker threads read
// the box column by column and write the swizzled K-major operands themselves)
bool make_map_tn(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BM, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct Maps {
  CUtensorMap vtn;           // V as the operand of umma_gram_tn_kernel
  CUtensorMap raw[UM_COUNT], ut;
  CUtensorMap split[2][2];   // v2: [0] = L^-1, [1] = X ; [.][0] = hi, [.][1] = lo
};

int fail(std::string* err, const char* what, cudaError_t e = cudaSuccess) {
  *err = std::string("tcgen05 path: ") + what + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string());
  return 2;  // AGP_ERR_CUDA
}

}  // namespace

bool umma_shape_ok(int m, int Bcap) { return m >= 128 && m % 128 == 0 && Bcap >= 128 && Bcap % 128 == 0; }

int umma_latent_alloc(std::string* err, UmmaLatent& u, int m, int ldm, int Bcap, const float* Knm, const float* V, const float* Linv,
                      const float* X, cudaStream_t st) {
  u.m = m; u.ldm = ldm; u.Bcap = Bcap;
  const size_t rows[UM_COUNT] = {(size_t)Bcap, (size_t)Bcap, (size_t)m, (size_t)m};
  const float* ptr[UM_COUNT] = {Knm, V, Linv, X};
  cudaError_t e;
  if ((e = cudaMalloc(&u.UT, (size_t)m * Bcap * sizeof(float))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  cudaMemsetAsync(u.UT, 0, (size_t)m * Bcap * sizeof(float), st);
  Maps* mp = new Maps();
  bool ok = true;
  for (int i = 0; i < UM_COUNT; ++i) ok = ok && make_map(&mp->raw[i], ptr[i], rows[i], m, ldm);
  ok = ok && make_map(&mp->ut, u.UT, m, Bcap, Bcap);
  ok = ok && make_map_tn(&mp->vtn, V, Bcap, m, ldm);
  u.tmaps = mp;
  if (!ok) return fail(err, "cuTensorMapEncodeTiled failed");
  if ((e = cudaFuncSetAttribute(umma_gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute", e);
  if ((e = cudaFuncSetAttribute(umma_gram_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tn::SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute (gram tn)", e);
  u.gram_tn = 1;
  if (const char* env = getenv("AGP_GRAM_TN")) u.gram_tn = atoi(env) != 0;
  if (getenv("AGP_UMMA_V2") || getenv("AGP_UMMA_V3")) u.gram_tn = 0;     // the experimental main loops keep their own Gram path
  if (const char* env = getenv("AGP_UMMA_V2")) u.v2 = atoi(env);
  u.ps = 1;
  if (const char* env = getenv("AGP_UMMA_PS")) u.ps = atoi(env) != 0;
  if (u.v2 || getenv("AGP_UMMA_V3")) u.ps = 0;
  if (u.ps) {
    if ((e = cudaFuncSetAttribute(umma_gemm_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ps::SMEM_BYTES)) != cudaSuccess)
      return fail(err, "cudaFuncSetAttribute (ps)", e);
  }
  if (u.v2 || u.ps) {
    // pre-split copies of the two m x m right operands: [L^-1 hi | L^-1 lo | X hi | X lo], each [m][ldm]
    const size_t each = (size_t)m * ldm;
    if (each % 4) return fail(err, "v2: m * ldm must be a multiple of 4");
    if ((e = cudaMalloc(&u.Bsplit, 4 * each * sizeof(float))) != cudaSuccess) return fail(err, "cudaMalloc", e);
    cudaMemsetAsync(u.Bsplit, 0, 4 * each * sizeof(float), st);
    for (int w = 0; w < 2; ++w)
      for (int h = 0; h < 2; ++h) ok = ok && make_map(&mp->split[w][h], u.Bsplit + (size_t)(2 * w + h) * each, m, m, ldm);
    if (!ok) return fail(err, "cuTensorMapEncodeTiled failed (v2)");
    if (u.v2 && (e = cudaFuncSetAttribute(umma_gemm_nt_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_BYTES)) != cudaSuccess)
      return fail(err, "cudaFuncSetAttribute (v2)", e);
  }
  return 0;
}

float* umma_split_ptr(const UmmaLatent& u, int which, int lo) {
  if (!u.Bsplit || (which != UM_LINV && which != UM_X)) return nullptr;
  return u.Bsplit + (size_t)(2 * (which == UM_X ? 1 : 0) + (lo ? 1 : 0)) * (size_t)u.m * u.ldm;
}

int umma_presplit(std::string* err, UmmaLatent& u, int which, const float* src, cudaStream_t st) {
  if (!u.v2 && !u.ps) return 0;
  float* hi = umma_split_ptr(u, which, 0);
  float* lo = umma_split_ptr(u, which, 1);
  if (!hi) return fail(err, "v2: no pre-split buffer for this operand");
  const int64_t n4 = (int64_t)u.m * u.ldm / 4;
  split_tf32_kernel<<<(unsigned)((n4 + 255) / 256 < 592 ? (n4 + 255) / 256 : 592), 256, 0, st>>>(src, hi, lo, n4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "split_tf32_kernel", e);
  return 0;
}

void umma_latent_free(UmmaLatent& u) {
  cudaFree(u.UT);
  u.UT = nullptr;
  cudaFree(u.Bsplit);
  u.Bsplit = nullptr;
  delete (Maps*)u.tmaps;
  u.tmaps = nullptr;
}

// launches of the per-step chain carry the programmatic-stream-serialization attribute (umma_set_pdl)
static bool g_pdl = false;
void umma_set_pdl(bool on) { g_pdl = on; }
template <typename K, typename... Args>
static void launch_chain(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    if (const char* e = getenv("AGP_GEMM_GRID")) { int v = atoi(e); if (v > 0 && v < n) n = v; }   // tuning / experiments
  }
  return n;
}

// persistent-grid cap for GEMMs launched next to a resident kernel that needs its own SMs (the prefetch of the next minibatch's
// V = Knm L^-T runs on a side stream while the persistent m x m tail occupies one SM per CTA): 0 = all SMs
static int g_grid_cap = 0;
void umma_set_grid_cap(int n) { g_grid_cap = n; }
// N-tile range [lo, hi] of the NEXT triangular-B umma_gemm_nt launches (lo < 0: all tiles)
static int g_tn_lo = -1, g_tn_hi = -1;
void umma_set_tile_range(int lo, int hi) { g_tn_lo = lo; g_tn_hi = hi; }
// third-generation main loop (two conversions in flight): opt-in with AGP_UMMA_V3=1 (measured slower than the first generation)
static bool v3_on() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("AGP_UMMA_V3");
    on = (e && e[0] == '1') ? 1 : 0;
    if (on) {
      cudaFuncSetAttribute(umma_gemm3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::SMEM_BYTES);
      cudaFuncSetAttribute(umma_gemm3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::SMEM_BYTES);
    }
  }
  return on == 1;
}
static int grid_cap() { const int n = sm_count(); return (g_grid_cap > 0 && g_grid_cap < n) ? g_grid_cap : n; }

int umma_gemm_nt(std::string* err, UmmaLatent& u, int a_which, int b_which, float* C, int M, int N, const UmmaEpilogue& ep,
                 cudaStream_t st) {
  Maps* mp = (Maps*)u.tmaps;
  if (M % BM || N % BN || u.m % BK) return fail(err, "shape not a multiple of the 128 x 128 x 32 tile");
  GemmWork w{};
  w.ntm = M / BM; w.ntn = N / BN; w.nsplit = 1; w.total = w.ntm * w.ntn;
  w.total_kb = u.m / BK; w.kb_per_split = w.total_kb;
  w.tri_mode = (b_which == UM_LINV || b_which == UM_X) ? 1 : 0;
  if (g_tn_lo >= 0) {
    // tri_mode 1 enumerates the units N tile by N tile, last tile first: tile_n = ntn - 1 - u / ntm
    if (w.tri_mode != 1 || g_tn_hi >= w.ntn || g_tn_lo > g_tn_hi) return fail(err, "bad N-tile range");
    w.u0 = (w.ntn - 1 - g_tn_hi) * w.ntm;
    w.total = (g_tn_hi - g_tn_lo + 1) * w.ntm;
  }
  const int grid = w.total < grid_cap() ? w.total : grid_cap();
  if (ep.fin && !(u.ps && b_which == UM_X)) return fail(err, "fused row finish needs the pre-split kernel (V X^T product)");
  if (u.v2) {
    // v2 & 2: keep the in-kernel split of the right operand (A/B experiment); otherwise L^-1 / X arrive pre-split
    const int sp = (b_which == UM_LINV) ? 0 : (b_which == UM_X) ? 1 : -1;
    const int presplit = (sp >= 0 && !(u.v2 & 2)) ? 1 : 0;
    launch_chain(umma_gemm_nt_v2_kernel, dim3(grid), dim3(v2::NUM_THREADS), v2::SMEM_BYTES, st, mp->raw[a_which],
                 presplit ? mp->split[sp][0] : mp->raw[b_which], presplit ? mp->split[sp][1] : mp->raw[b_which], C, (int64_t)u.ldm,
                 (int64_t)0, w, ep, presplit);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return fail(err, "umma_gemm_nt_v2_kernel", e2);
    return 0;
  }
  if (u.ps && (b_which == UM_LINV || b_which == UM_X)) {
    const int sp = b_which == UM_LINV ? 0 : 1;
    UmmaRowFinish fin{};
    UmmaEpilogue epk = ep;
    if (ep.fin) {
      if (ep.mode != UMMA_EPI_STATS_ONLY || g_tn_lo >= 0 || w.tri_mode != 1) return fail(err, "fused row finish: statistics product over all N tiles only");
      fin = *ep.fin;
    }
    epk.fin = nullptr;
    launch_chain(umma_gemm_ps_kernel, dim3(grid), dim3(NUM_THREADS), ps::SMEM_BYTES, st, mp->raw[a_which], mp->split[sp][0], mp->split[sp][1], C,
                 (int64_t)u.ldm, (int64_t)0, w, epk, fin);
    cudaError_t e3 = cudaGetLastError();
    if (e3 != cudaSuccess) return fail(err, "umma_gemm_ps_kernel", e3);
    return 0;
  }
  if (v3_on())
    launch_chain(umma_gemm3_kernel<false>, dim3(grid), dim3(NUM_THREADS), v3::SMEM_BYTES, st, mp->raw[a_which], mp->raw[b_which], C, ep,
                 (const GemmGroup*)nullptr, 1, (int64_t)u.ldm, (int64_t)0, w, 0);
  else
  launch_chain(umma_gemm_nt_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, mp->raw[a_which], mp->raw[b_which], C, (int64_t)u.ldm, (int64_t)0, w, ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma_gemm_nt_kernel", e);
  return 0;
}

int umma_scale_transpose(std::string* err, UmmaLatent& u, const float* V, const double* w, double rho, const double* g, double* v1,
                         int B, int m, cudaStream_t st) {
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  launch_chain(scale_transpose_kernel, dim3(m / 32, B / 128), dim3(32, 8), 0, st, V, (int64_t)u.ldm, w, rho, u.UT, (int64_t)u.Bcap, g, v1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "scale_transpose_kernel", e);
  return 0;
}

int umma_gram(std::string* err, UmmaLatent& u, float* Gpart, int B, int m, int* n_split, cudaStream_t st) {
  Maps* mp = (Maps*)u.tmaps;
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  const int total_kb = B / BK;
  const int nt = m / BN, upper_tiles = nt * (nt + 1) / 2;
  int S = 148 / upper_tiles;     // one wave: tiles x splits <= SM count (a few CTAs spilling into a second wave doubled the time)
  S = S < 1 ? 1 : S;
  if (S > *n_split) S = *n_split;
  if (S > total_kb) S = total_kb;
  int per = (total_kb + S - 1) / S;
  S = (total_kb + per - 1) / per;
  GemmWork w{};
  w.ntm = nt; w.ntn = nt; w.nsplit = S; w.total = upper_tiles * S;
  w.total_kb = total_kb; w.kb_per_split = per; w.tri_mode = 2;
  const int grid = w.total < sm_count() ? w.total : sm_count();
  UmmaEpilogue ep{};
  ep.mode = UMMA_EPI_STORE_MIRROR;
  if (u.v2)
    launch_chain(umma_gemm_nt_v2_kernel, dim3(grid), dim3(v2::NUM_THREADS), v2::SMEM_BYTES, st, mp->ut, mp->ut, mp->ut, Gpart, (int64_t)u.ldm,
                 (int64_t)m * u.ldm, w, ep, 0);
  else if (v3_on())
    launch_chain(umma_gemm3_kernel<false>, dim3(grid), dim3(NUM_THREADS), v3::SMEM_BYTES, st, mp->ut, mp->ut, Gpart, ep, (const GemmGroup*)nullptr, 1,
                 (int64_t)u.ldm, (int64_t)m * u.ldm, w, 0);
  else
    launch_chain(umma_gemm_nt_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, mp->ut, mp->ut, Gpart, (int64_t)u.ldm, (int64_t)m * u.ldm, w, ep);
  *n_split = S;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma gram", e);
  return 0;
}

// The Gram product straight from V (umma_gram_tn_kernel): Gpart[s] = split-K partials of V^T diag(rho w) V, and v1 += V^T g
int umma_gram_tn(std::string* err, UmmaLatent& u, float* Gpart, const double* w, double rho, const double* g, double* v1, int B, int m,
                 int* n_split, cudaStream_t st) {
  Maps* mp = (Maps*)u.tmaps;
  if (!mp || !u.gram_tn) return fail(err, "umma_gram_tn: not available for this latent");
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  const int total_kb = B / BK;
  const int nt = m / BN, upper_tiles = nt * (nt + 1) / 2;
  int S = 148 / upper_tiles;     // one wave (see umma_gram)
  S = S < 1 ? 1 : S;
  if (S > *n_split) S = *n_split;
  if (S > total_kb) S = total_kb;
  int per = (total_kb + S - 1) / S;
  S = (total_kb + per - 1) / per;
  GemmWork wk{};
  wk.ntm = nt; wk.ntn = nt; wk.nsplit = S; wk.total = upper_tiles * S;
  wk.total_kb = total_kb; wk.kb_per_split = per; wk.tri_mode = 2;
  const int grid = wk.total < sm_count() ? wk.total : sm_count();
  launch_chain(umma_gram_tn_kernel<false>, dim3(grid), dim3(NUM_THREADS), tn::SMEM_BYTES, st, mp->vtn, Gpart, (int64_t)u.ldm, (int64_t)m * u.ldm, wk, w, rho, g, v1,
               (const GramTnGroup*)nullptr, 1);
  *n_split = S;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma gram tn", e);
  return 0;
}

// One part of the Gram product as its own launch: part 0 = the first diagonal tile (0, 0) over S split-K slices, part 1 = every
// other upper tile over S slices; at most grid_max persistent CTAs.  The engine runs part 0 on the critical chain (the tail's first
// kernel only needs that tile) and part 1 beside it (agp_engine.cu, split Gram).
int umma_gram_part(std::string* err, UmmaLatent& u, float* Gpart, int B, int m, int part, int S, int grid_max, cudaStream_t st) {
  Maps* mp = (Maps*)u.tmaps;
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  if (u.v2 || v3_on()) return fail(err, "split Gram launches use the first-generation kernel");
  const int total_kb = B / BK;
  const int nt = m / BN, upper_tiles = nt * (nt + 1) / 2;
  if (S < 1 || S > total_kb || upper_tiles < 2) return fail(err, "bad split count");
  const int per = (total_kb + S - 1) / S;
  if ((total_kb + per - 1) / per != S) return fail(err, "split count does not divide the k range evenly enough");
  GemmWork w{};
  w.ntm = nt; w.ntn = nt; w.nsplit = S; w.total_kb = total_kb; w.kb_per_split = per; w.tri_mode = 2;
  w.u0 = part == 0 ? 0 : S;
  w.total = part == 0 ? S : (upper_tiles - 1) * S;
  int grid = w.total < sm_count() ? w.total : sm_count();
  if (grid_max > 0 && grid > grid_max) grid = grid_max;
  UmmaEpilogue ep{};
  ep.mode = UMMA_EPI_STORE_MIRROR;
  launch_chain(umma_gemm_nt_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, mp->ut, mp->ut, Gpart, (int64_t)u.ldm, (int64_t)m * u.ldm, w, ep);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma gram part", e);
  return 0;
}
// largest S <= cap whose slices cover the k range without an empty one
int umma_gram_splits(int B, int cap) {
  const int total_kb = B / BK;
  for (int S = cap < total_kb ? cap : total_kb; S >= 1; --S) {
    const int per = (total_kb + S - 1) / S;
    if ((total_kb + per - 1) / per == S) return S;
  }
  return 1;
}

// ---- grouped launches (several latent GPs per launch) ----
int umma_groups_build(std::string* err, UmmaGroups& gs, UmmaLatent* const* lats, int n, int a_which, int b_which, float* const* C,
                      double* const* acc0, double* const* acc1, const double* const* tvec, cudaStream_t st) {
  std::vector<GemmGroup> h((size_t)n);
  for (int q = 0; q < n; ++q) {
    Maps* mp = (Maps*)lats[q]->tmaps;
    if (!mp) return fail(err, "grouped launch: latent without tensor maps");
    h[q].tmA = a_which < 0 ? mp->ut : mp->raw[a_which];
    h[q].tmB = b_which < 0 ? mp->ut : mp->raw[b_which];
    const bool psq = lats[q]->ps && (b_which == UM_LINV || b_which == UM_X);
    if (q == 0) gs.ps = psq ? 1 : 0;
    else if ((gs.ps != 0) != psq) return fail(err, "grouped launch: latents disagree on the pre-split operand");
    if (psq) { h[q].tmBhi = mp->split[b_which == UM_LINV ? 0 : 1][0]; h[q].tmBlo = mp->split[b_which == UM_LINV ? 0 : 1][1]; }
    h[q].C = C ? C[q] : nullptr;
    h[q].acc0 = acc0 ? acc0[q] : nullptr; h[q].acc1 = acc1 ? acc1[q] : nullptr; h[q].tvec = tvec ? tvec[q] : nullptr;
  }
  cudaError_t e;
  if (!gs.dev || gs.n != n) {
    if (gs.dev) cudaFree(gs.dev);
    gs.dev = nullptr;
    if ((e = cudaMalloc(&gs.dev, (size_t)n * sizeof(GemmGroup))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  }
  gs.n = n;
  if ((e = cudaMemcpyAsync(gs.dev, h.data(), (size_t)n * sizeof(GemmGroup), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(err, "cudaMemcpy", e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(err, "cudaStreamSynchronize", e);
  if ((e = cudaFuncSetAttribute(umma_gemm_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute", e);
  if ((e = cudaFuncSetAttribute(umma_gemm_grouped_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ps::SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute (grouped ps)", e);
  return 0;
}
void umma_groups_free(UmmaGroups& gs) { if (gs.dev) cudaFree(gs.dev); gs.dev = nullptr; gs.n = 0; }

int umma_gemm_nt_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& u, int b_tri, int M, int N, int epi_mode, cudaStream_t st) {
  if (!gs.dev || gs.n < 1) return fail(err, "grouped launch without groups");
  if (M % BM || N % BN || u.m % BK) return fail(err, "shape not a multiple of the 128 x 128 x 32 tile");
  GemmWork w{};
  w.ntm = M / BM; w.ntn = N / BN; w.nsplit = 1; w.total = w.ntm * w.ntn;
  w.total_kb = u.m / BK; w.kb_per_split = w.total_kb;
  w.tri_mode = b_tri ? 1 : 0;
  const int all = w.total * gs.n;
  const int grid = all < grid_cap() ? all : grid_cap();
  if (gs.ps) {
    launch_chain(umma_gemm_grouped_ps_kernel, dim3(grid), dim3(NUM_THREADS), ps::SMEM_BYTES, st, (const GemmGroup*)gs.dev, gs.n, (int64_t)u.ldm, (int64_t)0, w, epi_mode);
  } else if (v3_on()) {
    static const CUtensorMap none{};
    launch_chain(umma_gemm3_kernel<true>, dim3(grid), dim3(NUM_THREADS), v3::SMEM_BYTES, st, none, none, (float*)nullptr, UmmaEpilogue{}, (const GemmGroup*)gs.dev, gs.n,
                 (int64_t)u.ldm, (int64_t)0, w, epi_mode);
  } else
  launch_chain(umma_gemm_grouped_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, (const GemmGroup*)gs.dev, gs.n, (int64_t)u.ldm, (int64_t)0, w, epi_mode);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma_gemm_grouped_kernel", e);
  return 0;
}

int umma_gram_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& u, int B, int m, int* n_split, cudaStream_t st) {
  if (!gs.dev || gs.n < 1) return fail(err, "grouped launch without groups");
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  const int total_kb = B / BK;
  const int nt = m / BN, upper_tiles = nt * (nt + 1) / 2;
  // split-K: wave-quantised cost of S slices per tile over all groups (+ a small charge per partial that combine_kernel re-reads)
  const int sms = sm_count();
  int bestS = 1; long best = -1;
  for (int S = 1; S <= *n_split && S <= total_kb; ++S) {
    const int per = (total_kb + S - 1) / S;
    const int Su = (total_kb + per - 1) / per;
    const long units = (long)upper_tiles * gs.n * Su;
    const long cost = ((units + sms - 1) / sms) * (long)per + 2L * Su;
    if (best < 0 || cost < best) { best = cost; bestS = Su; }
  }
  const int per = (total_kb + bestS - 1) / bestS;
  const int S = (total_kb + per - 1) / per;
  GemmWork w{};
  w.ntm = nt; w.ntn = nt; w.nsplit = S; w.total = upper_tiles * S;
  w.total_kb = total_kb; w.kb_per_split = per; w.tri_mode = 2;
  const int all = w.total * gs.n;
  const int grid = all < sms ? all : sms;
  if (v3_on()) {
    static const CUtensorMap none{};
    launch_chain(umma_gemm3_kernel<true>, dim3(grid), dim3(NUM_THREADS), v3::SMEM_BYTES, st, none, none, (float*)nullptr, UmmaEpilogue{}, (const GemmGroup*)gs.dev, gs.n,
                 (int64_t)u.ldm, (int64_t)m * u.ldm, w, (int)UMMA_EPI_STORE_MIRROR);
  } else
  launch_chain(umma_gemm_grouped_kernel, dim3(grid), dim3(NUM_THREADS), SMEM_BYTES, st, (const GemmGroup*)gs.dev, gs.n, (int64_t)u.ldm, (int64_t)m * u.ldm, w,
               (int)UMMA_EPI_STORE_MIRROR);
  *n_split = S;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma gram (grouped)", e);
  return 0;
}

// Grouped Gram product straight from V (umma_gram_tn_kernel<true>): no scale-transpose pass per latent
int umma_gram_tn_groups_build(std::string* err, UmmaGroups& gs, UmmaLatent* const* lats, int n, float* const* Gpart, const double* const* w,
                              const double* const* g, double* const* v1, cudaStream_t st) {
  std::vector<GramTnGroup> h((size_t)n);
  for (int q = 0; q < n; ++q) {
    Maps* mp = (Maps*)lats[q]->tmaps;
    if (!mp || !lats[q]->gram_tn) return fail(err, "grouped Gram from V: not available for this latent");
    h[q].tmV = mp->vtn; h[q].C = Gpart[q]; h[q].w = w[q]; h[q].g = g[q]; h[q].v1 = v1[q];
  }
  cudaError_t e;
  if (!gs.dev || gs.n != n) {
    if (gs.dev) cudaFree(gs.dev);
    gs.dev = nullptr;
    if ((e = cudaMalloc(&gs.dev, (size_t)n * sizeof(GramTnGroup))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  }
  gs.n = n;
  if ((e = cudaMemcpyAsync(gs.dev, h.data(), (size_t)n * sizeof(GramTnGroup), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(err, "cudaMemcpy", e);
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(err, "cudaStreamSynchronize", e);
  if ((e = cudaFuncSetAttribute(umma_gram_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tn::SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute (grouped gram tn)", e);
  return 0;
}

int umma_gram_tn_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& u, double rho, int B, int m, int* n_split, cudaStream_t st) {
  if (!gs.dev || gs.n < 1) return fail(err, "grouped launch without groups");
  if (B % BM || m % BN) return fail(err, "shape not a multiple of the tile");
  const int total_kb = B / BK;
  const int nt = m / BN, upper_tiles = nt * (nt + 1) / 2;
  // split-K as in umma_gram_grouped: wave-quantised k-blocks per CTA, plus the unit's drain (the worker groups of this kernel convert AND
  // drain, so an epilogue is not hidden behind the next unit: ~4 k-blocks' worth) and the partial that combine_kernel re-reads
  const int sms = sm_count();
  int bestS = 1; long best = -1;
  for (int S = 1; S <= *n_split && S <= total_kb; ++S) {
    const int per = (total_kb + S - 1) / S;
    const int Su = (total_kb + per - 1) / per;
    const long units = (long)upper_tiles * gs.n * Su;
    const long cost = ((units + sms - 1) / sms) * (long)(per + 4) + 2L * Su;
    if (best < 0 || cost < best) { best = cost; bestS = Su; }
  }
  const int per = (total_kb + bestS - 1) / bestS;
  const int S = (total_kb + per - 1) / per;
  GemmWork w{};
  w.ntm = nt; w.ntn = nt; w.nsplit = S; w.total = upper_tiles * S;
  w.total_kb = total_kb; w.kb_per_split = per; w.tri_mode = 2;
  const int all = w.total * gs.n;
  const int grid = all < sms ? all : sms;
  static const CUtensorMap none{};
  launch_chain(umma_gram_tn_kernel<true>, dim3(grid), dim3(NUM_THREADS), tn::SMEM_BYTES, st, none, (float*)nullptr, (int64_t)u.ldm, (int64_t)m * u.ldm, w,
               (const double*)nullptr, rho, (const double*)nullptr, (double*)nullptr, (const GramTnGroup*)gs.dev, gs.n);
  *n_split = S;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma gram tn (grouped)", e);
  return 0;
}

// ---- experimental Newton-Schulz refinement (see agp_umma.h) ----
struct NsMaps { CUtensorMap P, Y, W[2], T; };

// T = I - Y P with the product in fp64 on DMMA (mma.sync.m8n8k4.f64): Y fp32 [m][ldy], P fp64 [m][ldp] symmetric, T fp32 [m][ldt];
// *resid2 += |T|_F^2.  The residual is where the accuracy of the refinement is decided (an fp32 / 3xTF32 residual floors at
// eps_fp32 * cond(P), profiles/r1/studies/newton_schulz_precision_study.txt), the correction product may be low precision.
// One CTA (8 warps) per 64 x 32 tile of T, 32-deep k-steps staged in shared memory (leading dimension 36 = 4 mod 16: conflict-free
// 8x4 fragment loads, as in agp_tail2.cuh); warp w owns rows 8w..8w+7 of the tile.
constexpr int NSLD = 36, NSK = 32, NSTN = 32;   // 64 x 32 tiles of T: 128 CTAs at m = 512 (a 64 x 64 tiling would leave 84 SMs idle)
__global__ void __launch_bounds__(256) ns_resid_f64_kernel(const float* __restrict__ Y, int64_t ldy, const double* __restrict__ P, int64_t ldp,
                                                           float* __restrict__ T, int64_t ldt, int m, double* __restrict__ resid2) {
  __shared__ __align__(16) double sA[64 * NSLD], sB[NSTN * NSLD];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * NSTN;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5, r = lane >> 2, kk = lane & 3;
  double acc[NSTN / 8][2];
#pragma unroll
  for (int nb = 0; nb < NSTN / 8; ++nb) { acc[nb][0] = 0.0; acc[nb][1] = 0.0; }
  for (int k0 = 0; k0 < m; k0 += NSK) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {                 // Y tile: 64 x 32 elements as 1024 pairs
      const int e = t + u * 256, row = e >> 4, c2 = (e & 15) * 2;
      const float2 y = *reinterpret_cast<const float2*>(Y + (int64_t)(i0 + row) * ldy + k0 + c2);
      *reinterpret_cast<double2*>(sA + row * NSLD + c2) = make_double2((double)y.x, (double)y.y);
    }
#pragma unroll
    for (int u = 0; u < NSTN / 16; ++u) {         // P tile: NSTN x 32 elements
      const int e = t + u * 256, row = e >> 4, c2 = (e & 15) * 2;
      *reinterpret_cast<double2*>(sB + row * NSLD + c2) = *reinterpret_cast<const double2*>(P + (int64_t)(j0 + row) * ldp + k0 + c2);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < NSK; k += 4) {
      const double a = sA[(8 * w + r) * NSLD + k + kk];
#pragma unroll
      for (int nb = 0; nb < NSTN / 8; ++nb) {
        const double b = sB[(8 * nb + r) * NSLD + k + kk];   // P[j][k] = P[k][j]
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(acc[nb][0]), "+d"(acc[nb][1]) : "d"(a), "d"(b));
      }
    }
  }
  double sq = 0.0;
  const int row = i0 + 8 * w + r;
#pragma unroll
  for (int nb = 0; nb < NSTN / 8; ++nb) {
    const int col = j0 + 8 * nb + 2 * kk;
    const double v0 = (row == col ? 1.0 : 0.0) - acc[nb][0], v1 = (row == col + 1 ? 1.0 : 0.0) - acc[nb][1];
    sq = fma(v0, v0, fma(v1, v1, sq));
    *reinterpret_cast<float2*>(T + (int64_t)row * ldt + col) = make_float2((float)v0, (float)v1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) atomicAdd(resid2, sq);
}

// Out = (S + S^T) / 2, 32 x 32 tiles, one CTA per tile pair (bi <= bj); Out != S.  Rounding leaves an antisymmetric part in the
// iterate that the Newton-Schulz map doubles every pass; removed once per refinement it never gets past ~1e-6.
__global__ void __launch_bounds__(256) ns_symmetrize_kernel(const float* __restrict__ S, float* __restrict__ Out, int64_t ld, int m) {
  __shared__ float a[32][33], b[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    a[rr][tx] = S[(int64_t)(bi * 32 + rr) * ld + bj * 32 + tx];
    b[rr][tx] = S[(int64_t)(bj * 32 + rr) * ld + bi * 32 + tx];
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    Out[(int64_t)(bi * 32 + rr) * ld + bj * 32 + tx] = 0.5f * (a[rr][tx] + b[tx][rr]);
    Out[(int64_t)(bj * 32 + rr) * ld + bi * 32 + tx] = 0.5f * (b[rr][tx] + a[tx][rr]);
  }
}

int umma_ns_alloc(std::string* err, UmmaNs& ns, int m, cudaStream_t st) {
  if (m < 128 || m % 128) return fail(err, "Newton-Schulz path needs m % 128 == 0");
  ns.m = m; ns.ldm = m;
  const size_t each = (size_t)m * ns.ldm;
  cudaError_t e;
  if ((e = cudaMalloc(&ns.buf, 5 * each * sizeof(float))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  if ((e = cudaMalloc(&ns.resid, 64 * sizeof(double))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  if ((e = cudaMalloc(&ns.P64, each * sizeof(double))) != cudaSuccess) return fail(err, "cudaMalloc", e);
  cudaMemsetAsync(ns.buf, 0, 5 * each * sizeof(float), st);
  cudaMemsetAsync(ns.resid, 0, 64 * sizeof(double), st);
  NsMaps* mp = new NsMaps();
  ns.maps = mp;
  bool ok = make_map(&mp->P, ns.P(), m, m, ns.ldm) && make_map(&mp->Y, ns.Y(), m, m, ns.ldm) &&
            make_map(&mp->W[0], ns.W(0), m, m, ns.ldm) && make_map(&mp->W[1], ns.W(1), m, m, ns.ldm) &&
            make_map(&mp->T, ns.T(), m, m, ns.ldm);
  if (!ok) return fail(err, "cuTensorMapEncodeTiled failed (ns)");
  if ((e = cudaFuncSetAttribute(umma_gemm_nt_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM_BYTES)) != cudaSuccess)
    return fail(err, "cudaFuncSetAttribute (v2)", e);
  return 0;
}

void umma_ns_free(UmmaNs& ns) {
  cudaFree(ns.buf); ns.buf = nullptr;
  cudaFree(ns.resid); ns.resid = nullptr;
  cudaFree(ns.P64); ns.P64 = nullptr;
  delete (NsMaps*)ns.maps; ns.maps = nullptr;
}

int umma_gemm_sigma(std::string* err, UmmaLatent& u, UmmaNs& ns, int a_which, int M, const UmmaEpilogue& ep, cudaStream_t st) {
  Maps* mp = (Maps*)u.tmaps;
  NsMaps* nm = (NsMaps*)ns.maps;
  if (!mp || !nm || !u.v2) return fail(err, "umma_gemm_sigma needs the v2 kernel (AGP_UMMA_V2) and an allocated Newton-Schulz state");
  if (M % BM || ns.m != u.m || ns.ldm != u.ldm) return fail(err, "umma_gemm_sigma: shape mismatch");
  GemmWork w{};
  w.ntm = M / BM; w.ntn = u.m / BN; w.nsplit = 1; w.total = w.ntm * w.ntn;
  w.total_kb = u.m / BK; w.kb_per_split = w.total_kb; w.tri_mode = 0;
  const int grid = w.total < sm_count() ? w.total : sm_count();
  launch_chain(umma_gemm_nt_v2_kernel, dim3(grid), dim3(v2::NUM_THREADS), v2::SMEM_BYTES, st, mp->raw[a_which], nm->Y, nm->Y, (float*)nullptr,
               (int64_t)u.ldm, (int64_t)0, w, ep, 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma_gemm_sigma", e);
  return 0;
}

int umma_ns_iterate(std::string* err, UmmaNs& ns, int iters, int mode, const double* P64, int64_t ldp, cudaStream_t st) {
  NsMaps* mp = (NsMaps*)ns.maps;
  if (!mp || iters < 0 || iters > 64) return fail(err, "umma_ns_iterate: bad state / iteration count");
  if (!P64) { P64 = ns.P64; ldp = ns.m; }
  GemmWork w{};
  w.ntm = ns.m / BM; w.ntn = ns.m / BN; w.nsplit = 1; w.total = w.ntm * w.ntn;
  w.total_kb = ns.m / BK; w.kb_per_split = w.total_kb; w.tri_mode = 0;
  const int grid = w.total < sm_count() ? w.total : sm_count();
  cudaMemsetAsync(ns.resid, 0, 64 * sizeof(double), st);
  // the iterate travels Y -> W0 -> W1 -> W0 ... and returns to Y at the end: every launch sees the same addresses on every call
  const float* src = ns.Y();
  const CUtensorMap* src_map = &mp->Y;
  for (int it = 0; it < iters; ++it) {
    float* dst = ns.W(it & 1);
    // T = I - Y P   (= (I - P Y)^T for symmetric P, Y: exactly the [n][k] operand the second product needs)
    if (mode & 1) {
      ns_resid_f64_kernel<<<dim3(ns.m / NSTN, ns.m / 64), 256, 0, st>>>(src, (int64_t)ns.ldm, P64, ldp, ns.T(), (int64_t)ns.ldm, ns.m, ns.resid + it);
    } else {
      UmmaEpilogue e1{};
      e1.mode = UMMA_EPI_EYE_MINUS; e1.acc0 = ns.resid + it;
      launch_chain(umma_gemm_nt_v2_kernel, dim3(grid), dim3(v2::NUM_THREADS), v2::SMEM_BYTES, st, *src_map, mp->P, mp->P, ns.T(),
                   (int64_t)ns.ldm, (int64_t)0, w, e1, 0);
    }
    UmmaEpilogue e2{};   // Y' = Y + Y (I - P Y)
    e2.mode = UMMA_EPI_ADD; e2.cin = src;
    launch_chain(umma_gemm_nt_v2_kernel, dim3(grid), dim3(v2::NUM_THREADS), v2::SMEM_BYTES, st, *src_map, mp->T, mp->T, dst,
                 (int64_t)ns.ldm, (int64_t)0, w, e2, 0);
    src = dst;
    src_map = &mp->W[it & 1];
  }
  if (iters > 0) {
    if (mode & 2) ns_symmetrize_kernel<<<dim3(ns.m / 32, ns.m / 32), 256, 0, st>>>(src, ns.Y(), (int64_t)ns.ldm, ns.m);
    else cudaMemcpyAsync(ns.Y(), src, (size_t)ns.m * ns.ldm * sizeof(float), cudaMemcpyDeviceToDevice, st);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(err, "umma_ns_iterate", e);
  return 0;
}

}  // namespace agp
