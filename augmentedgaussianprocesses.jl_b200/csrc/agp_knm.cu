// agp_knm.cu -- kernel-matrix construction K_nm = k(X_b, Z) (gpblocks/latentgp.jl:210 of the reference; kernel forms of
// KernelFunctions.jl: SqExponential, Matern-3/2, Matern-5/2 with ScaleTransform and a variance factor) for sm_100a.
//
// The kernel is bound by its 4*B*m output bytes (C2: 16.8 MB of the 17.96 MB algorithmic traffic), so everything is
// organised around streaming that tile out of the SM at full width:
//   * one CTA = one 128 x 128 output tile, two or three CTAs resident per SM (gather / MMA / store phases of different
//     tiles overlap);
//   * A operand: 128 minibatch rows gathered straight from the resident X (one 16-byte-vector row segment per thread,
//     no shared-memory staging), split into TF32 hi / lo in registers and written to TENSOR MEMORY (tcgen05.st);
//   * B operand: the inducing points Z, pre-split once into hi / lo (they only change with the hyper-parameters), loaded
//     by TMA (cp.async.bulk.tensor.2d, 128B swizzle) into shared memory;
//   * x.z in 3xTF32 on the tensor core (tcgen05.mma kind::tf32, A from TMEM, accumulator in TMEM): lo*hi + hi*lo + hi*hi;
//   * epilogue: tcgen05.ld -> d2 = s^2 (|x|^2 + |z|^2 - 2 x.z) -> kernel function -> 128B-swizzled shared staging ->
//     TMA store (cp.async.bulk.tensor.2d.global.shared::cta) of full 128-byte row segments.
// Supports D <= 128 (all k-blocks of a tile stay resident); the engine keeps the SIMT kernel for larger D.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "agp_gemm_simt.cuh"  // kfn_eval
#include "agp_tc.cuh"
#include "agp_umma.h"

namespace agp {

namespace {
using namespace tc;

constexpr int KBM = 128, KBN = 128, KBK = 32;
constexpr int KTILE = KBN * KBK * 4;                 // 16 KB: 128 rows of Z x 32 k (one swizzle atom wide)
constexpr int K_STAGE_BYTES = 8 * 4096;              // one 32 x 32 fp32 box (128B-swizzled) per worker warp
constexpr int K_WORKERS = 256;                       // warps 2..9
constexpr int K_THREADS = 64 + K_WORKERS;
constexpr int K_MAX_KB = 4;                          // D <= 128
constexpr int K_TB_STRIDE = 144;                     // bytes per row of the warp-private transposition buffer (conflict-free)

__host__ __device__ constexpr int knm_smem_bytes(int nkb) { return nkb * 2 * KTILE + K_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
#define TMEM_ST8(taddr, r)                                                                                      \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"                  \
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) \
               : "memory")
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrt_approx(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#ifdef AGP_KNM_TIMING
__device__ unsigned long long agp_knm_t[1024 * 8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define KNM_T(slot) do { if (lane == 0 && warp == 2) agp_knm_t[((blockIdx.y * gridDim.x + blockIdx.x) & 1023) * 8 + (slot)] = gtimer(); } while (0)
#else
#define KNM_T(slot) do {} while (0)
#endif

struct KnmArgs {
  const float* X; int64_t ldx;       // resident inputs, row-major [n][ldx]
  const int64_t* gather;             // minibatch row -> physical row (nullptr: identity)
  const float* xx;                   // squared norms by minibatch row
  const float* zz;                   // squared norms of the inducing points [m]
  int Dp, nkb;                       // padded feature count (multiple of 4), k-blocks of 32
  int tiles_per_cta;                 // consecutive 128-column tiles handled by one CTA (the gathered A tile is reused)
  float scale2, variance;
};

// per-thread constants of the kernel-function epilogue.  KIND 0 (SqExponential): k = exp2(lv + c1 (|x|^2 + |z|^2) + c2 x.z),
// c1 = -s^2 log2(e) / 2, c2 = s^2 log2(e), lv = log2(variance); d2 >= 0 clamps the exponent at lv.
// KIND 1 / 2 (Matern 3/2, 5/2): d2 = max(s^2 (|x|^2 + |z|^2 - 2 x.z), 0), a = sqrt(3 | 5) sqrt(d2), k = variance poly(a) exp(-a).
template <int KIND>
struct KfnFast {
  float c1, c2, base, lv, var;
  __device__ __forceinline__ KfnFast(float scale2, float variance, float xr) {
    var = variance;
    if (KIND == 0) {
      c1 = -0.5f * 1.4426950408889634f * scale2; c2 = 1.4426950408889634f * scale2; lv = log2f(variance);
      base = fmaf(xr, c1, lv);
    } else {
      c1 = scale2; c2 = -2.f * scale2; lv = 0.f;
      base = xr * scale2;
    }
  }
  __device__ __forceinline__ float operator()(float acc, float zc) const {
    const float u = fmaf(zc, c1, base);
    if (KIND == 0) return ex2_approx(fminf(fmaf(acc, c2, u), lv));
    const float d2 = fmaxf(fmaf(acc, c2, u), 0.f);
    const float d = sqrt_approx(d2);
    if (KIND == 1) { const float t = 1.7320508075688772f * d; return var * (1.f + t) * ex2_approx(-1.4426950408889634f * t); }
    const float t = 2.23606797749979f * d;
    return var * (1.f + t + 1.6666666666666667f * d2) * ex2_approx(-1.4426950408889634f * t);
  }
};

// grid = (column groups, row tiles).  One CTA: gather 128 minibatch rows once, then loop over its column tiles.
template <int KIND>
__global__ void __launch_bounds__(K_THREADS, 2) knm_umma_kernel(const __grid_constant__ CUtensorMap tmZhi, const __grid_constant__ CUtensorMap tmZlo,
                                                               const __grid_constant__ CUtensorMap tmOut, const KnmArgs a) {
  const int tile_m = blockIdx.y, tile_n0 = blockIdx.x * a.tiles_per_cta, ntile = a.tiles_per_cta;
  const int nkb = a.nkb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_bytes = nkb * 2 * KTILE;
  const uint32_t stage_base = smem_base + b_bytes;
  uint8_t* stage_gen = smem_gen + b_bytes;
  const uint32_t bars = stage_base + K_STAGE_BYTES;
  const uint32_t b_full = bars, a_ready = bars + 8, tmem_full = bars + 16, acc_free = bars + 24;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(stage_gen + K_STAGE_BYTES + 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (KBN + 64 * nkb <= 256) ? 256u : 512u;
  KNM_T(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmZhi); tma_prefetch_desc(&tmZlo); tma_prefetch_desc(&tmOut);
    mbar_init(b_full, 1); mbar_init(a_ready, 128); mbar_init(tmem_full, 1); mbar_init(acc_free, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  constexpr uint32_t TMEM_A0 = KBN;   // A_hi / A_lo of k-block kb: columns TMEM_A0 + 64 kb (+32)

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: pre-split inducing points of one column tile (all k-blocks); the single buffer is free again
      // as soon as the MMAs of the previous tile have retired (tmem_full), i.e. during that tile's epilogue =====
      for (int t = 0; t < ntile; ++t) {
        if (t > 0) mbar_wait(tmem_full, (t - 1) & 1);
        mbar_expect_tx(b_full, (uint32_t)b_bytes);
        for (int kb = 0; kb < nkb; ++kb) {
          tma_load_2d(smem_base + (2 * kb + 0) * KTILE, &tmZhi, b_full, kb * KBK, (tile_n0 + t) * KBN);
          tma_load_2d(smem_base + (2 * kb + 1) * KTILE, &tmZlo, b_full, kb * KBK, (tile_n0 + t) * KBN);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    mbar_wait(a_ready, 0);
    for (int t = 0; t < ntile; ++t) {
      mbar_wait(b_full, t & 1);
      if (t > 0) mbar_wait(acc_free, (t - 1) & 1);      // the epilogue has read the previous accumulator out of TMEM
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_tf32(KBM, KBN);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t b_hi = smem_base + (2 * kb) * KTILE, b_lo = b_hi + KTILE;
          const uint32_t a_hi = tmem_base + TMEM_A0 + kb * 64, a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < KBK / 8; ++kk) {
            const uint64_t dbh = make_desc(b_hi + kk * 32), dbl = make_desc(b_lo + kk * 32);
            tc_mma_tf32_ts(tmem_base, a_lo + kk * 8, dbh, idesc, (kb > 0 || kk > 0) ? 1u : 0u);   // small terms first
            tc_mma_tf32_ts(tmem_base, a_hi + kk * 8, dbl, idesc, 1u);
            tc_mma_tf32_ts(tmem_base, a_hi + kk * 8, dbh, idesc, 1u);
          }
        }
        tc_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ===== workers: warps 2..9.  q = TMEM lane quarter this warp may touch, half = which pair of 32-column chunks =====
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row0 = tile_m * KBM + q * 32;
    const float xr = a.xx[row0 + lane];
    if (half == 0) {
      // gather + split: minibatch rows row0 .. row0+31 -> TF32 hi / lo -> tensor memory.  Loads are coalesced (8 lanes x
      // 16 B = one 128-byte row segment, 4 rows per instruction) and transposed to one-row-per-thread (= one TMEM lane
      // per thread) through a warp-private padded buffer.
      const int64_t prow_lane = a.gather ? a.gather[row0 + lane] : (int64_t)(row0 + lane);
      const int sub = lane >> 3, ch = lane & 7, nvec = a.Dp >> 2;
      uint8_t* tb = stage_gen + q * 8192;                  // staging boxes of warps (q, half 0/1): unused until the epilogue
      const float* rowp[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rowp[i] = a.X + __shfl_sync(0xffffffffu, prow_lane, 4 * i + sub) * a.ldx;
      KNM_T(1);
      for (int kb = 0; kb < nkb; ++kb) {
        float4 v[8];
        const bool inb = kb * 8 + ch < nvec;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = inb ? __ldg(reinterpret_cast<const float4*>(rowp[i]) + kb * 8 + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(tb + (4 * i + sub) * K_TB_STRIDE + ch * 16) = v[i];
        __syncwarp();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + TMEM_A0 + kb * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x0 = *reinterpret_cast<const float4*>(tb + lane * K_TB_STRIDE + (2 * c) * 16);
          const float4 x1 = *reinterpret_cast<const float4*>(tb + lane * K_TB_STRIDE + (2 * c + 1) * 16);
          const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
          uint32_t h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = tf32_rna(xs[e]);
            h[e] = __float_as_uint(t); l[e] = __float_as_uint(tf32_rna(xs[e] - t));
          }
          TMEM_ST8(ta + 8 * c, h);
          TMEM_ST8(ta + 32 + 8 * c, l);
        }
        __syncwarp();
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      fence_proxy_async();      // the transposition buffer is later rewritten as TMA-store staging
      mbar_arrive(a_ready);
      KNM_T(2);
    }
    const KfnFast<KIND> kfn(a.scale2, a.variance, xr);
    uint8_t* box = stage_gen + (q * 2 + half) * 4096;       // 32 rows x 128 B, 128B-swizzled (chunk ^= row & 7)
    const uint32_t box_s = stage_base + (q * 2 + half) * 4096;
    for (int t = 0; t < ntile; ++t) {
      mbar_wait(tmem_full, t & 1);
      tc_fence_after();
      if (t == 0) KNM_T(3);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;                          // 32-column chunk of the tile
        const int col0 = (tile_n0 + t) * KBN + c * 32;
        uint32_t r[32];
        TMEM_LD32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const float zc_lane = __ldg(a.zz + col0 + lane);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc == 1) {                                        // this warp is done with the accumulator of tile t
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_free);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has left the box
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = kfn(__uint_as_float(r[4 * j + e]), __shfl_sync(0xffffffffu, zc_lane, 4 * j + e));
          *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, box_s, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    KNM_T(4);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
    KNM_T(5);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// Z [m][ldz] fp32 -> hi / lo [m][Kp] (zero padded to a multiple of 32 columns)
__global__ void split_z_kernel(const float* __restrict__ Z, int64_t ldz, int m, int D, float* __restrict__ hi, float* __restrict__ lo, int Kp) {
  const int i = blockIdx.x, k = threadIdx.x;
  if (i >= m || k >= Kp) return;
  const float v = (k < D) ? Z[(int64_t)i * ldz + k] : 0.f;
  const float h = tf32_rna(v);
  hi[(int64_t)i * Kp + k] = h;
  lo[(int64_t)i * Kp + k] = tf32_rna(v - h);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn knm_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
bool knm_make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = knm_get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct KnmMaps { CUtensorMap zhi, zlo, out; };

int kfail(std::string* err, const char* what, cudaError_t e = cudaSuccess) {
  *err = std::string("tcgen05 K_nm path: ") + what + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string());
  return 2;  // AGP_ERR_CUDA
}

}  // namespace

bool umma_knm_shape_ok(int m, int Bcap, int D) { return umma_shape_ok(m, Bcap) && D >= 1 && D <= KBK * K_MAX_KB; }

int umma_knm_setup(std::string* err, UmmaKnm& k, const float* Z, int64_t ldz, int m, int D, float* Knm, int64_t ldk, int Bcap, cudaStream_t st) {
  k.m = m; k.D = D; k.Kp = (D + KBK - 1) / KBK * KBK;
  cudaError_t e;
  if (!k.Zhi) {
    if ((e = cudaMalloc(&k.Zhi, (size_t)m * k.Kp * sizeof(float))) != cudaSuccess) return kfail(err, "cudaMalloc", e);
    if ((e = cudaMalloc(&k.Zlo, (size_t)m * k.Kp * sizeof(float))) != cudaSuccess) return kfail(err, "cudaMalloc", e);
  }
  split_z_kernel<<<m, k.Kp, 0, st>>>(Z, ldz, m, D, k.Zhi, k.Zlo, k.Kp);
  if ((e = cudaGetLastError()) != cudaSuccess) return kfail(err, "split_z_kernel", e);
  if (!k.maps) {
    KnmMaps* mp = new KnmMaps();
    k.maps = mp;
    bool ok = knm_make_map(&mp->zhi, k.Zhi, m, k.Kp, k.Kp, KBK, KBN) && knm_make_map(&mp->zlo, k.Zlo, m, k.Kp, k.Kp, KBK, KBN) &&
              knm_make_map(&mp->out, Knm, Bcap, m, ldk, 32, 32);
    if (!ok) return kfail(err, "cuTensorMapEncodeTiled failed");
    if ((e = cudaFuncSetAttribute(knm_umma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, knm_smem_bytes(K_MAX_KB))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(knm_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, knm_smem_bytes(K_MAX_KB))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(knm_umma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, knm_smem_bytes(K_MAX_KB))) != cudaSuccess)
      return kfail(err, "cudaFuncSetAttribute", e);
  }
  return 0;
}

void umma_knm_free(UmmaKnm& k) {
  cudaFree(k.Zhi); cudaFree(k.Zlo);
  k.Zhi = k.Zlo = nullptr;
  delete (KnmMaps*)k.maps;
  k.maps = nullptr;
}

int umma_knm(std::string* err, UmmaKnm& k, const float* X, int64_t ldx, int Dp, const int64_t* gather, const float* xx, const float* zz,
             int B, int kind, double scale2, double variance, cudaStream_t st) {
  KnmMaps* mp = (KnmMaps*)k.maps;
  if (!mp) return kfail(err, "not set up");
  if (B % KBM || k.m % KBN) return kfail(err, "shape not a multiple of the 128 x 128 tile");
  KnmArgs a{};
  a.X = X; a.ldx = ldx; a.gather = gather; a.xx = xx; a.zz = zz; a.Dp = Dp; a.nkb = k.Kp / KBK;
  a.scale2 = (float)scale2; a.variance = (float)variance;
  // column tiles per CTA: the gathered A tile is reused across them; split the columns only as far as needed to fill the chip
  const int nct = k.m / KBN, nrt = B / KBM;
  int groups = 1;
  while (groups < nct && nrt * groups < 148 && nct % (groups * 2) == 0) groups *= 2;
  if (k.force_groups > 0 && nct % k.force_groups == 0) groups = k.force_groups;
  a.tiles_per_cta = nct / groups;
  const dim3 grid(groups, nrt);
  const int smem = knm_smem_bytes(a.nkb);
  if (kind == 0) knm_umma_kernel<0><<<grid, K_THREADS, smem, st>>>(mp->zhi, mp->zlo, mp->out, a);
  else if (kind == 1) knm_umma_kernel<1><<<grid, K_THREADS, smem, st>>>(mp->zhi, mp->zlo, mp->out, a);
  else knm_umma_kernel<2><<<grid, K_THREADS, smem, st>>>(mp->zhi, mp->zlo, mp->out, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return kfail(err, "knm_umma_kernel", e);
  return 0;
}

}  // namespace agp
