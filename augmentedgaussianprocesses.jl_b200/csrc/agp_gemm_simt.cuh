// SIMT (CUDA-core) tiled GEMM used for (a) the fp64 "exact" precision mode, (b) the fp32 SIMT mode and
// (c) every fp64 m x m contraction of the tail (blocked Cholesky / triangular inverse / Sigma = X^T X).
// The tensor-core (tcgen05) path for the B x m contractions lives in agp_umma.cu.
//
// Layout rules shared by every matrix in the engine: row-major, leading dimension a multiple of 4
// elements, base pointers 16-byte aligned, padding columns hold zeros (never written).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace agp {

template <typename T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int W = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int W = 2; };

enum GemmEpi : int { EPI_PLAIN = 0, EPI_KERNELFN = 1 };

template <typename T>
struct GemmParams {
  // A_T == false: A is [M][K] (K contiguous);  A_T == true: A is [K][M] (M contiguous)
  const T* A; int64_t lda;
  // B_T == false: B is [N][K] (K contiguous, "NT");  B_T == true: B is [K][N] (N contiguous, "NN")
  const T* B; int64_t ldb;
  T* C; int64_t ldc;
  int M, N, K;
  const int64_t* a_gather;   // optional (A_T == false): logical row r reads physical row a_gather[r]
  const double* k_scale;     // optional (A_T == true): row k of A is multiplied by k_scale[k] * k_scale_mul
  double k_scale_mul;
  int k_chunk;               // > 0: split-K, blockIdx.z handles k in [z*k_chunk, (z+1)*k_chunk)
  int64_t zs_a, zs_b, zs_c;  // per-blockIdx.z element offsets (batched GEMM / split-K partial outputs)
  double alpha, beta;        // C = alpha * acc + beta * C
  int lower_only;            // skip tiles that lie strictly above the diagonal
  int k_from_diag;           // TN product of lower-triangular factors: start k at max(row0, col0)
  int k_to_diag;             // NT with a lower-triangular B ([N][K], zero for k > n): stop k at the tile's last column
  // EPI_KERNELFN: the k-loop accumulates sum_k (a_k - b_k)^2 instead of the product, C = variance * base(scale2 * acc).
  // (The differences keep a relative rounding error on the squared distance; |x|^2 + |z|^2 - 2 x.z has an absolute one of
  // eps (|x|^2 + |z|^2), which V = K_nm L^-T amplifies by sqrt(cond K_mm) - the tcgen05 K_nm kernel accepts that for D > 16.)
  const T* xx; const T* zz;  // unused by this kernel since the difference form (kept: callers fill them for the tcgen05 path)
  double scale2, variance; int kernel_kind;
  int xx_direct;
};

__device__ __forceinline__ float kfn_eval(int kind, float d2, float var) {
  d2 = fmaxf(d2, 0.f);
  if (kind == 0) return var * __expf(-0.5f * d2);
  float d = sqrtf(d2);
  if (kind == 1) { float a = 1.7320508075688772f * d; return var * (1.f + a) * __expf(-a); }
  float a = 2.23606797749979f * d;
  return var * (1.f + a + 1.6666666666666667f * d2) * __expf(-a);
}
__device__ __forceinline__ double kfn_eval(int kind, double d2, double var) {
  d2 = fmax(d2, 0.0);
  if (kind == 0) return var * exp(-0.5 * d2);
  double d = sqrt(d2);
  if (kind == 1) { double a = 1.7320508075688772 * d; return var * (1.0 + a) * exp(-a); }
  double a = 2.23606797749979 * d;
  return var * (1.0 + a + (5.0 / 3.0) * d2) * exp(-a);
}

// 256 threads (16 x 16).  Thread (ty, tx) owns rows {c*BM/2 + ty*W + i} and cols {c*BN/2 + tx*W + j},
// c in {0,1}, i,j < W (W = 16-byte vector width), so every shared-memory read is a conflict-free
// 16-byte vector.  BM = BN = 32*W (128 for float, 64 for double), BK = 16.
template <typename T, bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams<T> p) {
  using V = typename VecOf<T>::type;
  constexpr int W = VecOf<T>::W;
  constexpr int BM = 32 * W, BN = 32 * W, BK = 16;
  constexpr int TM = 2 * W, TN = 2 * W;
  constexpr int A_VECS = BM * BK / W / 256;  // vectors of the A tile each thread stages
  constexpr int B_VECS = BN * BK / W / 256;
  __shared__ __align__(16) T As[BK][BM];
  __shared__ __align__(16) T Bs[BK][BN];

  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  if (p.lower_only && bn >= bm + BM) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int z = blockIdx.z;
  const T* __restrict__ A = p.A + (int64_t)z * p.zs_a;
  const T* __restrict__ B = p.B + (int64_t)z * p.zs_b;
  T* __restrict__ C = p.C + (int64_t)z * p.zs_c;

  int k0 = 0, k1 = p.K;
  if (p.k_chunk > 0) { k0 = z * p.k_chunk; k1 = min(p.K, k0 + p.k_chunk); }
  if (p.k_from_diag) { int ks = max(bm, bn); k0 = max(k0, (ks / BK) * BK); }
  if (p.k_to_diag) k1 = min(k1, bn + BN);
  const int K4 = (p.K + 3) & ~3;  // padded extent of a K-contiguous row (padding holds zeros)

  T acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  V ra[A_VECS], rb[B_VECS];

  auto load_a = [&](int kt) {
#pragma unroll
    for (int v = 0; v < A_VECS; ++v) {
      int e = tid + v * 256;
      V val;
      if constexpr (W == 4) val = make_float4(0.f, 0.f, 0.f, 0.f); else val = make_double2(0.0, 0.0);
      if constexpr (!A_T) {
        int r = e / (BK / W), kv = (e % (BK / W)) * W;
        int row = bm + r, k = kt + kv;
        if (row < p.M && k < k1 && k < K4) {
          int64_t prow = p.a_gather ? p.a_gather[row] : (int64_t)row;
          val = *reinterpret_cast<const V*>(A + prow * p.lda + k);
        }
      } else {
        int kk = e / (BM / W), mv = (e % (BM / W)) * W;
        int k = kt + kk, row = bm + mv;
        if (k < k1 && row < p.M) {
          val = *reinterpret_cast<const V*>(A + (int64_t)k * p.lda + row);
          if (p.k_scale) {
            T s = (T)(p.k_scale[k] * p.k_scale_mul);
            val.x *= s; val.y *= s;
            if constexpr (W == 4) { val.z *= s; val.w *= s; }
          }
        }
      }
      ra[v] = val;
    }
  };
  auto load_b = [&](int kt) {
#pragma unroll
    for (int v = 0; v < B_VECS; ++v) {
      int e = tid + v * 256;
      V val;
      if constexpr (W == 4) val = make_float4(0.f, 0.f, 0.f, 0.f); else val = make_double2(0.0, 0.0);
      if constexpr (!B_T) {
        int r = e / (BK / W), kv = (e % (BK / W)) * W;
        int col = bn + r, k = kt + kv;
        if (col < p.N && k < k1 && k < K4) val = *reinterpret_cast<const V*>(B + (int64_t)col * p.ldb + k);
      } else {
        int kk = e / (BN / W), nv = (e % (BN / W)) * W;
        int k = kt + kk, col = bn + nv;
        if (k < k1 && col < p.N) val = *reinterpret_cast<const V*>(B + (int64_t)k * p.ldb + col);
      }
      rb[v] = val;
    }
  };
  auto store_a = [&]() {
#pragma unroll
    for (int v = 0; v < A_VECS; ++v) {
      int e = tid + v * 256;
      if constexpr (!A_T) {
        int r = e / (BK / W), kv = (e % (BK / W)) * W;
        As[kv + 0][r] = ra[v].x; As[kv + 1][r] = ra[v].y;
        if constexpr (W == 4) { As[kv + 2][r] = ra[v].z; As[kv + 3][r] = ra[v].w; }
      } else {
        int kk = e / (BM / W), mv = (e % (BM / W)) * W;
        *reinterpret_cast<V*>(&As[kk][mv]) = ra[v];
      }
    }
  };
  auto store_b = [&]() {
#pragma unroll
    for (int v = 0; v < B_VECS; ++v) {
      int e = tid + v * 256;
      if constexpr (!B_T) {
        int r = e / (BK / W), kv = (e % (BK / W)) * W;
        Bs[kv + 0][r] = rb[v].x; Bs[kv + 1][r] = rb[v].y;
        if constexpr (W == 4) { Bs[kv + 2][r] = rb[v].z; Bs[kv + 3][r] = rb[v].w; }
      } else {
        int kk = e / (BN / W), nv = (e % (BN / W)) * W;
        *reinterpret_cast<V*>(&Bs[kk][nv]) = rb[v];
      }
    }
  };

  if (k0 < k1) {
    load_a(k0); load_b(k0);
    store_a(); store_b();
    __syncthreads();
    for (int kt = k0; kt < k1; kt += BK) {
      const bool more = kt + BK < k1;
      if (more) { load_a(kt + BK); load_b(kt + BK); }
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        T a[TM], b[TN];
        V a0 = *reinterpret_cast<const V*>(&As[k][ty * W]);
        V a1 = *reinterpret_cast<const V*>(&As[k][BM / 2 + ty * W]);
        V b0 = *reinterpret_cast<const V*>(&Bs[k][tx * W]);
        V b1 = *reinterpret_cast<const V*>(&Bs[k][BN / 2 + tx * W]);
        a[0] = a0.x; a[1] = a0.y; a[W] = a1.x; a[W + 1] = a1.y;
        b[0] = b0.x; b[1] = b0.y; b[W] = b1.x; b[W + 1] = b1.y;
        if constexpr (W == 4) {
          a[2] = a0.z; a[3] = a0.w; a[W + 2] = a1.z; a[W + 3] = a1.w;
          b[2] = b0.z; b[3] = b0.w; b[W + 2] = b1.z; b[W + 3] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            if constexpr (EPI == EPI_KERNELFN) {   // squared distance from the differences: relative, not absolute, rounding error
              T d = a[i] - b[j];
              acc[i][j] = fma(d, d, acc[i][j]);
            } else {
              acc[i][j] = fma(a[i], b[j], acc[i][j]);
            }
          }
      }
      __syncthreads();
      if (more) { store_a(); store_b(); }
      __syncthreads();
    }
  }

  // epilogue
  const T alpha = (T)p.alpha, beta = (T)p.beta;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int row = bm + (i / W) * (BM / 2) + ty * W + (i % W);
    if (row >= p.M) continue;
#pragma unroll
    for (int jc = 0; jc < 2; ++jc) {
      int col0 = bn + jc * (BN / 2) + tx * W;
      T out[W];
#pragma unroll
      for (int j = 0; j < W; ++j) {
        T v = acc[i][jc * W + j];
        int col = col0 + j;
        if constexpr (EPI == EPI_KERNELFN) {
          T d2 = (T)p.scale2 * v;
          v = kfn_eval(p.kernel_kind, d2, (T)p.variance);
        } else {
          v = alpha * v;
          if (p.beta != 0.0 && col < p.N) v += beta * C[(int64_t)row * p.ldc + col];
        }
        out[j] = v;
      }
      T* dst = C + (int64_t)row * p.ldc + col0;
      if (col0 + W <= p.N) {
        V o;
        o.x = out[0]; o.y = out[1];
        if constexpr (W == 4) { o.z = out[2]; o.w = out[3]; }
        *reinterpret_cast<V*>(dst) = o;
      } else {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (col0 + j < p.N) dst[j] = out[j];
      }
    }
  }
}

template <typename T> constexpr int gemm_tile() { return 32 * VecOf<T>::W; }

template <typename T, bool A_T, bool B_T, int EPI>
inline void gemm_simt_launch(const GemmParams<T>& p, int nz, cudaStream_t st) {
  constexpr int BT = 32 * VecOf<T>::W;
  dim3 grid((p.N + BT - 1) / BT, (p.M + BT - 1) / BT, nz);
  gemm_simt_kernel<T, A_T, B_T, EPI><<<grid, 256, 0, st>>>(p);
}

}  // namespace agp
