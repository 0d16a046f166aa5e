// agp_hyper.cuh -- gradient of the ELBO w.r.t. the kernel parameters and the inducing points, SURVEY 8 row f3.
//
// The reference differentiates ELBO(model, x, y, pr_means, kernels, Zs, state) (functions/ELBO.jl:15-21) with Zygote inside
// update_hyperparameters! (hyperparameter/autotuning.jl:86-140): kernel matrices recomputed from (kernels, Zs), posterior and
// local variables fixed.  The same gradient in closed form (the test suite checks the formula against finite differences of the
// CPU restatement), with K = K_mm + jitter I, kappa = K_nm K^-1, (a, b) = d E / d (mu_f, var_f), M = a mu^T + diag(b)(2 kappa Sigma - K_nm):
//     A_nm = rho (M K^-1 - diag(b) kappa)
//     A_mm = -rho sym(kappa^T M K^-1) - K^-1 / 2 + K^-1 (Sigma + (mu - mu0)(mu - mu0)^T) K^-1 / 2
//     dELBO/dtheta = <A_nm, dK_nm/dtheta> + <A_mm, dK_mm/dtheta> + rho sum_i b_i dk_ii/dtheta
// Everything here is fp64 and off the per-iteration hot path (it runs every `atfrequency` iterations); the B x m x m products
// use the SIMT fp64 GEMM of agp_gemm_simt.cuh, this file holds the element-wise pieces.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "agp_kernels.cuh"

namespace agp {

// rows of the current minibatch in fp64 (+ nothing else: distances are formed from differences, no cancellation)
template <typename T>
__global__ void hg_gather_x_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ idx, int B, int D, double* __restrict__ out,
                                   int64_t ldo) {
  const int b = blockIdx.x * blockDim.y + threadIdx.y;
  if (b >= B) return;
  const int64_t src = idx ? idx[b] : (int64_t)b;
  for (int d = threadIdx.x; d < (int)ldo; d += blockDim.x) out[(int64_t)b * ldo + d] = d < D ? (double)X[src * ldx + d] : 0.0;
}

// (a, b) = d expec_loglikelihood / d (mu_f, var_f) per OWNED latent, from the local variables of the last step and the moments
// under the updated posterior (derivatives of the reference's own formulas, quirks included: likelihood/*.jl expec_loglikelihood)
__device__ __forceinline__ void lik_elbo_grad_single(int kind, double p0, double y, double mu, double th, double gam, double& a, double& b) {
  b = -0.5 * th;
  if (kind == 0) { a = (y - mu) / p0; b = -0.5 / p0; }
  else if (kind == 1) a = 0.5 * (y - th);                       // Q1: dot(theta, mu)
  else if (kind == 2 || kind == 4) a = th * (y - mu);
  else if (kind == 5) a = y - 2.0 * th * y * (1.0 - y * mu);
  else if (kind == 6) a = 0.5 * (y - p0) - 0.5 * th;
  else a = 0.5 * (y - gam) - th * mu;                           // poisson
}
__global__ void hg_ab_kernel(const LikParams p_in, double* __restrict__ aout, double* __restrict__ bout) {
  const LikParams p = lik_resolve(p_in);
  const int bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= p.B) return;
  const int64_t ld = p.ldB;
  if (p.model_kind == 0 && p.lik_kind[0] == 3) {          // LogisticSoftMax
    const int cls = p.ycls[bi];
    for (int ql = 0; ql < p.n_latent_local; ++ql) {
      const int k = p.latent_begin + ql;
      const double th = p.theta[k * ld + bi];
      aout[ql * ld + bi] = 0.5 * ((k == cls ? 1.0 : 0.0) - p.gamma[k * ld + bi]) - th * p.mean_f[k * ld + bi];
      bout[ql * ld + bi] = -0.5 * th;
    }
    return;
  }
  if (p.model_kind == 0 && p.lik_kind[0] == 8) {          // heteroscedastic (rows: c[1] = phi unused here, gamma[0] = gamma)
    const double lam = p.lam[0], y = p.yb[bi], m1 = p.mean_f[bi], v1 = p.var_f[bi], m2 = p.mean_f[ld + bi];
    const double g = p.gamma[bi], th = p.theta[bi];
    const double lam0 = 0.5 * lam * ((y - m1) * (y - m1) + v1), w = 1.0 - g / lam0;
    for (int ql = 0; ql < p.n_latent_local; ++ql) {
      const int q = p.latent_begin + ql;
      aout[ql * ld + bi] = q == 0 ? -lam * (m1 - y) * w : 0.5 * (0.5 - g) - th * m2;
      bout[ql * ld + bi] = q == 0 ? -0.5 * lam * w : -0.5 * th;
    }
    return;
  }
  if (p.model_kind == 0) {
    double a, b;
    lik_elbo_grad_single(p.lik_kind[0], p.p0[0], p.yb[bi], p.mean_f[bi], p.theta[bi], p.gamma[bi], a, b);
    aout[bi] = a; bout[bi] = b;
    return;
  }
  // MOSVGP: mu_t = sum_q A_tq mu_q, var_t = sum_q A_tq^2 var_q  (tmu / tvar were refreshed by lik_update_kernel(update = 0))
  for (int ql = 0; ql < p.n_latent_local; ++ql) {
    const int q = p.latent_begin + ql;
    double sa = 0.0, sb = 0.0;
    for (int t = 0; t < p.n_task; ++t) {
      double a, b;
      lik_elbo_grad_single(p.lik_kind[t], p.p0[t], p.yb[t * ld + bi], p.tmu[t * ld + bi], p.theta[t * ld + bi], p.gamma[t * ld + bi], a, b);
      const double w = p.A[t * p.Q + q];
      sa += w * a; sb += w * w * b;
    }
    aout[ql * ld + bi] = sa; bout[ql * ld + bi] = sb;
  }
}

// M = a mu^T + diag(b)(2 T - Knm)   (T = kappa Sigma, overwritten)
__global__ void hg_M_kernel(double* __restrict__ T, const double* __restrict__ Knm, int64_t ld, int B, int m, const double* __restrict__ a,
                            const double* __restrict__ b, const double* __restrict__ mu) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= m || i >= B) return;
  const int64_t o = (int64_t)i * ld + j;
  T[o] = a[i] * mu[j] + b[i] * (2.0 * T[o] - Knm[o]);
}
// A_nm = rho (MK - diag(b) kappa)   (written over M)
__global__ void hg_Anm_kernel(double* __restrict__ out, const double* __restrict__ MK, const double* __restrict__ kappa, int64_t ld, int B, int m,
                              const double* __restrict__ b, double rho) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= m || i >= B) return;
  const int64_t o = (int64_t)i * ld + j;
  out[o] = rho * (MK[o] - b[i] * kappa[o]);
}
// A_mm = sym(Acc) - Kinv / 2 + (KSK + v v^T) / 2   with Acc = -rho kappa^T MK, KSK = Kinv Sigma Kinv, v = Kinv (mu - mu0)
__global__ void hg_Amm_kernel(double* __restrict__ Amm, const double* __restrict__ Acc, const double* __restrict__ Kinv, const double* __restrict__ KSK,
                              const double* __restrict__ v, int64_t ld, int m) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= m || i >= m) return;
  const int64_t o = (int64_t)i * ld + j, ot = (int64_t)j * ld + i;
  Amm[o] = 0.5 * (Acc[o] + Acc[ot]) - 0.5 * Kinv[o] + 0.5 * (0.5 * (KSK[o] + KSK[ot]) + v[i] * v[j]);
}

// kernel and its derivative w.r.t. d2 = scale^2 |x - z|^2  (KernelFunctions: SqExponential, Matern32, Matern52)
__device__ __forceinline__ void kfn_with_deriv(int kind, double d2, double var, double& K, double& W) {
  if (kind == 0) { const double e = exp(-0.5 * d2); K = var * e; W = -0.5 * var * e; return; }
  const double d = sqrt(d2);
  if (kind == 1) { const double c = 1.7320508075688772, e = exp(-c * d); K = var * (1.0 + c * d) * e; W = -1.5 * var * e; return; }
  const double c = 2.23606797749979, e = exp(-c * d);
  K = var * (1.0 + c * d + (5.0 / 3.0) * d2) * e; W = -(5.0 / 6.0) * var * (1.0 + c * d) * e;
}
// Contraction of A (n x m) with dK/dscale and K (for d/dvariance), and G = A .* dK/d(d2) written back for the dZ products.
//   rows of P: n x D (ld ldp), rows of Z: m x D (ld ldz).  out[0] += sum A dK/dscale, out[1] += sum A K,
//   colsum[j] += sum_i G_ij, rowsum[i] += sum_j G_ij (rowsum may be null).   One block = 8 rows x 128 columns.
__global__ void __launch_bounds__(128) hg_contract_kernel(double* __restrict__ A, int64_t lda, int n, int m, const double* __restrict__ P, int64_t ldp,
                                                          const double* __restrict__ Z, int64_t ldz, int D, int kind, double scale, double var,
                                                          double* __restrict__ out, double* __restrict__ colsum, double* __restrict__ rowsum) {
  extern __shared__ double sh[];               // [8][D] rows of P
  const int j = blockIdx.x * 128 + threadIdx.x, i0 = blockIdx.y * 8;
  for (int e = threadIdx.x; e < 8 * D; e += 128) { const int r = e / D, d = e % D; sh[e] = (i0 + r < n) ? P[(int64_t)(i0 + r) * ldp + d] : 0.0; }
  __syncthreads();
  double s_scale = 0.0, s_var = 0.0, cs = 0.0;
  double rs[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) rs[r] = 0.0;
  if (j < m) {
    double diff2[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) diff2[r] = 0.0;
    for (int d = 0; d < D; ++d) {
      const double z = Z[(int64_t)j * ldz + d];
#pragma unroll
      for (int r = 0; r < 8; ++r) { const double t = sh[r * D + d] - z; diff2[r] = fma(t, t, diff2[r]); }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (i0 + r >= n) continue;
      double K, W;
      kfn_with_deriv(kind, scale * scale * diff2[r], var, K, W);
      const int64_t o = (int64_t)(i0 + r) * lda + j;
      const double av = A[o], g = av * W;
      s_scale += g * 2.0 * scale * diff2[r];
      s_var += av * K;
      A[o] = g;
      cs += g; rs[r] = g;
    }
    atomicAdd(colsum + j, cs);
  }
  __shared__ double red[4][10];
  s_scale = warp_sum(s_scale); s_var = warp_sum(s_var);
#pragma unroll
  for (int r = 0; r < 8; ++r) rs[r] = warp_sum(rs[r]);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[w][0] = s_scale; red[w][1] = s_var; for (int r = 0; r < 8; ++r) red[w][2 + r] = rs[r]; }
  __syncthreads();
  if (threadIdx.x < 10) {
    const double t = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    if (threadIdx.x < 2) atomicAdd(out + threadIdx.x, t);
    else if (rowsum && i0 + (int)threadIdx.x - 2 < n) atomicAdd(rowsum + i0 + threadIdx.x - 2, t);
  }
}
// dZ = -2 scale^2 (P1 + P2 + P3 - (cs1 + cs2 + rs2) .* Z)
__global__ void hg_dz_kernel(double* __restrict__ dZ, int D, int m, const double* __restrict__ P1, const double* __restrict__ P2,
                             const double* __restrict__ P3, int64_t ldp, const double* __restrict__ cs1, const double* __restrict__ cs2,
                             const double* __restrict__ rs2, const double* __restrict__ Z, int64_t ldz, double scale) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * D) return;
  const int j = e / D, d = e % D;
  const int64_t o = (int64_t)j * ldp + d;
  dZ[e] = -2.0 * scale * scale * (P1[o] + P2[o] + P3[o] - (cs1[j] + cs2[j] + rs2[j]) * Z[(int64_t)j * ldz + d]);
}
__global__ void hg_rownorm_kernel(const double* __restrict__ X, int64_t ld, int B, int D, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) { const double v = X[(int64_t)b * ld + d]; s = fma(v, v, s); }
  out[b] = s;
}
__global__ void hg_sub_kernel(const double* __restrict__ a, const double* __restrict__ b, int m, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < m) out[j] = a[j] - b[j];
}
// sum of b (for d/dvariance through k_ii = variance)
__global__ void hg_sum_kernel(const double* __restrict__ b, int B, double* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += b[i];
  __shared__ double sh[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w]; *out = t; }
}

}  // namespace agp
