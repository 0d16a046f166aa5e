// Elementwise / reduction / small-matrix kernels of the CAVI step (everything that is not a big GEMM).
// Reference formulas are cited per kernel (paths relative to /root/reference/src).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "agp_gemm_simt.cuh"
#include "agp_lik.cuh"

namespace agp {

// ------------------------------------------------------------------------------------------------
// data preparation
// ------------------------------------------------------------------------------------------------
// convert a chunk of host-layout X (staged on device) to row-major T with leading dimension ldx
// and compute the squared row norms used by the GEMM-form distance.
template <typename TS, typename T>
__global__ void convert_rows_kernel(const TS* __restrict__ src, int layout, int64_t src_ld /*n for colmajor, D for rowmajor*/,
                                    int64_t rows, int D, T* __restrict__ dst, int64_t ldx, T* __restrict__ xx) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    double v = layout == 0 ? (double)src[r + src_ld * d] : (double)src[r * src_ld + d];
    dst[r * ldx + d] = (T)v;
    T q = (T)v;
    s += (double)q * (double)q;
  }
  if (xx) xx[r] = (T)s;
}

template <typename TS>
__global__ void convert_vec_kernel(const TS* __restrict__ src, double* __restrict__ dst, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (double)src[i];
}

__global__ void idx_rebase_kernel(const int64_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n, int base) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] - base;
}

// copy the index list of the current cursor position into the fixed per-step buffer (so that every
// later kernel of the step - and a captured CUDA graph - reads one fixed address)
// padding rows [B, Bk) of a ragged minibatch repeat its first sample (their weights are zero: see Engine::rowsK)
__global__ void idx_pad_kernel(int64_t* __restrict__ idx, int B, int Bk) {
  for (int b = B + threadIdx.x; b < Bk; b += blockDim.x) idx[b] = idx[0];
}

template <typename T>
__global__ void idx_select_kernel(const int64_t* __restrict__ pool, int64_t n_lists, int B,
                                  const int64_t* __restrict__ counters /*[0]=t,[1]=cursor*/, int cursor_offset,
                                  int64_t* __restrict__ dst, const T* __restrict__ xx_all, T* __restrict__ xx_cur) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  int64_t cur = (counters[1] + cursor_offset) % n_lists;
  int64_t r = pool[cur * B + i];
  dst[i] = r;
  xx_cur[i] = xx_all[r];   // gather the squared norms once (the K_nm epilogue then reads them by minibatch row)
}

// same for a host-provided (already rebased) list
template <typename T>
__global__ void xx_gather_kernel(const int64_t* __restrict__ idx, int B, const T* __restrict__ xx_all, T* __restrict__ xx_cur) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) xx_cur[i] = xx_all[idx[i]];
}

// ------------------------------------------------------------------------------------------------
// per-row statistics in the whitened basis (V = Knm L^-T, q(v) = N(mu_v, Sigma_v), u = L v):
//   Ktilde = kdiag + jitter - rowsum(V .* V)        == kdiag + jitter - diag_ABt(kappa, Knm)   latentgp.jl:212
//   mean_f = V mu_v = (V X^T) t,  t = X eta1_v         == kappa * mu                              latentgp.jl:179
//   var_f  = rowsum((V X^T).^2) + Ktilde, Sigma_v = X^T X  == diag_ABt(kappa*Sigma, kappa) + Ktilde   latentgp.jl:189
//            (VS below holds V X^T, X = inverse Cholesky factor of P_v = -2 eta2_v)
// one warp per minibatch row; sums accumulate in fp64 whatever T is.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void rowstats_kernel(const T* __restrict__ V, const T* __restrict__ VS, const double* __restrict__ mu, int B, int m,
                                int64_t ld, double kdiag_jit, double* __restrict__ Ktilde, double* __restrict__ mean_f,
                                double* __restrict__ var_f, int* __restrict__ status, int compute_ktilde,
                                const int64_t* __restrict__ xepoch, int64_t par_stride) {
  if (xepoch) { const int64_t off = ((*xepoch + 1) & 1) * par_stride; mean_f += off; var_f += off; }   // next exchange's buffer
  using VT = typename VecOf<T>::type;
  constexpr int W = VecOf<T>::W;
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const T* vp = V + (int64_t)warp * ld;
  const T* sp = VS + (int64_t)warp * ld;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int m4 = (m + 3) & ~3;  // padding columns are zero
  for (int j = lane * W; j < m4; j += 32 * W) {
    VT a = *reinterpret_cast<const VT*>(vp + j);
    VT c = *reinterpret_cast<const VT*>(sp + j);
    double av[W], cv[W];
    av[0] = a.x; av[1] = a.y; cv[0] = c.x; cv[1] = c.y;
    if constexpr (W == 4) { av[2] = a.z; av[3] = a.w; cv[2] = c.z; cv[3] = c.w; }
#pragma unroll
    for (int q = 0; q < W; ++q) {
      double muj = (j + q < m) ? mu[j + q] : 0.0;   // mu holds t = X eta1_v
      s1 += av[q] * av[q];
      s2 += cv[q] * muj;
      s3 += cv[q] * cv[q];
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
  if (lane == 0) {
    double kt;
    if (compute_ktilde == 2) {   // full VGP (kappa = I): var_f = diag(Sigma), no Ktilde (latentgp.jl:182-186)
      kt = 0.0;
      Ktilde[warp] = 0.0;
    } else if (compute_ktilde) {
      kt = kdiag_jit - s1;
      Ktilde[warp] = kt;
      if (!(kt > 0.0)) atomicOr(status, ST_KTILDE);  // latentgp.jl:213
    } else {
      kt = Ktilde[warp];
    }
    mean_f[warp] = s2;
    var_f[warp] = s3 + kt;
  }
}

// finishing step of the fused row statistics (tensor-core path): acc = {sum V^2, sum (VX^T)^2, sum (VX^T) t}
__global__ void rowfinish_kernel(const double* __restrict__ sumsq_v, const double* __restrict__ sumsq_vs, const double* __restrict__ dot_vs,
                                 int B, double kdiag_jit, double* __restrict__ Ktilde, double* __restrict__ mean_f,
                                 double* __restrict__ var_f, int* __restrict__ status, int compute_ktilde,
                                 const int64_t* __restrict__ xepoch, int64_t par_stride) {
  pdl_prologue();
  if (xepoch) { const int64_t off = ((*xepoch + 1) & 1) * par_stride; mean_f += off; var_f += off; }   // next exchange's buffer
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double kt;
  if (compute_ktilde == 2) {     // full VGP (kappa = I)
    kt = 0.0;
    Ktilde[b] = 0.0;
  } else if (compute_ktilde) {
    kt = kdiag_jit - sumsq_v[b];
    Ktilde[b] = kt;
    if (!(kt > 0.0)) atomicOr(status, ST_KTILDE);  // latentgp.jl:213
  } else {
    kt = Ktilde[b];
  }
  mean_f[b] = dot_vs[b];
  var_f[b] = sumsq_vs[b] + kt;
}

// ---- batched forms of the small per-latent kernels of multi-latent steps: up to SMALL_NB latents per launch, the latent of a block
// from blockIdx.y / .z, its pointers from arrays passed BY VALUE in the launch (no device-side table to keep in sync).  At C5 (64 latents
// on one GPU) 64 launches of 4-11 us each, fanned over four streams, become four. ----
constexpr int SMALL_NB = 16;
struct RowFinishBatch { const double* racc[SMALL_NB]; double* Ktilde[SMALL_NB]; double kdiag[SMALL_NB]; };
// rowfinish_kernel for the latents of a batch: racc[z] = {sum V^2 | sum (VX^T)^2 | sum (VX^T) t}, each ldB long; mean_f / var_f point at
// the first latent of the batch, rows out_ld apart
__global__ void rowfinish_batched_kernel(const RowFinishBatch bt, int64_t ldB, int B, double* __restrict__ mean_f, double* __restrict__ var_f,
                                         int64_t out_ld, int* __restrict__ status, int compute_ktilde, const int64_t* __restrict__ xepoch,
                                         int64_t par_stride) {
  pdl_prologue();
  const int z = blockIdx.y;
  if (xepoch) { const int64_t off = ((*xepoch + 1) & 1) * par_stride; mean_f += off; var_f += off; }   // next exchange's buffer
  mean_f += (int64_t)z * out_ld; var_f += (int64_t)z * out_ld;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* __restrict__ racc = bt.racc[z];
  double* __restrict__ Kt = bt.Ktilde[z];
  double kt;
  if (compute_ktilde) {
    kt = bt.kdiag[z] - racc[b];
    Kt[b] = kt;
    if (!(kt > 0.0)) atomicOr(status, ST_KTILDE);  // latentgp.jl:213
  } else {
    kt = Kt[b];
  }
  mean_f[b] = racc[2 * ldB + b];
  var_f[b] = racc[ldB + b] + kt;
}

// out[j] += sum_b V[b][j] * g[b]   (transpose(kappa) * grad_mu of analyticVI.jl:168, whitened; rho applied later)
// block = 32 columns x 8 row lanes; grid = (ceil(m/32), row_chunks); one atomicAdd per column per block
template <typename T>
__global__ void gemv_t_kernel(const T* __restrict__ V, int64_t ld, const double* __restrict__ g, int B, int m, int rows_per_block,
                              double* __restrict__ out) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int b0 = blockIdx.y * rows_per_block, b1 = min(B, b0 + rows_per_block);
  double s = 0.0;
  if (j < m) {
#pragma unroll 4
    for (int b = b0 + threadIdx.y; b < b1; b += 8) s += (double)V[(int64_t)b * ld + j] * g[b];
  }
  __shared__ double sh[8][33];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < m) {
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) a += sh[r][threadIdx.x];
    atomicAdd(out + j, a);
  }
}


__global__ void lik_update_kernel(const LikParams p_in) {
  pdl_prologue();
  const LikParams p = lik_resolve(p_in);
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  double r0 = 0.0, r1 = 0.0;
  if (b < p.B) lik_update_sample(p, b, r0, r1);
  if (!p.need_reduce || p.update != 1 || p.model_kind != 0) return;   // uniform
  __shared__ double s0[8], s1[8];
  r0 = warp_sum(r0); r1 = warp_sum(r1);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s0[w] = r0; s1[w] = r1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s0[i]; c += s1[i]; }
    atomicAdd(p.lamacc + 0, a);
    atomicAdd(p.lamacc + 1, c);
  }
}

// MOSVGP local updates in two dimensions (the loops of lik_update_sample's multi-output branch, same summation order, so the
// results are bit-identical): one thread per (sample, task) for the task moments + local variables + per-task gradients, then one
// thread per (sample, owned latent) for the latent gradients.  One thread per sample looping over T x Q (64 x 64 at BASELINE C5) left
// 64 CTAs latency-bound for ~580 us per step, replicated on every rank of a sharded run.
constexpr int LIK_MO_TPT = 8;   // tasks (latents) per thread: every latent moment (per-task term) is loaded once for 8 accumulators
__global__ void lik_mo_task_kernel(const LikParams p_in) {
  pdl_prologue();
  const LikParams p = lik_resolve(p_in);
  const int b = blockIdx.x * blockDim.x + threadIdx.x, t0 = blockIdx.y * LIK_MO_TPT;
  if (b >= p.B) return;
  const int64_t ld = p.ldB;
  const int Q = p.Q, T = p.n_task;
  const int64_t src = p.idx ? p.idx[b] : (int64_t)b;
  double mt[LIK_MO_TPT], vt[LIK_MO_TPT];
#pragma unroll
  for (int u = 0; u < LIK_MO_TPT; ++u) { mt[u] = 0.0; vt[u] = 0.0; }
  for (int q = 0; q < Q; ++q) {
    const double mq = p.mean_f[q * ld + b], vq = p.var_f[q * ld + b];
#pragma unroll
    for (int u = 0; u < LIK_MO_TPT; ++u) {
      const double a = (t0 + u < T) ? p.A[(t0 + u) * Q + q] : 0.0;
      mt[u] += a * mq;
      vt[u] += a * a * vq;
    }
  }
#pragma unroll
  for (int u = 0; u < LIK_MO_TPT; ++u) {
    const int t = t0 + u;
    if (t >= T) break;
    double y = p.idx ? p.y_all[(int64_t)t * p.n + src] : p.yb[t * ld + b];
    if (p.idx) p.yb[t * ld + b] = y;
    p.tmu[t * ld + b] = mt[u]; p.tvar[t * ld + b] = vt[u];
    const int kind = p.lik_kind[t];
    if (p.update == 1 || (p.update == 2 && kind == 0)) {
      double c, th, gam, gm, gs;
      lik_single(kind, p.p0[t], p.p1[t], p.lam[t], y, mt[u], vt[u], c, th, gam, gm, gs);
      p.c[t * ld + b] = c; p.theta[t * ld + b] = th;
      p.gm[t * ld + b] = gm; p.gs[t * ld + b] = gs;
      if (kind == 0 && p.noise_opt && p.noise_opt[t] && p.update == 1) atomicAdd(p.lamacc + 2 * t, (y - mt[u]) * (y - mt[u]) + vt[u]);
      if (kind == 7 && p.update == 1) {
        p.gamma[t * ld + b] = gam;
        atomicAdd(p.lamacc + 2 * t, y);
        atomicAdd(p.lamacc + 2 * t + 1, expect_logistic(p.qnodes, p.qweights, p.nq, mt[u], vt[u]));
      }
    }
  }
}
__global__ void lik_mo_grad_kernel(const LikParams p_in) {
  pdl_prologue();
  const LikParams p = lik_resolve(p_in);
  const int b = blockIdx.x * blockDim.x + threadIdx.x, l0 = blockIdx.y * LIK_MO_TPT;
  if (b >= p.B) return;
  const int64_t ld = p.ldB;
  const int T = p.n_task, Q = p.Q, nl = p.n_latent_local;
  double muq[LIK_MO_TPT], a1[LIK_MO_TPT], a2[LIK_MO_TPT];
#pragma unroll
  for (int u = 0; u < LIK_MO_TPT; ++u) {
    muq[u] = (l0 + u < nl) ? p.mean_f[(p.latent_begin + l0 + u) * ld + b] : 0.0;
    a1[u] = 0.0; a2[u] = 0.0;
  }
  for (int t = 0; t < T; ++t) {
    const double tm = p.tmu[t * ld + b], gmt = p.gm[t * ld + b], gst = p.gs[t * ld + b];
#pragma unroll
    for (int u = 0; u < LIK_MO_TPT; ++u) {
      const double a = (l0 + u < nl) ? p.A[t * Q + p.latent_begin + l0 + u] : 0.0;
      const double others = tm - a * muq[u];
      a1[u] += a * (gmt - 2.0 * gst * others);
      a2[u] += a * a * gst;
    }
  }
#pragma unroll
  for (int u = 0; u < LIK_MO_TPT; ++u) {
    if (l0 + u >= nl) break;
    p.gmu[(l0 + u) * ld + b] = a1[u];
    p.gS[(l0 + u) * ld + b] = a2[u];
  }
}

// ------------------------------------------------------------------------------------------------
// update_A! (models/single_and_multi_output_utils.jl:87-118) for MOSVGP with an A optimiser: runs between the latent
// moments and local_updates!, i.e. with the local variables of the PREVIOUS iteration and the labels of the current batch.
// ------------------------------------------------------------------------------------------------
// expectation gradients from stored local variables (grad_E_mu / grad_E_Sigma of likelihood/*.jl)
__device__ __forceinline__ void lik_grads_from_state(int kind, double p0, double y, double th, double gam, double& gm, double& gs) {
  if (kind == 0) { th = 1.0 / p0; gm = y / p0; }       // gaussian.jl:74-80 (theta = 1/sigma^2 from init_local_vars on)
  else if (kind == 1) gm = 0.5 * y;                    // logistic.jl:64-66
  else if (kind == 2 || kind == 4) gm = th * y;        // studentt.jl:96, laplace.jl:87-89
  else if (kind == 5) gm = y * (th + 1.0);             // bayesiansvm.jl:54-58
  else if (kind == 6) gm = 0.5 * (y - p0);             // negativebinomial.jl:94-96
  else gm = 0.5 * (y - gam);                           // poisson.jl:98-102
  gs = 0.5 * th;
}
// one block per (task t, latent q): gradA[t][q] = x1 - 2 A_tq x2 (:96-107), deterministic in-block reduction
__global__ void update_A_grad_kernel(const LikParams p_in, double* __restrict__ gradA) {
  const LikParams p = lik_resolve(p_in);
  const int t = blockIdx.x / p.Q, q = blockIdx.x % p.Q;
  const int64_t ld = p.ldB;
  const double a = p.A[t * p.Q + q];
  const int kind = p.lik_kind[t];
  const double p0 = p.p0[t];
  double x1 = 0.0, x2 = 0.0;
  for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
    const double y = p.yb[t * ld + b];                  // gathered by the preceding lik_update_kernel(update = 0) pass
    double gm, gs;
    lik_grads_from_state(kind, p0, y, p.theta[t * ld + b], p.gamma[t * ld + b], gm, gs);
    const double mq = p.mean_f[q * ld + b], vq = p.var_f[q * ld + b];
    const double others = p.tmu[t * ld + b] - a * mq;   // sum over q' != q of A_tq' mu_q'
    x1 += gm * mq - 2.0 * gs * mq * others;
    x2 += gs * (mq * mq + vq);
  }
  __shared__ double s1[8], s2[8];
  x1 = warp_sum(x1); x2 = warp_sum(x2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s1[w] = x1; s2[w] = x2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double u = 0.0, v = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { u += s1[i]; v += s2[i]; }
    gradA[t * p.Q + q] = u - 2.0 * a * v;
  }
}
// ADAM step of Optimisers.jl (apply: mt, vt, beta_t) followed by A_t += step and the projection on the unit circle (:109-113);
// one block per task, thread q
__global__ void update_A_adam_kernel(double* __restrict__ A, const double* __restrict__ gradA, double* __restrict__ mt, double* __restrict__ vt,
                                     double* __restrict__ bt /*[T][2]*/, int Q, double eta, double b1, double b2, double eps) {
  const int t = blockIdx.x, q = threadIdx.x;
  __shared__ double red[32];
  double an = 0.0;
  if (q < Q) {
    const double g = gradA[t * Q + q];
    const double m_ = b1 * mt[t * Q + q] + (1.0 - b1) * g;
    const double v_ = b2 * vt[t * Q + q] + (1.0 - b2) * g * g;
    mt[t * Q + q] = m_; vt[t * Q + q] = v_;
    const double step = m_ / (1.0 - bt[2 * t]) / (sqrt(v_ / (1.0 - bt[2 * t + 1])) + eps) * eta;
    an = A[t * Q + q] + step;
  }
  double ss = warp_sum(an * an);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  double tot = 0.0;
  for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) tot += red[i];
  if (q < Q) A[t * Q + q] = an / sqrt(tot);
  __syncthreads();
  if (q == 0) { bt[2 * t] *= b1; bt[2 * t + 1] *= b2; }
}

// Full VGP (models/VGP.jl; natural_gradient!(::VarLatent) analyticVI.jl:126-140): with Z = X and kappa = I the whitened
// "kappa" is V = K L^-T = L itself, so the step's V buffer is just the T shadow of chol(K) (zeros above the diagonal)
template <typename T>
__global__ void vgp_fill_v_kernel(const double* __restrict__ Lc, int64_t ldl, int n, T* __restrict__ V, int64_t ldv) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= (int)ldv || i >= n) return;
  V[(int64_t)i * ldv + j] = (j <= i && j < n) ? (T)Lc[(int64_t)i * ldl + j] : T(0);
}

// rowfinish_kernel + lik_update_kernel in one launch for the common single-latent case (SVGP, one latent, no lambda
// re-estimation, not sharded): the thread that finishes (mu_f, sigma2_f) of a sample runs its local update right away
__global__ void rowfinish_lik_kernel(const double* __restrict__ sumsq_v, const double* __restrict__ sumsq_vs, const double* __restrict__ dot_vs,
                                     int B, double kdiag_jit, double* __restrict__ Ktilde, int* __restrict__ status, const LikParams p) {
  pdl_prologue();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double kt = kdiag_jit - sumsq_v[b];
  Ktilde[b] = kt;
  if (!(kt > 0.0)) atomicOr(status, ST_KTILDE);  // latentgp.jl:213
  const_cast<double*>(p.mean_f)[b] = dot_vs[b];
  const_cast<double*>(p.var_f)[b] = sumsq_vs[b] + kt;
  double r0 = 0.0, r1 = 0.0;
  lik_update_sample(p, b, r0, r1);
}

// re-estimation of the link parameter lambda at the end of local_updates! (poisson.jl:80, heteroscedastic.jl:98);
// one thread per task, accumulators cleared for the next step
__global__ void lik_lambda_kernel(const LikParams p) {
  pdl_prologue();
  int t = threadIdx.x;
  if (t >= p.n_task) return;
  int kind = p.lik_kind[t];
  if (kind == 7) p.lam[t] = p.lamacc[2 * t] / p.lamacc[2 * t + 1];
  else if (kind == 8) p.lam[t] = fmax((double)p.B / (2.0 * p.lamacc[2 * t]), p.lam[t]);
  else if (kind == 0 && p.noise_opt && p.noise_opt[t]) {
    // gaussian.jl:62-68: grad = ((sum (y - mu)^2 + sum var_f) / sigma^2 - B) / 2, ADAM step, sigma^2 <- exp(log sigma^2 + step)
    double* st = p.noise_state + 4 * t;
    double* s2 = const_cast<double*>(p.p0) + t;
    const double g = 0.5 * (p.lamacc[2 * t] / *s2 - (double)p.B);
    st[0] = p.n_b1 * st[0] + (1.0 - p.n_b1) * g;
    st[1] = p.n_b2 * st[1] + (1.0 - p.n_b2) * g * g;
    const double step = st[0] / (1.0 - st[2]) / (sqrt(st[1] / (1.0 - st[3])) + p.n_eps) * p.n_eta;
    st[2] *= p.n_b1; st[3] *= p.n_b2;
    *s2 = exp(log(*s2) + step);
  }
  p.lamacc[2 * t] = 0.0; p.lamacc[2 * t + 1] = 0.0;
}

// grad_E_mu / grad_E_Sigma of the heteroscedastic likelihood (heteroscedastic.jl:111-127) with the NEW lambda
__global__ void hetero_grad_kernel(const LikParams p) {
  pdl_prologue();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int64_t ld = p.ldB;
  const double lam = p.lam[0], y = p.yb[b], sg = p.gamma[ld + b], gam = p.gamma[b], th = p.theta[b];
  for (int ql = 0; ql < p.n_latent_local; ++ql) {
    int q = p.latent_begin + ql;
    p.gmu[ql * ld + b] = q == 0 ? 0.5 * y * lam * sg : 0.5 * (0.5 - gam);
    p.gS[ql * ld + b] = q == 0 ? 0.5 * lam * sg : 0.5 * th;
  }
}

// ------------------------------------------------------------------------------------------------
// Latent-sharded exchange of the per-sample moments over NVLink peer memory (the only data that crosses GPUs).
// Every rank owns rows [qbeg, qbeg+Ql) of the [Q][ldB] arrays mean_f / var_f.  The arrays live in one exported allocation
// per rank:  [mean: 2 x Q x ldB][var: 2 x Q x ldB][flags: 64 x int64], double-buffered by the exchange counter's parity so
// a rank that runs ahead cannot overwrite moments a slower peer is still reading.
//   peer_publish_kernel: store my rows of the NEXT parity into every peer's arrays (plain st.global on mapped peer pointers)
//   peer_sync_kernel   : release-store the new exchange number into my slot of every peer's flag array, bump my counter,
//                        then acquire-spin until every peer's number has arrived in MY flag array (bounded: ~2 s)
// No host involvement, no NCCL call on the step path: both kernels are ordinary nodes of the step's CUDA graph.
// ------------------------------------------------------------------------------------------------
__global__ void peer_publish_kernel(double* const* __restrict__ peers, int world, int rank, const int64_t* __restrict__ xepoch,
                                    int64_t par_stride, int qbeg, int Ql, int64_t ldB, int B) {
  const int64_t par = ((*xepoch + 1) & 1) * par_stride;
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)Ql * B) return;
  const int64_t off = par + (int64_t)(qbeg + e / B) * ldB + e % B;
  const double* mine = peers[rank];
  const double mu = mine[off], var = mine[2 * par_stride + off];
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    double* dst = peers[p];
    dst[off] = mu;
    dst[2 * par_stride + off] = var;
  }
}
__global__ void peer_sync_kernel(double* const* __restrict__ peers, int world, int rank, int64_t* __restrict__ xepoch, int64_t flag_off,
                                 int* __restrict__ status) {
  const int p = threadIdx.x;
  const int64_t e = *xepoch + 1;
  if (p < world && p != rank) {
    __threadfence_system();
    int64_t* f = reinterpret_cast<int64_t*>(peers[p] + flag_off) + rank;
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(f), "l"(e) : "memory");
  }
  __syncthreads();
  if (p == 0) *xepoch = e;
  if (p < world && p != rank) {
    const int64_t* f = reinterpret_cast<const int64_t*>(peers[rank] + flag_off) + p;
    const long long t0 = clock64();
    int64_t v;
    do {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v < e && clock64() - t0 > 4000000000LL) { atomicOr(status, ST_PEER_TIMEOUT); break; }
    } while (v < e);
  }
}

// ------------------------------------------------------------------------------------------------
// ELBO likelihood terms (inference/analyticVI.jl:255-297): out[0] += expec_loglikelihood (un-scaled),
// out[2] += AugmentedKL (un-scaled).  Block-reduced, one atomicAdd pair per block.
// ------------------------------------------------------------------------------------------------
__global__ void elbo_lik_kernel(const LikParams p_in, double* __restrict__ out) {
  const LikParams p = lik_resolve(p_in);
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0, kl = 0.0;
  const int64_t ld = p.ldB;
  if (b < p.B) {
    if (p.model_kind == 0 && p.lik_kind[0] == 3) {
      // logisticsoftmax.jl:106-140 ; KLdivergences.jl:83-98 (Q2, Q10)
      const int K = p.Q;
      int cls = p.ycls[b];
      double alpha = p.alpha[b], beta = (double)K;
      double psi = digamma_pos(alpha), lb = log(beta);
      e += -(double)K * 0.6931471805599453;
      for (int k = 0; k < K; ++k) {
        double mu = p.mean_f[k * ld + b], var = p.var_f[k * ld + b];
        double g = p.gamma[k * ld + b], th = p.theta[k * ld + b], c = p.c[k * ld + b];
        double yk = (k == cls) ? 1.0 : 0.0;
        e += -(g + yk) * 0.6931471805599453 + 0.5 * (mu * (yk - g) - th * mu * mu - th * var);
        kl += (yk + g) * logcosh_d(0.5 * c) - 0.5 * c * c * th;          // PolyaGammaKL
        kl += alpha / beta - g + xlogx_d(g) - g * (psi - lb);           // PoissonKL
      }
      kl += -alpha - lgamma(alpha) - (1.0 - alpha) * psi;               // GammaEntropy (per-sample part)
      if (b == 0) kl += lb;                                             // Q2: log(beta[1]) once
    } else if (p.model_kind == 0 && p.lik_kind[0] == 8) {
      // heteroscedastic.jl:143-179 ; KLdivergences.jl:83-89, 96-98
      const double lam = p.lam[0], y = p.yb[b];
      double m1 = p.mean_f[b], v1 = p.var_f[b], m2 = p.mean_f[ld + b], v2 = p.var_f[ld + b];
      double c = p.c[b], gam = p.gamma[b], th = p.theta[b];
      e += 0.5 * log(lam) - log(2.0 * sqrt(6.283185307179586));
      e += 0.5 * (m2 * (0.5 - gam) - m2 * m2 * th - v2 * th);
      double lam0 = 0.5 * lam * ((y - m1) * (y - m1) + v1);
      e -= lam0 - gam + xlogx_d(gam) - gam * log(lam0);                      // PoissonKL(gamma, lam0, log lam0)
      kl += (0.5 + gam) * logcosh_d(0.5 * c) - 0.5 * c * c * th;            // PolyaGammaKL
    } else {
      int T = p.model_kind == 0 ? 1 : p.n_task;
      for (int t = 0; t < T; ++t) {
        double mu, var;
        if (p.model_kind == 0) { mu = p.mean_f[b]; var = p.var_f[b]; }
        else { mu = p.tmu[t * ld + b]; var = p.tvar[t * ld + b]; }
        double y = p.yb[t * ld + b], th = p.theta[t * ld + b], c = p.c[t * ld + b];
        int kind = p.lik_kind[t];
        double p0 = p.p0[t], p1 = p.p1[t];
        if (kind == 1) {  // logistic.jl:73-92 (Q1: theta*mu, not theta*mu^2)
          e += -0.5 * 0.6931471805599453 + 0.5 * (mu * y - th * var - th * mu);
          kl += logcosh_d(0.5 * c) - 0.5 * c * c * th;
        } else if (kind == 2) {  // studentt.jl:103-127 ; KLdivergences.jl:62-67
          double al = 0.5 * (p0 + 1.0), ap = 0.5 * p0, bp = ap * p1 * p1;
          e += -0.5 * log(6.283185307179586 * p1 * p1) - (log(c) - digamma_pos(al)) -
               0.5 * (th * var + th * mu * mu - 2.0 * th * mu * y + th * y * y);
          kl += (al - ap) * digamma_pos(al) - log(tgamma(al)) + log(tgamma(ap)) + ap * (log(c) - log(bp)) + al * (bp - c) / c;
        } else if (kind == 4) {  // laplace.jl:95-125 ; GIGEntropy (KLdivergences.jl:105-114) with p = 1/2 in closed form:
          // K_{1/2}(s) = K_{-1/2}(s) = sqrt(pi / 2s) e^-s,  K_{3/2}(s) = K_{1/2}(s) (1 + 1/s)
          double a = 1.0 / (p0 * p0), bb = c;   // c holds b
          e += -0.5 * log(6.283185307179586) + 0.5 * log(th) - 0.5 * (th * var + th * mu * mu - 2.0 * th * mu * y + th * y * y);
          double sq = sqrt(a) * bb;
          // as written in the reference (scalar a and p in sum / mapreduce, quirk Q13): log(a) / 2 and log(2 K_p(sqrt(ab))) enter once,
          // for the first sample of the batch; -log(b^2) / 2 and the Bessel-ratio term are summed over every sample
          double ent = -log(bb) + (sq + 0.5);
          if (b == 0) ent += 0.5 * log(a) + 0.6931471805599453 + 0.5 * log(3.141592653589793 / (2.0 * sq)) - sq;
          double ex = -log(2.0 * p0 * p0) - (a * bb + bb * bb * sqrt(a)) / (a * bb * bb * p0 * p0) / 2.0;
          kl += ent - ex;
        } else if (kind == 5) {  // bayesiansvm.jl:68-89 (signs as written in the reference)
          double d = 1.0 - y * mu, sc = sqrt(c);
          e += -0.5 * 0.6931471805599453 + mu * y - 0.5 * th * var + th * d * d;
          kl += 0.5 * log(c) + (0.6931471805599453 + 0.5 * log(3.141592653589793 / (2.0 * sc)) - sc) - 0.5 * sc;
        } else if (kind == 6) {  // negativebinomial.jl:102-128 (dot(theta, mu) as written)
          e += lgamma(y + p0) - lgamma(y + 1.0) - lgamma(p0) - 0.6931471805599453 * (y + p0);
          e += 0.5 * mu * (y - p0) - 0.5 * th * mu - 0.5 * th * var;
          kl += (y + p0) * logcosh_d(0.5 * c) - 0.5 * c * c * th;
        } else if (kind == 7) {  // poisson.jl:110-136 ; KLdivergences.jl:74-76
          double lam = p.lam[t], g = p.gamma[t * ld + b];
          e += 0.5 * (mu * (y - g) - th * mu * mu - th * var) + y * log(lam) - lgamma(y + 1.0) - 0.6931471805599453 * (y + g);
          kl += lam - (1.0 + log(lam)) * g + xlogx_d(g);
          kl += (y + g) * logcosh_d(0.5 * c) - 0.5 * c * c * th;
        } else {  // gaussian.jl:82-95
          double d = y - mu;
          e += -0.5 * (log(6.283185307179586) + log(p0) + (d * d + var) / p0);
        }
      }
    }
  }
  __shared__ double se[8], sk[8];
  e = warp_sum(e); kl = warp_sum(kl);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { se[w] = e; sk[w] = kl; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c2 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += se[i]; c2 += sk[i]; }
    atomicAdd(out + 0, a);
    atomicAdd(out + 2, c2);
  }
}

// GaussianKL (functions/KLdivergences.jl:11-18) in the whitened basis: with Sigma = L Sigma_v L^T, mu = L mu_v,
//   logdet K - logdet Sigma = logdet P_v,  tr(K \ Sigma) = tr(Sigma_v),  invquad(K, mu - mu0) = |mu_v - L^-1 mu0|^2.
// out[0] += tr(Sigma_v), out[1] += |mu_v - mu0_v|^2        (single block)
__global__ void gauss_kl_kernel(const double* __restrict__ SigmaV, int64_t ld, int m, const double* __restrict__ muv,
                                const double* __restrict__ mu0v, double* __restrict__ out) {
  double tr = 0.0, q = 0.0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    tr += SigmaV[(int64_t)j * ld + j];
    double d = muv[j] - mu0v[j];
    q += d * d;
  }
  __shared__ double s1[8], s2[8];
  tr = warp_sum(tr); q = warp_sum(q);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s1[w] = tr; s2[w] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += s1[i]; c += s2[i]; }
    out[0] += a;
    out[1] += c;
  }
}

// ------------------------------------------------------------------------------------------------
// m x m tail (always fp64)
// ------------------------------------------------------------------------------------------------
struct TailParams {
  int m, mp;            // logical / padded (power-of-two multiple of 64) size
  int64_t ld;           // = mp : leading dimension of every fp64 m x m matrix
  int n_split; int64_t gpart_stride; int64_t gpart_ld;  // G partials [n_split][m][gpart_ld]
  int g_mirrored;       // 1: both triangles hold identical values (read coalesced); 0: take the upper triangle (Q5)
  const double* v1;     // V^T grad_mu (un-scaled by rho)
  const double* mu0v;   // L^-1 mu0 (whitened prior mean)
  double* eta1; double* eta2; double* P;
  const int64_t* counters;  // [0] = Robbins-Monro t (starts at 1)
  const double* lr;         // step size of this iteration (written by lik_update_kernel)
  double* v1_zero;          // non-null: clear V^T grad_mu after use (the next step accumulates into it without a memset)
  int stochastic; double rm_kappa, rm_tau, rho;
  double* logdet; int* status;
  // OnlineSVGP (analyticVI.jl:183-203): constant terms of the natural gradient carried over from the previous inducing set,
  // whitened: eta1_off = (kappa_a L)^T prev_eta1, eta2_off = (kappa_a L)^T invD_a (kappa_a L) / 2.  Null for every other model.
  const double* eta1_off; const double* eta2_off;
  // split Gram (agp_engine.cu): 0 = every element, 1 = only the first 128 x 128 block (grid (1, 128)), 2 = everything else
  int blk_mode;
  // non-null (whole-matrix launches in front of the multi-launch tail): the 64 blocks that hold rows 0..63 of column block 0 count
  // themselves here once their part of P_v is written, so that the tail's first kernel can start on tile (0, 0) while the rest of
  // this grid is still running (tail2_potf2_first_kernel, TailStepParams::early_flag)
  int* tile0_flag;
};

// natural gradient + global update of the natural parameters (inference/analyticVI.jl:160-180, 229-246;
// inference/optimisers.jl:14-19) in the whitened basis (eta1_v = L^T eta1, eta2_v = L^T eta2 L, so that the prior
// precision K^-1 becomes I and K \ mu0 becomes L^-1 mu0), and P_v = -2 eta2_v whose inverse is Sigma_v
// (inference.jl:26).  G is symmetrised from its upper triangle like Julia's Symmetric() (analyticVI.jl:238, Q5).
template <typename TG>
__device__ __forceinline__ void combine_body(const TailParams& p, const TG* __restrict__ Gpart) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= p.mp) return;
  if (p.blk_mode == 2 && i < 128 && j < 128) return;
  const double lr = *p.lr;
  if (i >= p.m || j >= p.m) {
    p.P[(int64_t)i * p.ld + j] = (i == j) ? 1.0 : 0.0;
    return;
  }
  int a = p.g_mirrored ? i : min(i, j), b = p.g_mirrored ? j : max(i, j);
  double g = 0.0;
  const TG* gp = Gpart + (int64_t)a * p.gpart_ld + b;
  for (int s0 = 0; s0 < p.n_split; s0 += 16) {   // up to 16 independent loads in flight (the split-K partials are 1 MB apart)
    TG v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (s0 + u < p.n_split) ? gp[(int64_t)(s0 + u) * p.gpart_stride] : TG(0);
#pragma unroll
    for (int u = 0; u < 16; ++u) g += (double)v[u];
  }
  int64_t o = (int64_t)i * p.ld + j;
  double e2 = p.eta2[o];
  if (p.eta2_off) g += p.eta2_off[o];
  double d2 = -(g + (i == j ? 0.5 : 0.0)) - e2;
  e2 += lr * d2;
  p.eta2[o] = e2;
  p.P[o] = -2.0 * e2;
  if (i == 0) {
    double e1 = p.eta1[j];
    double d1 = p.rho * p.v1[j] + p.mu0v[j] + (p.eta1_off ? p.eta1_off[j] : 0.0) - e1;
    p.eta1[j] = e1 + lr * d1;
    if (p.v1_zero) p.v1_zero[j] = 0.0;
    if (j == 0) *p.logdet = 0.0;
  }
}
template <typename TG>
__global__ void combine_kernel(const TailParams p, const TG* __restrict__ Gpart) {
  pdl_prologue();
  combine_body<TG>(p, Gpart);
  if (p.tile0_flag && blockIdx.x == 0 && blockIdx.y < 64) {     // block-uniform
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p.tile0_flag, 1);
  }
}
// the same update for the latents of a batch (blockIdx.z): `common` holds everything the latents share, the per-latent pointers follow
struct CombineBatch {
  const double* v1[SMALL_NB]; const double* mu0v[SMALL_NB]; double* eta1[SMALL_NB]; double* eta2[SMALL_NB]; double* P[SMALL_NB];
  double* v1_zero[SMALL_NB]; double* logdet[SMALL_NB]; const float* G[SMALL_NB];
};
__global__ void combine_batched_kernel(const TailParams common, const CombineBatch bt) {
  pdl_prologue();
  const int z = blockIdx.z;
  TailParams p = common;
  p.v1 = bt.v1[z]; p.mu0v = bt.mu0v[z]; p.eta1 = bt.eta1[z]; p.eta2 = bt.eta2[z]; p.P = bt.P[z]; p.v1_zero = bt.v1_zero[z]; p.logdet = bt.logdet[z];
  combine_body<float>(p, bt.G[z]);
}

// The same update, four consecutive columns per thread (16-byte loads of the fp32 split-K partials, all slices in flight at once): the
// tf32x3 path of a whole matrix with m = mp a multiple of 4 (the partials are mirrored there, so rows are read as stored).  One block
// per row.  Opt-in (AGP_COMBINE4=1): measured slower than the scalar kernel at C2 (step 185.2 vs 181.3 us) - a quarter of the threads, so fewer loads in flight.
__global__ void __launch_bounds__(128) combine4_kernel(const TailParams p, const float* __restrict__ Gpart) {
  pdl_prologue();
  const int i = blockIdx.y;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j >= p.m) return;
  const double lr = *p.lr;
  const float4* gp = reinterpret_cast<const float4*>(Gpart + (int64_t)i * p.gpart_ld + j);
  const int64_t st4 = p.gpart_stride / 4;
  double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
  for (int s0 = 0; s0 < p.n_split; s0 += 16) {
    float4 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = (s0 + u < p.n_split) ? gp[(int64_t)(s0 + u) * st4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 16; ++u) { g0 += (double)v[u].x; g1 += (double)v[u].y; g2 += (double)v[u].z; g3 += (double)v[u].w; }
  }
  const int64_t o = (int64_t)i * p.ld + j;
  double2 ea = *reinterpret_cast<const double2*>(p.eta2 + o), eb = *reinterpret_cast<const double2*>(p.eta2 + o + 2);
  if (p.eta2_off) {
    const double2 fa = *reinterpret_cast<const double2*>(p.eta2_off + o), fb = *reinterpret_cast<const double2*>(p.eta2_off + o + 2);
    g0 += fa.x; g1 += fa.y; g2 += fb.x; g3 += fb.y;
  }
  ea.x += lr * (-(g0 + (i == j ? 0.5 : 0.0)) - ea.x);
  ea.y += lr * (-(g1 + (i == j + 1 ? 0.5 : 0.0)) - ea.y);
  eb.x += lr * (-(g2 + (i == j + 2 ? 0.5 : 0.0)) - eb.x);
  eb.y += lr * (-(g3 + (i == j + 3 ? 0.5 : 0.0)) - eb.y);
  *reinterpret_cast<double2*>(p.eta2 + o) = ea; *reinterpret_cast<double2*>(p.eta2 + o + 2) = eb;
  *reinterpret_cast<double2*>(p.P + o) = make_double2(-2.0 * ea.x, -2.0 * ea.y);
  *reinterpret_cast<double2*>(p.P + o + 2) = make_double2(-2.0 * eb.x, -2.0 * eb.y);
  if (i == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double e1 = p.eta1[j + u];
      const double d1 = p.rho * p.v1[j + u] + p.mu0v[j + u] + (p.eta1_off ? p.eta1_off[j + u] : 0.0) - e1;
      p.eta1[j + u] = e1 + lr * d1;
      if (p.v1_zero) p.v1_zero[j + u] = 0.0;
    }
    if (j == 0) *p.logdet = 0.0;
  }
}

// Unblocked Cholesky of one 64 x 64 diagonal block held in shared memory, fused with the inverse of
// its factor (the rank-1 update that eliminates column j is applied to [A | W], W starting as I, so
// W ends as L^-1).  In: lower triangle of Ablk.  Out: L in the lower triangle of Ablk, L^-1 (lower,
// explicit zeros above the diagonal) in Xblk, logdet += 2 sum log L_jj.
constexpr int POTF2_NB = 64;
__global__ void __launch_bounds__(256) potf2_inv_kernel(double* __restrict__ Ablk, int64_t lda, double* __restrict__ Xblk,
                                                        int64_t ldx, double* __restrict__ logdet, int* __restrict__ status) {
  constexpr int NB = POTF2_NB, LDS = NB + 1;
  extern __shared__ double sm[];
  double* a = sm;              // [NB][LDS]
  double* w = sm + NB * LDS;   // [NB][LDS]
  const int tid = threadIdx.x;
  for (int e = tid; e < NB * NB; e += 256) {
    int i = e / NB, j = e % NB;
    a[i * LDS + j] = (j <= i) ? Ablk[(int64_t)i * lda + j] : 0.0;
    w[i * LDS + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  __shared__ double s_r;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int j = 0; j < NB; ++j) {
    // one thread forms 1/sqrt(pivot); everybody else waits at the barrier (the pivot chain is serial anyway)
    if (tid == 0) {
      double d = a[j * LDS + j];
      if (!(d > 0.0)) { s_bad = 1; d = 1.0; }
      s_r = rsqrt(d);
    }
    __syncthreads();
    const double r = s_r;
    // scale column j of A (rows >= j) and row j of W (cols <= j)
    if (tid < NB) {
      if (tid >= j) a[tid * LDS + j] *= r;
    } else if (tid < 2 * NB) {
      int c = tid - NB;
      if (c <= j) w[j * LDS + c] *= r;
    }
    __syncthreads();
    // eliminate: rows i > j.  cols c > j (c <= i): A update;  cols c <= j: W update.
    int nrow = NB - 1 - j;
    for (int e = tid; e < nrow * NB; e += 256) {
      int i = j + 1 + e / NB, c = e % NB;
      double lij = a[i * LDS + j];
      if (c > j) { if (c <= i) a[i * LDS + c] -= lij * a[c * LDS + j]; }
      else w[i * LDS + c] -= lij * w[j * LDS + c];
    }
    __syncthreads();
  }
  // logdet += 2 sum_j log L_jj, one warp-parallel pass at the end
  if (tid < 32) {
    double ld_acc = 2.0 * (log(a[tid * LDS + tid]) + log(a[(tid + 32) * LDS + tid + 32]));
    ld_acc = warp_sum(ld_acc);
    if (tid == 0) {
      atomicAdd(logdet, ld_acc);
      if (s_bad) atomicOr(status, ST_NOT_POSDEF);
    }
  }
  for (int e = tid; e < NB * NB; e += 256) {
    int i = e / NB, j = e % NB;
    if (j <= i) Ablk[(int64_t)i * lda + j] = a[i * LDS + j];
    Xblk[(int64_t)i * ldx + j] = (j <= i) ? w[i * LDS + j] : 0.0;
  }
}

// Sigma lower triangle (from the lower-only X^T X product) -> full symmetric fp64 master + T shadow
template <typename T>
__global__ void symmetrize_shadow_kernel(double* __restrict__ S, int64_t ld, int m, T* __restrict__ shadow, int64_t lds) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= m || i >= m) return;
  double v = (j <= i) ? S[(int64_t)i * ld + j] : S[(int64_t)j * ld + i];
  if (j > i) S[(int64_t)i * ld + j] = v;
  if (shadow) shadow[(int64_t)i * lds + j] = (T)v;
}

// fp64 matrix (ld) -> T shadow copy (lds)
template <typename T>
__global__ void shadow_kernel(const double* __restrict__ S, int64_t ld, int m, T* __restrict__ shadow, int64_t lds) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= m || i >= m) return;
  shadow[(int64_t)i * lds + j] = (T)S[(int64_t)i * ld + j];
}

// y = S x  (mu = Sigma * eta1, inference/inference.jl:27); one warp per row
__global__ void symv_kernel(const double* __restrict__ S, int64_t ld, int m, const double* __restrict__ x,
                            double* __restrict__ y) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= m) return;
  double s = 0.0;
  for (int j = lane; j < m; j += 32) s += S[(int64_t)row * ld + j] * x[j];
  s = warp_sum(s);
  if (lane == 0) y[row] = s;
}

// y = S^T x ; one thread per output column (S is m x m, tiny, off the hot path)
__global__ void matvec_t_kernel(const double* __restrict__ S, int64_t ld, int m, const double* __restrict__ x,
                                double* __restrict__ y) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  double s = 0.0;
  for (int i = 0; i < m; ++i) s += S[(int64_t)i * ld + j] * x[i];
  y[j] = s;
}

// lower triangle (incl. diagonal) of src -> dst, zeros above; used to keep copies of L and L^-1
__global__ void copy_lower_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int mp) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= mp) return;
  dst[(int64_t)i * ld + j] = (j <= i) ? src[(int64_t)i * ld + j] : 0.0;
}

// dst = alpha * src with identity padding outside the m x m block
__global__ void scale_pad_kernel(const double* __restrict__ src, double* __restrict__ dst, int64_t ld, int m, int mp, double alpha) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= mp) return;
  dst[(int64_t)i * ld + j] = (i < m && j < m) ? alpha * src[(int64_t)i * ld + j] : (i == j ? 1.0 : 0.0);
}

__global__ void bump_counters_kernel(int64_t* counters, int bump_t, int bump_cursor) {
  pdl_prologue();
  if (threadIdx.x == 0 && blockIdx.x == 0) { counters[0] += bump_t; counters[1] += bump_cursor; }
}

// small rectangular fp64 mat-vec products of the OnlineSVGP carry-over (off the hot path): S is [rows][ld]
__global__ void rect_matvec_t_kernel(const double* __restrict__ S, int64_t ld, int rows, int cols, const double* __restrict__ x, double* __restrict__ y) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;     // y = S^T x
  if (j >= cols) return;
  double a = 0.0;
  for (int i = 0; i < rows; ++i) a = fma(S[(int64_t)i * ld + j], x[i], a);
  y[j] = a;
}
__global__ void rect_matvec_kernel(const double* __restrict__ S, int64_t ld, int rows, int cols, const double* __restrict__ x, double* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;     // y = S x
  if (i >= rows) return;
  double a = 0.0;
  for (int j = 0; j < cols; ++j) a = fma(S[(int64_t)i * ld + j], x[j], a);
  y[i] = a;
}
// K_mm finishing touch in fp64 (gpblocks/latentgp.jl:206): exact diagonal variance + jitter, identity padding
__global__ void kmm_fix_kernel(double* __restrict__ K, int64_t ld, int m, int mp, double diag) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= mp) return;
  if (i >= m || j >= m) K[(int64_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
  else if (i == j) K[(int64_t)i * ld + j] = diag;
}

__global__ void set_identity_kernel(double* __restrict__ S, int64_t ld, int mp, double diag) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j >= mp) return;
  S[(int64_t)i * ld + j] = (i == j) ? diag : 0.0;
}

// Bernoulli predictive by Gauss-Hermite quadrature (likelihood/classification.jl:14-26)
__global__ void proba_logistic_kernel(const double* __restrict__ mu, const double* __restrict__ var, int64_t n,
                                      const double* __restrict__ nodes, const double* __restrict__ weights, int nn,
                                      double* __restrict__ p, double* __restrict__ pv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sd = sqrt(fmax(var[i], 0.0)), m = mu[i];
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nn; ++k) {
    double l = logistic_d(nodes[k] * sd + m);
    s1 += weights[k] * l;
    s2 += weights[k] * l * l;
  }
  p[i] = s1;
  pv[i] = fmax(s2 - s1 * s1, 0.0);
}

// compute_proba of the count likelihoods by quadrature (poisson.jl:43-55, negativebinomial.jl:47-62, classification.jl:14-26)
__global__ void proba_link_kernel(int link, double p0, const double* __restrict__ mu, const double* __restrict__ var, int64_t n,
                                  const double* __restrict__ nodes, const double* __restrict__ weights, int nn,
                                  double* __restrict__ p, double* __restrict__ pv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sd = sqrt(fmax(var[i], 0.0)), m = mu[i];
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nn; ++k) {
    double l = logistic_d(nodes[k] * sd + m);
    double v = link == 1 ? p0 * l : (link == 2 ? l * p0 / (1.0 - l) : l);
    if (link == 3) {  // svmlikelihood (bayesiansvm.jl:27-35)
      double f = nodes[k] * sd + m, pos = exp(-2.0 * fmax(1.0 - f, 0.0)), neg = exp(-2.0 * fmax(1.0 + f, 0.0));
      v = pos / (pos + neg);
    }
    s1 += weights[k] * v;
    s2 += weights[k] * v * v;
  }
  p[i] = s1;
  pv[i] = (link == 0 || link == 3) ? fmax(s2 - s1 * s1, 0.0) : s2 - s1 * s1;
}

}  // namespace agp
