// agp_lik.cuh -- device-side building blocks shared by the step kernels (agp_kernels.cuh) and the tcgen05 GEMM epilogues (agp_umma.cu):
// status bits, programmatic-dependent-launch helpers, special functions, and the likelihood local updates + expectation gradients of
// one minibatch sample (LikParams / lik_update_sample).  Only inline device functions and plain structs: safe in every translation unit.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace agp {

// sticky device status bits (read back by agp_sync)
enum : int { ST_KTILDE = 1, ST_NOT_POSDEF = 2, ST_PEER_TIMEOUT = 4, ST_NS_NOCONV = 8, ST_TAIL_TIMEOUT = 16 };

// Programmatic dependent launch (griddepcontrol).  Kernels on the per-step critical chain call pdl_prologue() first:
// launch_dependents lets the NEXT kernel of the chain become resident while this one runs, wait blocks until the
// PREVIOUS kernel has completed and its writes are visible.  Both are no-ops for a launch without the PDL attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// special functions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double digamma_pos(double x) {  // x > 0
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}
__device__ __forceinline__ double logistic_d(double x) { return 1.0 / (1.0 + exp(-x)); }
// functions/utils.jl:84-86
__device__ __forceinline__ double safe_expcosh_d(double mu, double c) {
  double v = exp(mu) / cosh(c);
  return isfinite(v) ? v : 2.0 * logistic_d(2.0 * fmax(mu, c));
}
// functions/utils.jl:89-91
__device__ __forceinline__ double logcosh_d(double c) { return log(exp(-2.0 * c) + 1.0) + c - 0.6931471805599453; }
__device__ __forceinline__ double xlogx_d(double x) { return x > 0.0 ? x * log(x) : 0.0; }

// ------------------------------------------------------------------------------------------------
// likelihood local updates + expectation gradients (one thread per minibatch sample, fp64)
// ------------------------------------------------------------------------------------------------
struct LikParams {
  int model_kind, n_task, Q, B;
  int64_t ldB;
  int latent_begin, n_latent_local;
  const int* lik_kind; const double* p0; const double* p1;  // [T] device
  const double* A;                                           // [T*Q] device (MOSVGP)
  const double* mean_f; const double* var_f;                 // [Q][ldB] latent moments (all latents)
  // labels: resident data + gather list, or (idx == nullptr) already in yb / ycls
  const double* y_all; int64_t n; const int* ycls_all; const int64_t* idx;
  double* yb; int* ycls;                                     // [T][ldB], [ldB]  minibatch labels (kept for the ELBO)
  double* c; double* theta; double* gamma; double* alpha;    // local variables [R][ldB] (R = T or K), alpha [ldB]
  double* tmu; double* tvar;                                 // [T][ldB] task moments (MOSVGP scratch / ELBO)
  double* gm; double* gs;                                    // [T][ldB] per-task gradients (MOSVGP scratch)
  double* gmu; double* gS;                                   // [n_latent_local][ldB] outputs
  int update;                                                // 1: local_updates!, 0: only (re)compute task moments
  // link parameters re-estimated by local_updates! (Poisson / Heteroscedastic lambda): value [T], accumulators [T][2]
  double* lam; double* lamacc;
  const double* qnodes; const double* qweights; int nq;      // Gauss-Hermite rule of `expectation` (functions/utils.jl:16-19)
  int need_reduce;                                           // some task accumulates into lamacc
  // GaussianLikelihood(opt_noise) (gaussian.jl:56-72): per-task ADAM state [T][4] = (mt, vt, beta1^t, beta2^t), flags [T], and
  // (eta, beta1, beta2, eps); the re-estimated sigma^2 is written back into p0[t] (every kernel reads p0 live)
  const int* noise_opt; double* noise_state; double n_eta, n_b1, n_b2, n_eps;
  // latent-sharded peer exchange: the moment arrays are double-buffered by exchange parity (see peer_sync_kernel)
  const int64_t* xepoch; int64_t par_stride;
};
__device__ __forceinline__ LikParams lik_resolve(const LikParams& in) {
  LikParams p = in;
  if (p.xepoch) { const int64_t off = (*p.xepoch & 1) * p.par_stride; p.mean_f += off; p.var_f += off; }
  return p;
}

// E[logistic(f)], f ~ N(mu, var), by the Gauss-Hermite rule (functions/utils.jl:16-19)
__device__ __forceinline__ double expect_logistic(const double* __restrict__ nodes, const double* __restrict__ w, int nq, double mu, double var) {
  const double sd = sqrt(fmax(var, 0.0));
  double s = 0.0;
  for (int k = 0; k < nq; ++k) s += w[k] * logistic_d(nodes[k] * sd + mu);
  return s;
}

// one single-latent likelihood: c, theta (gamma for Poisson) and the two expectation gradients
__device__ __forceinline__ void lik_single(int kind, double p0, double p1, double lam, double y, double mu, double var, double& c,
                                           double& th, double& gam, double& gm, double& gs) {
  gam = 0.0;
  if (kind == 1) {  // likelihood/logistic.jl:39-51, 64-69
    c = sqrt(mu * mu + var);
    th = tanh(0.5 * c) / (2.0 * c);
    gm = 0.5 * y;
  } else if (kind == 2) {  // likelihood/studentt.jl:68-82, 96-99
    double d = mu - y;
    c = 0.5 * (d * d + var + p1 * p1 * p0);
    th = 0.5 * (p0 + 1.0) / c;
    gm = th * y;
  } else if (kind == 4) {  // likelihood/laplace.jl:61-74, 87-92 : c holds b = sqrt(E[(f-y)^2]), theta = sqrt(a) / b, a = beta^-2
    double d = mu - y;
    c = sqrt(d * d + var);
    th = (1.0 / p0) / c;
    gm = th * y;
  } else if (kind == 5) {  // likelihood/bayesiansvm.jl:40-64 : c = E[(1 - y f)^2] (not its root), theta = c^-1/2
    double d = 1.0 - y * mu;
    c = d * d + var;
    th = 1.0 / sqrt(c);
    gm = y * (th + 1.0);
  } else if (kind == 6) {  // likelihood/negativebinomial.jl:69-99
    c = sqrt(mu * mu + var);
    th = (p0 + y) * tanh(0.5 * c) / c;
    gm = 0.5 * (y - p0);
  } else if (kind == 7) {  // likelihood/poisson.jl:65-108 (lambda = the value BEFORE this call's re-estimation)
    c = sqrt(mu * mu + var);
    gam = lam * safe_expcosh_d(-0.5 * mu, 0.5 * c) / 2.0;
    th = (y + gam) / c * tanh(0.5 * c);
    gm = 0.5 * (y - gam);
  } else {  // likelihood/gaussian.jl:56-80
    c = 0.0;
    th = 1.0 / p0;
    gm = y / p0;
  }
  gs = 0.5 * th;
}

// body of local_updates! for sample b; acc[2t], acc[2t+1]: this sample's contribution to the lambda statistics of task t
// (only the FIRST reducing task is accumulated in registers: r0, r1; MOSVGP Poisson tasks use atomics directly)
__device__ __forceinline__ void lik_update_sample(const LikParams& p, int b, double& r0, double& r1) {
  const int64_t ld = p.ldB;
  const int64_t src = p.idx ? p.idx[b] : (int64_t)b;
  if (p.model_kind == 0 && p.lik_kind[0] == 3) {
    // ---- LogisticSoftMax (likelihood/logisticsoftmax.jl:55-79, 98-103) ----
    const int K = p.Q;
    int cls = p.idx ? p.ycls_all[src] : p.ycls[b];
    if (p.idx) p.ycls[b] = cls;
    if (!p.update) return;
    double alpha = p.alpha[b];
    const double beta = (double)K;  // beta stays = K forever (Q6)
    for (int k = 0; k < K; ++k) {
      double mu = p.mean_f[k * ld + b], var = p.var_f[k * ld + b];
      p.c[k * ld + b] = sqrt(mu * mu + var);
    }
    for (int pass = 0; pass < 2; ++pass) {
      double e = exp(digamma_pos(alpha));
      double s = 0.0;
      for (int k = 0; k < K; ++k) {
        double mu = p.mean_f[k * ld + b], c = p.c[k * ld + b];
        double g = e * safe_expcosh_d(-0.5 * mu, 0.5 * c) / (2.0 * beta);
        p.gamma[k * ld + b] = g;
        s += g;
      }
      alpha = 1.0 + s;
    }
    p.alpha[b] = alpha;
    for (int k = 0; k < K; ++k) {
      double c = p.c[k * ld + b], g = p.gamma[k * ld + b];
      double yk = (k == cls) ? 1.0 : 0.0;
      double th = (yk + g) * tanh(0.5 * c) / (2.0 * c);
      p.theta[k * ld + b] = th;
      int ql = k - p.latent_begin;
      if (ql >= 0 && ql < p.n_latent_local) {
        p.gmu[ql * ld + b] = 0.5 * (yk - g);
        p.gS[ql * ld + b] = 0.5 * th;
      }
    }
    return;
  }
  if (p.model_kind == 0 && p.lik_kind[0] == 8) {
    // ---- Heteroscedastic Gaussian (likelihood/heteroscedastic.jl:73-100): latent 0 = f, latent 1 = g.
    // rows: c[0] = c, c[1] = phi, gamma[0] = gamma, gamma[1] = sigma_g, theta[0] = theta.  The gradients need the
    // lambda re-estimated from this minibatch (:98, :111-127): hetero_grad_kernel, after lik_lambda_kernel.
    double y = p.idx ? p.y_all[src] : p.yb[b];
    if (p.idx) p.yb[b] = y;
    if (!p.update) return;
    const double lam = p.lam[0];
    double m1 = p.mean_f[b], v1 = p.var_f[b], m2 = p.mean_f[ld + b], v2 = p.var_f[ld + b];
    double d = m1 - y;
    double phi = 0.5 * (d * d + v1);
    double c = sqrt(m2 * m2 + v2);
    double sg = safe_expcosh_d(-0.5 * m2, 0.5 * c) / 2.0;
    double gam = lam * phi * sg;
    double th = (0.5 + gam) * tanh(0.5 * c) / (2.0 * c);
    p.c[b] = c; p.c[ld + b] = phi; p.gamma[b] = gam; p.gamma[ld + b] = sg; p.theta[b] = th;
    r0 = phi * (1.0 - sg);
    return;
  }
  if (p.model_kind == 0) {
    // ---- single-latent SVGP ----
    double y = p.idx ? p.y_all[src] : p.yb[b];
    if (p.idx) p.yb[b] = y;
    if (!p.update) return;
    double c, th, gam, gm, gs;
    const int kind = p.lik_kind[0];
    const double mu = p.mean_f[b], var = p.var_f[b];
    lik_single(kind, p.p0[0], p.p1[0], p.lam[0], y, mu, var, c, th, gam, gm, gs);
    p.c[b] = c; p.theta[b] = th;
    p.gmu[b] = gm; p.gS[b] = gs;
    if (kind == 0 && p.noise_opt && p.noise_opt[0] && p.update == 1) r0 = (y - mu) * (y - mu) + var;   // gaussian.jl:63
    if (kind == 7 && p.update == 1) {  // poisson.jl:80 : lambda = sum(y) / sum(E[logistic(f)])
      p.gamma[b] = gam;
      r0 = y;
      r1 = expect_logistic(p.qnodes, p.qweights, p.nq, mu, var);
    }
    return;
  }
  // ---- MOSVGP (models/single_and_multi_output_utils.jl:24-84) ----
  const int T = p.n_task, Q = p.Q;
  for (int t = 0; t < T; ++t) {
    double y = p.idx ? p.y_all[(int64_t)t * p.n + src] : p.yb[t * ld + b];
    if (p.idx) p.yb[t * ld + b] = y;
    double mt = 0.0, vt = 0.0;
    for (int q = 0; q < Q; ++q) {
      double a = p.A[t * Q + q];
      mt += a * p.mean_f[q * ld + b];
      vt += a * a * p.var_f[q * ld + b];
    }
    p.tmu[t * ld + b] = mt; p.tvar[t * ld + b] = vt;
    const int kind_t = p.lik_kind[t];
    if (p.update == 1 || (p.update == 2 && kind_t == 0)) {   // update == 2: second pass after the noise re-estimation, Gaussian tasks only
      double c, th, gam, gm, gs;
      const int kind = kind_t;
      lik_single(kind, p.p0[t], p.p1[t], p.lam[t], y, mt, vt, c, th, gam, gm, gs);
      p.c[t * ld + b] = c; p.theta[t * ld + b] = th;
      p.gm[t * ld + b] = gm; p.gs[t * ld + b] = gs;
      if (kind == 0 && p.noise_opt && p.noise_opt[t] && p.update == 1) atomicAdd(p.lamacc + 2 * t, (y - mt) * (y - mt) + vt);
      if (kind == 7 && p.update == 1) {
        p.gamma[t * ld + b] = gam;
        atomicAdd(p.lamacc + 2 * t, y);
        atomicAdd(p.lamacc + 2 * t + 1, expect_logistic(p.qnodes, p.qweights, p.nq, mt, vt));
      }
    }
  }
  if (!p.update) return;
  for (int ql = 0; ql < p.n_latent_local; ++ql) {
    int q = p.latent_begin + ql;
    double muq = p.mean_f[q * ld + b];
    double a1 = 0.0, a2 = 0.0;
    for (int t = 0; t < T; ++t) {
      double a = p.A[t * Q + q];
      double others = p.tmu[t * ld + b] - a * muq;
      a1 += a * (p.gm[t * ld + b] - 2.0 * p.gs[t * ld + b] * others);
      a2 += a * a * p.gs[t * ld + b];
    }
    p.gmu[ql * ld + b] = a1;
    p.gS[ql * ld + b] = a2;
  }
}

}  // namespace agp
