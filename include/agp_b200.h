/*
 * agp_b200.h -- C ABI of the B200-native sparse-variational-GP CAVI engine.
 *
 * The reference (theogf/AugmentedGaussianProcesses.jl, pure Julia) has NO FFI boundary for this
 * path; the seam is Julia multiple dispatch.  Each entry point below names the reference method
 * (file:line under /root/reference/src) whose work it replaces.  A Julia shim overrides those
 * methods and `ccall`s this library (julia/AGPB200.jl).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.
 *   - caller owns every host buffer; the library owns device memory behind opaque handles.
 *   - every function returns an int status (AGP_OK == 0) unless documented otherwise; the text of
 *     the last error is available from agp_last_error().
 *   - one calling thread per agp_ctx (the reference is single-threaded).
 *   - all host floating-point buffers are double unless a dtype argument says otherwise
 *     (the reference path is Float64-only: models/SVGP.jl:43, training/states.jl:15).
 *   - there is NO CPU fallback: without a CUDA device every call fails with AGP_ERR_CUDA.
 */
#ifndef AGP_B200_H
#define AGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AGP_ABI_VERSION 1

/* ---- status codes (reference error sites) ------------------------------------------------ */
#define AGP_OK 0
#define AGP_ERR_BAD_ARG 1       /* training/training.jl:27-29, models/SVGP.jl:45-49 (argument checks) */
#define AGP_ERR_CUDA 2          /* CUDA runtime failure (no reference counterpart)                    */
#define AGP_ERR_KTILDE_NONPOS 3 /* gpblocks/latentgp.jl:213  error("K̃ has negative values")          */
#define AGP_ERR_NOT_POSDEF 4    /* PosDefException from cholesky(): latentgp.jl:206, inference.jl:26  */
#define AGP_ERR_STATE 5         /* call order violated (e.g. step before data upload)                 */

/* ---- enums -------------------------------------------------------------------------------- */
/* KernelFunctions.jl kernels used by the reference call sites (gpblocks/latentgp.jl:206,210,212) */
#define AGP_KERNEL_SQEXP 0    /* SqExponentialKernel: exp(-d^2/2)               */
#define AGP_KERNEL_MATERN32 1 /* Matern32Kernel: (1+sqrt3 d) exp(-sqrt3 d)      */
#define AGP_KERNEL_MATERN52 2 /* Matern52Kernel                                 */

/* the files under likelihood/ (AnalyticVI methods only) */
#define AGP_LIK_GAUSSIAN 0        /* likelihood/gaussian.jl:56-95        p0 = sigma^2            */
#define AGP_LIK_LOGISTIC 1        /* likelihood/logistic.jl:39-92                                  */
#define AGP_LIK_STUDENTT 2        /* likelihood/studentt.jl:68-127       p0 = nu, p1 = sigma      */
#define AGP_LIK_LOGISTICSOFTMAX 3 /* likelihood/logisticsoftmax.jl:43-140  (n_latent = #classes)  */
#define AGP_LIK_LAPLACE 4         /* likelihood/laplace.jl:57-125        p0 = beta                */
#define AGP_LIK_BAYESIANSVM 5     /* likelihood/bayesiansvm.jl:40-89                               */
#define AGP_LIK_NEGBINOMIAL 6     /* likelihood/negativebinomial.jl:65-128  p0 = r                 */
#define AGP_LIK_POISSON 7         /* likelihood/poisson.jl:61-136  p0 = initial lambda (re-estimated
                                     by every local_updates!, poisson.jl:80); needs agp_set_quadrature */
#define AGP_LIK_HETEROSCEDASTIC 8 /* likelihood/heteroscedastic.jl:50-180  p0 = initial lambda
                                     (re-estimated, :98); n_latent = 2 (f and the noise GP g)      */

#define AGP_MODEL_SVGP 0   /* models/SVGP.jl:22-80: one likelihood, n_latent(likelihood) latents */
#define AGP_MODEL_MOSVGP 1 /* models/MOSVGP.jl:22-115: T single-latent tasks mixed from Q latents */
#define AGP_MODEL_VGP 2    /* models/VGP.jl: full variational GP (natural_gradient!(::VarLatent), analyticVI.jl:126-140):
                              pass Z = the n training inputs (m = n), AnalyticVI (stochastic = 0), and step with the full
                              index list (B = n); the engine then runs the SVGP algebra with kappa = I, Ktilde = 0 */
#define AGP_MODEL_MOVGP 3  /* models/MOVGP.jl: the multi-output full GP = AGP_MODEL_MOSVGP with the same Z = X convention  */

/* arithmetic of the B x m contractions; the m x m tail (eta update, Cholesky, inverse) is always f64 */
#define AGP_PREC_F64 0    /* everything in fp64 SIMT: bit-for-bit algorithm of the oracle        */
#define AGP_PREC_F32 1    /* fp32 SIMT contractions                                              */
#define AGP_PREC_TF32X3 2 /* tcgen05 tensor-core contractions, 3xTF32 error-compensated split    */

#define AGP_DTYPE_F64 0
#define AGP_DTYPE_F32 1
#define AGP_LAYOUT_COLMAJOR 0 /* Julia Matrix n x D (obsdim = 1): X[i + n*d]  (data/datacontainer.jl:64-66) */
#define AGP_LAYOUT_ROWMAJOR 1 /* C / NumPy n x D: X[i*D + d]                                               */

#define AGP_Y_REAL 0  /* double[n] per task: +-1 labels (classification.jl:29-39) or reals   */
#define AGP_Y_CLASS 1 /* int32[n] 0-based class index (one-hot of multiclass.jl:81-94)       */

typedef struct agp_ctx agp_ctx;
typedef struct agp_model agp_model;

/* Everything the constructors SVGP(...) / MOSVGP(...) + AnalyticVI()/AnalyticSVI(B) fix
 * (models/SVGP.jl:33-80, models/MOSVGP.jl:40-115, inference/analyticVI.jl:44-52). */
typedef struct agp_model_desc {
  int32_t model_kind;       /* AGP_MODEL_*                                                        */
  int32_t n_latent_global;  /* Q: latent GPs of the whole model (all ranks)                       */
  int32_t latent_begin;     /* first latent owned by this process (0 when not sharded)            */
  int32_t n_latent_local;   /* latents owned by this process                                      */
  int32_t m;                /* inducing points per latent                                         */
  int32_t D;                /* input dimension                                                    */
  int32_t batch_capacity;   /* largest B any step / ELBO call will use                            */
  int32_t precision;        /* AGP_PREC_*                                                         */
  int32_t stochastic;       /* 1: AnalyticSVI (optimisers.jl RobbinsMonro), 0: AnalyticVI (lr=1)  */
  double rm_kappa, rm_tau;  /* RobbinsMonro(kappa=0.51, tau=1)  inference/optimisers.jl:6-19      */
  double jitter;            /* functions/utils.jl:8 (1e-4 for Float64)                            */
  int32_t n_task;           /* T: 1 for SVGP, #likelihoods for MOSVGP                             */
  const int32_t* lik_kind;  /* [T] AGP_LIK_*                                                      */
  const double* lik_p0;     /* [T]                                                                */
  const double* lik_p1;     /* [T]                                                                */
  const double* A;          /* [T*Q] row-major mixing weights (MOSVGP.jl:103), NULL for SVGP      */
  const int32_t* kernel_kind;    /* [n_latent_local]                                              */
  const double* kernel_scale;    /* [n_latent_local] ScaleTransform(s) (= 1/lengthscale)          */
  const double* kernel_variance; /* [n_latent_local] sigma^2 * k                                  */
  const double* Z;               /* [n_latent_local][m][D] row-major inducing points              */
  const double* mu0;             /* [n_latent_local][m] prior mean at Z, or NULL (ZeroMean)       */
} agp_model_desc;

/* ---- context ------------------------------------------------------------------------------ */
int agp_abi_version(void);
/* cuda_stream: a cudaStream_t to launch on (e.g. the host framework's current stream), or NULL to
 * let the library create its own. */
int agp_ctx_create(int device, void* cuda_stream, agp_ctx** out);
void agp_ctx_destroy(agp_ctx* ctx);
const char* agp_last_error(const agp_ctx* ctx);

/* ---- model (SVGP.jl:33-80 / MOSVGP.jl:40-115; posterior init gpblocks/posterior.jl:29-37) -- */
int agp_model_create(agp_ctx* ctx, const agp_model_desc* desc, agp_model** out);
void agp_model_destroy(agp_model* model);

/* wrap_X / wrap_data (data/datacontainer.jl:64-74, data/utils.jl:20-28): upload once, keep resident.
 * y: array of T pointers (one per task). */
int agp_data_upload(agp_model* model, const void* X, int x_dtype, int x_layout, int64_t n,
                    const void* const* y, int y_kind);

/* Resident minibatch index lists (training/training.jl:51-53 draws them with StatsBase.sample; the
 * RNG stream cannot be reproduced outside Julia, so the lists cross the ABI).  idx: [n_lists][B].
 * Steps called with idx == NULL consume the lists in order (wrapping around). */
int agp_minibatches_upload(agp_model* model, const int64_t* idx, int64_t n_lists, int32_t B,
                           int32_t idx_base /* 0 or 1 */);

/* init_state (training/states.jl:1-9, 61-71): what a train! call without `state` does -- Robbins-Monro
 * counters back to 1, local variables back to init_local_vars (LogisticSoftMax alpha = K).  The
 * posterior (mu, Sigma, eta1, eta2) is NOT touched (it lives in the model, not in the state). */
int agp_state_reset(agp_model* model);

/* compute_K (gpblocks/latentgp.jl:205-207) for every local latent: K_mm + jitter I, its Cholesky,
 * K^-1, logdet K, K^-1 mu0.  Called once per train! (training/training.jl:41-43, quirk Q3). */
int agp_refresh_K(agp_model* model);
/* setkernel! (hyper-parameter change from the host side); follow with agp_refresh_K. */
int agp_set_kernel(agp_model* model, int32_t latent_local, int32_t kind, double scale, double variance);

/* ---- the hot path: update_parameters!(model::SVGP, state, x, y)  (training/training.jl:140-157)
 *   = compute_kernel_matrices (training.jl:187-208) -> compute_kappa (latentgp.jl:209-215)
 *   + variational_updates (inference/analyticVI.jl:62-111): mean_f/var_f (latentgp.jl:171-189),
 *     local_updates! + grad_E_mu/grad_E_Sigma (likelihood/<lik>.jl), natural_gradient!
 *     (analyticVI.jl:143-180), global_update! (analyticVI.jl:229-246, inference/inference.jl:25-28).
 * idx: host list of B row indices into the uploaded data (idx_base 0 or 1), or NULL to take the next
 * resident list.  rho = n / B (training.jl:30).
 * AGP_PREC_TF32X3: batch_capacity is a multiple of 128 and m > 64; m is padded to the 128-wide tile inside the engine (zero rows /
 * columns of L^-1 and X), and B itself may be ragged on the host-list path (the extra rows of the B x m products repeat sample 0 and
 * carry zero weights; host-row batches are padded with zero rows); resident minibatch lists keep B % 128 == 0. */
int agp_step(agp_model* model, const int64_t* idx, int32_t B, int32_t idx_base, double rho);
/* same, without the trailing synchronisation / error read-back (errors surface at agp_sync). */
int agp_step_async(agp_model* model, const int64_t* idx, int32_t B, int32_t idx_base, double rho);
/* same step for a minibatch handed over as HOST arrays (the x, y views update_parameters! receives):
 * xb: B x D, yb: array of T pointers of length-B vectors.  Copies host->device inside the call. */
int agp_step_batch(agp_model* model, const void* xb, int x_dtype, int x_layout, const void* const* yb,
                   int y_kind, int32_t B, double rho);
/* agp_step_batch without the trailing synchronisation (NOT yet run on a GPU -- written after the round's GPU budget was spent;
 * bench.py uses it only with --e2e-async).  The host->device copy of the batch runs on a copy stream into one of two
 * pre-staging slots, so the copy of batch i+1 overlaps the computation of step i; the host buffers may be reused once the
 * ticket of that step has been waited for (or after agp_sync).  *ticket identifies the step; agp_result_wait(ticket, mu)
 * blocks until that step is done, returns its sticky status like agp_sync and the posterior mean (length m) of latent 0
 * after that step.  Only the last two tickets are kept. */
int agp_step_batch_async(agp_model* model, const void* xb, int x_dtype, int x_layout, const void* const* yb,
                         int y_kind, int32_t B, double rho, int64_t* ticket);
int agp_result_wait(agp_model* model, int64_t ticket, double* mu);
/* wait for the stream and return the sticky device status (AGP_ERR_KTILDE_NONPOS / NOT_POSDEF). */
int agp_sync(agp_model* model);

/* ---- latent-sharded variant (one process per GPU) ---------------------------------
 * phase 1: kernel matrices + per-latent predictive moments of the owned latents (latentgp.jl:171-215)
 * written into rows [latent_begin, latent_begin+n_latent_local) of two [Q][ldB] double device arrays.
 * The host all-gathers those rows (NCCL) in place, then phase 2 runs local updates for all latents
 * (single_and_multi_output_utils.jl:24-84, logisticsoftmax.jl:55-79) and the natural-gradient /
 * global update of the owned latents. */
int agp_step_moments_async(agp_model* model, const int64_t* idx, int32_t B, int32_t idx_base);
int agp_step_update_async(agp_model* model, double rho);
/* device pointers + leading dimension of the moment arrays (which: 0 = mean_f, 1 = var_f). */
void* agp_moments_devptr(agp_model* model, int32_t which, int64_t* ld_out);

/* AnalyticSVI(B; optimiser = Descent(eta)) (inference/analyticVI.jl:28-52, global_update! :229-246): a constant step eta in (0, 1]
 * for the stochastic natural-gradient update instead of the default RobbinsMonro schedule; eta = 0 restores Robbins-Monro. */
int agp_set_step_size(agp_model* model, double eta);

/* GaussianLikelihood(sigma2; opt_noise) (likelihood/gaussian.jl:18-24, 56-72): kind 1 = ADAM(eta, (beta1, beta2)) on log sigma^2
 * inside every local update (the reference default for opt_noise = true is ADAM(0.05)); 0 = fixed noise.  The live sigma^2 is
 * read with agp_get_lik_param(task). */
int agp_set_noise_optimiser(agp_model* model, int32_t task, int32_t kind, double eta, double beta1, double beta2, double eps);

/* Gradient of ELBO(model, x, y, pr_means, kernels, Zs, state) (functions/ELBO.jl:15-21) w.r.t. each owned latent's kernel scale
 * (ScaleTransform s), kernel variance and inducing points, on the last minibatch with the posterior and local variables fixed:
 * what update_hyperparameters! (hyperparameter/autotuning.jl:86-140) obtains from Zygote.  d_scale, d_variance: [n_latent_local];
 * dZ: [n_latent_local][m][D] or NULL.  The host applies its optimiser (update_kernel! / update_Z!, autotuning_utils.jl:47-82),
 * pushes the new values with agp_set_kernel / agp_set_Z and calls agp_refresh_K.  Collective on a sharded model. */
int agp_hyper_grads(agp_model* model, double rho, double* d_scale, double* d_variance, double* dZ);
int agp_set_Z(agp_model* model, int32_t latent_local, const double* Z /* [m][D] row-major */);
/* Reference quirk Q3 switch.  on = 1 (what the reference does): the sparse update_hyperparameters! never raises
 * HyperParametersUpdated (its only setHPupdated!(.., true) call site, hyperparameter/autotuning.jl:45, is commented out, and
 * compute_kernel_matrices clears the flag, training/training.jl:187-208), so after agp_set_kernel / agp_set_Z the steps keep the
 * K_mm factor of the last agp_refresh_K next to a K_nm built from the new kernel / Z, and agp_hyper_grads evaluates its ELBO with
 * a fresh factorisation (functions/ELBO.jl:15-21) that it discards afterwards.  on = 0 (default of the bare ABI): a kernel / Z
 * change invalidates the factor and agp_refresh_K must follow. */
int agp_keep_stale_K(agp_model* model, int32_t on);

/* update_A! (models/single_and_multi_output_utils.jl:87-118; MOSVGP `Aoptimiser`, MOSVGP.jl:51,79-81): kind 0 = A fixed
 * (Aoptimiser = false), 1 = ADAM(eta, (beta1, beta2)) of Optimisers.jl with epsilon.  When on, every step first moves each
 * task's mixing row along the ADAM step of its ELBO gradient (local variables of the previous iteration) and renormalises
 * it, then runs variational_updates (training/training.jl:153-158).  agp_state_reset re-initialises the ADAM state
 * (init_state_A, training/states.jl:100-105).  agp_get_A: the current T x Q matrix, row-major. */
int agp_set_A_optimiser(agp_model* model, int32_t kind, double eta, double beta1, double beta2, double eps);
int agp_get_A(agp_model* model, double* A);

/* Device-side exchange over NVLink peer memory (same node, one process per GPU): every rank exports the IPC handle of its
 * moment block (64 bytes), the host gathers the handles over its process group (rank order) and attaches them.  Afterwards
 * every agp_step* / agp_elbo_moments_async call publishes the owned rows into all peers' arrays and waits for theirs inside
 * the step (two small kernels, CUDA-graph capturable): no NCCL call and no host round trip on the step path, and
 * agp_step / agp_step_async (incl. resident lists + agp_use_graph) become usable on a sharded model. */
int agp_peer_export(agp_model* model, void* handle64);
int agp_peer_attach(agp_model* model, int32_t world, int32_t rank, const void* handles /* [world][64] */);
/* undo agp_peer_attach (all ranks must agree: a rank that failed to map its peers makes everybody fall back) */
int agp_peer_detach(agp_model* model);

/* ---- ELBO(model, state, y) (inference/analyticVI.jl:255-297) on the last minibatch -----------
 * out[0] = rho * expec_loglikelihood, out[1] = GaussianKL summed over OWNED latents
 * (functions/KLdivergences.jl:2-18), out[2] = rho * AugmentedKL;  ELBO = out[0] - out[1] - out[2]
 * (a sharded host all-reduces out[1]). */
int agp_elbo(agp_model* model, double rho, double* out3);
/* sharded ELBO: recompute the owned latents' moments under the updated posterior (the host then
 * all-gathers them before calling agp_elbo; a non-sharded agp_elbo does this internally). */
int agp_elbo_moments_async(agp_model* model);

/* ---- posterior access / checkpoint (gpblocks/posterior.jl:21-27; train!(...; state) re-entry) - */
int agp_get_posterior(agp_model* model, int32_t latent_local, double* mu, double* Sigma, double* eta1,
                      double* eta2); /* any pointer may be NULL; matrices are m x m row-major */
int agp_set_posterior(agp_model* model, int32_t latent_local, const double* eta1, const double* eta2);
/* Robbins-Monro step counters (states.jl:67-68) and index-list cursor */
int agp_get_counters(agp_model* model, int64_t* rm_t, int64_t* cursor);
int agp_set_counters(agp_model* model, int64_t rm_t, int64_t cursor);
/* local variables of the last step: name in {"c","theta","gamma","alpha","mean_f","var_f",
 * "Ktilde","grad_mu","grad_Sigma"} (+ "b" for Laplace, "phi" / "sigma_g" for Heteroscedastic);
 * row = task / class / latent index; out: double[B]. */
int agp_get_local(agp_model* model, const char* name, int32_t row, double* out, int32_t B);
/* kernel matrices of the last step for one owned latent (state.kernel_matrices): B x m row-major */
int agp_get_kernel_matrices(agp_model* model, int32_t latent_local, double* Knm, double* kappa, int32_t B);
/* K^-1 and logdet K from the last agp_refresh_K */
int agp_get_Kinv(agp_model* model, int32_t latent_local, double* Kinv, double* logdetK);

/* ---- _predict_f (training/predictions.jl:25-50, diag = true) for the owned latents ------------
 * mu, var: [n_latent_local][nt] (var may be NULL when want_var == 0). */
int agp_predict_f(agp_model* model, const void* Xt, int x_dtype, int x_layout, int64_t nt, int want_var,
                  double* mu, double* var);
/* compute_proba(::BernoulliLikelihood) (likelihood/classification.jl:14-26): Gauss-Hermite
 * expectation of the logistic link; nodes/weights as in predictions.jl:4 (already scaled). */
int agp_proba_logistic(agp_model* model, const double* mu, const double* var, int64_t n, const double* nodes,
                       const double* weights, int32_t n_nodes, double* p, double* p_var);

/* Gauss-Hermite rule of `expectation` (functions/utils.jl:16-19; pred_nodes / pred_weights of
 * training/predictions.jl:4, i.e. gausshermite(100) nodes * sqrt2 and weights / sqrt(pi)).  Needed by the Poisson
 * local update (poisson.jl:80) and by agp_proba_link; n_nodes <= 128. */
int agp_set_quadrature(agp_model* model, const double* nodes, const double* weights, int32_t n_nodes);
/* link parameters that local_updates! re-estimates (PoissonLikelihood / HeteroscedasticLikelihood lambda:
 * l.invlink.lambda[1]); task = 0 for SVGP.  The value lives on the device between steps. */
int agp_get_lik_param(agp_model* model, int32_t task, double* value);
int agp_set_lik_param(agp_model* model, int32_t task, double value);
/* compute_proba by quadrature for the count likelihoods: link 0 = logistic (classification.jl:14-26), 1 = scaled
 * logistic p0 * sigma(f) (poisson.jl:43-55), 2 = negative-binomial mean sigma(f) p0 / (1 - sigma(f))
 * (negativebinomial.jl:47-62), 3 = svmlikelihood (bayesiansvm.jl:27-35).  Uses the rule of agp_set_quadrature.  pred, pred_var: double[n]. */
int agp_proba_link(agp_model* model, int32_t link, double p0, const double* mu, const double* var, int64_t n, double* pred,
                   double* pred_var);

/* _predict_f(...; cov = true, diag = false) (training/predictions.jl:45-49): posterior mean [n_latent_local][nt] and FULL predictive
 * covariance [n_latent_local][nt][nt] of the owned latents at nt <= 16384 test points (Xt: row-major fp64 [nt][D]); fp64 throughout. */
int agp_predict_f_cov(agp_model* model, const double* Xt, int64_t nt, double* mu, double* cov);

/* ---- measurement hooks ------------------------------------------------------------------------ */
/* per-phase CUDA-event timers around the kernels of a step (off by default; adds event records). */
int agp_profile_enable(agp_model* model, int on);
/* number of phases; names[i] points to a static string, ms[i] / launches[i] accumulate since enable */
int agp_profile_read(agp_model* model, int32_t max_phases, const char** names, double* ms, int64_t* launches);
/* kernels launched by this model since creation (the `gpu_launches` claim of bench.py). */
int64_t agp_launch_count(agp_model* model);
/* measurement hook for the roofline lines of bench.py: `reps` back-to-back launches of ONE hot kernel of the step on
 * the model's stream between two CUDA events (so the event overhead is amortised), on latent 0 with the resident
 * minibatch lists; *ms_per_launch = elapsed / reps.  which: 0 = K_nm construction (a new minibatch per launch),
 * 1 = V = K_nm L^-T, 2 = V X^T (+ row statistics), 3 = Gram product U^T U.  Needs a completed resident-list step
 * (AGP_ERR_STATE otherwise).  Does not change the posterior; the step pipeline is re-primed afterwards. */
int agp_time_kernel(agp_model* model, int32_t which, int32_t reps, double* ms_per_launch);
/* capture the step into a CUDA graph and replay it on later agp_step*(idx == NULL) calls. */
int agp_use_graph(agp_model* model, int on);

/* ---- OnlineSVGP (models/OnlineSVGP.jl, training/onlinetraining.jl; AnalyticVI only: OnlineSVGP.jl:46) ----------------------
 * The streaming model is an SVGP-kind model whose natural gradient carries two constant terms from the previous inducing set
 * Z_a (natural_gradient!(::OnlineVarLatent), inference/analyticVI.jl:183-203):
 *   eta1 = K^-1 mu0 + kappa^T grad_mu + kappa_a^T prev_eta1
 *   eta2 = -(kappa^T diag(grad_Sigma) kappa + kappa_a^T invD_a kappa_a / 2 + K^-1 / 2),   kappa_a = K_ab K^-1.
 * The inducing set itself is chosen by the caller (the reference draws it with the un-vendored InducingPoints.jl from Julia's
 * global RNG, onlinetraining.jl:157,175,193): a new set = a new model of that m, then agp_online_carry.
 *
 * agp_online_carry: replaces save_old_gp! (onlinetraining.jl:171-183; its outputs invD_a [ma][ma], prev_eta1 [ma], prev_L are
 * passed in, canonical form) and the K_ab / kappa_a / Ktilde_a part of compute_kappa(::OnlineVarLatent)
 * (gpblocks/latentgp.jl:217-230).  Za: previous inducing points, row-major [ma][D].  ma = 0: first batch (kappa_a = I, invD_a = I,
 * prev_eta1 = 0, Ktilde_a = 0: training/states.jl:86-98, latentgp.jl:220-223).  Call after agp_refresh_K; every later step of
 * this model adds the two terms (stochastic models are refused, like the reference's constructor). */
int agp_online_carry(agp_model* model, int32_t latent_local, const double* Za, int32_t ma, const double* invDa, const double* prev_eta1,
                     double prev_L);
/* extraKL(model::OnlineSVGP, state) (functions/KLdivergences.jl:37-54) of the owned latents: ELBO = agp_elbo parts - this. */
int agp_online_extra_kl(agp_model* model, double* out);
/* local_updates! and the expectation gradients of the minibatch whose moments agp_step_moments_async just computed, WITHOUT the
 * natural gradient (first iteration on a new batch, onlinetraining.jl:81-92: the local updates run under the previous model);
 * read the results with agp_get_local("grad_mu" / "grad_Sigma"). */
int agp_local_updates_async(agp_model* model);
/* kernel matrices of the batch, natural_gradient! and global_update! with SUPPLIED expectation gradients grad_mu / grad_Sigma
 * (host, [n_latent][B]) instead of this model's own local updates (onlinetraining.jl:93-104), rho = 1. */
int agp_step_with_gradients(agp_model* model, const int64_t* idx, int32_t B, int32_t base, const double* grad_mu, const double* grad_Sigma);

/* EXPERIMENTAL -- NOT on the product path and not yet run on a GPU (written after the round's GPU budget was spent).
 * Building block of the planned replacement of the fp64 Cholesky tail of global_update! (inference/inference.jl:25-28,
 * Sigma = -1/2 eta2^-1): Newton-Schulz refinement  Y <- Y + Y (I - P Y)  of an approximate inverse Y of the SPD m x m
 * matrix P, `iters` times, as 3xTF32 tcgen05 products (m % 128 == 0).  P, Y: row-major fp64 host [m][m], Y is updated in
 * place (through an fp32 round trip).  mode bit 0: residual I - Y P in fp64 on DMMA instead of 3xTF32 (an fp32-class
 * residual floors at eps_fp32 * cond(P)); bit 1: symmetrise Y at the end (the iteration doubles the antisymmetric
 * rounding residue every pass).  resid[i] (may be NULL) = |I - Y P|_F at the start of iteration i; *ms (may be
 * NULL) = device time of the whole refinement.  Self-contained: allocates and frees its own buffers, touches no model.
 * Feasibility numbers: profiles/r1/studies/. */
int agp_experimental_ns_refine(agp_ctx* ctx, int32_t m, const double* P, double* Y, int32_t iters, int32_t mode, double* resid, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* AGP_B200_H */
