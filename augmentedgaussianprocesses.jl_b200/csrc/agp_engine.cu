// agp_engine.cu -- host orchestration of the CAVI step and the C ABI declared in include/agp_b200.h.
// No CPU fallback: every entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/agp_b200.h"
#include "agp_kernels.cuh"
#include "agp_tail.cuh"
#include "agp_tail2.cuh"
#include "agp_tail3.cuh"
#include "agp_hyper.cuh"
#include "agp_umma.h"

using namespace agp;

struct agp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +     \
                 std::to_string(__LINE__) + ")";                                                   \
      return AGP_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)
#define CKS(expr)                      \
  do {                                 \
    int s__ = (expr);                  \
    if (s__ != AGP_OK) return s__;     \
  } while (0)
#define BAD(msg)                  \
  do {                            \
    ctx->err = (msg);             \
    return AGP_ERR_BAD_ARG;       \
  } while (0)

static inline int64_t rup(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

enum Phase {
  PH_IDX = 0, PH_KMAT, PH_KAPPA, PH_KSIGMA, PH_ROWSTATS, PH_LIK, PH_GRADMU, PH_GRAM, PH_COMBINE, PH_CHOL, PH_TRTRI,
  PH_SIGMA, PH_FINAL, PH_SPLIT, PH_COUNT
};
static const char* kPhaseNames[PH_COUNT] = {"idx_select",  "kmat_knm",   "gemm_v",     "gemm_v_sigma", "rowstats",
                                            "lik_update",  "gemv_grad1", "gemm_gram",  "combine_eta",      "chol_blocked",
                                            "trtri",       "gemm_sigma", "finalize",   "scale_transpose"};

// ---------------------------------------------------------------------------------------------------
struct EngineBase {
  agp_ctx* ctx = nullptr;
  virtual ~EngineBase() {}
  virtual int data_upload(const void* X, int x_dtype, int x_layout, int64_t n, const void* const* y, int y_kind) = 0;
  virtual int minibatches_upload(const int64_t* idx, int64_t n_lists, int B, int base) = 0;
  virtual int refresh_K() = 0;
  virtual int state_reset() = 0;
  virtual int set_kernel(int ql, int kind, double scale, double variance) = 0;
  virtual int step_moments(const int64_t* idx, int B, int base, bool from_batch) = 0;
  virtual int step_update(double rho) = 0;
  virtual int step_full(const int64_t* idx, int B, int base, double rho) = 0;
  virtual int step_batch(const void* xb, int x_dtype, int x_layout, const void* const* yb, int y_kind, int B, double rho) = 0;
  virtual int step_batch_async(const void* xb, int x_dtype, int x_layout, const void* const* yb, int y_kind, int B, double rho, int64_t* ticket) = 0;
  virtual int result_wait(int64_t ticket, double* mu) = 0;
  virtual int sync_status() = 0;
  virtual void* moments_ptr(int which, int64_t* ld) = 0;
  virtual int elbo_moments() = 0;
  virtual int elbo(double rho, double* out3) = 0;
  virtual int get_posterior(int ql, double* mu, double* Sigma, double* eta1, double* eta2) = 0;
  virtual int set_posterior(int ql, const double* eta1, const double* eta2) = 0;
  virtual int get_counters(int64_t* t, int64_t* cur) = 0;
  virtual int set_counters(int64_t t, int64_t cur) = 0;
  virtual int get_local(const char* name, int row, double* out, int B) = 0;
  virtual int get_kernel_matrices(int ql, double* Knm, double* kappa, int B) = 0;
  virtual int get_Kinv(int ql, double* Kinv, double* logdetK) = 0;
  virtual int predict_f(const void* Xt, int x_dtype, int x_layout, int64_t nt, int want_var, double* mu, double* var) = 0;
  virtual int proba_logistic(const double* mu, const double* var, int64_t n, const double* nodes, const double* w, int nn,
                             double* p, double* pv) = 0;
  virtual int set_quadrature(const double* nodes, const double* w, int nn) = 0;
  virtual int get_lik_param(int task, double* v) = 0;
  virtual int set_lik_param(int task, double v) = 0;
  virtual int proba_link(int link, double p0, const double* mu, const double* var, int64_t n, double* p, double* pv) = 0;
  virtual int set_step_size(double eta) = 0;
  virtual int set_noise_optimiser(int task, int kind, double eta, double b1, double b2, double eps) = 0;
  virtual int hyper_grads(double rho, double* d_scale, double* d_variance, double* dZ) = 0;
  virtual int set_Z(int ql, const double* Z) = 0;
  virtual int keep_stale_K(int on) = 0;
  virtual int set_A_optimiser(int kind, double eta, double b1, double b2, double eps) = 0;
  virtual int get_A(double* A) = 0;
  virtual int peer_export(void* handle64) = 0;
  virtual int peer_attach(int world, int rank, const void* handles) = 0;
  virtual int peer_detach() = 0;
  virtual int profile_enable(int on) = 0;
  virtual int profile_read(int maxp, const char** names, double* ms, int64_t* launches) = 0;
  virtual int64_t launch_count() = 0;
  virtual int time_kernel(int which, int reps, double* ms) = 0;
  virtual int use_graph(int on) = 0;
  virtual int online_carry(int ql, const double* Za, int ma, const double* invDa, const double* prev_eta1, double prev_L) = 0;
  virtual int online_extra_kl(double* out) = 0;
  virtual int predict_f_cov(const double* Xt, int64_t nt, double* mu, double* cov) = 0;
  virtual int local_updates_only() = 0;
  virtual int step_with_gradients(const int64_t* idx, int B, int base, const double* gmu_h, const double* gS_h) = 0;
  virtual void join_async() {}   // order the main stream behind the result stream of the pipelined asynchronous host-batch steps
};

struct agp_model {
  EngineBase* eng = nullptr;
};

// ---------------------------------------------------------------------------------------------------
template <typename T>
struct Engine : EngineBase {
  // ---- configuration ----
  int model_kind, Qg, qbeg, Ql, m, D, Dp, mp, Bcap, prec, stochastic, nT;
  int mk = 0;   // columns of the fp32 B x m operands: m, or m rounded up to the 128-wide tensor-core tile (tf32x3; the padding columns of
                // L^-1 and X are zero, so whatever the padding inducing points (all-zero rows of Z) put into K_nm never reaches V)
  int64_t ldm, ldB;
  double rm_kappa, rm_tau, jitter;
  std::vector<int> h_lik_kind;
  std::vector<double> h_p0, h_p1, h_A;
  bool is_vgp = false;   // AGP_MODEL_VGP: full variational GP = the SVGP algebra with Z = X, kappa = I, Ktilde = 0
  bool vgp_identity = false;   // set around the step / ELBO moment passes (prediction uses the ordinary sparse formulas)
  bool is_lsm = false;   // class-index labels (LogisticSoftMax)
  bool is_het = false;   // HeteroscedasticLikelihood: 2 latents (f, g), one real-valued target
  bool need_lam = false; // some likelihood re-estimates its link parameter lambda in local_updates! (Poisson, Heteroscedastic)
  bool need_quad = false;
  int R = 1;  // rows of the local-variable arrays
  double *d_lam = nullptr, *d_lamacc = nullptr, *d_qnodes = nullptr, *d_qw = nullptr; int nq = 0;

  // One latent GP.  The device works in the WHITENED basis u = L v (K = L L^T):
  //   V = Knm L^-T,  mu = L mu_v,  Sigma = L Sigma_v L^T,  eta1_v = L^T eta1,  eta2_v = L^T eta2 L.
  // It is the same iteration as the reference's (every map is linear), but the prior precision becomes I
  // and P_v = -2 eta2_v >= lr*I, so fp32 / TF32 contractions lose sqrt(cond K) digits instead of cond K.
  // The canonical (mu, Sigma, eta1, eta2) are produced on request (get_posterior) in fp64.
  struct Latent {
    int kind; double scale, variance;
    std::vector<double> hZ, hmu0;
    bool has_mu0 = false;
    T* Z = nullptr; T* zz = nullptr;              // [m][Dp], [m]
    double* Zd = nullptr; double* zzd = nullptr;  // fp64 copies for K_mm
    double *Lc = nullptr, *Linv = nullptr, *Kinv = nullptr;  // chol(K) lower, its inverse, K^-1   [mp][mp]
    double *mu0 = nullptr, *mu0v = nullptr;                  // prior mean at Z, L^-1 mu0
    T* Linv_T = nullptr;
    double logdetK = 0.0;
    double *eta1c = nullptr, *eta2c = nullptr;               // canonical natural parameters (valid when !white_valid or after sync)
    double *eta1v = nullptr, *eta2v = nullptr, *muv = nullptr;  // whitened state
    double* tvec = nullptr; bool muv_valid = false;             // t = X eta1_v ; mu_v = X^T t computed lazily
    double *Xv = nullptr, *Dinv = nullptr;  // X = chol(P_v)^-1 (lower triangular; Sigma_v = X^T X), its diagonal blocks
    bool white_valid = false;                                // whitened state is the live one
    T* Xv_T = nullptr;
    T *Knm = nullptr, *V = nullptr, *VS = nullptr;           // [Bcap][ldm]
    double* Ktilde = nullptr;
    double* racc = nullptr;                                  // [3][ldB] fused row-statistic accumulators (tensor-core path)
    T* Gpart = nullptr;
    double* v1 = nullptr;
    double *P = nullptr, *X = nullptr, *W = nullptr;  // tail workspaces [mp][mp]
    double* logdetP = nullptr;                        // device scalars: [0] logdet P_v, [1] scratch for K
    UmmaLatent um;                                    // tcgen05 path
    UmmaKnm uk; bool knm_tc = false;                  // tcgen05 K_nm construction (16 < D <= 128)
    int gram_splits = 1;                              // split-K partials of the last Gram product
    // OnlineSVGP carry-over (agp_online_carry): whitened offsets of the natural gradient + what extraKL needs
    bool online = false; int on_ma = 0;
    double *on_c1v = nullptr, *on_C2v = nullptr, *on_Va = nullptr;    // [mp], [mp][mp], [ma_p][mp]  (V_a = kappa_a L = K_ab L^-T)
    std::vector<double> on_invD, on_preveta1; double on_prevL = 0.0, on_trDK = 0.0;
    // experimental Newton-Schulz tail (AGP_TAIL_NS): ns.Y() = Sigma_v (fp32, full symmetric) refined from step to step
    UmmaNs ns; bool ns_alloc = false;
    bool ns_seeded = false;      // ns.Y() is the covariance of the current eta2_v
    bool factor_valid = true;    // Xv / Xv_T / tvec describe the current natural parameters
    // quirk Q3 (agp_keep_stale_K): the factor of the last agp_refresh_K, parked while agp_hyper_grads works with a fresh one
    double *bkLc = nullptr, *bkLinv = nullptr, *bkKinv = nullptr, *bkmu0v = nullptr; T* bkLinvT = nullptr; double bklogdetK = 0.0;
  };
  std::vector<Latent> lat;

  // ---- data ----
  int64_t n = 0;
  T* X = nullptr; T* xx = nullptr;
  double* y_all = nullptr; int* ycls_all = nullptr;
  T* Xb = nullptr; T* xxb = nullptr;  // host-batch path
  T *pKS = nullptr, *pXb = nullptr, *pxxb = nullptr;  // scratch: kappa getter, prediction input rows
  bool kernel_matrices_stale = false;  // predict_f reused Knm / V / VS: recompute them before the next ELBO
  void* stage = nullptr; size_t stage_bytes = 0;
  int64_t* idx_pool = nullptr; int64_t n_lists = 0; int pool_B = 0;
  int64_t* idx_cur = nullptr;
  T* xx_cur = nullptr;   // squared norms of the current minibatch rows
  int64_t* counters = nullptr;  // [0] RM t (starts 1), [1] cursor
  int* status = nullptr;
  int *d_lik_kind = nullptr; double *d_p0 = nullptr, *d_p1 = nullptr, *d_A = nullptr;
  double *mean_f = nullptr, *var_f = nullptr, *gmu = nullptr, *gS = nullptr;
  double *lc = nullptr, *ltheta = nullptr, *lgamma_ = nullptr, *lalpha = nullptr;
  double *tmu = nullptr, *tvar = nullptr, *gm = nullptr, *gs = nullptr, *yb = nullptr;
  int* ycls = nullptr;
  double* d_out = nullptr;  // [8] scratch scalars
  int n_split = 1, k_chunk = 0;
  bool tail_pdl = true;  // AGP_TAIL_PDL=0 disables programmatic dependent launch along the per-step kernel chain
  // latent-sharded peer exchange (agp_peer_export / agp_peer_attach): one exported block [mean 2Q ldB | var 2Q ldB | flags]
  double* xchg = nullptr; int64_t par_stride = 0; bool peer = false; int peer_world = 1, peer_rank = 0;
  int64_t* d_xepoch = nullptr; double** d_peers = nullptr; std::vector<void*> peer_opened;
  // GaussianLikelihood(opt_noise): ADAM on log sigma^2 inside local_updates! (gaussian.jl:56-72)
  bool noise_any = false; std::vector<int> h_noise_opt; int* d_noise_opt = nullptr; double* d_noise_state = nullptr;
  double n_eta = 0.05, n_b1 = 0.9, n_b2 = 0.999, n_eps = 1e-8;
  // update_A! (MOSVGP Aoptimiser): ADAM state on the device
  bool a_opt = false; double a_eta = 0.01, a_b1 = 0.9, a_b2 = 0.999, a_eps = 1e-8;
  double *d_gradA = nullptr, *d_Amt = nullptr, *d_Avt = nullptr, *d_Abt = nullptr;
  double* d_lr = nullptr;          // Robbins-Monro step size of the current iteration (lik_update_kernel -> combine_kernel)
  bool fuse_lik_next = false, fuse_from_batch = false, lik_fused = false;   // rowfinish + local-update fusion (set by the step paths)
  bool stats_use_early = false;    // moments stage 2 of this step only adds the last N tile (set by step_pool)
  bool racc2_precleared = false;   // the V X^T row-statistic accumulators were cleared off the critical chain (side stream)
  bool rowfin_on = false, rowfin_now = false;  // AGP_ROWFIN=1: row finish + local update inside the statistics product's epilogue (UmmaRowFinish) instead of
                                               // rowfinish_lik_kernel: device-resident C2 +0.8 % (5.44 k vs 5.39 k it/s) but end to end -2.5 % -> opt-in
  unsigned* d_rowcnt = nullptr;               // per-sample arrival counters of the fused row finish (self-resetting)
  bool combine4_on = false;  // AGP_COMBINE4=1: four columns per thread in the natural-parameter update (measured SLOWER: 5.40 k vs 5.51 k it/s at C2 - fewer loads in flight)
  int chain_variant = 0; // AGP_CHAIN: pivot-chain implementation of the multi-launch tail (agp_tail2.cuh chain16_*): 0 = scalar (default), 1 = look-ahead (opt-in: 132 vs 126.7 us at C2)
  int tail_variant = 3;  // AGP_TAIL_VARIANT: 0 = agp_tail.cuh (generation 1, SIMT tile products), 2 = agp_tail2.cuh (DMMA, panel potf2, one
                         // launch per block step and latent), 3 (default) = agp_tail3.cuh (one persistent launch for all owned latents)
  Tail3Lat* d_t3lat = nullptr; u64* d_t3flags = nullptr;
  int t3_sm_budget = 148;   // AGP_TAIL3_SMS: CTAs (= SMs) the persistent tail may occupy

  // ---- step state ----
  int curB = 0; bool cur_from_batch = false; bool have_K = false; bool have_data = false; bool have_step = false;
  int64_t launches = 0;
  // profiling
  bool prof = false;
  struct Ev { cudaEvent_t a, b; int ph; };
  std::vector<Ev> pending;
  std::vector<cudaEvent_t> ev_pool;
  double ph_ms[PH_COUNT] = {0}; int64_t ph_launch[PH_COUNT] = {0};
  int cur_phase = -1; cudaEvent_t cur_ev = nullptr; int64_t cur_phase_l0 = 0;
  // graph
  bool want_graph = false; cudaGraphExec_t gexec = nullptr; int gB = -1; double grho = -1; int64_t g_launches = 0;
  bool capturing = false;
  // host-batch steps also leave the canonical posterior mean of latent 0 in a pinned host buffer (computed and copied inside
  // the step's graph), so the per-step read-back of the end-to-end loop is a host memcpy after the step's one synchronisation
  double* h_mu = nullptr; bool h_mu_valid = false;
  cudaGraphExec_t gexec_b = nullptr; int gB_b = -1, gkey_b = -1; double grho_b = -1; int64_t g_launches_b = 0;   // host-batch step graph

  // launch stream: the context's stream, or the side stream while the next minibatch is being prefetched
  cudaStream_t cur_stream = nullptr;
  cudaStream_t st() const { return cur_stream ? cur_stream : ctx->stream; }
  // software pipeline (resident-list steps): while the fp64 tail of step t runs on the main stream, the side stream
  // builds Knm and V = Knm L^-T of minibatch t+1 (V only depends on the fixed L^-1, not on the posterior)
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool pipeline = true;      // AGP_PIPELINE=0 disables
  bool prefetched = false;   // Knm / V / sum V^2 / idx_cur / xx_cur hold the minibatch at the device cursor
  // early statistics (single latent, tf32x3, multi-launch tail): row block j of X = chol(P_v)^-1 is final after block step 2j+1 of the
  // tail, so the side stream runs N tile j of the NEXT step's V X^T statistics (and that block's share of x_finalize) while the tail is
  // still factorising the later blocks; only the last N tile (and the last row block's finalize) stay on the next step's chain
  // split Gram (single latent, tf32x3, multi-launch tail; opt-in AGP_SPLIT_GRAM=1 -- measured SLOWER, 196 vs 189 us per C2 step: the
  // cross-stream join costs the programmatic edges of the chain): the tail's first kernel only needs tile (0, 0) of
  // P_v, so that tile's Gram slices + its share of combine_kernel + the first 64 x 64 factorisation run on the main stream while the
  // other nine upper tiles and their combine run beside them on the side stream; the block steps wait for both
  bool split_gram_on = false, split_gram_now = false, allow_split_gram = false;
  int split_SA = 0, split_SB = 0;
  cudaEvent_t ev_g0 = nullptr, ev_gb = nullptr;
  bool stats_early = false;     // racc[1..2] already hold the N tiles 0 .. ntn-2 of the prefetched minibatch's statistics
  bool early_on = false;        // AGP_EARLY_STATS=1 enables (measured: no gain)
  bool early_now = false;       // this step issues the early tiles (set by step_pool around step_update_b)
  cudaEvent_t ev_blk[16] = {nullptr};
  int64_t* idx_prev = nullptr;  // indices of the minibatch the last step consumed (ELBO / getters after a prefetch)

  template <typename U>
  int dalloc(U** p, size_t count) {
    CK(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(U)));
    CK(cudaMemsetAsync(*p, 0, std::max<size_t>(count, 1) * sizeof(U), st()));
    return AGP_OK;
  }

  // ---- phase timers -------------------------------------------------------------------------------
  cudaEvent_t get_ev() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void ph_begin(int ph) {
    cur_phase = ph; cur_phase_l0 = launches;
    if (prof && !capturing) { cur_ev = get_ev(); cudaEventRecord(cur_ev, st()); }
  }
  void ph_end() {
    if (cur_phase < 0) return;
    ph_launch[cur_phase] += launches - cur_phase_l0;
    if (prof && !capturing) {
      cudaEvent_t b = get_ev(); cudaEventRecord(b, st());
      pending.push_back({cur_ev, b, cur_phase});
    }
    cur_phase = -1;
  }
  void resolve_pending() {
    for (auto& e : pending) {
      float ms = 0.f;
      if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) ph_ms[e.ph] += ms;
      ev_pool.push_back(e.a); ev_pool.push_back(e.b);
    }
    pending.clear();
  }

  // ---- construction ---------------------------------------------------------------------------------
  int init(agp_ctx* c, const agp_model_desc* d) {
    ctx = c;
    model_kind = d->model_kind;
    if (model_kind == AGP_MODEL_VGP) { is_vgp = true; model_kind = AGP_MODEL_SVGP; }
    if (model_kind == AGP_MODEL_MOVGP) { is_vgp = true; model_kind = AGP_MODEL_MOSVGP; }
    Qg = d->n_latent_global; qbeg = d->latent_begin; Ql = d->n_latent_local;
    m = d->m; D = d->D; Bcap = d->batch_capacity; prec = d->precision; stochastic = d->stochastic;
    rm_kappa = d->rm_kappa; rm_tau = d->rm_tau; jitter = d->jitter; nT = d->n_task;
    if (Qg < 1 || Ql < 1 || qbeg < 0 || qbeg + Ql > Qg || m < 1 || D < 1 || Bcap < 1 || nT < 1) BAD("bad model sizes");
    if (!d->lik_kind || !d->kernel_kind || !d->kernel_scale || !d->kernel_variance || !d->Z) BAD("null descriptor array");
    if (is_vgp && stochastic) BAD("VGP is a full-batch model: use AnalyticVI (models/VGP.jl)");
    if (stochastic && !(rm_kappa > 0.5 && rm_kappa <= 1.0 && rm_tau > 0)) BAD("kappa should be in the interval (0.5,1], tau positive");
    mk = prec == AGP_PREC_TF32X3 ? (int)rup(m, 128) : m;
    Dp = (int)rup(D, 4); ldm = rup(mk, 4); ldB = rup(Bcap, 4);
    int nblk = (m + POTF2_NB - 1) / POTF2_NB, pw = 1;
    while (pw < nblk) pw *= 2;
    mp = pw * POTF2_NB;
    h_lik_kind.assign(d->lik_kind, d->lik_kind + nT);
    h_p0.assign(nT, 0.0); h_p1.assign(nT, 0.0);
    for (int t = 0; t < nT; ++t) { if (d->lik_p0) h_p0[t] = d->lik_p0[t]; if (d->lik_p1) h_p1[t] = d->lik_p1[t]; }
    for (int t = 0; t < nT; ++t) {
      int k = h_lik_kind[t];
      if (k < 0 || k > 8) BAD("unknown likelihood kind");
      if (k == AGP_LIK_LAPLACE && !(h_p0[t] > 0)) BAD("Laplace scale beta must be positive");
      if (k == AGP_LIK_NEGBINOMIAL && !(h_p0[t] > 0)) BAD("r must be positive");
      if ((k == AGP_LIK_POISSON || k == AGP_LIK_HETEROSCEDASTIC) && !(h_p0[t] > 0)) BAD("lambda must be positive");
      if (k == AGP_LIK_POISSON || k == AGP_LIK_HETEROSCEDASTIC) need_lam = true;
      if (k == AGP_LIK_POISSON) need_quad = true;
      if (k == AGP_LIK_GAUSSIAN && !(h_p0[t] > 0)) BAD("Gaussian noise variance must be positive");
      if (k == AGP_LIK_STUDENTT && !(h_p0[t] > 0.5)) BAD("nu should be greater than 0.5");
    }
    if (model_kind == AGP_MODEL_SVGP) {
      if (nT != 1) BAD("SVGP takes exactly one likelihood");
      is_lsm = h_lik_kind[0] == AGP_LIK_LOGISTICSOFTMAX;
      is_het = h_lik_kind[0] == AGP_LIK_HETEROSCEDASTIC;
      if (is_het && Qg != 2) BAD("HeteroscedasticLikelihood needs exactly two latent GPs");
      if (!is_lsm && !is_het && Qg != 1) BAD("single-latent likelihood needs exactly one latent GP");
      if (is_lsm && Qg < 2) BAD("LogisticSoftMax needs at least 2 classes");
      R = Qg;
    } else if (model_kind == AGP_MODEL_MOSVGP) {
      if (!d->A) BAD("MOSVGP needs the mixing matrix A");
      for (int t = 0; t < nT; ++t)
        if (h_lik_kind[t] == AGP_LIK_LOGISTICSOFTMAX || h_lik_kind[t] == AGP_LIK_HETEROSCEDASTIC) BAD("MOSVGP tasks must be single-latent likelihoods");
      h_A.assign(d->A, d->A + (size_t)nT * Qg);
      R = nT;
    } else BAD("unknown model kind");
    if (is_vgp && Ql != Qg) BAD("VGP latents are not sharded");
    if (prec < 0 || prec > 2) BAD("unknown precision");
    if (prec == AGP_PREC_TF32X3 && (m <= 64 || !umma_shape_ok(mk, Bcap))) BAD("TF32X3 precision needs m > 64 and batch_capacity % 128 == 0 (m itself is padded to a multiple of 128 internally)");

    CK(cudaSetDevice(ctx->device));
    // split-K of the Gram product kappa^T diag(w) kappa: enough CTAs to fill the chip
    int bt = gemm_tile<T>();
    int tiles = (int)(((m + bt - 1) / bt) * ((m + bt - 1) / bt));
    int want = std::max(1, (2 * 148 + tiles - 1) / tiles);
    k_chunk = (int)rup((Bcap + want - 1) / want, 64);
    n_split = (Bcap + k_chunk - 1) / k_chunk;
    if (prec == AGP_PREC_TF32X3 && Ql == 1) n_split = std::max(n_split, 32);   // split Gram: tile (0, 0) alone over up to 32 slices

    lat.resize(Ql);
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      L.kind = d->kernel_kind[q]; L.scale = d->kernel_scale[q]; L.variance = d->kernel_variance[q];
      if (L.kind < 0 || L.kind > 2 || !(L.scale > 0) || !(L.variance > 0)) BAD("bad kernel parameters");
      L.hZ.assign(d->Z + (size_t)q * m * D, d->Z + (size_t)(q + 1) * m * D);
      if (d->mu0) { L.hmu0.assign(d->mu0 + (size_t)q * m, d->mu0 + (size_t)(q + 1) * m); L.has_mu0 = true; }
      CKS(dalloc(&L.Z, (size_t)mk * Dp)); CKS(dalloc(&L.zz, mk));
      CKS(dalloc(&L.Zd, (size_t)m * Dp)); CKS(dalloc(&L.zzd, m));
      CKS(dalloc(&L.Lc, (size_t)mp * mp)); CKS(dalloc(&L.Linv, (size_t)mp * mp)); CKS(dalloc(&L.Kinv, (size_t)mp * mp));
      CKS(dalloc(&L.mu0, 2 * (size_t)mp)); CKS(dalloc(&L.mu0v, mp));   // mu0: [0, mp) prior mean, [mp, 2mp) scratch for the canonical mean
      CKS(dalloc(&L.Linv_T, (size_t)mk * ldm));
      CKS(dalloc(&L.eta1c, mp)); CKS(dalloc(&L.eta2c, (size_t)mp * mp));
      CKS(dalloc(&L.eta1v, mp)); CKS(dalloc(&L.eta2v, (size_t)mp * mp)); CKS(dalloc(&L.muv, mp)); CKS(dalloc(&L.tvec, mp));
      CKS(dalloc(&L.Xv, (size_t)mp * mp)); CKS(dalloc(&L.Dinv, (size_t)mp * POTF2_NB));
      CKS(dalloc(&L.Xv_T, (size_t)mk * ldm));
      CKS(dalloc(&L.Knm, (size_t)Bcap * ldm)); CKS(dalloc(&L.V, (size_t)Bcap * ldm)); CKS(dalloc(&L.VS, (size_t)Bcap * ldm));
      CKS(dalloc(&L.Ktilde, ldB)); CKS(dalloc(&L.racc, 3 * ldB));
      CKS(dalloc(&L.Gpart, (size_t)n_split * mk * ldm));
      CKS(dalloc(&L.v1, mp));
      CKS(dalloc(&L.P, (size_t)mp * mp)); CKS(dalloc(&L.X, (size_t)mp * mp)); CKS(dalloc(&L.W, (size_t)mp * mp));
      CKS(dalloc(&L.logdetP, 2));
      // posterior init (gpblocks/posterior.jl:29-37): mu = 0, Sigma = I, eta1 = 0, eta2 = -I/2 (canonical; whitened at refresh_K)
      set_identity_kernel<<<dim3((mp + 127) / 128, mp), 128, 0, st()>>>(L.eta2c, mp, mp, -0.5);
      ++launches;
      // inducing points
      std::vector<double> zp((size_t)m * Dp, 0.0), zn(m, 0.0);
      std::vector<T> zt((size_t)m * Dp, T(0)), znt(m, T(0));
      for (int i = 0; i < m; ++i) {
        double s = 0, s_t = 0;
        for (int k = 0; k < D; ++k) {
          double v = L.hZ[(size_t)i * D + k];
          zp[(size_t)i * Dp + k] = v; zt[(size_t)i * Dp + k] = (T)v;
          s += v * v; double vt = (double)(T)v; s_t += vt * vt;
        }
        zn[i] = s; znt[i] = (T)s_t;
      }
      CK(cudaMemcpyAsync(L.Zd, zp.data(), zp.size() * sizeof(double), cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(L.zzd, zn.data(), zn.size() * sizeof(double), cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(L.Z, zt.data(), zt.size() * sizeof(T), cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(L.zz, znt.data(), znt.size() * sizeof(T), cudaMemcpyHostToDevice, st()));
      if (L.has_mu0) CK(cudaMemcpyAsync(L.mu0, L.hmu0.data(), m * sizeof(double), cudaMemcpyHostToDevice, st()));
      CK(cudaStreamSynchronize(st()));
      if (prec == AGP_PREC_TF32X3)
        CKS(umma_latent_alloc(ctx_err(), L.um, mk, (int)ldm, Bcap, (const float*)(const void*)L.Knm, (const float*)(const void*)L.V,
                              (const float*)(const void*)L.Linv_T, (const float*)(const void*)L.Xv_T, st()));
      // D <= 16: K_nm from the coordinate differences on the CUDA cores (one 16-deep k-block of gemm_simt_kernel, a few us) instead of
      // |x|^2 + |z|^2 - 2 x.z on the tensor cores: low-dimensional inputs are where K_mm is ill conditioned, and V = K_nm L^-T amplifies
      // the absolute error of the product form by sqrt(cond K_mm) (tools/shape_sweep.py: D = 1..3, m ~ 500, cond 1e6).
      if (prec == AGP_PREC_TF32X3 && umma_knm_shape_ok(mk, Bcap, D) && (D > 16 || getenv("AGP_KNM_TC_SMALL_D")) && !getenv("AGP_KNM_SIMT")) {
        CKS(umma_knm_setup(ctx_err(), L.uk, (const float*)(const void*)L.Z, Dp, mk, D, (float*)(void*)L.Knm, ldm, Bcap, st()));
        L.knm_tc = true;
      }
    }
    CKS(groups_init());
    CKS(fan_init());
    CKS(dalloc(&Xb, (size_t)Bcap * Dp)); CKS(dalloc(&xxb, Bcap));
    CKS(dalloc(&idx_cur, Bcap)); CKS(dalloc(&xx_cur, Bcap)); CKS(dalloc(&idx_prev, Bcap));
    CK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    { const char* e = getenv("AGP_PIPELINE"); if (e && e[0] == '0') pipeline = false; }
    { const char* e = getenv("AGP_EARLY_STATS"); if (e && e[0] == '1') early_on = true; }
    { const char* e = getenv("AGP_SPLIT_GRAM"); if (e && e[0] == '1') split_gram_on = true; }
    CK(cudaEventCreateWithFlags(&ev_g0, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ev_gb, cudaEventDisableTiming));
    for (int j = 0; j < 16; ++j) CK(cudaEventCreateWithFlags(&ev_blk[j], cudaEventDisableTiming));
    CKS(dalloc(&counters, 2)); CKS(dalloc(&status, 1));
    int64_t c0[2] = {1, 0};
    CK(cudaMemcpyAsync(counters, c0, sizeof(c0), cudaMemcpyHostToDevice, st()));
    CKS(dalloc(&d_lik_kind, nT)); CKS(dalloc(&d_p0, nT)); CKS(dalloc(&d_p1, nT)); CKS(dalloc(&d_A, (size_t)nT * Qg));
    CK(cudaMemcpyAsync(d_lik_kind, h_lik_kind.data(), nT * sizeof(int), cudaMemcpyHostToDevice, st()));
    CK(cudaMemcpyAsync(d_p0, h_p0.data(), nT * sizeof(double), cudaMemcpyHostToDevice, st()));
    CK(cudaMemcpyAsync(d_p1, h_p1.data(), nT * sizeof(double), cudaMemcpyHostToDevice, st()));
    if (!h_A.empty()) CK(cudaMemcpyAsync(d_A, h_A.data(), h_A.size() * sizeof(double), cudaMemcpyHostToDevice, st()));
    par_stride = (int64_t)Qg * ldB;
    CKS(dalloc(&xchg, 4 * (size_t)par_stride + 64));
    mean_f = xchg; var_f = xchg + 2 * par_stride;   // parity 0 halves; parity 1 (peer mode only) follows each at + par_stride
    CKS(dalloc(&d_xepoch, 1));
    CKS(dalloc(&gmu, (size_t)Ql * ldB)); CKS(dalloc(&gS, (size_t)Ql * ldB));
    CKS(gram_tn_group_init());
    CKS(dalloc(&lc, (size_t)R * ldB)); CKS(dalloc(&ltheta, (size_t)R * ldB)); CKS(dalloc(&lgamma_, (size_t)R * ldB));
    CKS(dalloc(&lalpha, ldB));
    CKS(dalloc(&tmu, (size_t)nT * ldB)); CKS(dalloc(&tvar, (size_t)nT * ldB));
    CKS(dalloc(&gm, (size_t)nT * ldB)); CKS(dalloc(&gs, (size_t)nT * ldB)); CKS(dalloc(&yb, (size_t)nT * ldB));
    CKS(dalloc(&ycls, ldB));
    CKS(dalloc(&d_out, 8));
    CKS(dalloc(&d_lr, 1));
    CKS(dalloc(&d_tile0_flag, 1));
    CK(cudaMemsetAsync(d_tile0_flag, 0, sizeof(int), st()));
    { const char* e = getenv("AGP_EARLY_POTF2"); if (e && e[0] == '0') early_potf2 = false; }
    CK(cudaStreamSynchronize(st()));
    CKS(upload_lr(1));
    CKS(dalloc(&d_lam, nT)); CKS(dalloc(&d_lamacc, 2 * (size_t)nT)); CKS(dalloc(&d_qnodes, 128)); CKS(dalloc(&d_qw, 128));
    CK(cudaMemcpyAsync(d_lam, h_p0.data(), nT * sizeof(double), cudaMemcpyHostToDevice, st()));
    CKS(reset_local_vars());
    CK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, potf2_smem()));
    CK(cudaFuncSetAttribute(tail_potf2_first_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM));
    CK(cudaFuncSetAttribute(tail_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM));
    CK(cudaFuncSetAttribute(tail2_potf2_first_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM));
    CK(cudaFuncSetAttribute(tail2_potf2_first_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM));
    CK(cudaFuncSetAttribute(tail2_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM));
    CK(cudaFuncSetAttribute(tail2_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM));
    { const char* e = getenv("AGP_TAIL_PDL"); if (e && e[0] == '0') tail_pdl = false; }
    { const char* e = getenv("AGP_TAIL_NS"); if (e) ns_iters = std::max(0, std::min(8, atoi(e))); }
    { const char* e = getenv("AGP_TAIL_NS_AFTER"); if (e) ns_after = std::max(1, atoi(e)); }
    { const char* e = getenv("AGP_TAIL_NS_TOL"); if (e && atof(e) > 0.0) ns_tol = atof(e); }
    if (ns_iters > 0 && ns_eligible()) {
      CKS(umma_ns_alloc(ctx_err(), lat[0].ns, m, st()));
      lat[0].ns_alloc = true;
    } else ns_iters = 0;
    // default: the persistent tail for two or more owned latents (one launch, the chains run side by side), the multi-launch tail2
    // chain for a single latent (126 vs 138 us at m = 512: see agp_tail3.cuh)
    tail_variant = Ql >= 2 ? 3 : 2;
    { const char* e = getenv("AGP_CHAIN"); if (e) chain_variant = atoi(e) ? 1 : 0; }
    { const char* e = getenv("AGP_COMBINE4"); if (e) combine4_on = e[0] != '0'; }
    { const char* e = getenv("AGP_ROWFIN"); if (e) rowfin_on = e[0] != '0'; }
    if (prec == AGP_PREC_TF32X3) CKS(dalloc(&d_rowcnt, (size_t)ldB + 128));
    { const char* e = getenv("AGP_TAIL_VARIANT"); if (e) { tail_variant = atoi(e); if (tail_variant != 0 && tail_variant != 2) tail_variant = 3; } }
    { const char* e = getenv("AGP_TAIL3_SMS"); if (e && atoi(e) >= 2) t3_sm_budget = std::min(148, atoi(e)); }
    {   // per-latent descriptors + dependency words of the persistent tail (zero-initialised: epoch 0)
      const int nblk = mp / TNB;
      const size_t fw = tail3_flag_words(nblk);
      CKS(dalloc(&d_t3flags, fw * Ql)); CKS(dalloc(&d_t3lat, Ql));
      std::vector<Tail3Lat> h(Ql);
      for (int q = 0; q < Ql; ++q) {
        Latent& L = lat[q];
        h[q].P = L.P; h[q].W = L.W; h[q].Xout = L.Xv; h[q].Dinv = L.Dinv; h[q].logdet = L.logdetP; h[q].flags = d_t3flags + fw * q;
      }
      CK(cudaMemcpyAsync(d_t3lat, h.data(), Ql * sizeof(Tail3Lat), cudaMemcpyHostToDevice, st()));
      CK(cudaStreamSynchronize(st()));
      CK(cudaFuncSetAttribute(tail3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL3_SMEM));
    }
    CK(cudaStreamSynchronize(st()));
    return AGP_OK;
  }
  std::string* ctx_err() { return &ctx->err; }
  static int potf2_smem() { return 2 * POTF2_NB * (POTF2_NB + 1) * (int)sizeof(double); }

  // init_local_vars (likelihood/logisticsoftmax.jl:43-53): alpha = K (beta = K is a constant)
  int reset_local_vars() {
    if (is_lsm) {
      std::vector<double> a(ldB, (double)Qg);
      CK(cudaMemcpyAsync(lalpha, a.data(), ldB * sizeof(double), cudaMemcpyHostToDevice, st()));
      CK(cudaStreamSynchronize(st()));
    }
    return AGP_OK;
  }

  ~Engine() override {
    resolve_pending();
    for (auto e : ev_pool) cudaEventDestroy(e);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (gexec_b) cudaGraphExecDestroy(gexec_b);
    if (h_status) cudaFreeHost(h_status);
    if (h_mu) cudaFreeHost(h_mu);
    for (auto& L : lat) {
      void* ps[] = {L.Z, L.zz, L.Zd, L.zzd, L.Lc, L.Linv, L.Kinv, L.mu0, L.mu0v, L.Linv_T, L.eta1c, L.eta2c, L.eta1v, L.eta2v,
                    L.muv, L.tvec, L.Xv, L.Dinv, L.Xv_T, L.Knm, L.V, L.VS, L.Ktilde, L.racc, L.Gpart, L.v1, L.P, L.X, L.W, L.logdetP, L.on_c1v, L.on_C2v, L.on_Va, L.bkLc, L.bkLinv, L.bkKinv, L.bkmu0v, L.bkLinvT};
      for (void* p : ps) cudaFree(p);
      umma_latent_free(L.um);
      if (L.ns_alloc) umma_ns_free(L.ns);
      umma_knm_free(L.uk);
    }
    umma_groups_free(grpV); umma_groups_free(grpS); umma_groups_free(grpG); umma_groups_free(grpGT);
    for (int i = 0; i < NAUX; ++i) { if (aux[i]) cudaStreamDestroy(aux[i]); if (ev_aux_join[i]) cudaEventDestroy(ev_aux_join[i]); }
    if (ev_aux_fork) cudaEventDestroy(ev_aux_fork);
    for (void* q : peer_opened) cudaIpcCloseMemHandle(q);
    if (side) cudaStreamDestroy(side);
    drop_graph_p();
    if (copy_stream) {
      cudaStreamDestroy(copy_stream);
      if (res_stream) cudaStreamDestroy(res_stream);
      if (ev_vfree) cudaEventDestroy(ev_vfree);
      if (ev_p1) cudaEventDestroy(ev_p1);
      if (ev_res) cudaEventDestroy(ev_res);
      if (h_stat) cudaFreeHost(h_stat);
      for (int s = 0; s < 2; ++s) {
        cudaFree(pre_x[s]); cudaFree(pre_y[s]); cudaFree(pre_ycls[s]);
        if (h_res[s]) cudaFreeHost(h_res[s]);
        if (ev_h2d[s]) cudaEventDestroy(ev_h2d[s]);
        if (ev_free[s]) cudaEventDestroy(ev_free[s]);
        if (ev_done[s]) cudaEventDestroy(ev_done[s]);
        if (ev_xfree[s]) cudaEventDestroy(ev_xfree[s]);
        if (ev_step[s]) cudaEventDestroy(ev_step[s]);
      }
    }
    for (int j = 0; j < 16; ++j) if (ev_blk[j]) cudaEventDestroy(ev_blk[j]);
    if (ev_g0) cudaEventDestroy(ev_g0);
    if (ev_gb) cudaEventDestroy(ev_gb);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    void* ps[] = {idx_prev, xx_cur, pKS, pXb, pxxb, X, xx, y_all, ycls_all, Xb, xxb, stage, idx_pool, idx_cur, counters, status, d_lik_kind, d_p0, d_p1, d_A,
                  xchg, d_xepoch, d_peers, gmu, gS, lc, ltheta, lgamma_, lalpha, tmu, tvar, gm, gs, yb, ycls, d_out, d_lam, d_lamacc, d_qnodes, d_qw, d_lr, d_gradA, d_Amt, d_Avt, d_Abt, d_noise_opt, d_noise_state,
                  hgH[0], hgH[1], hgH[2], hgH[3], hgX, hgA, hgB, hgM1, hgM2, hgV, hgP, d_t3lat, d_t3flags, d_rowcnt, d_tile0_flag};
    for (void* p : ps) cudaFree(p);
  }

  int ensure_stage(size_t bytes) {
    if (bytes <= stage_bytes) return AGP_OK;
    if (stage) cudaFree(stage);
    if (gexec_b) { cudaGraphExecDestroy(gexec_b); gexec_b = nullptr; }   // the host-batch graph reads from the staging buffer
    drop_graph_p();
    stage = nullptr; stage_bytes = 0;
    CK(cudaMalloc(&stage, bytes));
    stage_bytes = bytes;
    return AGP_OK;
  }

  // host matrix rows [r0, r0+rows) -> device row-major T (+ squared norms)
  int upload_rows(const void* Xh, int x_dtype, int x_layout, int64_t ntot, int64_t r0, int64_t rows, T* dst, T* dxx) {
    size_t es = x_dtype == AGP_DTYPE_F64 ? 8 : 4;
    CKS(ensure_stage((size_t)rows * D * es));
    const char* src = (const char*)Xh;
    if (x_layout == AGP_LAYOUT_ROWMAJOR) {
      CK(cudaMemcpyAsync(stage, src + (size_t)r0 * D * es, (size_t)rows * D * es, cudaMemcpyHostToDevice, st()));
    } else {
      CK(cudaMemcpy2DAsync(stage, (size_t)rows * es, src + (size_t)r0 * es, (size_t)ntot * es, (size_t)rows * es, D,
                           cudaMemcpyHostToDevice, st()));
    }
    int64_t sld = x_layout == AGP_LAYOUT_ROWMAJOR ? D : rows;
    int bl = (int)((rows + 255) / 256);
    if (x_dtype == AGP_DTYPE_F64)
      convert_rows_kernel<double, T><<<bl, 256, 0, st()>>>((const double*)stage, x_layout, sld, rows, D, dst, Dp, dxx);
    else
      convert_rows_kernel<float, T><<<bl, 256, 0, st()>>>((const float*)stage, x_layout, sld, rows, D, dst, Dp, dxx);
    ++launches;
    CK(cudaGetLastError());
    return AGP_OK;
  }

  int data_upload(const void* Xh, int x_dtype, int x_layout, int64_t n_, const void* const* y, int y_kind) override {
    if (!Xh || !y || n_ < 1) BAD("null data");
    if ((x_dtype != 0 && x_dtype != 1) || (x_layout != 0 && x_layout != 1)) BAD("bad dtype/layout");
    if (is_lsm != (y_kind == AGP_Y_CLASS)) BAD("label kind does not match the likelihood");
    cudaFree(X); cudaFree(xx); cudaFree(y_all); cudaFree(ycls_all);
    X = nullptr; xx = nullptr; y_all = nullptr; ycls_all = nullptr;
    n = n_;
    CKS(dalloc(&X, (size_t)n * Dp)); CKS(dalloc(&xx, n));
    const int64_t chunk = 1 << 20;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
      int64_t rows = std::min(chunk, n - r0);
      CKS(upload_rows(Xh, x_dtype, x_layout, n, r0, rows, X + r0 * Dp, xx + r0));
      CK(cudaStreamSynchronize(st()));
    }
    if (y_kind == AGP_Y_CLASS) {
      const int32_t* yc = (const int32_t*)y[0];
      for (int64_t i = 0; i < n; ++i) if (yc[i] < 0 || yc[i] >= Qg) BAD("Some labels of y are not part of the expect labels");
      CKS(dalloc(&ycls_all, n));
      CK(cudaMemcpyAsync(ycls_all, yc, n * sizeof(int), cudaMemcpyHostToDevice, st()));
    } else {
      CKS(dalloc(&y_all, (size_t)nT * n));
      for (int t = 0; t < nT; ++t) {
        if (!y[t]) BAD("null y");
        if (h_lik_kind[t] == AGP_LIK_LOGISTIC || h_lik_kind[t] == AGP_LIK_BAYESIANSVM) {
          const double* yt = (const double*)y[t];
          for (int64_t i = 0; i < n; ++i) if (yt[i] != 1.0 && yt[i] != -1.0) BAD("Labels of y should be binary {-1,1} or {0,1}");
        }
        if (h_lik_kind[t] == AGP_LIK_POISSON || h_lik_kind[t] == AGP_LIK_NEGBINOMIAL) {  // likelihood/event.jl:7-13
          const double* yt = (const double*)y[t];
          for (int64_t i = 0; i < n; ++i) if (!(yt[i] >= 0.0) || yt[i] != std::floor(yt[i])) BAD("For event count target(s) should be integers");
        }
        CK(cudaMemcpyAsync(y_all + (size_t)t * n, y[t], n * sizeof(double), cudaMemcpyHostToDevice, st()));
      }
    }
    CK(cudaStreamSynchronize(st()));
    have_data = true;
    return AGP_OK;
  }

  int minibatches_upload(const int64_t* idx, int64_t nl, int B, int base) override {
    if (!have_data) { ctx->err = "upload data first"; return AGP_ERR_STATE; }
    if (!idx || nl < 1 || B < 1 || B > Bcap || (base != 0 && base != 1)) BAD("bad minibatch lists");
    for (int64_t i = 0; i < nl * B; ++i) if (idx[i] - base < 0 || idx[i] - base >= n) BAD("minibatch index out of range");
    cudaFree(idx_pool); idx_pool = nullptr;
    CKS(dalloc(&idx_pool, (size_t)nl * B));
    CKS(ensure_stage((size_t)nl * B * 8));
    CK(cudaMemcpyAsync(stage, idx, (size_t)nl * B * 8, cudaMemcpyHostToDevice, st()));
    idx_rebase_kernel<<<(int)((nl * B + 255) / 256), 256, 0, st()>>>((const int64_t*)stage, idx_pool, nl * B, base);
    ++launches;
    int64_t zero = 0;
    CK(cudaMemcpyAsync(counters + 1, &zero, 8, cudaMemcpyHostToDevice, st()));
    CK(cudaStreamSynchronize(st()));
    n_lists = nl; pool_B = B; prefetched = false; stats_early = false;
    drop_graph();
    return AGP_OK;
  }

  // ---- SPD inverse: blocked Cholesky (right-looking, nb = 64) + recursive triangular inverse + X^T X --------
  // P: in = SPD (lower triangle read), out = L in the lower triangle.  X = L^-1.  Out lower triangle = P^-1.
  void spd_inverse(double* P, double* Xw, double* Ww, double* Out, double* logdet) {
    const int nb = POTF2_NB, nblk = mp / nb;
    const int64_t ld = mp;
    ph_begin(PH_CHOL);
    // The recursive-doubling inverse below reads whole s x s diagonal blocks of Xw (s = 128, 256, ...) but only ever writes its 64 x 64
    // diagonal blocks and the blocks below them: the blocks ABOVE the block diagonal must be zero.  They are after allocation, but Xw
    // (L.X) is also the scratch matrix of the hyper-parameter gradients / the Newton-Schulz seed (dense Sigma), after which the next
    // factorisation of a matrix with mp >= 256 silently picked the leftovers up (K-tilde < 0 on the step after agp_hyper_grads).
    if (nblk > 2) cudaMemsetAsync(Xw, 0, (size_t)mp * mp * sizeof(double), st());
    for (int k = 0; k < nblk; ++k) {
      double* dk = P + (int64_t)k * nb * (ld + 1);
      double* xk = Xw + (int64_t)k * nb * (ld + 1);
      potf2_inv_kernel<<<1, 256, potf2_smem(), st()>>>(dk, ld, xk, ld, logdet, status);
      ++launches;
      int rem = mp - (k + 1) * nb;
      if (rem > 0) {
        GemmParams<double> g{};
        double* panel = P + (int64_t)(k + 1) * nb * ld + (int64_t)k * nb;
        g.A = panel; g.lda = ld; g.B = xk; g.ldb = ld; g.C = panel; g.ldc = ld;
        g.M = rem; g.N = nb; g.K = nb; g.alpha = 1.0; g.beta = 0.0;
        gemm_simt_launch<double, false, false, EPI_PLAIN>(g, 1, st());
        GemmParams<double> u{};
        u.A = panel; u.lda = ld; u.B = panel; u.ldb = ld;
        u.C = P + (int64_t)(k + 1) * nb * (ld + 1); u.ldc = ld;
        u.M = rem; u.N = rem; u.K = nb; u.alpha = -1.0; u.beta = 1.0; u.lower_only = 1;
        gemm_simt_launch<double, false, false, EPI_PLAIN>(u, 1, st());
        launches += 2;
      }
    }
    ph_end();
    ph_begin(PH_TRTRI);
    for (int s = nb; s < mp; s *= 2) {
      int npairs = mp / (2 * s);
      int64_t zs = (int64_t)2 * s * (ld + 1);
      GemmParams<double> a{};  // T = L21 * X11
      a.A = P + (int64_t)s * ld; a.lda = ld; a.B = Xw; a.ldb = ld; a.C = Ww + (int64_t)s * ld; a.ldc = ld;
      a.M = s; a.N = s; a.K = s; a.alpha = 1.0; a.beta = 0.0; a.zs_a = zs; a.zs_b = zs; a.zs_c = zs;
      gemm_simt_launch<double, false, true, EPI_PLAIN>(a, npairs, st());
      GemmParams<double> b{};  // X21 = -X22 * T
      b.A = Xw + (int64_t)s * (ld + 1); b.lda = ld; b.B = Ww + (int64_t)s * ld; b.ldb = ld; b.C = Xw + (int64_t)s * ld; b.ldc = ld;
      b.M = s; b.N = s; b.K = s; b.alpha = -1.0; b.beta = 0.0; b.zs_a = zs; b.zs_b = zs; b.zs_c = zs;
      gemm_simt_launch<double, false, true, EPI_PLAIN>(b, npairs, st());
      launches += 2;
    }
    ph_end();
    ph_begin(PH_SIGMA);
    GemmParams<double> c{};  // Out = X^T X (lower tiles; k starts at the diagonal because X is lower triangular)
    c.A = Xw; c.lda = ld; c.B = Xw; c.ldb = ld; c.C = Out; c.ldc = ld;
    c.M = mp; c.N = mp; c.K = mp; c.alpha = 1.0; c.beta = 0.0; c.lower_only = 1; c.k_from_diag = 1;
    gemm_simt_launch<double, true, true, EPI_PLAIN>(c, 1, st());
    ++launches;
    ph_end();
  }

  // ---- small fp64 m x m helpers (off the hot path) ------------------------------------------------------
  void dgemm(bool at, bool bt, const double* A, const double* B, double* Cc, double alpha, double beta) {
    GemmParams<double> g{};
    g.A = A; g.lda = mp; g.B = B; g.ldb = mp; g.C = Cc; g.ldc = mp; g.M = mp; g.N = mp; g.K = mp; g.alpha = alpha; g.beta = beta;
    if (!at && !bt) gemm_simt_launch<double, false, false, EPI_PLAIN>(g, 1, st());
    else if (!at && bt) gemm_simt_launch<double, false, true, EPI_PLAIN>(g, 1, st());
    else gemm_simt_launch<double, true, true, EPI_PLAIN>(g, 1, st());
    ++launches;
  }
  dim3 grid_mp() const { return dim3((mp + 127) / 128, mp); }
  // rows of the B x m tensor-core products: the tcgen05 kernels tile the samples by 128, so a ragged minibatch (host index list path)
  // runs with its row count rounded up -- the extra rows repeat sample 0 (prep_idx) and carry zero weights (natgrad_products)
  int rowsK(int B) const { return prec == AGP_PREC_TF32X3 ? (int)rup(B, 128) : B; }

  // whitened -> canonical natural parameters: eta1 = L^-T eta1_v, eta2 = L^-T eta2_v L^-1   (X = L^-1)
  void canonicalize(Latent& L) {
    matvec_t_kernel<<<(mp + 127) / 128, 128, 0, st()>>>(L.Linv, mp, mp, L.eta1v, L.eta1c);
    ++launches;
    dgemm(false, true, L.eta2v, L.Linv, L.W, 1.0, 0.0);   // W = eta2_v * X
    dgemm(true, true, L.Linv, L.W, L.eta2c, 1.0, 0.0);    // eta2 = X^T * W
  }
  // canonical -> whitened: eta1_v = L^T eta1, P_v = L^T (-2 eta2) L, then Sigma_v = P_v^-1, mu_v = Sigma_v eta1_v
  int whiten(Latent& L) {
    matvec_t_kernel<<<(mp + 127) / 128, 128, 0, st()>>>(L.Lc, mp, mp, L.eta1c, L.eta1v);
    ++launches;
    dgemm(false, true, L.eta2c, L.Lc, L.W, 1.0, 0.0);     // W = eta2 * L
    dgemm(true, true, L.Lc, L.W, L.eta2v, 1.0, 0.0);      // eta2_v = L^T * W
    scale_pad_kernel<<<grid_mp(), 128, 0, st()>>>(L.eta2v, L.P, mp, m, mp, -2.0);
    ++launches;
    CK(cudaMemsetAsync(L.logdetP, 0, sizeof(double), st()));
    CKS(eta_to_moments(L));
    L.white_valid = true;
    L.factor_valid = true; L.ns_seeded = false;   // new basis / new natural parameters: the refined covariance is stale
    return AGP_OK;
  }

  int refresh_K() override {
    h_mu_valid = false;
    for (auto& L : lat) {
      if (L.white_valid) { canonicalize(L); L.white_valid = false; }  // K changes: carry the posterior over in canonical form
      // K_mm in fp64 (GEMM form of the squared distance is exact enough in fp64), then the exact diagonal
      GemmParams<double> g{};
      g.A = L.Zd; g.lda = Dp; g.B = L.Zd; g.ldb = Dp; g.C = L.P; g.ldc = mp; g.M = m; g.N = m; g.K = D;
      g.alpha = 1.0; g.xx = L.zzd; g.zz = L.zzd; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
      gemm_simt_launch<double, false, false, EPI_KERNELFN>(g, 1, st());
      kmm_fix_kernel<<<grid_mp(), 128, 0, st()>>>(L.P, mp, m, mp, L.variance + jitter);
      launches += 2;
      CK(cudaMemsetAsync(L.logdetP + 1, 0, sizeof(double), st()));
      spd_inverse(L.P, L.X, L.W, L.Kinv, L.logdetP + 1);
      copy_lower_kernel<<<grid_mp(), 128, 0, st()>>>(L.P, L.Lc, mp, mp);
      copy_lower_kernel<<<grid_mp(), 128, 0, st()>>>(L.X, L.Linv, mp, mp);
      symmetrize_shadow_kernel<T><<<dim3((m + 127) / 128, m), 128, 0, st()>>>(L.Kinv, mp, m, (T*)nullptr, ldm);
      shadow_kernel<T><<<dim3((m + 127) / 128, m), 128, 0, st()>>>(L.Linv, mp, m, L.Linv_T, ldm);
      if (L.um.v2 || L.um.ps) { CKS(umma_presplit(ctx_err(), L.um, UM_LINV, (const float*)(const void*)L.Linv_T, st())); ++launches; }
      symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Linv, mp, m, L.mu0, L.mu0v);  // L^-1 mu0
      launches += 5;
      CK(cudaMemcpyAsync(&L.logdetK, L.logdetP + 1, sizeof(double), cudaMemcpyDeviceToHost, st()));
      int s0 = sync_status();   // a failed Cholesky of K must not be hidden by the next factorisation
      if (s0 != AGP_OK) { have_K = false; return s0; }
      CKS(whiten(L));
    }
    CK(cudaGetLastError());
    int s = sync_status();
    have_K = (s == AGP_OK);
    K_dirty = false;
    prefetched = false; stats_early = false;
    drop_graph();
    return s;
  }

  // d_lr holds the step size of the UPCOMING iteration: lr = (tau + t)^-kappa (optimisers.jl:14-19), 1 for AnalyticVI
  double fixed_lr = 0.0;   // > 0: Descent(eta) instead of Robbins-Monro for the stochastic natural-gradient step (agp_set_step_size)
  int set_step_size(double eta) override {
    if (!(eta >= 0.0) || eta > 1.0) BAD("step size must be in [0, 1] (0 = back to Robbins-Monro)");
    fixed_lr = eta;
    drop_graph();
    int64_t c[2];
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(c, counters, 16, cudaMemcpyDeviceToHost));
    return upload_lr(c[0]);
  }
  int upload_lr(int64_t t) {
    double lr = stochastic ? (fixed_lr > 0.0 ? fixed_lr : std::pow(rm_tau + (double)t, -rm_kappa)) : 1.0;
    CK(cudaMemcpy(d_lr, &lr, 8, cudaMemcpyHostToDevice));
    return AGP_OK;
  }
  int state_reset() override {
    int64_t c[2];
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(c, counters, 16, cudaMemcpyDeviceToHost));
    c[0] = 1;
    h_steps = 0;
    CK(cudaMemcpy(counters, c, 16, cudaMemcpyHostToDevice));
    CKS(upload_lr(1));
    curB = 0; have_step = false;
    CKS(reset_A_state());     // init_state_A (training/states.jl:100-105)
    CKS(reset_noise_state()); // init_local_vars(::GaussianLikelihood) re-creates state_sigma2 (gaussian.jl:47-54)
    return reset_local_vars();
  }

  int set_kernel(int ql, int kind, double scale, double variance) override {
    if (ql < 0 || ql >= Ql || kind < 0 || kind > 2 || !(scale > 0) || !(variance > 0)) BAD("bad kernel parameters");
    lat[ql].kind = kind; lat[ql].scale = scale; lat[ql].variance = variance;
    if (stale_K_ok && have_K) K_dirty = true; else have_K = false;
    prefetched = false; stats_early = false;
    drop_graph();
    return AGP_OK;
  }

  // ---- the step -------------------------------------------------------------------------------------
  int prep_idx(const int64_t* idx, int B, int base, int cursor_offset = 0) {
    ph_begin(PH_IDX);
    if (idx) {
      for (int i = 0; i < B; ++i) if (idx[i] - base < 0 || idx[i] - base >= n) { ph_end(); BAD("minibatch index out of range"); }
      CKS(ensure_stage((size_t)B * 8));
      CK(cudaMemcpyAsync(stage, idx, (size_t)B * 8, cudaMemcpyHostToDevice, st()));
      idx_rebase_kernel<<<(B + 255) / 256, 256, 0, st()>>>((const int64_t*)stage, idx_cur, B, base);
      const int Bk = rowsK(B);
      if (Bk != B) { idx_pad_kernel<<<1, 128, 0, st()>>>(idx_cur, B, Bk); ++launches; }
      xx_gather_kernel<T><<<(Bk + 255) / 256, 256, 0, st()>>>(idx_cur, Bk, xx, xx_cur);
      ++launches;
    } else {
      if (!idx_pool || pool_B != B) { ph_end(); ctx->err = "no resident minibatch lists for this batch size"; return AGP_ERR_STATE; }
      idx_select_kernel<T><<<(B + 255) / 256, 256, 0, st()>>>(idx_pool, n_lists, B, counters, cursor_offset, idx_cur, xx, xx_cur);
    }
    ++launches;
    ph_end();
    return AGP_OK;
  }

  // kernel matrices + predictive moments of the owned latents for B rows of Xsrc (gathered through `gather` when given).
  // Also the whole of _predict_f (training/predictions.jl:25-50): mu* = k* (K \ mu) = V* mu_v and
  // sigma2* = kdiag + jitter - diag(k* A k*^T) = Ktilde* + rowsum((V* Sigma_v) .* V*).
  // stages: 1 = kernel matrices (Knm, V [+ sum V^2]), 2 = V X^T + row statistics, 3 = both
  // ---- fan-out of the per-latent small kernels (K_nm, row statistics, scale-transpose, combine, finalize) of multi-latent steps over
  // NAUX auxiliary streams: they are independent across latents and each is too short (4-10 us, partly launch latency) to fill the
  // GPU alone.  Inside a graph capture the streams become parallel branches.  AGP_FAN=0 disables. ----
  static constexpr int NAUX = 4;
  cudaStream_t aux[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_aux_fork = nullptr, ev_aux_join[NAUX] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t fan_base = nullptr; bool fan_active = false, fan_on = true;
  int fan_init() {
    if (const char* e = getenv("AGP_FAN")) if (e[0] == '0') fan_on = false;
    if (Ql < 2) fan_on = false;
    if (!fan_on) return AGP_OK;
    CK(cudaEventCreateWithFlags(&ev_aux_fork, cudaEventDisableTiming));
    for (int i = 0; i < NAUX; ++i) {
      CK(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev_aux_join[i], cudaEventDisableTiming));
    }
    return AGP_OK;
  }
  void fan_begin() {
    if (!fan_on || prof || fan_active) return;
    fan_base = st();
    cudaEventRecord(ev_aux_fork, fan_base);
    for (int i = 0; i < NAUX; ++i) cudaStreamWaitEvent(aux[i], ev_aux_fork, 0);
    fan_active = true;
  }
  void fan_select(int q) { if (fan_active) cur_stream = aux[q % NAUX]; }
  void fan_end() {
    if (!fan_active) return;
    for (int i = 0; i < NAUX; ++i) {
      cudaEventRecord(ev_aux_join[i], aux[i]);
      cudaStreamWaitEvent(fan_base, ev_aux_join[i], 0);
    }
    cur_stream = (fan_base == ctx->stream) ? nullptr : fan_base;
    fan_active = false;
  }

  // ---- grouped tcgen05 launches: with >= 2 owned latents (tf32x3) each of the three B x m x m products of a step is ONE persistent
  // launch over all owned latents (umma_gemm_grouped_kernel) instead of one launch per latent ----
  UmmaGroups grpV, grpS, grpG, grpGT;
  bool use_groups = false;
  bool grp_gram_tn = false;      // grouped Gram product straight from V (no scale-transpose launches); AGP_GRAM_TN_GROUPED=0 disables
  bool batch_small = false;      // row statistics / natural-parameter update / finalize of several latents per launch; AGP_BATCH_SMALL=0 disables
  int groups_init() {
    use_groups = false; grp_gram_tn = false; batch_small = false;
    if (prec != AGP_PREC_TF32X3 || Ql < 2 || is_vgp || getenv("AGP_NO_GROUPED")) return AGP_OK;
    std::vector<UmmaLatent*> lp; std::vector<float*> cV, cG; std::vector<double*> a0V, a0S, a1S; std::vector<const double*> tv;
    for (auto& L : lat) {
      lp.push_back(&L.um); cV.push_back((float*)(void*)L.V); cG.push_back((float*)(void*)L.Gpart);
      a0V.push_back(L.racc); a0S.push_back(L.racc + ldB); a1S.push_back(L.racc + 2 * ldB); tv.push_back(L.tvec);
    }
    CKS(umma_groups_build(ctx_err(), grpV, lp.data(), Ql, UM_KNM, UM_LINV, cV.data(), a0V.data(), nullptr, nullptr, st()));
    CKS(umma_groups_build(ctx_err(), grpS, lp.data(), Ql, UM_V, UM_X, nullptr, a0S.data(), a1S.data(), tv.data(), st()));
    CKS(umma_groups_build(ctx_err(), grpG, lp.data(), Ql, -1, -1, cG.data(), nullptr, nullptr, nullptr, st()));
    use_groups = true;
    // measured on one B200: C5 (64 latents) 13.35 k -> 13.59 k latent-it/s, C4 (8 latents) 24.5 k -> 24.2 k: with few latents the four
    // auxiliary streams already run the per-latent kernels side by side, so the batched form is the default above 16 latents only
    batch_small = Ql > 16;
    if (const char* env = getenv("AGP_BATCH_SMALL")) batch_small = atoi(env) != 0;
    return AGP_OK;
  }
  // needs the gradient buffers (gmu, gS), which the constructor allocates after groups_init
  int gram_tn_group_init() {
    grp_gram_tn = false;
    if (!use_groups) return AGP_OK;
    bool tn_ok = true;
    for (auto& L : lat) tn_ok = tn_ok && L.um.gram_tn;
    if (const char* env = getenv("AGP_GRAM_TN_GROUPED")) tn_ok = tn_ok && atoi(env) != 0;
    if (!tn_ok) return AGP_OK;
    std::vector<UmmaLatent*> lp; std::vector<float*> cG; std::vector<const double*> wq, gq; std::vector<double*> v1q;
    for (int q = 0; q < Ql; ++q) {
      lp.push_back(&lat[q].um); cG.push_back((float*)(void*)lat[q].Gpart);
      wq.push_back(gS + (size_t)q * ldB); gq.push_back(gmu + (size_t)q * ldB); v1q.push_back(lat[q].v1);
    }
    CKS(umma_gram_tn_groups_build(ctx_err(), grpGT, lp.data(), Ql, cG.data(), wq.data(), gq.data(), v1q.data(), st()));
    grp_gram_tn = true;
    return AGP_OK;
  }
  int moments_rows_grouped(const T* Xsrc, const T* xsrc, const int64_t* gather, int B, bool fresh_kernel_matrices, double* mean_out,
                           double* var_out, int64_t out_ld, int stages) {
    const int Bk = rowsK(B);
    if (stages & 1) {
      ph_begin(PH_KMAT);
      fan_begin();
      for (int q = 0; q < Ql; ++q) {
        Latent& L = lat[q];
        fan_select(q);
        if (L.knm_tc) {
          CKS(umma_knm(ctx_err(), L.uk, (const float*)(const void*)Xsrc, Dp, Dp, gather, (const float*)(const void*)(gather ? xx_cur : xsrc),
                       (const float*)(const void*)L.zz, Bk, L.kind, L.scale * L.scale, L.variance, st()));
        } else {
          GemmParams<T> g{};
          g.A = Xsrc; g.lda = Dp; g.a_gather = gather; g.B = L.Z; g.ldb = Dp; g.C = L.Knm; g.ldc = ldm;
          g.M = Bk; g.N = m; g.K = D; g.alpha = 1.0;
          g.xx = gather ? xx_cur : xsrc; g.xx_direct = 1; g.zz = L.zz; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
          gemm_simt_launch<T, false, false, EPI_KERNELFN>(g, 1, st());
        }
        ++launches;
        CK(cudaMemsetAsync(L.racc, 0, ldB * sizeof(double), st()));
      }
      fan_end();
      ph_end();
      ph_begin(PH_KAPPA);
      CKS(umma_gemm_nt_grouped(ctx_err(), grpV, lat[0].um, 1, Bk, mk, UMMA_EPI_STORE_SUMSQ, st()));
      ++launches;
      ph_end();
    }
    if (!(stages & 2)) { CK(cudaGetLastError()); return AGP_OK; }
    ph_begin(PH_KSIGMA);
    for (int q = 0; q < Ql; ++q) CK(cudaMemsetAsync(lat[q].racc + ldB, 0, 2 * ldB * sizeof(double), st()));
    racc2_precleared = false;
    CKS(umma_gemm_nt_grouped(ctx_err(), grpS, lat[0].um, 1, Bk, mk, UMMA_EPI_STATS_ONLY, st()));
    ++launches;
    ph_end();
    ph_begin(PH_ROWSTATS);
    fan_begin();
    if (batch_small) {
      // up to SMALL_NB latents per launch (pointers by value), the launches over the auxiliary streams
      for (int q0 = 0, nb_ = 0; q0 < Ql; q0 += SMALL_NB, ++nb_) {
        const int cnt = std::min(SMALL_NB, Ql - q0);
        RowFinishBatch bt{};
        for (int z = 0; z < cnt; ++z) { bt.racc[z] = lat[q0 + z].racc; bt.Ktilde[z] = lat[q0 + z].Ktilde; bt.kdiag[z] = lat[q0 + z].variance + jitter; }
        fan_select(nb_);
        launch_chain(rowfinish_batched_kernel, dim3((B + 255) / 256, cnt), dim3(256), 0, bt, (int64_t)ldB, B, mean_out + (size_t)q0 * out_ld,
                     var_out + (size_t)q0 * out_ld, out_ld, status, fresh_kernel_matrices ? 1 : 0,
                     (const int64_t*)((peer && mean_out == mean_f + (size_t)qbeg * ldB) ? d_xepoch : nullptr), par_stride);
        ++launches;
      }
    } else
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      fan_select(q);
      launch_chain(rowfinish_kernel, dim3((B + 255) / 256), dim3(256), 0, (const double*)L.racc, (const double*)(L.racc + ldB),
                   (const double*)(L.racc + 2 * ldB), B, L.variance + jitter, L.Ktilde, mean_out + (size_t)q * out_ld,
                   var_out + (size_t)q * out_ld, status, fresh_kernel_matrices ? 1 : 0,
                   (const int64_t*)((peer && mean_out == mean_f + (size_t)qbeg * ldB) ? d_xepoch : nullptr), par_stride);
      ++launches;
    }
    fan_end();
    ph_end();
    CK(cudaGetLastError());
    return AGP_OK;
  }
  bool groups_now() const {
    if (!use_groups || vgp_identity) return false;
    for (auto& L : lat) if (!L.factor_valid) return false;
    return true;
  }
  int moments_rows(const T* Xsrc, const T* xsrc, const int64_t* gather, int B, bool fresh_kernel_matrices, double* mean_out,
                   double* var_out, int64_t out_ld, bool need_var, int stages = 3) {
    if (groups_now()) return moments_rows_grouped(Xsrc, xsrc, gather, B, fresh_kernel_matrices, mean_out, var_out, out_ld, stages);
    const int Bk = rowsK(B);
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      if ((stages & 1) && vgp_identity) {
        // full VGP: V = chol(K) (see vgp_fill_v_kernel); Ktilde = 0 is forced in the row statistics below
        ph_begin(PH_KMAT);
        if (B != m) { ph_end(); BAD("VGP steps are full-batch: B must equal the number of training points"); }
        vgp_fill_v_kernel<T><<<dim3((int)((ldm + 127) / 128), B), 128, 0, st()>>>(L.Lc, mp, m, L.V, ldm);
        ++launches;
        if (prec == AGP_PREC_TF32X3) CK(cudaMemsetAsync(L.racc, 0, ldB * sizeof(double), st()));
        ph_end();
      } else if (stages & 1) {
        ph_begin(PH_KMAT);
        if (L.knm_tc) {
          CKS(umma_knm(ctx_err(), L.uk, (const float*)(const void*)Xsrc, Dp, Dp, gather, (const float*)(const void*)(gather ? xx_cur : xsrc),
                       (const float*)(const void*)L.zz, Bk, L.kind, L.scale * L.scale, L.variance, st()));
        } else {
        GemmParams<T> g{};
        g.A = Xsrc; g.lda = Dp; g.a_gather = gather; g.B = L.Z; g.ldb = Dp; g.C = L.Knm; g.ldc = ldm;
        g.M = Bk; g.N = m; g.K = D; g.alpha = 1.0;
        g.xx = gather ? xx_cur : xsrc; g.xx_direct = 1; g.zz = L.zz; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
        gemm_simt_launch<T, false, false, EPI_KERNELFN>(g, 1, st());
        }
        ++launches;
        ph_end();
        if (prec == AGP_PREC_TF32X3) {
          ph_begin(PH_KAPPA);
          CK(cudaMemsetAsync(L.racc, 0, ldB * sizeof(double), st()));
          UmmaEpilogue ep{};
          ep.mode = UMMA_EPI_STORE_SUMSQ; ep.acc0 = L.racc;
          CKS(umma_gemm_nt(ctx_err(), L.um, UM_KNM, UM_LINV, (float*)(void*)L.V, Bk, mk, ep, st()));
          ++launches;
          ph_end();
        } else {
          ph_begin(PH_KAPPA);
          GemmParams<T> k{};  // V = Knm L^-T  (the whitened kappa of latentgp.jl:211); L^-1 is lower triangular
          k.A = L.Knm; k.lda = ldm; k.B = L.Linv_T; k.ldb = ldm; k.C = L.V; k.ldc = ldm; k.M = B; k.N = m; k.K = m; k.alpha = 1.0;
          k.k_to_diag = 1;
          gemm_simt_launch<T, false, false, EPI_PLAIN>(k, 1, st());
          ++launches;
          ph_end();
        }
      }
      if (!(stages & 2)) continue;
      {
        ph_begin(PH_KSIGMA);
        if (prec == AGP_PREC_TF32X3) {
          if (!((racc2_precleared || stats_use_early) && Ql == 1)) CK(cudaMemsetAsync(L.racc + ldB, 0, 2 * ldB * sizeof(double), st()));
          racc2_precleared = false;
          UmmaEpilogue ep{};
          ep.mode = UMMA_EPI_STATS_ONLY; ep.acc0 = L.racc + ldB; ep.acc1 = L.racc + 2 * ldB; ep.tvec = L.tvec;
          // single-latent SVGP step on the pre-split kernel: the row finish + local update of a sample runs in the epilogue thread that
          // adds its last N-tile contribution (UmmaRowFinish), no rowfinish_lik_kernel launch
          UmmaRowFinish fin{};
          if (fuse_lik_next && rowfin_on && L.um.ps && L.factor_valid && !stats_use_early && d_rowcnt) {
            fin.cnt = d_rowcnt; fin.B = B; fin.sumsq_v = L.racc; fin.kdiag_jit = L.variance + jitter; fin.Ktilde = L.Ktilde; fin.status = status;
            fin.lp = lik_params(B, fuse_from_batch, 1);
            ep.fin = &fin;
          }
          if (!L.factor_valid) {
            // the previous tail was a Newton-Schulz refinement: no factor, statistics against the full Sigma_v = ns.Y():
            // var_f - Ktilde = rowsum((V Sigma_v) o V),  mean_f = (V Sigma_v) eta1_v
            ep.mode = UMMA_EPI_STATS_SIGMA; ep.cin = (const float*)(const void*)L.V; ep.tvec = L.eta1v;
            CKS(umma_gemm_sigma(ctx_err(), L.um, L.ns, UM_V, Bk, ep, st()));
          } else {
            // with early statistics the N tiles 0 .. ntn-2 were accumulated behind the previous step's tail: only the last one is left
            if (stats_use_early) umma_set_tile_range(m / 128 - 1, m / 128 - 1);
            int sg = umma_gemm_nt(ctx_err(), L.um, UM_V, UM_X, (float*)(void*)L.VS, Bk, mk, ep, st());
            umma_set_tile_range(-1, -1);
            CKS(sg);
            rowfin_now = ep.fin != nullptr;
          }
        } else {
          GemmParams<T> s{};  // V X^T with Sigma_v = X^T X  (kappa * Sigma of latentgp.jl:189); X is lower triangular
          s.A = L.V; s.lda = ldm; s.B = L.Xv_T; s.ldb = ldm; s.C = L.VS; s.ldc = ldm; s.M = B; s.N = m; s.K = m; s.alpha = 1.0;
          s.k_to_diag = 1;
          gemm_simt_launch<T, false, false, EPI_PLAIN>(s, 1, st());
        }
        ++launches;
        ph_end();
      }
      ph_begin(PH_ROWSTATS);
      if (prec == AGP_PREC_TF32X3 && fuse_lik_next && rowfin_now) {
        lik_fused = true;      // done inside the statistics product
        rowfin_now = false;
      } else if (prec == AGP_PREC_TF32X3 && fuse_lik_next) {
        // single-latent SVGP step: local updates fused into the row-statistics kernel (step_update_a skips its lik launch)
        launch_chain(rowfinish_lik_kernel, dim3((B + 255) / 256), dim3(256), 0, (const double*)L.racc, (const double*)(L.racc + ldB),
                     (const double*)(L.racc + 2 * ldB), B, L.variance + jitter, L.Ktilde, status, lik_params(B, fuse_from_batch, 1));
        lik_fused = true;
      } else if (prec == AGP_PREC_TF32X3)
        launch_chain(rowfinish_kernel, dim3((B + 255) / 256), dim3(256), 0, (const double*)L.racc, (const double*)(L.racc + ldB),
                     (const double*)(L.racc + 2 * ldB), B, L.variance + jitter, L.Ktilde, mean_out + (size_t)q * out_ld,
                     var_out + (size_t)q * out_ld, status, vgp_identity ? 2 : (fresh_kernel_matrices ? 1 : 0),
                     (const int64_t*)((peer && mean_out == mean_f + (size_t)qbeg * ldB) ? d_xepoch : nullptr), par_stride);
      else
      rowstats_kernel<T><<<(B * 32 + 255) / 256, 256, 0, st()>>>(L.V, L.VS, L.tvec, B, m, ldm, L.variance + jitter,
                                                                 L.Ktilde, mean_out + (size_t)q * out_ld, var_out + (size_t)q * out_ld,
                                                                 status, vgp_identity ? 2 : (fresh_kernel_matrices ? 1 : 0),
                                                                 (const int64_t*)((peer && mean_out == mean_f + (size_t)qbeg * ldB) ? d_xepoch : nullptr), par_stride);
      ++launches;
      ph_end();
    }
    CK(cudaGetLastError());
    return AGP_OK;
  }
  int moments_impl(bool from_batch, int B, bool fresh, int stages = 3) {
    vgp_identity = is_vgp;
    int sm_ = moments_rows(from_batch ? Xb : X, from_batch ? xxb : xx, from_batch ? nullptr : idx_cur, B, fresh,
                           mean_f + (size_t)qbeg * ldB, var_f + (size_t)qbeg * ldB, ldB, true, stages);
    vgp_identity = false;
    CKS(sm_);
    if (peer && (stages & 2)) {   // publish the owned rows to every peer, then wait for theirs (device-side, graph-capturable)
      ph_begin(PH_ROWSTATS);
      peer_publish_kernel<<<(int)(((int64_t)Ql * B + 255) / 256), 256, 0, st()>>>(d_peers, peer_world, peer_rank, d_xepoch, par_stride, qbeg, Ql, ldB, B);
      peer_sync_kernel<<<1, 32 * ((peer_world + 31) / 32), 0, st()>>>(d_peers, peer_world, peer_rank, d_xepoch, 4 * par_stride, status);
      launches += 2;
      ph_end();
      CK(cudaGetLastError());
    }
    return AGP_OK;
  }

  // ---- hyper-parameter / inducing-point gradients of the ELBO (agp_hyper.cuh) -----------------------------------
  double *hgH[4] = {nullptr, nullptr, nullptr, nullptr}, *hgX = nullptr, *hgA = nullptr, *hgB = nullptr, *hgM1 = nullptr, *hgM2 = nullptr,
         *hgV = nullptr, *hgP = nullptr;
  void dgemm_bm(bool a_t, bool b_t, const double* A, int64_t lda, const double* Bm, int64_t ldb, double* Cc, int64_t ldc, int M, int N, int K,
                double alpha) {
    GemmParams<double> g{};
    g.A = A; g.lda = lda; g.B = Bm; g.ldb = ldb; g.C = Cc; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = 0.0;
    if (!a_t && !b_t) gemm_simt_launch<double, false, false, EPI_PLAIN>(g, 1, st());
    else if (!a_t && b_t) gemm_simt_launch<double, false, true, EPI_PLAIN>(g, 1, st());
    else gemm_simt_launch<double, true, true, EPI_PLAIN>(g, 1, st());
    ++launches;
  }
  int hyper_grads(double rho, double* d_scale, double* d_variance, double* dZ) override {
    if (is_vgp) BAD("hyper-parameter gradients of full (non-sparse) models are outside the accelerated path (autotuning.jl:48-84 differentiates a different ELBO)");
    if (have_K && K_dirty && stale_K_ok) {
      CKS(swap_in_fresh_K());
      int rc = hyper_grads_impl(rho, d_scale, d_variance, dZ);
      int rb = swap_back_stale_K();
      return rc != AGP_OK ? rc : rb;
    }
    return hyper_grads_impl(rho, d_scale, d_variance, dZ);
  }
  int hyper_grads_impl(double rho, double* d_scale, double* d_variance, double* dZ) {
    if (!d_scale || !d_variance) BAD("null output");
    if (curB < 1 || !have_step) { ctx->err = "hyper-parameter gradients need a completed step (the last minibatch is differentiated)"; return AGP_ERR_STATE; }
    CKS(ensure_factors());
    const int B = curB;
    const int64_t ldh = rup(m, 4);
    // moments under the updated posterior (like ELBO(model, state, y)); sharded models exchange them here (collective call)
    CKS(elbo_moments());
    LikParams lp = lik_params(B, true, 0);
    if (model_kind == AGP_MODEL_MOSVGP) launch_lik(lp, false);
    if (!hgA) {
      for (int i = 0; i < 4; ++i) CKS(dalloc(&hgH[i], (size_t)Bcap * ldh));
      CKS(dalloc(&hgX, (size_t)Bcap * Dp)); CKS(dalloc(&hgA, (size_t)Ql * ldB)); CKS(dalloc(&hgB, (size_t)Ql * ldB));
      CKS(dalloc(&hgM1, (size_t)mp * mp)); CKS(dalloc(&hgM2, (size_t)mp * mp)); CKS(dalloc(&hgV, 8 * (size_t)mp + 16)); CKS(dalloc(&hgP, 2 * (size_t)mp * Dp));
    }
    hg_ab_kernel<<<(B + 127) / 128, 128, 0, st()>>>(lp, hgA, hgB);
    hg_gather_x_kernel<T><<<(B + 7) / 8, dim3(32, 8), 0, st()>>>(cur_from_batch ? Xb : X, Dp, cur_from_batch ? nullptr : idx_cur, B, D, hgX, Dp);
    launches += 2;
    double *Knm = hgH[0], *kap = hgH[1], *Tm = hgH[2], *MK = hgH[3];
    double *mu_c = hgV, *dvec = hgV + mp, *vvec = hgV + 2 * mp, *cs1 = hgV + 3 * mp, *cs2 = hgV + 4 * mp, *rs2 = hgV + 5 * mp, *sc = hgV + 6 * mp;  // sc: scalars
    std::vector<double> hz((size_t)m * D);
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      // canonical mu, Sigma (fp64): mu = L mu_v, Sigma = L (X^T X) L^T
      ensure_muv(L);
      symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Lc, mp, m, L.muv, mu_c); ++launches;
      dgemm(true, true, L.Xv, L.Xv, L.X, 1.0, 0.0);
      dgemm(false, false, L.X, L.Lc, L.W, 1.0, 0.0);
      dgemm(false, true, L.Lc, L.W, L.X, 1.0, 0.0);                       // L.X = Sigma
      // K_nm in fp64 from row differences is not needed for the products: GEMM form with exact fp64 norms
      { GemmParams<double> g{};
        g.A = hgX; g.lda = Dp; g.B = L.Zd; g.ldb = Dp; g.C = Knm; g.ldc = ldh; g.M = B; g.N = m; g.K = D; g.alpha = 1.0;
        CK(cudaMemsetAsync(sc, 0, 16 * sizeof(double), st()));
        hg_rownorm_kernel<<<(B + 127) / 128, 128, 0, st()>>>(hgX, Dp, B, D, Tm);                 // |x|^2 (Tm as scratch vector)
        g.xx = Tm; g.xx_direct = 1; g.zz = L.zzd; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
        gemm_simt_launch<double, false, false, EPI_KERNELFN>(g, 1, st()); launches += 2; }
      dgemm_bm(false, true, Knm, ldh, L.Kinv, mp, kap, ldh, B, m, m, 1.0);                         // kappa = Knm K^-1
      dgemm_bm(false, true, kap, ldh, L.X, mp, Tm, ldh, B, m, m, 1.0);                             // T = kappa Sigma
      hg_M_kernel<<<dim3((m + 127) / 128, B), 128, 0, st()>>>(Tm, Knm, ldh, B, m, hgA + (size_t)q * ldB, hgB + (size_t)q * ldB, mu_c); ++launches;
      dgemm_bm(false, true, Tm, ldh, L.Kinv, mp, MK, ldh, B, m, m, 1.0);                           // MK = M K^-1
      dgemm_bm(true, true, kap, ldh, MK, ldh, hgM1, mp, m, m, B, -rho);                            // Acc = -rho kappa^T MK
      hg_Anm_kernel<<<dim3((m + 127) / 128, B), 128, 0, st()>>>(Tm, MK, kap, ldh, B, m, hgB + (size_t)q * ldB, rho); ++launches;   // Tm = A_nm
      // KL part: Kinv Sigma Kinv and v = Kinv (mu - mu0)
      dgemm(false, true, L.Kinv, L.X, L.W, 1.0, 0.0);                     // W = Kinv Sigma
      dgemm(false, true, L.W, L.Kinv, L.P, 1.0, 0.0);                     // P = Kinv Sigma Kinv
      hg_sub_kernel<<<(m + 127) / 128, 128, 0, st()>>>(mu_c, L.mu0, m, dvec);
      symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Kinv, mp, m, dvec, vvec);
      hg_Amm_kernel<<<dim3((m + 127) / 128, m), 128, 0, st()>>>(hgM2, hgM1, L.Kinv, L.P, vvec, mp, m);
      launches += 3;
      // contractions with the kernel derivatives
      CK(cudaMemsetAsync(cs1, 0, 3 * (size_t)mp * sizeof(double), st()));
      const size_t shb = 8 * (size_t)D * sizeof(double);
      hg_contract_kernel<<<dim3((m + 127) / 128, (B + 7) / 8), 128, shb, st()>>>(Tm, ldh, B, m, hgX, Dp, L.Zd, Dp, D, L.kind, L.scale, L.variance, sc, cs1, nullptr);
      hg_contract_kernel<<<dim3((m + 127) / 128, (m + 7) / 8), 128, shb, st()>>>(hgM2, mp, m, m, L.Zd, Dp, L.Zd, Dp, D, L.kind, L.scale, L.variance, sc + 2, cs2, rs2);
      hg_sum_kernel<<<1, 256, 0, st()>>>(hgB + (size_t)q * ldB, B, sc + 4);
      launches += 3;
      double h[6];
      CK(cudaMemcpyAsync(h, sc, 6 * sizeof(double), cudaMemcpyDeviceToHost, st()));
      if (dZ) {
        dgemm_bm(true, true, Tm, ldh, hgX, Dp, hgP, Dp, m, D, B, 1.0);                              // G1^T X_b
        dgemm_bm(true, true, hgM2, mp, L.Zd, Dp, hgP + (size_t)mp * Dp, Dp, m, D, m, 1.0);          // G^T Z  (G symmetric: counted twice)
        hg_dz_kernel<<<(m * D + 255) / 256, 256, 0, st()>>>(hgM1, D, m, hgP, hgP + (size_t)mp * Dp, hgP + (size_t)mp * Dp, Dp, cs1, cs2, cs2, L.Zd, Dp, L.scale);
        ++launches;
        CK(cudaMemcpyAsync(hz.data(), hgM1, (size_t)m * D * sizeof(double), cudaMemcpyDeviceToHost, st()));
      }
      CK(cudaStreamSynchronize(st()));
      d_scale[q] = h[0] + h[2];
      d_variance[q] = (h[1] + h[3]) / L.variance + rho * h[4];
      if (dZ) memcpy(dZ + (size_t)q * m * D, hz.data(), (size_t)m * D * sizeof(double));
    }
    CK(cudaGetLastError());
    return AGP_OK;
  }
  // setZ! (gpblocks/latentgp.jl:197): new inducing points for one owned latent; follow with agp_refresh_K
  int set_Z(int ql, const double* Zn) override {
    if (ql < 0 || ql >= Ql || !Zn) BAD("bad inducing-point update");
    Latent& L = lat[ql];
    L.hZ.assign(Zn, Zn + (size_t)m * D);
    std::vector<double> zp((size_t)m * Dp, 0.0), zn(m, 0.0);
    std::vector<T> zt((size_t)m * Dp, T(0)), znt(m, T(0));
    for (int i = 0; i < m; ++i) {
      double s_ = 0, s_t = 0;
      for (int k = 0; k < D; ++k) {
        double v = L.hZ[(size_t)i * D + k];
        zp[(size_t)i * Dp + k] = v; zt[(size_t)i * Dp + k] = (T)v;
        s_ += v * v; double vt = (double)(T)v; s_t += vt * vt;
      }
      zn[i] = s_; znt[i] = (T)s_t;
    }
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(L.Zd, zp.data(), zp.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.zzd, zn.data(), zn.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.Z, zt.data(), zt.size() * sizeof(T), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(L.zz, znt.data(), znt.size() * sizeof(T), cudaMemcpyHostToDevice));
    if (L.knm_tc) CKS(umma_knm_setup(ctx_err(), L.uk, (const float*)(const void*)L.Z, Dp, mk, D, (float*)(void*)L.Knm, ldm, Bcap, st()));
    if (stale_K_ok && have_K) K_dirty = true; else have_K = false;
    prefetched = false; stats_early = false;
    drop_graph();
    return AGP_OK;
  }

  // ---- quirk Q3 of the reference (agp_keep_stale_K) -------------------------------------------------------------------------
  // The reference's sparse update_hyperparameters! never raises the HyperParametersUpdated flag (hyperparameter/autotuning.jl:45
  // is commented out; compute_kernel_matrices clears it, training/training.jl:187-208): after a kernel / Z update the steps of
  // the same train! call keep the K_mm factor of the call's first iteration next to a K_nm built from the new kernel and Z, while
  // the next gradient evaluates ELBO(model, x, y, mu0, ks, Zs, state) with everything recomputed (functions/ELBO.jl:15-21).
  // With the policy on, agp_set_kernel / agp_set_Z leave the factor in place (K_dirty) and agp_hyper_grads swaps a fresh
  // factorisation in for the duration of the call: the posterior travels through its canonical form (fp64) both ways.
  bool stale_K_ok = false, K_dirty = false;
  int keep_stale_K(int on) override { stale_K_ok = on != 0; return AGP_OK; }
  int swap_in_fresh_K() {
    const size_t mm = (size_t)mp * mp;
    for (auto& L : lat) {
      if (!L.bkLc) {
        CKS(dalloc(&L.bkLc, mm)); CKS(dalloc(&L.bkLinv, mm)); CKS(dalloc(&L.bkKinv, mm)); CKS(dalloc(&L.bkmu0v, mp));
        CKS(dalloc(&L.bkLinvT, (size_t)mk * ldm));
      }
      CK(cudaMemcpyAsync(L.bkLc, L.Lc, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.bkLinv, L.Linv, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.bkKinv, L.Kinv, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.bkmu0v, L.mu0v, (size_t)mp * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.bkLinvT, L.Linv_T, (size_t)mk * ldm * sizeof(T), cudaMemcpyDeviceToDevice, st()));
      L.bklogdetK = L.logdetK;
    }
    CKS(ensure_factors());
    CKS(refresh_K());               // canonicalises with the parked factor, factorises the current kernel / Z, whitens again
    kernel_matrices_stale = true;   // V = Knm L^-T of the last minibatch belongs to the parked factor
    return AGP_OK;
  }
  int swap_back_stale_K() {
    const size_t mm = (size_t)mp * mp;
    for (auto& L : lat) {
      if (L.white_valid) { canonicalize(L); L.white_valid = false; }
      CK(cudaMemcpyAsync(L.Lc, L.bkLc, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.Linv, L.bkLinv, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.Kinv, L.bkKinv, mm * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.mu0v, L.bkmu0v, (size_t)mp * 8, cudaMemcpyDeviceToDevice, st()));
      CK(cudaMemcpyAsync(L.Linv_T, L.bkLinvT, (size_t)mk * ldm * sizeof(T), cudaMemcpyDeviceToDevice, st()));
      if (L.um.v2 || L.um.ps) { CKS(umma_presplit(ctx_err(), L.um, UM_LINV, (const float*)(const void*)L.Linv_T, st())); ++launches; }
      L.logdetK = L.bklogdetK;
      CKS(whiten(L));
    }
    kernel_matrices_stale = true; prefetched = false; stats_early = false;
    drop_graph();
    int s_ = sync_status();
    have_K = (s_ == AGP_OK);
    K_dirty = true;                 // the kernel / Z still differ from the factor
    return s_;
  }

  // ---- GaussianLikelihood(opt_noise = ADAM(..)) switch ------------------------------------------------------------------
  int reset_noise_state() {
    if (!noise_any) return AGP_OK;
    std::vector<double> st(4 * (size_t)nT, 0.0);
    for (int t = 0; t < nT; ++t) { st[4 * t + 2] = n_b1; st[4 * t + 3] = n_b2; }
    CK(cudaMemcpy(d_noise_state, st.data(), st.size() * 8, cudaMemcpyHostToDevice));
    return AGP_OK;
  }
  int set_noise_optimiser(int task, int kind, double eta, double b1, double b2, double eps) override {
    if (task < 0 || task >= nT || h_lik_kind[task] != AGP_LIK_GAUSSIAN) BAD("opt_noise applies to a GaussianLikelihood task");
    if (kind != 0 && kind != 1) BAD("unknown noise optimiser (0 = none, 1 = ADAM)");
    if (kind == 1 && !(eta > 0 && b1 > 0 && b1 < 1 && b2 > 0 && b2 < 1 && eps > 0)) BAD("bad ADAM parameters");
    CK(cudaStreamSynchronize(st()));
    drop_graph();
    if (!d_noise_opt) { CKS(dalloc(&d_noise_opt, nT)); CKS(dalloc(&d_noise_state, 4 * (size_t)nT)); h_noise_opt.assign(nT, 0); CK(cudaStreamSynchronize(st())); }
    h_noise_opt[task] = kind;
    if (kind == 1) { n_eta = eta; n_b1 = b1; n_b2 = b2; n_eps = eps; }
    noise_any = false;
    for (int v : h_noise_opt) noise_any = noise_any || v != 0;
    CK(cudaMemcpy(d_noise_opt, h_noise_opt.data(), nT * sizeof(int), cudaMemcpyHostToDevice));
    return reset_noise_state();
  }

  // ---- update_A! switch (MOSVGP.jl:51,79-81 `Aoptimiser`); kind 0 = off, 1 = ADAM ------------------------------------------
  int reset_A_state() {
    if (!a_opt) return AGP_OK;
    std::vector<double> bt(2 * (size_t)nT);
    for (int t = 0; t < nT; ++t) { bt[2 * t] = a_b1; bt[2 * t + 1] = a_b2; }
    CK(cudaMemsetAsync(d_Amt, 0, (size_t)nT * Qg * 8, st())); CK(cudaMemsetAsync(d_Avt, 0, (size_t)nT * Qg * 8, st()));
    CK(cudaMemcpyAsync(d_Abt, bt.data(), bt.size() * 8, cudaMemcpyHostToDevice, st()));
    CK(cudaStreamSynchronize(st()));
    return AGP_OK;
  }
  int set_A_optimiser(int kind, double eta, double b1, double b2, double eps) override {
    if (model_kind != AGP_MODEL_MOSVGP) BAD("the mixing matrix A only exists for MOSVGP");
    if (kind != 0 && kind != 1) BAD("unknown A optimiser (0 = none, 1 = ADAM)");
    if (kind == 1 && !(eta > 0 && b1 > 0 && b1 < 1 && b2 > 0 && b2 < 1 && eps > 0)) BAD("bad ADAM parameters");
    drop_graph();
    a_opt = kind == 1;
    if (!a_opt) return AGP_OK;
    a_eta = eta; a_b1 = b1; a_b2 = b2; a_eps = eps;
    if (!d_gradA) {
      CKS(dalloc(&d_gradA, (size_t)nT * Qg)); CKS(dalloc(&d_Amt, (size_t)nT * Qg)); CKS(dalloc(&d_Avt, (size_t)nT * Qg)); CKS(dalloc(&d_Abt, 2 * (size_t)nT));
    }
    return reset_A_state();
  }
  int get_A(double* A) override {
    if (model_kind != AGP_MODEL_MOSVGP || !A) BAD("no mixing matrix");
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(A, d_A, (size_t)nT * Qg * 8, cudaMemcpyDeviceToHost));
    return AGP_OK;
  }

  // ---- peer exchange set-up (one process per GPU, same node; handles travel through the host's process group) -----------
  int peer_export(void* handle64) override {
    if (!handle64) BAD("null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, xchg));
    memcpy(handle64, &h, 64);
    return AGP_OK;
  }
  int peer_attach(int world, int rank, const void* handles) override {
    if (world < 2 || world > 64 || rank < 0 || rank >= world || !handles) BAD("bad peer group");
    if (peer) BAD("peers already attached");
    std::vector<double*> ptrs(world, nullptr);
    for (int p = 0; p < world; ++p) {
      if (p == rank) { ptrs[p] = xchg; continue; }
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + 64 * (size_t)p, 64);
      void* q = nullptr;
      CK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      peer_opened.push_back(q);
      ptrs[p] = (double*)q;
    }
    CKS(dalloc(&d_peers, world));
    CK(cudaMemcpy(d_peers, ptrs.data(), world * sizeof(double*), cudaMemcpyHostToDevice));
    peer_world = world; peer_rank = rank; peer = true;
    drop_graph();
    return AGP_OK;
  }
  int peer_detach() override {   // back to the host-driven (NCCL) exchange, e.g. when another rank could not map the peers
    CK(cudaStreamSynchronize(st()));
    for (void* q : peer_opened) cudaIpcCloseMemHandle(q);
    peer_opened.clear();
    if (d_peers) { cudaFree(d_peers); d_peers = nullptr; }
    peer = false; peer_world = 1; peer_rank = 0;
    drop_graph();
    return AGP_OK;
  }

  // ---- OnlineSVGP (models/OnlineSVGP.jl, training/onlinetraining.jl) ---------------------------------------------------------
  // Carry-over of the previous inducing set Z_a into the natural gradient of the current one (analyticVI.jl:183-203):
  //   eta1 = K^-1 mu0 + kappa^T grad_mu + kappa_a^T prev_eta1,   eta2 = -(kappa^T diag(grad_Sigma) kappa + kappa_a^T invD_a kappa_a / 2 + K^-1 / 2)
  // with kappa_a = K_ab K^-1 (gpblocks/latentgp.jl:225-226).  In the whitened basis (eta_v = L^T eta [L]) the two constant terms become
  //   c1_v = V_a^T prev_eta1,  C2_v = V_a^T invD_a V_a / 2,  V_a = kappa_a L = K_ab L^-T,
  // added by combine_kernel.  invD_a, prev_eta1, prev_L are the outputs of save_old_gp! (onlinetraining.jl:171-183), passed in
  // canonical form; ma = 0 is the first batch (kappa_a = I, invD_a = I, prev_eta1 = 0: states.jl:86-98, latentgp.jl:220-223).
  int online_carry(int ql, const double* Za, int ma, const double* invDa, const double* prev_eta1, double prev_L) override {
    if (ql < 0 || ql >= Ql || ma < 0 || (ma > 0 && (!Za || !invDa || !prev_eta1))) BAD("bad online carry-over");
    if (stochastic) BAD("The inference object should be of type `AnalyticVI`");   // models/OnlineSVGP.jl:46
    if (!have_K) { ctx->err = "agp_refresh_K must be called before agp_online_carry"; return AGP_ERR_STATE; }
    Latent& L = lat[ql];
    drop_graph();
    CK(cudaStreamSynchronize(st()));
    cudaFree(L.on_c1v); cudaFree(L.on_C2v); cudaFree(L.on_Va);
    L.on_c1v = L.on_C2v = L.on_Va = nullptr;
    const int rows = ma > 0 ? (int)rup(ma, 4) : mp;
    CKS(dalloc(&L.on_c1v, (size_t)mp)); CKS(dalloc(&L.on_C2v, (size_t)mp * mp)); CKS(dalloc(&L.on_Va, (size_t)rows * mp));
    L.on_ma = ma; L.on_prevL = ma > 0 ? prev_L : 0.0; L.on_trDK = 0.0;
    L.on_invD.clear(); L.on_preveta1.clear();
    if (ma == 0) {
      CK(cudaMemcpyAsync(L.on_Va, L.Lc, (size_t)mp * mp * sizeof(double), cudaMemcpyDeviceToDevice, st()));   // V_a = I L
      dgemm(true, true, L.Lc, L.Lc, L.on_C2v, 0.5, 0.0);                                                        // L^T I L / 2
    } else {
      L.on_invD.assign(invDa, invDa + (size_t)ma * ma);
      L.on_preveta1.assign(prev_eta1, prev_eta1 + ma);
      const int64_t lda_ = rows;      // leading dimension of the ma x ma matrices
      std::vector<double> zp((size_t)rows * Dp, 0.0), zn(rows, 0.0), dpad((size_t)rows * lda_, 0.0), e1(rows, 0.0);
      for (int i = 0; i < ma; ++i) {
        double s_ = 0;
        for (int k = 0; k < D; ++k) { double v = Za[(size_t)i * D + k]; zp[(size_t)i * Dp + k] = v; s_ += v * v; }
        zn[i] = s_; e1[i] = prev_eta1[i];
        for (int j = 0; j < ma; ++j) dpad[(size_t)i * lda_ + j] = invDa[(size_t)i * ma + j];
      }
      double *dZa = nullptr, *dzn = nullptr, *dKab = nullptr, *dKa = nullptr, *dD = nullptr, *dTmp = nullptr, *de1 = nullptr;
      CKS(dalloc(&dZa, zp.size())); CKS(dalloc(&dzn, (size_t)rows)); CKS(dalloc(&dKab, (size_t)rows * mp)); CKS(dalloc(&dKa, (size_t)rows * lda_));
      CKS(dalloc(&dD, (size_t)rows * lda_)); CKS(dalloc(&dTmp, (size_t)rows * mp)); CKS(dalloc(&de1, (size_t)rows));
      CK(cudaMemcpyAsync(dZa, zp.data(), zp.size() * 8, cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(dzn, zn.data(), zn.size() * 8, cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(dD, dpad.data(), dpad.size() * 8, cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(de1, e1.data(), e1.size() * 8, cudaMemcpyHostToDevice, st()));
      GemmParams<double> g{};   // K_ab = k(Z_a, Z)
      g.A = dZa; g.lda = Dp; g.B = L.Zd; g.ldb = Dp; g.C = dKab; g.ldc = mp; g.M = ma; g.N = m; g.K = D;
      g.alpha = 1.0; g.xx = dzn; g.xx_direct = 1; g.zz = L.zzd; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
      gemm_simt_launch<double, false, false, EPI_KERNELFN>(g, 1, st());
      GemmParams<double> ga = g;   // K_a = k(Z_a, Z_a) + jitter I (exact diagonal)
      ga.B = dZa; ga.C = dKa; ga.ldc = lda_; ga.N = ma; ga.zz = dzn;
      gemm_simt_launch<double, false, false, EPI_KERNELFN>(ga, 1, st());
      kmm_fix_kernel<<<dim3((rows + 127) / 128, rows), 128, 0, st()>>>(dKa, lda_, ma, rows, L.variance + jitter);
      launches += 3;
      dgemm_bm(false, false, dKab, mp, L.Linv, mp, L.on_Va, mp, ma, m, m, 1.0);         // V_a = K_ab L^-T
      rect_matvec_t_kernel<<<(m + 127) / 128, 128, 0, st()>>>(L.on_Va, mp, ma, m, de1, L.on_c1v);   // c1_v = V_a^T prev_eta1
      ++launches;
      dgemm_bm(false, true, dD, lda_, L.on_Va, mp, dTmp, mp, ma, m, ma, 1.0);           // invD_a V_a
      dgemm_bm(true, true, L.on_Va, mp, dTmp, mp, L.on_C2v, mp, m, m, ma, 0.5);         // C2_v = V_a^T (invD_a V_a) / 2
      GemmParams<double> kt{};   // Ktilde_a = K_a - kappa_a K_ab^T = K_a - V_a V_a^T
      kt.A = L.on_Va; kt.lda = mp; kt.B = L.on_Va; kt.ldb = mp; kt.C = dKa; kt.ldc = lda_; kt.M = ma; kt.N = ma; kt.K = m; kt.alpha = -1.0; kt.beta = 1.0;
      gemm_simt_launch<double, false, false, EPI_PLAIN>(kt, 1, st());
      ++launches;
      std::vector<double> hKt((size_t)rows * lda_);
      CK(cudaMemcpyAsync(hKt.data(), dKa, hKt.size() * 8, cudaMemcpyDeviceToHost, st()));
      CK(cudaStreamSynchronize(st()));
      double tr = 0.0;   // trace_ABt(invD_a, Ktilde_a) (KLdivergences.jl:46-48)
      for (int i = 0; i < ma; ++i) for (int j = 0; j < ma; ++j) tr += invDa[(size_t)i * ma + j] * hKt[(size_t)i * lda_ + j];
      L.on_trDK = tr;
      cudaFree(dZa); cudaFree(dzn); cudaFree(dKab); cudaFree(dKa); cudaFree(dD); cudaFree(dTmp); cudaFree(de1);
    }
    L.online = true;
    CK(cudaGetLastError());
    return sync_status();
  }
  // extraKL(model::OnlineSVGP, state) (functions/KLdivergences.jl:37-54), summed over the owned latents; kappa_a mu = V_a mu_v and
  // kappa_a Sigma kappa_a^T = (V_a X^T)(V_a X^T)^T with Sigma_v = X^T X
  int online_extra_kl(double* out) override {
    if (!out) BAD("null output");
    double tot = 0.0;
    for (auto& L : lat) {
      if (!L.online) continue;
      CKS(ensure_factor(L));
      ensure_muv(L);
      const int ma = L.on_ma > 0 ? L.on_ma : m;
      const int rows = L.on_ma > 0 ? (int)rup(ma, 4) : mp;
      const int64_t lds = rows;
      double *dM1 = nullptr, *dS = nullptr, *dk = nullptr;
      CKS(dalloc(&dM1, (size_t)rows * mp)); CKS(dalloc(&dS, (size_t)rows * lds)); CKS(dalloc(&dk, (size_t)rows));
      rect_matvec_kernel<<<(ma + 127) / 128, 128, 0, st()>>>(L.on_Va, mp, ma, m, L.muv, dk);
      ++launches;
      dgemm_bm(false, false, L.on_Va, mp, L.Xv, mp, dM1, mp, ma, m, m, 1.0);     // V_a X^T
      dgemm_bm(false, false, dM1, mp, dM1, mp, dS, lds, ma, ma, m, 1.0);          // (V_a X^T)(V_a X^T)^T
      std::vector<double> hS((size_t)rows * lds), hk(rows);
      CK(cudaMemcpyAsync(hS.data(), dS, hS.size() * 8, cudaMemcpyDeviceToHost, st()));
      CK(cudaMemcpyAsync(hk.data(), dk, hk.size() * 8, cudaMemcpyDeviceToHost, st()));
      CK(cudaStreamSynchronize(st()));
      cudaFree(dM1); cudaFree(dS); cudaFree(dk);
      double tr = 0.0, lin = 0.0, quad = 0.0;
      for (int i = 0; i < ma; ++i) {
        double row = 0.0;
        for (int j = 0; j < ma; ++j) {
          const double d = L.on_ma > 0 ? L.on_invD[(size_t)i * ma + j] : (i == j ? 1.0 : 0.0);
          tr += d * hS[(size_t)i * lds + j];
          row += d * hk[j];
        }
        quad += hk[i] * row;
        if (L.on_ma > 0) lin += L.on_preveta1[i] * hk[i];
      }
      tot += L.on_prevL - (L.on_trDK + tr) / 2.0 + lin - quad / 2.0;
    }
    *out = tot;
    return AGP_OK;
  }
  // _predict_f(...; cov = true, diag = false) (training/predictions.jl:45-49): full predictive covariance of every owned latent,
  //   Sigma_f = K** + jitter I - k* A k*^T,  A = K^-1 (I - Sigma K^-1)   ==   K** + jitter I - V* V*^T + (V* X^T)(V* X^T)^T
  // in the whitened basis (V* = k* L^-T, Sigma_v = X^T X).  O(nt^2 m): fp64 SIMT throughout, off the hot path; Xt row-major [nt][D].
  int predict_f_cov(const double* Xt, int64_t nt64, double* mu_out, double* cov_out) override {
    if (!Xt || !mu_out || !cov_out || nt64 < 1) BAD("bad predict arguments");
    if (nt64 > 16384) BAD("full predictive covariance: at most 16384 test points per call (the result is nt x nt)");
    const int nt = (int)nt64;
    if (!have_K) CKS(refresh_K());
    CKS(ensure_factors());
    const int64_t ldn = rup(nt, 4);
    std::vector<double> xp((size_t)nt * Dp, 0.0), xn(nt, 0.0);
    for (int i = 0; i < nt; ++i) {
      double s_ = 0;
      for (int k = 0; k < D; ++k) { double v = Xt[(size_t)i * D + k]; xp[(size_t)i * Dp + k] = v; s_ += v * v; }
      xn[i] = s_;
    }
    double *dX = nullptr, *dxn = nullptr, *dKs = nullptr, *dV = nullptr, *dVS = nullptr, *dC = nullptr, *dmu = nullptr;
    CKS(dalloc(&dX, xp.size())); CKS(dalloc(&dxn, (size_t)nt)); CKS(dalloc(&dKs, (size_t)nt * mp)); CKS(dalloc(&dV, (size_t)nt * mp));
    CKS(dalloc(&dVS, (size_t)nt * mp)); CKS(dalloc(&dC, (size_t)nt * ldn)); CKS(dalloc(&dmu, (size_t)nt));
    CK(cudaMemcpyAsync(dX, xp.data(), xp.size() * 8, cudaMemcpyHostToDevice, st()));
    CK(cudaMemcpyAsync(dxn, xn.data(), xn.size() * 8, cudaMemcpyHostToDevice, st()));
    int rc = AGP_OK;
    for (int q = 0; q < Ql && rc == AGP_OK; ++q) {
      Latent& L = lat[q];
      GemmParams<double> g{};   // k* = k(X*, Z)
      g.A = dX; g.lda = Dp; g.B = L.Zd; g.ldb = Dp; g.C = dKs; g.ldc = mp; g.M = nt; g.N = m; g.K = D;
      g.alpha = 1.0; g.xx = dxn; g.xx_direct = 1; g.zz = L.zzd; g.scale2 = L.scale * L.scale; g.variance = L.variance; g.kernel_kind = L.kind;
      gemm_simt_launch<double, false, false, EPI_KERNELFN>(g, 1, st());
      GemmParams<double> gs = g;   // K** + jitter I (exact diagonal)
      gs.B = dX; gs.C = dC; gs.ldc = ldn; gs.N = nt; gs.zz = dxn;
      gemm_simt_launch<double, false, false, EPI_KERNELFN>(gs, 1, st());
      kmm_fix_kernel<<<dim3((int)((ldn + 127) / 128), nt), 128, 0, st()>>>(dC, ldn, nt, nt, L.variance + jitter);
      launches += 3;
      dgemm_bm(false, false, dKs, mp, L.Linv, mp, dV, mp, nt, m, m, 1.0);      // V* = k* L^-T
      dgemm_bm(false, false, dV, mp, L.Xv, mp, dVS, mp, nt, m, m, 1.0);        // V* X^T
      rect_matvec_kernel<<<(nt + 127) / 128, 128, 0, st()>>>(dVS, mp, nt, m, L.tvec, dmu);   // mu* = (V* X^T) t
      ++launches;
      GemmParams<double> c1{};
      c1.A = dV; c1.lda = mp; c1.B = dV; c1.ldb = mp; c1.C = dC; c1.ldc = ldn; c1.M = nt; c1.N = nt; c1.K = m; c1.alpha = -1.0; c1.beta = 1.0;
      gemm_simt_launch<double, false, false, EPI_PLAIN>(c1, 1, st());
      GemmParams<double> c2 = c1;
      c2.A = dVS; c2.B = dVS; c2.alpha = 1.0;
      gemm_simt_launch<double, false, false, EPI_PLAIN>(c2, 1, st());
      launches += 2;
      if (cudaMemcpy2DAsync(cov_out + (size_t)q * nt * nt, (size_t)nt * 8, dC, (size_t)ldn * 8, (size_t)nt * 8, nt, cudaMemcpyDeviceToHost, st()) != cudaSuccess) rc = AGP_ERR_CUDA;
      if (cudaMemcpyAsync(mu_out + (size_t)q * nt, dmu, (size_t)nt * 8, cudaMemcpyDeviceToHost, st()) != cudaSuccess) rc = AGP_ERR_CUDA;
      if (cudaStreamSynchronize(st()) != cudaSuccess) rc = AGP_ERR_CUDA;
    }
    cudaFree(dX); cudaFree(dxn); cudaFree(dKs); cudaFree(dV); cudaFree(dVS); cudaFree(dC); cudaFree(dmu);
    if (rc == AGP_ERR_CUDA && ctx->err.empty()) ctx->err = "CUDA failure in predict_f_cov";
    return rc;
  }
  // first iteration on a new batch (onlinetraining.jl:75-104): the expectation gradients come from the PREVIOUS model's local
  // updates on this batch; kernel matrices, natural gradient and global update with the current inducing set
  int step_with_gradients(const int64_t* idx, int B, int base, const double* gmu_h, const double* gS_h) override {
    if (!gmu_h || !gS_h) BAD("null gradients");
    if (peer || Ql != Qg) BAD("step_with_gradients: latent-sharded models are not supported");
    CKS(step_moments(idx, B, base, false));
    for (int q = 0; q < Ql; ++q) {
      CK(cudaMemcpyAsync(gmu + (size_t)q * ldB, gmu_h + (size_t)q * B, (size_t)B * 8, cudaMemcpyHostToDevice, st()));
      CK(cudaMemcpyAsync(gS + (size_t)q * ldB, gS_h + (size_t)q * B, (size_t)B * 8, cudaMemcpyHostToDevice, st()));
    }
    CK(cudaStreamSynchronize(st()));
    skip_lik = true;
    return step_update(1.0);
  }

  // local_updates! launch: multi-output models take the two-dimensional pair of kernels (AGP_LIK_1D=1: the one-thread-per-sample kernel)
  void launch_lik(const LikParams& lp, bool chain) {
    const int B = lp.B;
    static const bool one_d = getenv("AGP_LIK_1D") != nullptr;
    if (lp.model_kind == AGP_MODEL_MOSVGP && !one_d) {
      if (chain) launch_chain(lik_mo_task_kernel, dim3((B + 127) / 128, (lp.n_task + LIK_MO_TPT - 1) / LIK_MO_TPT), dim3(128), 0, lp);
      else lik_mo_task_kernel<<<dim3((B + 127) / 128, (lp.n_task + LIK_MO_TPT - 1) / LIK_MO_TPT), 128, 0, st()>>>(lp);
      ++launches;
      if (lp.update) {
        if (chain) launch_chain(lik_mo_grad_kernel, dim3((B + 127) / 128, (lp.n_latent_local + LIK_MO_TPT - 1) / LIK_MO_TPT), dim3(128), 0, lp);
        else lik_mo_grad_kernel<<<dim3((B + 127) / 128, (lp.n_latent_local + LIK_MO_TPT - 1) / LIK_MO_TPT), 128, 0, st()>>>(lp);
        ++launches;
      }
      return;
    }
    if (chain) launch_chain(lik_update_kernel, dim3((B + 127) / 128), dim3(128), 0, lp);
    else lik_update_kernel<<<(B + 127) / 128, 128, 0, st()>>>(lp);
    ++launches;
  }
  bool can_fuse_lik() const {
    return !noise_any && !is_vgp && prec == AGP_PREC_TF32X3 && model_kind == AGP_MODEL_SVGP && Qg == 1 && Ql == 1 && !need_lam && !peer && !a_opt && !prof &&
           !getenv("AGP_NO_FUSE_LIK");
  }
  LikParams lik_params(int B, bool from_batch, int update) {
    LikParams p{};
    p.model_kind = model_kind; p.n_task = nT; p.Q = Qg; p.B = B; p.ldB = ldB; p.latent_begin = qbeg; p.n_latent_local = Ql;
    p.lik_kind = d_lik_kind; p.p0 = d_p0; p.p1 = d_p1; p.A = d_A; p.mean_f = mean_f; p.var_f = var_f;
    p.y_all = y_all; p.n = n; p.ycls_all = ycls_all; p.idx = from_batch ? nullptr : idx_cur;
    p.yb = yb; p.ycls = ycls; p.c = lc; p.theta = ltheta; p.gamma = lgamma_; p.alpha = lalpha;
    p.tmu = tmu; p.tvar = tvar; p.gm = gm; p.gs = gs; p.gmu = gmu; p.gS = gS; p.update = update;
    p.xepoch = peer ? d_xepoch : nullptr; p.par_stride = par_stride;
    p.noise_opt = noise_any ? d_noise_opt : nullptr; p.noise_state = d_noise_state; p.n_eta = n_eta; p.n_b1 = n_b1; p.n_b2 = n_b2; p.n_eps = n_eps;
    p.lam = d_lam; p.lamacc = d_lamacc; p.qnodes = d_qnodes; p.qweights = d_qw; p.nq = nq; p.need_reduce = (need_lam || noise_any) ? 1 : 0;
    return p;
  }

  int step_moments(const int64_t* idx, int B, int base, bool from_batch) override {
    h_mu_valid = false;
    if (!have_K) { ctx->err = "agp_refresh_K must be called before a step"; return AGP_ERR_STATE; }
    if (!from_batch && !have_data) { ctx->err = "upload data first"; return AGP_ERR_STATE; }
    if (B < 1 || B > Bcap) BAD("The size of mini-batch is incorrect (negative or bigger than the batch capacity)");
    // tf32x3: a ragged B is padded to the next multiple of 128 on the host index list path and for host-row batches (rowsK); the
    // resident list pool and the pipelined asynchronous batches keep the multiple-of-128 rule
    if (prec == AGP_PREC_TF32X3 && (B % 128) && !from_batch && !idx) BAD("TF32X3 precision needs B % 128 == 0 for resident minibatch lists (host index lists and host-row batches are padded internally)");
    if (!from_batch) CKS(prep_idx(idx, B, base));
    curB = B; cur_from_batch = from_batch; kernel_matrices_stale = false; prefetched = false; stats_early = false;
    fuse_lik_next = fuse_in_step_moments && can_fuse_lik(); fuse_from_batch = from_batch; lik_fused = false;
    int sm = moments_impl(from_batch, B, true);
    fuse_lik_next = false;
    return sm;
  }
  bool fuse_in_step_moments = false;   // true only when step_update follows immediately (step_full / step_batch, not the sharded API)

  int step_update(double rho) override {
    CKS(step_update_a(rho));
    return step_update_b(rho);
  }
  // local updates + everything that still reads V (V^T g, Gram product)
  bool skip_lik = false;   // step_with_gradients: the expectation gradients were supplied by the caller
  int step_update_a(double rho) {
    if (curB < 1) { ctx->err = "no minibatch in flight"; return AGP_ERR_STATE; }
    if (!skip_lik) CKS(local_updates_part());
    skip_lik = false; lik_fused = false;
    return natgrad_products(rho);
  }
  // local_updates! + the expectation gradients of the minibatch in flight (likelihood/*.jl); `update_A!` first when it is optimised
  int local_updates_only() override {
    if (curB < 1) { ctx->err = "no minibatch in flight"; return AGP_ERR_STATE; }
    return local_updates_part();
  }
  int local_updates_part() {
    const int B = curB;
    ph_begin(PH_LIK);
    if (a_opt) {   // update_A! precedes variational_updates (training/training.jl:153-158)
      LikParams lp0 = lik_params(B, cur_from_batch, 0);
      launch_lik(lp0, false);                                                           // labels of this batch + task means with the current A
      update_A_grad_kernel<<<nT * Qg, 256, 0, st()>>>(lik_params(B, true, 0), d_gradA);
      update_A_adam_kernel<<<nT, 32 * ((Qg + 31) / 32), 0, st()>>>(d_A, d_gradA, d_Amt, d_Avt, d_Abt, Qg, a_eta, a_b1, a_b2, a_eps);
      launches += 2;
    }
    if (need_quad && nq < 1) { ph_end(); ctx->err = "agp_set_quadrature must be called before a Poisson step"; return AGP_ERR_STATE; }
    if (!lik_fused) launch_lik(lik_params(B, cur_from_batch, 1), true);
    lik_fused = false;
    if (need_lam || noise_any) {  // lambda / noise re-estimation closes local_updates! (poisson.jl:80, heteroscedastic.jl:98, gaussian.jl:62-70)
      LikParams lp = lik_params(B, true, 1);
      launch_chain(lik_lambda_kernel, dim3(1), dim3(std::max(32, (int)rup(nT, 32))), 0, lp);
      ++launches;
      if (is_het) { launch_chain(hetero_grad_kernel, dim3((B + 127) / 128), dim3(128), 0, lp); ++launches; }
      if (noise_any) {   // theta = 1 / sigma^2 and the Gaussian gradients with the NEW noise (gaussian.jl:70-80)
        LikParams lp2 = lik_params(B, true, 2);
        launch_lik(lp2, true);
      }
    }
    ph_end();
    return AGP_OK;
  }
  int natgrad_products(double rho) {
    const int B = curB, Bk = rowsK(curB);
    const bool grp = use_groups && prec == AGP_PREC_TF32X3;
    if (Bk != B)    // padding rows of a ragged minibatch: zero weights, so they drop out of V^T grad_mu and of the Gram product
      for (int q = 0; q < Ql; ++q) {
        CK(cudaMemsetAsync(gS + (size_t)q * ldB + B, 0, (size_t)(Bk - B) * sizeof(double), st()));
        CK(cudaMemsetAsync(gmu + (size_t)q * ldB + B, 0, (size_t)(Bk - B) * sizeof(double), st()));
      }
    if (grp && grp_gram_tn) {
      // one launch: every owned latent's Gram partials and V^T grad_mu straight from V (scaling / transposition in the worker threads)
      ph_begin(PH_GRAM);
      int ns = n_split;
      umma_set_pdl(tail_pdl && !prof);
      CKS(umma_gram_tn_grouped(ctx_err(), grpGT, lat[0].um, rho, Bk, mk, &ns, st()));
      umma_set_pdl(false);
      ++launches;
      for (auto& L : lat) L.gram_splits = ns;
      ph_end();
      return AGP_OK;
    }
    if (grp) fan_begin();
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      if (grp) fan_select(q);
      ph_begin(PH_GRADMU);
      // tf32x3: V^T grad_mu is accumulated by the scale-transpose kernel into v1, which combine_kernel clears after use
      if (prec != AGP_PREC_TF32X3) CK(cudaMemsetAsync(L.v1, 0, m * sizeof(double), st()));
      if (prec != AGP_PREC_TF32X3)
      { int rpb = std::max(64, (int)rup((B + 15) / 16, 8));
        gemv_t_kernel<T><<<dim3((m + 31) / 32, (B + rpb - 1) / rpb), dim3(32, 8), 0, st()>>>(L.V, ldm, gmu + (size_t)q * ldB, B, m, rpb, L.v1); }
      ++launches;
      ph_end();
      // single launches: the Gram kernel reads V itself (scaling, transposition and V^T grad_mu inside its worker threads); grouped and
      // split launches keep the scale-transpose pass
      const bool gram_tn = prec == AGP_PREC_TF32X3 && !grp && L.um.gram_tn && !split_gram_ok();
      if (prec == AGP_PREC_TF32X3 && !gram_tn) {
        ph_begin(PH_SPLIT);
        umma_set_pdl(tail_pdl && !prof && !fan_active);
        CKS(umma_scale_transpose(ctx_err(), L.um, (const float*)(const void*)L.V, gS + (size_t)q * ldB, rho, gmu + (size_t)q * ldB, L.v1, Bk, mk, st()));
        ++launches;
        ph_end();
      }
      if (grp) { umma_set_pdl(false); continue; }     // the Gram products of all owned latents follow in one grouped launch
      if (split_gram_ok()) {
        // side stream: tiles 1.. of the Gram product + their natural-parameter update; main stream: tile (0, 0) (step_update_b goes on
        // with its share of the update and the tail's first kernel, and joins before the block steps)
        ph_begin(PH_GRAM);
        const int sms = 148, ut = (m / 128) * (m / 128 + 1) / 2;
        split_SA = umma_gram_splits(Bk, std::min(32, n_split));
        split_SB = umma_gram_splits(Bk, std::max(1, std::min(n_split, (sms - split_SA) / (ut - 1))));
        umma_set_pdl(false);
        CK(cudaEventRecord(ev_g0, ctx->stream));
        CK(cudaStreamWaitEvent(side, ev_g0, 0));
        cudaStream_t saved = cur_stream;
        cur_stream = side;
        int sg = umma_gram_part(ctx_err(), L.um, (float*)(void*)L.Gpart, Bk, mk, 1, split_SB, sms - split_SA, st());
        ++launches;
        if (sg == AGP_OK) { launch_combine(L, rho, split_SB, 2); }
        if (sg == AGP_OK && cudaEventRecord(ev_gb, side) != cudaSuccess) sg = AGP_ERR_CUDA;
        cur_stream = saved;
        CKS(sg);
        umma_set_pdl(tail_pdl && !prof);      // scale_transpose -> tile (0, 0): programmatic edge on the main stream
        int sa_ = umma_gram_part(ctx_err(), L.um, (float*)(void*)L.Gpart, Bk, mk, 0, split_SA, split_SA, st());
        umma_set_pdl(false);
        CKS(sa_);
        ++launches;
        L.gram_splits = split_SA;
        split_gram_now = true;
        ph_end();
        continue;
      }
      ph_begin(PH_GRAM);
      int ns = n_split;
      if (gram_tn) {
        umma_set_pdl(tail_pdl && !prof && !fan_active);
        int sg = umma_gram_tn(ctx_err(), L.um, (float*)(void*)L.Gpart, gS + (size_t)q * ldB, rho, gmu + (size_t)q * ldB, L.v1, Bk, mk, &ns, st());
        umma_set_pdl(false);
        CKS(sg);
        ++launches;
      } else if (prec == AGP_PREC_TF32X3) {
        CKS(umma_gram(ctx_err(), L.um, (float*)(void*)L.Gpart, Bk, mk, &ns, st()));
        umma_set_pdl(false);
        ++launches;
      } else {
        GemmParams<T> g{};  // rho * V^T diag(grad_Sigma) V  (functions/utils.jl:70-72, whitened), split over the minibatch
        g.A = L.V; g.lda = ldm; g.B = L.V; g.ldb = ldm; g.C = L.Gpart; g.ldc = ldm; g.M = m; g.N = m; g.K = B;
        g.k_scale = gS + (size_t)q * ldB; g.k_scale_mul = rho; g.k_chunk = k_chunk; g.zs_c = (int64_t)mk * ldm; g.alpha = 1.0;
        ns = (B + k_chunk - 1) / k_chunk;
        gemm_simt_launch<T, true, true, EPI_PLAIN>(g, ns, st());
        ++launches;
      }
      L.gram_splits = ns;
      ph_end();
    }
    if (grp) {
      fan_end();
      ph_begin(PH_GRAM);
      int ns = n_split;
      umma_set_pdl(tail_pdl && !prof);
      CKS(umma_gram_grouped(ctx_err(), grpG, lat[0].um, Bk, mk, &ns, st()));
      umma_set_pdl(false);
      ++launches;
      for (auto& L : lat) L.gram_splits = ns;
      ph_end();
    }
    CK(cudaGetLastError());
    return AGP_OK;
  }
  // combine_kernel: split-K reduction of the Gram partials + natural-parameter update; blk: 0 = whole matrix, 1 / 2 = split Gram parts
  // the tail's first kernel starts on tile (0, 0) while combine_kernel is still running (AGP_EARLY_POTF2=0 disables)
  int* d_tile0_flag = nullptr;
  bool early_potf2 = true, tile0_signal_now = false;
  TailParams combine_params(Latent& L, double rho, int ns, int blk) {
    TailParams tp{};
    tp.m = m; tp.mp = mp; tp.ld = mp; tp.n_split = ns; tp.gpart_stride = (int64_t)mk * ldm; tp.gpart_ld = ldm;
    tp.g_mirrored = (prec == AGP_PREC_TF32X3) ? 1 : 0;
    tp.v1 = L.v1; tp.mu0v = L.mu0v; tp.eta1 = L.eta1v; tp.eta2 = L.eta2v; tp.P = L.P;
    tp.counters = counters; tp.stochastic = stochastic; tp.rm_kappa = rm_kappa; tp.rm_tau = rm_tau; tp.rho = rho;
    tp.logdet = L.logdetP; tp.status = status;
    tp.lr = d_lr; tp.v1_zero = (prec == AGP_PREC_TF32X3) ? L.v1 : nullptr;
    tp.eta1_off = L.online ? L.on_c1v : nullptr; tp.eta2_off = L.online ? L.on_C2v : nullptr;
    tp.blk_mode = blk;
    return tp;
  }
  void launch_combine(Latent& L, double rho, int ns, int blk) {
    TailParams tp = combine_params(L, rho, ns, blk);
    const bool c4 = blk == 0 && combine4_on && prec == AGP_PREC_TF32X3 && mp == m && m % 4 == 0 && ldm % 4 == 0 && tp.gpart_stride % 4 == 0;
    tile0_signal_now = tile0_signal_now && blk == 0 && !c4 && d_tile0_flag != nullptr;
    if (tile0_signal_now) tp.tile0_flag = d_tile0_flag;
    // whole matrix on the tf32x3 path: four columns per thread (combine4_kernel); the summation order of the slices is the scalar kernel's
    if (blk == 0 && combine4_on && prec == AGP_PREC_TF32X3 && mp == m && m % 4 == 0 && ldm % 4 == 0 && tp.gpart_stride % 4 == 0)
      launch_chain(combine4_kernel, dim3((m / 4 + 127) / 128, m), dim3(128), 0, tp, (const float*)(const void*)L.Gpart);
    else
    launch_chain(combine_kernel<T>, blk == 1 ? dim3(1, 128) : grid_mp(), dim3(128), 0, tp, (const T*)L.Gpart);
    ++launches;
  }
  bool split_gram_ok() const {
    return split_gram_on && allow_split_gram && pipeline && !prof && prec == AGP_PREC_TF32X3 && Ql == 1 && Qg == 1 && tail_variant == 2 && !ns_tail_now &&
           !is_vgp && !peer && mp == m && m >= 256 && !lat[0].um.v2 && !lat[0].online && n_split >= 16;
  }
  // natural-parameter update + the m x m tail of every owned latent
  int step_update_b(double rho) {
    const bool fan3 = !ns_tail_now && tail_variant == 3 && Ql >= 2;
    // several latents per launch (combine_batched_kernel / x_finalize_batched_kernel): the grouped tcgen05 path of a multi-latent model
    // whose tail is the persistent kernel, no online terms
    bool batched = fan3 && batch_small && use_groups && prec == AGP_PREC_TF32X3 && !split_gram_now;
    for (int q = 0; q < Ql && batched; ++q) batched = !lat[q].online && lat[q].gram_splits == lat[0].gram_splits;
    if (batched) {
      fan_begin();
      ph_begin(PH_COMBINE);
      for (int q0 = 0, nb_ = 0; q0 < Ql; q0 += SMALL_NB, ++nb_) {
        const int cnt = std::min(SMALL_NB, Ql - q0);
        TailParams tp = combine_params(lat[q0], rho, lat[q0].gram_splits, 0);
        CombineBatch bt{};
        for (int z = 0; z < cnt; ++z) {
          Latent& L = lat[q0 + z];
          bt.v1[z] = L.v1; bt.mu0v[z] = L.mu0v; bt.eta1[z] = L.eta1v; bt.eta2[z] = L.eta2v; bt.P[z] = L.P; bt.v1_zero[z] = L.v1; bt.logdet[z] = L.logdetP;
          bt.G[z] = (const float*)(const void*)L.Gpart;
        }
        fan_select(nb_);
        const dim3 g2 = grid_mp();
        launch_chain(combine_batched_kernel, dim3(g2.x, g2.y, cnt), dim3(128), 0, tp, bt);
        ++launches;
      }
      ph_end();
      fan_end();
      chol_inv_many(0, Ql);
      fan_begin();
      ph_begin(PH_FINAL);
      for (int q0 = 0, nb_ = 0; q0 < Ql; q0 += FINALIZE_NB, ++nb_) {
        const int cnt = std::min(FINALIZE_NB, Ql - q0);
        const bool last = q0 + cnt == Ql;
        FinalizeBatch<T> bt{};
        for (int z = 0; z < cnt; ++z) {
          Latent& L = lat[q0 + z];
          bt.X[z] = L.Xv; bt.eta1v[z] = L.eta1v; bt.shadow[z] = L.Xv_T; bt.hi[z] = umma_split_ptr(L.um, UM_X, 0); bt.lo[z] = umma_split_ptr(L.um, UM_X, 1);
          bt.tvec[z] = L.tvec;
          L.muv_valid = false; L.factor_valid = true; L.ns_seeded = false;
        }
        fan_select(nb_);
        launch_chain(x_finalize_batched_kernel<T>, dim3(m, cnt), dim3(128), 0, bt, (int64_t)mp, m, (int64_t)ldm,
                     (last && stochastic && !(fixed_lr > 0.0)) ? d_lr : (double*)nullptr, counters, rm_kappa, rm_tau, last ? 1 : 0);
        ++launches;
      }
      ph_end();
      fan_end();
      CK(cudaGetLastError());
      have_step = true;
      return AGP_OK;
    }
    if (fan3) fan_begin();
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      const int ns = L.gram_splits;
      if (fan3) fan_select(q);
      ph_begin(PH_COMBINE);
      tile0_signal_now = early_potf2 && tail_variant == 2 && !ns_tail_now && !split_gram_now && !prof;
      launch_combine(L, rho, ns, split_gram_now ? 1 : 0);
      ph_end();
      if (ns_tail_now || tail_variant == 3) tile0_signal_now = false;
      if (ns_tail_now) CKS(eta_to_moments_ns(L));
      else if (tail_variant != 3) {
        CKS(eta_to_moments(L, q == Ql - 1));   // the last latent's finalize kernel also prepares the next step size
        L.factor_valid = true; L.ns_seeded = false;
      }
    }
    if (fan3) fan_end();
    if (!ns_tail_now && tail_variant == 3) {   // every owned latent's Cholesky + inverse factor in one persistent launch
      chol_inv_many(0, Ql);
      if (fan3) fan_begin();
      for (int q = 0; q < Ql; ++q) {
        if (fan3) fan_select(q);
        int sf = finalize_factor(lat[q], q == Ql - 1);
        if (sf != AGP_OK) { fan_end(); return sf; }
        lat[q].factor_valid = true; lat[q].ns_seeded = false;
      }
      if (fan3) fan_end();
    }
    // (the counters are bumped by the last latent's finalize kernel)
    CK(cudaGetLastError());
    have_step = true;
    return AGP_OK;
  }

  // agp_tail2.cuh kernels are chained with programmatic dependent launch: block step k+1 becomes resident while step k
  // runs and waits in griddepcontrol.wait, which hides most of the ~2 us launch gap between the 9 dependent launches
  // kernels of the per-step critical chain whose in-stream predecessor is a kernel: launched with the programmatic
  // stream serialization attribute (they all start with pdl_prologue())
  template <typename K, typename... Args>
  void launch_chain(K kern, dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (tail_pdl && !prof && !chain_break && !fan_active) ? 1 : 0;
    chain_break = false;
    cudaLaunchKernelEx(&cfg, kern, args...);
  }
  template <typename K>
  void launch_tail2(K kern, int grid, const TailStepParams& tp) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TAIL_THREADS); cfg.dynamicSmemBytes = TAIL2_SMEM; cfg.stream = st();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = tail_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, tp);
  }

  // fused blocked Cholesky + inverse of the factor: P (lower tiles, destroyed) -> Xv = chol(P)^-1, for latents [q0, q0 + count)
  // Block steps of the m x m tail: the 64-wide blocks that hold rows of the logical matrix.  mp is padded to a power-of-two number of
  // blocks for the recursive inverse of refresh_K (spd_inverse); the padding blocks of P_v are identity and stay out of the tail (their
  // factor is the identity, their pivots add log 1 to the log-determinant, and nothing reads the rows of X beyond m): m = 640 runs 10
  // block steps instead of 16.
  int nblk_tail() const { return (m + TNB - 1) / TNB; }
  // tail_variant 3: one persistent launch per group of latents (agp_tail3.cuh); otherwise nblk + 1 launches per latent
  int tail3_ctas(int count) const {      // CTAs the persistent tail of `count` latents occupies (first launch)
    const int nblk = nblk_tail();
    if (nblk == 1) return std::min(count, t3_sm_budget);
    const int nl = std::min(count, std::max(1, t3_sm_budget / 4));
    return nl * std::max(2, std::min(tail3_team(nblk), t3_sm_budget / nl));
  }
  // CTAs per latent: the chain CTA + enough helpers that no helper has more than ~2 tile tasks per block step (a helper task takes
  // about half a chain step; measured with profiles/microbench/tail3_test.cu: 20 and 35 CTAs give the same time at m = 512)
  static int tail3_team(int nblk) { return (tail3_max_tasks(nblk) * 13 + 19) / 20 + 1; }
  void chol_inv_many(int q0, int count) {
    if (tail_variant != 3) { for (int q = q0; q < q0 + count; ++q) chol_inv(lat[q]); return; }
    ph_begin(PH_CHOL);
    const int nblk = nblk_tail();
    const int gcap = tail3_team(nblk);
    int done = 0;
    while (done < count) {
      const int nl = std::min(count - done, std::max(1, t3_sm_budget / (nblk == 1 ? 1 : 4)));   // at least 4 CTAs per latent
      const int G = nblk == 1 ? 1 : std::max(2, std::min(gcap, t3_sm_budget / nl));
      Tail3Params tp{};
      tp.lat = d_t3lat + q0 + done; tp.nlat = nl; tp.G = G; tp.nblk = nblk; tp.ld = mp; tp.status = status;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nl * G); cfg.blockDim = dim3(TAIL_THREADS); cfg.dynamicSmemBytes = TAIL3_SMEM; cfg.stream = st();
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = (tail_pdl && !prof) ? 1 : 0;
      cudaLaunchKernelEx(&cfg, tail3_kernel, tp);
      ++launches;
      done += nl;
    }
    ph_end();
  }
  void chol_inv(Latent& L) {
    if (tail_variant == 3) { chol_inv_many((int)(&L - lat.data()), 1); return; }
    ph_begin(PH_CHOL);
    TailStepParams tp{};
    tp.P = L.P; tp.W = L.W; tp.Xout = L.Xv; tp.Dinv = L.Dinv; tp.ld = mp; tp.nblk = nblk_tail(); tp.logdet = L.logdetP; tp.status = status;
    tp.early_flag = (tile0_signal_now && tail_variant == 2) ? d_tile0_flag : nullptr;   // the combine launch just before this one signals
    tile0_signal_now = false;
    if (tail_variant == 0) tail_potf2_first_kernel<0><<<1, TAIL_THREADS, TAIL_SMEM, st()>>>(tp);
    else if (chain_variant == 0) launch_tail2(tail2_potf2_first_kernel<0, 0>, 1, tp);
    else launch_tail2(tail2_potf2_first_kernel<0, 1>, 1, tp);
    ++launches;
    if (split_gram_now) { cudaStreamWaitEvent(st(), ev_gb, 0); split_gram_now = false; }   // the rest of P_v (side stream) is needed from the first block step on
    for (int k = 0; k < tp.nblk; ++k) {
      int r = tp.nblk - 1 - k;
      int tiles = r * (r + 1) / 2 + r * (k + 1) + k;
      if (tiles == 0) continue;
      tp.k = k;
      if (tail_variant == 0) tail_step_kernel<0><<<tiles, TAIL_THREADS, TAIL_SMEM, st()>>>(tp);
      else if (chain_variant == 0) launch_tail2(tail2_step_kernel<0>, tiles, tp);
      else launch_tail2(tail2_step_kernel<1>, tiles, tp);
      ++launches;
      // rows [64 k, 64 k + 64) of X are final now; every second block completes a 128-row N tile of the next step's statistics
      if (early_now && (k & 1) && (k >> 1) + 1 < m / 128) cudaEventRecord(ev_blk[k >> 1], st());
    }
    ph_end();
  }

  // global_update!(gp) (inference/inference.jl:25-28) in the whitened basis: Sigma_v = inv(P_v) = X^T X is kept in
  // factored form (X = chol(P_v)^-1), mu_v = Sigma_v eta1_v = X^T (X eta1_v)
  int eta_to_moments(Latent& L, bool in_step = false) {
    chol_inv(L);
    return finalize_factor(L, in_step);
  }
  // after the tail: fp32 shadow (+ TF32 split) of X, t = X eta1_v; the last latent's kernel also prepares the next step size
  int finalize_factor(Latent& L, bool in_step) {
    ph_begin(PH_FINAL);
    float* hi = umma_split_ptr(L.um, UM_X, 0);   // non-null only with the opt-in v2 GEMM: X leaves this kernel pre-split
    float* lo = umma_split_ptr(L.um, UM_X, 1);
    const int row0 = early_now ? (m / 128 - 1) * 128 : 0;    // the earlier row blocks are finalised on the side stream (step_pool)
    launch_chain(x_finalize_kernel<T>, dim3(m - row0), dim3(128), 0, (const double*)L.Xv, (int64_t)mp, m, (const double*)L.eta1v, L.Xv_T, (int64_t)ldm, hi, lo,
                 L.tvec, (in_step && stochastic && !(fixed_lr > 0.0)) ? d_lr : (double*)nullptr, counters, rm_kappa, rm_tau, in_step ? 1 : 0, row0);
    ++launches;
    ph_end();
    L.muv_valid = false;
    return AGP_OK;
  }
  // ---- experimental Newton-Schulz tail (AGP_TAIL_NS=<iterations per step>, needs AGP_UMMA_V2) ----
  int ns_iters = 0, ns_after = 8;
  double ns_tol = 0.3;
  int64_t h_steps = 0;          // steps taken since the posterior was last (re)initialised: the first ns_after steps keep the Cholesky tail
  bool ns_tail_now = false;     // this step's tail is the refinement (decided by ns_begin_step, constant during a capture)
  int ns_key = 0, g_nskey = -1, g_nskey_b = -1;   // bit 0: statistics against ns.Y(), bit 1: refinement tail -- part of the graph keys
  bool ns_eligible() const {
    return prec == AGP_PREC_TF32X3 && stochastic && model_kind == AGP_MODEL_SVGP && Ql == 1 && Qg == 1 && !is_vgp && !peer && !prof &&
           !lat.empty() && lat[0].um.v2 != 0 && mp == m;
  }
  // entry of every whole step (step_full / step_batch), outside any capture
  int ns_begin_step() {
    ns_tail_now = false; ns_key = 0;
    if (ns_iters <= 0) return AGP_OK;
    Latent& L = lat[0];
    if (ns_eligible() && h_steps >= ns_after) {
      if (!L.ns_seeded) {        // Y0 = Sigma_v of the current natural parameters = X^T X
        CKS(ensure_factor(L));
        dgemm(true, true, L.Xv, L.Xv, L.X, 1.0, 0.0);
        shadow_kernel<float><<<dim3((m + 127) / 128, m), 128, 0, st()>>>(L.X, mp, m, L.ns.Y(), (int64_t)L.ns.ldm);
        ++launches;
        L.ns_seeded = true;
      }
      ns_tail_now = true;
    }
    ns_key = (L.factor_valid ? 0 : 1) | (ns_tail_now ? 2 : 0);
    return AGP_OK;
  }
  void ns_end_step() {
    ++h_steps;
    if (ns_iters <= 0) return;
    Latent& L = lat[0];
    if (ns_tail_now) { L.factor_valid = false; L.ns_seeded = true; }
    else { L.factor_valid = true; L.ns_seeded = false; }
    ns_tail_now = false;
  }
  // the refinement tail: P_v = -2 eta2_v (written by combine_kernel, left intact) -> ns.Y() ~ P_v^-1
  int eta_to_moments_ns(Latent& L) {
    ph_begin(PH_CHOL);
    CKS(umma_ns_iterate(ctx_err(), L.ns, ns_iters, 3, L.P, (int64_t)mp, st()));
    launches += 2 * ns_iters + 1;
    ph_end();
    ph_begin(PH_FINAL);
    // accepted when |I - Y P|_F < ns_tol (0.3) at the START of the last iteration: |.|_2 <= |.|_F, so the map is still contracting
    // and the last pass squares the error (typically |.|_2 ~ |.|_F / 20 at m = 512: final error < 1e-3); a diverging run shows >= 1
    launch_chain(ns_finalize_kernel, dim3(1), dim3(32), 0, (const double*)L.ns.resid, ns_iters, ns_tol * ns_tol, status,
                 (stochastic && !(fixed_lr > 0.0)) ? d_lr : (double*)nullptr, counters, rm_kappa, rm_tau, 1);
    ++launches;
    ph_end();
    L.muv_valid = false; L.factor_valid = false; L.ns_seeded = true;
    return AGP_OK;
  }
  // everything off the hot step that reads the factor X (getters, ELBO, hyper-gradients, prediction) goes through here
  int ensure_factor(Latent& L) {
    if (L.factor_valid) return AGP_OK;
    CK(cudaMemsetAsync(L.logdetP, 0, sizeof(double), st()));
    CKS(eta_to_moments(L));        // Cholesky tail on the intact P_v; no counter bump
    L.factor_valid = true;
    return AGP_OK;
  }
  int ensure_factors() {
    for (auto& L : lat) CKS(ensure_factor(L));
    return AGP_OK;
  }

  // mu_v = X^T t (only getters / the ELBO need it)
  void ensure_muv(Latent& L) {
    if (L.muv_valid) return;
    if (!L.factor_valid && L.ns_seeded) {      // mu_v = Sigma_v eta1_v with the refined covariance
      cudaMemsetAsync(L.muv, 0, m * sizeof(double), st());
      gemv_t_kernel<float><<<dim3((m + 31) / 32, (m + 63) / 64), dim3(32, 8), 0, st()>>>(L.ns.Y(), (int64_t)L.ns.ldm, L.eta1v, m, m, 64, L.muv);
      ++launches;
      L.muv_valid = true;
      return;
    }
    cudaMemsetAsync(L.muv, 0, m * sizeof(double), st());
    gemv_t_kernel<double><<<dim3((m + 31) / 32, (m + 63) / 64), dim3(32, 8), 0, st()>>>(L.Xv, mp, L.tvec, m, m, 64, L.muv);
    ++launches;
    L.muv_valid = true;
  }

  void drop_graph() {
    if (gexec) { cudaGraphExecDestroy(gexec); gexec = nullptr; }
    if (gexec_b) { cudaGraphExecDestroy(gexec_b); gexec_b = nullptr; }
    drop_graph_p();
    gB = -1;
  }

  // one resident-list step, software-pipelined (see `side` above)
  int step_pool(int B, double rho) {
    h_mu_valid = false;
    if (!have_K) { ctx->err = "agp_refresh_K must be called before a step"; return AGP_ERR_STATE; }
    if (!have_data) { ctx->err = "upload data first"; return AGP_ERR_STATE; }
    if (B < 1 || B > Bcap) BAD("The size of mini-batch is incorrect (negative or bigger than the batch capacity)");
    if (prec == AGP_PREC_TF32X3 && (B % 128)) BAD("TF32X3 precision needs B % 128 == 0");
    const bool pipe = pipeline && !prof;
    if (!prefetched || curB != B) {          // cold start: kernel matrices of the minibatch at the cursor, on the main stream
      CKS(prep_idx(nullptr, B, 0, 0));
      CKS(moments_impl(false, B, true, 1));
    }
    const bool use_early = stats_early && curB == B;        // (stats_early implies prefetched)
    curB = B; cur_from_batch = false; kernel_matrices_stale = false; prefetched = false; stats_early = false;
    fuse_lik_next = can_fuse_lik(); fuse_from_batch = false; lik_fused = false;
    stats_use_early = use_early;
    int sm2 = moments_impl(false, B, true, 2);    // V X^T, row statistics (needs the posterior of the previous step)
    stats_use_early = false;
    fuse_lik_next = false;
    CKS(sm2);
    allow_split_gram = pipe;
    int sa = step_update_a(rho);             // local updates, V^T g, Gram product: last readers of V / idx_cur
    allow_split_gram = false;
    CKS(sa);
    const bool early = pipe && early_ok();
    if (pipe) {
      CK(cudaEventRecord(ev_fork, ctx->stream));
      CK(cudaStreamWaitEvent(side, ev_fork, 0));
      cur_stream = side;
      int s = AGP_OK;
      if (cudaMemcpyAsync(idx_prev, idx_cur, (size_t)B * 8, cudaMemcpyDeviceToDevice, st()) != cudaSuccess) s = AGP_ERR_CUDA;
      if (s == AGP_OK) s = prep_idx(nullptr, B, 0, 1);          // the cursor is bumped at the end of this step
      // the persistent tail holds one SM per CTA for its whole duration: keep the prefetch GEMM's persistent grid off those SMs
      // (single latent only: with several latents the tail is a small share of the step and wants every SM itself)
      if (tail_variant == 3 && Ql == 1 && !ns_tail_now) umma_set_grid_cap(std::max(32, 148 - tail3_ctas(Ql)));
      if (s == AGP_OK) s = moments_impl(false, B, true, 1);
      umma_set_grid_cap(0);
      if (s == AGP_OK && prec == AGP_PREC_TF32X3 && Ql == 1) {   // clear the next step's V X^T accumulators off the critical chain
        if (cudaMemsetAsync(lat[0].racc + ldB, 0, 2 * ldB * sizeof(double), st()) != cudaSuccess) s = AGP_ERR_CUDA;
        else racc2_precleared = true;
      }
      if (s == AGP_OK && !early && cudaEventRecord(ev_join, side) != cudaSuccess) s = AGP_ERR_CUDA;
      cur_stream = nullptr;
      if (s != AGP_OK) { if (ctx->err.empty()) ctx->err = "CUDA failure while prefetching"; return s; }
    }
    early_now = early;
    int sb = step_update_b(rho);
    early_now = false;
    CKS(sb);
    if (early) {
      // side stream, behind the prefetch: N tile j of the next step's statistics as soon as block step 2j+1 has finished rows
      // [128 j, 128 j + 128) of X (events recorded by chol_inv), preceded by that row block's fp32 shadow and t = X eta1_v entries
      Latent& L = lat[0];
      const int ntn = m / 128;
      cur_stream = side;
      int s = AGP_OK;
      for (int j = 0; j + 1 < ntn && s == AGP_OK; ++j) {
        if (cudaStreamWaitEvent(side, ev_blk[j], 0) != cudaSuccess) { s = AGP_ERR_CUDA; break; }
        float* hi = umma_split_ptr(L.um, UM_X, 0);
        float* lo = umma_split_ptr(L.um, UM_X, 1);
        x_finalize_kernel<T><<<128, 128, 0, st()>>>((const double*)L.Xv, (int64_t)mp, m, (const double*)L.eta1v, L.Xv_T, (int64_t)ldm, hi, lo, L.tvec,
                                                    (double*)nullptr, counters, rm_kappa, rm_tau, 0, 128 * j);
        ++launches;
        UmmaEpilogue ep{};
        ep.mode = UMMA_EPI_STATS_ONLY; ep.acc0 = L.racc + ldB; ep.acc1 = L.racc + 2 * ldB; ep.tvec = L.tvec;
        umma_set_tile_range(j, j);
        s = umma_gemm_nt(ctx_err(), L.um, UM_V, UM_X, (float*)(void*)L.VS, B, m, ep, st());
        umma_set_tile_range(-1, -1);
        ++launches;
      }
      if (s == AGP_OK && cudaEventRecord(ev_join, side) != cudaSuccess) s = AGP_ERR_CUDA;
      cur_stream = nullptr;
      if (s != AGP_OK) { if (ctx->err.empty()) ctx->err = "CUDA failure in the early statistics"; return s; }
    }
    if (pipe) {
      CK(cudaStreamWaitEvent(ctx->stream, ev_join, 0));
      prefetched = true;
      stats_early = early;
      if (early) racc2_precleared = false;   // the accumulators hold the early tiles, not zeros
      kernel_matrices_stale = true;          // Knm / V now belong to the NEXT minibatch
    }
    have_step = true;
    return AGP_OK;
  }
  // the early statistics need: one tf32x3 latent whose tail is the multi-launch chain (row blocks finish launch by launch), at
  // least two N tiles, 64-row tail blocks that pair up into 128-row tiles, no padding of m, and the first-generation GEMM kernel
  bool early_ok() const {
    return early_on && prec == AGP_PREC_TF32X3 && Ql == 1 && Qg == 1 && tail_variant == 2 && !ns_tail_now && !is_vgp && !peer && mp == m && m >= 256 &&
           m / 128 <= 16 && TNB == 64 && lat[0].factor_valid && !lat[0].um.v2;
  }

  // rebuild Knm / V of the minibatch the last step consumed (after a prefetch or a predict_f overwrote them)
  int ensure_current_kernel_matrices() {
    if (!kernel_matrices_stale) return AGP_OK;
    if (!cur_from_batch && prefetched) {
      CK(cudaMemcpyAsync(idx_cur, idx_prev, (size_t)curB * 8, cudaMemcpyDeviceToDevice, st()));
      xx_gather_kernel<T><<<(curB + 255) / 256, 256, 0, st()>>>(idx_cur, curB, xx, xx_cur);
      ++launches;
    }
    prefetched = false; stats_early = false;
    kernel_matrices_stale = false;
    return moments_impl(cur_from_batch, curB, true, 1);
  }

  int step_full(const int64_t* idx, int B, int base, double rho) override {
    CKS(ns_begin_step());
    int rc = step_full_impl(idx, B, base, rho);
    if (rc == AGP_OK) ns_end_step();
    else ns_tail_now = false;
    return rc;
  }
  int step_full_impl(const int64_t* idx, int B, int base, double rho) {
    if (idx) {
      fuse_in_step_moments = true;
      int s1 = step_moments(idx, B, base, false);
      fuse_in_step_moments = false;
      CKS(s1);
      return step_update(rho);
    }
    if (want_graph && !prof && !capturing) {
      const bool need_prime = pipeline && (!prefetched || curB != B || (prec == AGP_PREC_TF32X3 && Ql == 1 && !(racc2_precleared || stats_early)) || (early_ok() && !stats_early));
      if (!gexec || gB != B || grho != rho || g_nskey != ns_key || need_prime) {
        drop_graph();
        if (need_prime) return step_pool(B, rho);  // priming step (brings the pipeline to its steady state); later calls replay the graph
      }
      if (!gexec) {
        cudaGraph_t graph = nullptr;
        int64_t l0 = launches;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        capturing = true;
        int s = step_pool(B, rho);
        capturing = false;
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (s != AGP_OK) { if (graph) cudaGraphDestroy(graph); return s; }
        CK(ce);
        g_launches = launches - l0;
        launches = l0;
        CK(cudaGraphInstantiate(&gexec, graph, 0));
        cudaGraphDestroy(graph);
        gB = B; grho = rho; g_nskey = ns_key;
      }
      CK(cudaGraphLaunch(gexec, ctx->stream));
      launches += g_launches;
      for (auto& L : lat) L.muv_valid = false;
      curB = B; cur_from_batch = false; have_step = true;
      prefetched = pipeline; kernel_matrices_stale = pipeline;
      return AGP_OK;
    }
    return step_pool(B, rho);
  }

  // host-batch step: the host->device copies are issued eagerly (the source pointers change every call), everything after
  // them (row conversion, kernel matrices, moments, local updates, natural gradient, tail) is one CUDA graph per
  // (B, rho, dtype, layout) when agp_use_graph is on
  int batch_compute(int x_dtype, int x_layout, int B, double rho) {
    int64_t sld = x_layout == AGP_LAYOUT_ROWMAJOR ? D : B;
    int bl = (B + 255) / 256;
    if (x_dtype == AGP_DTYPE_F64) convert_rows_kernel<double, T><<<bl, 256, 0, st()>>>((const double*)stage, x_layout, sld, B, D, Xb, Dp, xxb);
    else convert_rows_kernel<float, T><<<bl, 256, 0, st()>>>((const float*)stage, x_layout, sld, B, D, Xb, Dp, xxb);
    ++launches;
    if (rowsK(B) != B) {   // padding rows of a ragged batch: x = 0 (K_nm row k(0, z): finite), zero weights (natgrad_products)
      CK(cudaMemsetAsync(Xb + (size_t)B * Dp, 0, (size_t)(rowsK(B) - B) * Dp * sizeof(T), st()));
      CK(cudaMemsetAsync(xxb + B, 0, (size_t)(rowsK(B) - B) * sizeof(T), st()));
    }
    fuse_in_step_moments = true;
    int s1 = step_moments(nullptr, B, 0, true);
    fuse_in_step_moments = false;
    CKS(s1);
    CKS(step_update(rho));
    {   // canonical mean of latent 0 -> pinned host buffer, in stream order behind the step
      Latent& L = lat[0];
      ensure_muv(L);
      symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Lc, mp, m, L.muv, L.mu0 + mp);
      ++launches;
      CK(cudaMemcpyAsync(h_mu, L.mu0 + mp, m * sizeof(double), cudaMemcpyDeviceToHost, st()));
    }
    return AGP_OK;
  }
  int step_batch(const void* xbh, int x_dtype, int x_layout, const void* const* ybh, int y_kind, int B, double rho) override {
    CKS(ns_begin_step());
    int rc = step_batch_impl(xbh, x_dtype, x_layout, ybh, y_kind, B, rho);
    if (rc == AGP_OK) ns_end_step();
    else ns_tail_now = false;
    return rc;
  }
  // kind = cudaMemcpyDeviceToDevice: the batch already sits in the device pre-staging buffers of step_batch_async
  int step_batch_impl(const void* xbh, int x_dtype, int x_layout, const void* const* ybh, int y_kind, int B, double rho,
                      cudaMemcpyKind kind = cudaMemcpyHostToDevice) {
    if (!xbh || !ybh) BAD("null batch");
    if (B < 1 || B > Bcap) BAD("The size of mini-batch is incorrect (negative or bigger than the batch capacity)");
    if (is_lsm != (y_kind == AGP_Y_CLASS)) BAD("label kind does not match the likelihood");
    if ((x_dtype != 0 && x_dtype != 1) || (x_layout != 0 && x_layout != 1)) BAD("bad dtype/layout");
    if (!have_K) { ctx->err = "agp_refresh_K must be called before a step"; return AGP_ERR_STATE; }
    const size_t es = x_dtype == AGP_DTYPE_F64 ? 8 : 4;
    if (!h_mu) CK(cudaMallocHost((void**)&h_mu, (size_t)mp * sizeof(double)));
    h_mu_valid = false;
    CKS(ensure_stage((size_t)Bcap * D * 8));     // fixed staging address: the captured graph stays valid
    CK(cudaMemcpyAsync(stage, xbh, (size_t)B * D * es, kind, st()));
    if (y_kind == AGP_Y_CLASS) CK(cudaMemcpyAsync(ycls, ybh[0], B * sizeof(int), kind, st()));
    else for (int t = 0; t < nT; ++t) CK(cudaMemcpyAsync(yb + (size_t)t * ldB, ybh[t], B * sizeof(double), kind, st()));
    const int key = x_dtype * 2 + x_layout;
    if (want_graph && !prof && !capturing) {
      if (gexec_b && (gB_b != B || grho_b != rho || gkey_b != key || g_nskey_b != ns_key)) { cudaGraphExecDestroy(gexec_b); gexec_b = nullptr; }
      if (!gexec_b) {
        cudaGraph_t graph = nullptr;
        int64_t l0 = launches;
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
        capturing = true;
        int s = batch_compute(x_dtype, x_layout, B, rho);
        capturing = false;
        cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
        if (s != AGP_OK) { if (graph) cudaGraphDestroy(graph); return s; }
        CK(ce);
        g_launches_b = launches - l0;
        launches = l0;
        CK(cudaGraphInstantiate(&gexec_b, graph, 0));
        cudaGraphDestroy(graph);
        gB_b = B; grho_b = rho; gkey_b = key; g_nskey_b = ns_key;
      }
      CK(cudaGraphLaunch(gexec_b, ctx->stream));
      launches += g_launches_b;
      for (auto& L : lat) L.muv_valid = false;
      lat[0].muv_valid = true;     // the graph recomputed mu_v of latent 0
      curB = B; cur_from_batch = true; kernel_matrices_stale = false; prefetched = false; stats_early = false; have_step = true;
      h_mu_valid = true;
      return AGP_OK;
    }
    CKS(batch_compute(x_dtype, x_layout, B, rho));
    h_mu_valid = true;
    return AGP_OK;
  }

  // ---- host-batch steps without a per-step synchronisation (untested on a GPU yet) ----
  // Two slots: the host->device copy of batch i+1 runs on a copy stream while step i computes; each step leaves its result
  // (canonical mean of latent 0 + the sticky status word) in a pinned slot that agp_result_wait(ticket) reads later.
  void* pre_x[2] = {nullptr, nullptr}; double* pre_y[2] = {nullptr, nullptr}; int* pre_ycls[2] = {nullptr, nullptr};
  double* h_res[2] = {nullptr, nullptr}; int* h_stat = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_xfree[2] = {nullptr, nullptr}, ev_step[2] = {nullptr, nullptr}, ev_vfree = nullptr, ev_p1 = nullptr, ev_res = nullptr;
  cudaStream_t copy_stream = nullptr, res_stream = nullptr;
  int64_t n_tickets = 0;
  int res_pending = -1;    // slot whose result kernels (result stream) the main stream has not been ordered behind yet
  // pipelined variant: two graphs -- [0] stage 1 on the side stream, [1] the rest of the step on the main stream
  cudaGraphExec_t gexec_p[3] = {nullptr, nullptr, nullptr}; int64_t g_launches_p[3] = {0, 0, 0};
  int gB_p = -1, gkey_p = -1; double grho_p = -1;
  bool chain_break = false;   // the next launch_chain call has no kernel predecessor in its graph: no programmatic edge
  void drop_graph_p() {
    for (int i = 0; i < 3; ++i) if (gexec_p[i]) { cudaGraphExecDestroy(gexec_p[i]); gexec_p[i] = nullptr; }
    gB_p = -1;
  }
  int async_init() {
    if (copy_stream) return AGP_OK;
    CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&res_stream, cudaStreamNonBlocking));
    CK(cudaMallocHost((void**)&h_stat, 2 * sizeof(int)));
    h_stat[0] = h_stat[1] = 0;
    CK(cudaEventCreateWithFlags(&ev_vfree, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_p1, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_res, cudaEventDisableTiming));
    for (int s = 0; s < 2; ++s) {
      CK(cudaMalloc(&pre_x[s], (size_t)Bcap * D * 8));
      CK(cudaMalloc((void**)&pre_y[s], (size_t)nT * ldB * sizeof(double)));
      CK(cudaMalloc((void**)&pre_ycls[s], (size_t)ldB * sizeof(int)));
      CK(cudaMallocHost((void**)&h_res[s], (size_t)mp * sizeof(double)));
      CK(cudaEventCreateWithFlags(&ev_h2d[s], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_free[s], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_done[s], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_xfree[s], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&ev_step[s], cudaEventDisableTiming));
    }
    return AGP_OK;
  }
  void join_async() override {
    if (res_pending < 0) return;
    cudaStreamWaitEvent(ctx->stream, ev_done[res_pending], 0);
    res_pending = -1;
  }
  // one piece of the pipelined step on stream s: replayed from its graph when agp_use_graph is on, launched eagerly otherwise
  template <typename F>
  int run_piece(int which, cudaStream_t s, F body) {
    cudaStream_t saved = cur_stream;
    cur_stream = (s == ctx->stream) ? nullptr : s;
    int rc = AGP_OK;
    if (want_graph && !prof) {
      if (!gexec_p[which]) {
        cudaGraph_t graph = nullptr;
        const int64_t l0 = launches;
        cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed);
        if (ce != cudaSuccess) { cur_stream = saved; ctx->err = std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(ce); return AGP_ERR_CUDA; }
        capturing = true;
        rc = body();
        capturing = false; chain_break = false;
        ce = cudaStreamEndCapture(s, &graph);
        if (rc == AGP_OK && ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); rc = AGP_ERR_CUDA; }
        if (rc == AGP_OK) {
          g_launches_p[which] = launches - l0;
          launches = l0;
          ce = cudaGraphInstantiate(&gexec_p[which], graph, 0);
          if (ce != cudaSuccess) { ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); rc = AGP_ERR_CUDA; gexec_p[which] = nullptr; }
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != AGP_OK) { cur_stream = saved; return rc; }
      }
      cudaError_t ce = cudaGraphLaunch(gexec_p[which], s);
      if (ce != cudaSuccess) { ctx->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(ce); rc = AGP_ERR_CUDA; }
      launches += g_launches_p[which];
    } else {
      rc = body();
    }
    cur_stream = saved;
    return rc;
  }
  // Pipelined form (default): the kernel matrices of batch i (row conversion, K_nm, V = K_nm L^-T: they do not depend on the
  // posterior) are built on the side stream while the m x m tail of step i-1 still runs on the main stream -- the host-batch
  // counterpart of step_pool's prefetch -- and the result read-back (mu_v = X^T t, mu = L mu_v, device -> pinned host) runs on a
  // third stream beside the next step's statistics product.  Ordering:
  //   copy stream : wait(x, y slot free) -> H2D x, y -> ev_h2d
  //   side stream : wait(ev_h2d, ev_vfree = step i-1 is done reading V, y) -> x slot -> staging, y slot -> yb -> [graph 0] -> ev_p1
  //   main stream : wait(ev_p1) -> [graph 1: statistics, local updates, Gram | record ev_vfree | wait ev_res = result kernels of step
  //                 i-1 | natural-parameter update, m x m tail] -> ev_step          (the two events are external-event graph nodes)
  //   result strm : wait(ev_step) -> mu -> h_res[slot], status -> h_stat[slot] -> ev_done[slot], ev_res
  int step_batch_async_pipelined(const void* xbh, int x_dtype, int x_layout, const void* const* ybh, int y_kind, int B, double rho, int64_t* ticket) {
    const int slot = (int)(n_tickets & 1);
    const size_t es = x_dtype == AGP_DTYPE_F64 ? 8 : 4;
    const int key = x_dtype * 2 + x_layout;
    if (!have_K) { ctx->err = "agp_refresh_K must be called before a step"; return AGP_ERR_STATE; }
    if (prec == AGP_PREC_TF32X3 && (B % 128)) BAD("TF32X3 precision needs B % 128 == 0");
    CKS(ensure_stage((size_t)Bcap * D * 8));
    if (gB_p != B || grho_p != rho || gkey_p != key) { drop_graph_p(); gB_p = B; grho_p = rho; gkey_p = key; }
    h_mu_valid = false;
    // copy stream
    CK(cudaStreamWaitEvent(copy_stream, ev_free[slot], 0));
    CK(cudaStreamWaitEvent(copy_stream, ev_xfree[slot], 0));
    CK(cudaMemcpyAsync(pre_x[slot], xbh, (size_t)B * D * es, cudaMemcpyHostToDevice, copy_stream));
    if (y_kind == AGP_Y_CLASS) CK(cudaMemcpyAsync(pre_ycls[slot], ybh[0], B * sizeof(int), cudaMemcpyHostToDevice, copy_stream));
    else for (int t = 0; t < nT; ++t) CK(cudaMemcpyAsync(pre_y[slot] + (size_t)t * ldB, ybh[t], B * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
    CK(cudaEventRecord(ev_h2d[slot], copy_stream));
    // side stream: stage 1
    CK(cudaStreamWaitEvent(side, ev_h2d[slot], 0));
    CK(cudaStreamWaitEvent(side, ev_vfree, 0));
    CK(cudaMemcpyAsync(stage, pre_x[slot], (size_t)B * D * es, cudaMemcpyDeviceToDevice, side));
    if (y_kind == AGP_Y_CLASS) CK(cudaMemcpyAsync(ycls, pre_ycls[slot], B * sizeof(int), cudaMemcpyDeviceToDevice, side));
    else for (int t = 0; t < nT; ++t) CK(cudaMemcpyAsync(yb + (size_t)t * ldB, pre_y[slot] + (size_t)t * ldB, B * sizeof(double), cudaMemcpyDeviceToDevice, side));
    CK(cudaEventRecord(ev_xfree[slot], side));
    CKS(run_piece(0, side, [&]() -> int {
      const int64_t sld = x_layout == AGP_LAYOUT_ROWMAJOR ? D : B;
      const int bl = (B + 255) / 256;
      if (x_dtype == AGP_DTYPE_F64) convert_rows_kernel<double, T><<<bl, 256, 0, st()>>>((const double*)stage, x_layout, sld, B, D, Xb, Dp, xxb);
      else convert_rows_kernel<float, T><<<bl, 256, 0, st()>>>((const float*)stage, x_layout, sld, B, D, Xb, Dp, xxb);
      ++launches;
      CKS(moments_impl(true, B, true, 1));
      if (prec == AGP_PREC_TF32X3 && Ql == 1) CK(cudaMemsetAsync(lat[0].racc + ldB, 0, 2 * ldB * sizeof(double), st()));   // the V X^T accumulators, off the critical chain
      return AGP_OK;
    }));
    CK(cudaEventRecord(ev_p1, side));
    // main stream: ONE graph -- statistics, local updates, Gram | ev_vfree | wait(result kernels of step i-1) | update + tail
    CK(cudaStreamWaitEvent(ctx->stream, ev_p1, 0));
    curB = B; cur_from_batch = true; kernel_matrices_stale = false; prefetched = false; stats_early = false;
    CKS(run_piece(1, ctx->stream, [&]() -> int {
      fuse_lik_next = can_fuse_lik(); fuse_from_batch = true; lik_fused = false;
      racc2_precleared = (prec == AGP_PREC_TF32X3 && Ql == 1);
      int s2 = moments_impl(true, B, true, 2);
      fuse_lik_next = false;
      CKS(s2);
      allow_split_gram = true;
      int sa = step_update_a(rho);
      allow_split_gram = false;
      CKS(sa);
      // inside a capture these become event-record / event-wait NODES (external events); eagerly they are ordinary stream operations
      CK(cudaEventRecordWithFlags(ev_vfree, st(), capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
      CK(cudaStreamWaitEvent(st(), ev_res, capturing ? cudaEventWaitExternal : cudaEventWaitDefault));   // they still read X / t
      chain_break = true;                              // combine_kernel's in-graph predecessor is not a kernel
      return step_update_b(rho);
    }));
    res_pending = -1;
    CK(cudaEventRecord(ev_step[slot], ctx->stream));
    for (auto& L : lat) L.muv_valid = false;
    have_step = true;
    // result stream
    CK(cudaStreamWaitEvent(res_stream, ev_step[slot], 0));
    {
      Latent& L = lat[0];
      cudaStream_t saved = cur_stream;
      cur_stream = res_stream;
      ensure_muv(L);
      symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Lc, mp, m, L.muv, L.mu0 + mp);
      ++launches;
      cur_stream = saved;
      CK(cudaMemcpyAsync(h_res[slot], L.mu0 + mp, m * sizeof(double), cudaMemcpyDeviceToHost, res_stream));
      CK(cudaMemcpyAsync(h_stat + slot, status, sizeof(int), cudaMemcpyDeviceToHost, res_stream));
    }
    CK(cudaEventRecord(ev_done[slot], res_stream));
    CK(cudaEventRecord(ev_res, res_stream));
    res_pending = slot;
    CK(cudaGetLastError());
    *ticket = n_tickets++;
    ++h_steps;
    return AGP_OK;
  }
  bool async_pipelined_ok() const { return pipeline && !prof && ns_iters <= 0 && !peer && !getenv("AGP_ASYNC_SERIAL"); }
  int step_batch_async(const void* xbh, int x_dtype, int x_layout, const void* const* ybh, int y_kind, int B, double rho, int64_t* ticket) override {
    if (!xbh || !ybh || !ticket) BAD("null batch / ticket");
    if (B < 1 || B > Bcap) BAD("The size of mini-batch is incorrect (negative or bigger than the batch capacity)");
    if (is_lsm != (y_kind == AGP_Y_CLASS)) BAD("label kind does not match the likelihood");
    if ((x_dtype != 0 && x_dtype != 1) || (x_layout != 0 && x_layout != 1)) BAD("bad dtype/layout");
    CKS(async_init());
    // a ragged batch on the tcgen05 path takes the in-order variant below (its step pads the rows itself)
    if (async_pipelined_ok() && !(prec == AGP_PREC_TF32X3 && (B % 128))) return step_batch_async_pipelined(xbh, x_dtype, x_layout, ybh, y_kind, B, rho, ticket);
    join_async();
    const int slot = (int)(n_tickets & 1);
    const size_t es = x_dtype == AGP_DTYPE_F64 ? 8 : 4;
    // copy stream: this slot's previous batch has been consumed (its step finished), then host -> pre-staging
    CK(cudaStreamWaitEvent(copy_stream, ev_free[slot], 0));
    CK(cudaStreamWaitEvent(copy_stream, ev_xfree[slot], 0));
    CK(cudaMemcpyAsync(pre_x[slot], xbh, (size_t)B * D * es, cudaMemcpyHostToDevice, copy_stream));
    std::vector<const void*> yp((size_t)std::max(nT, 1));
    if (y_kind == AGP_Y_CLASS) {
      CK(cudaMemcpyAsync(pre_ycls[slot], ybh[0], B * sizeof(int), cudaMemcpyHostToDevice, copy_stream));
      yp[0] = pre_ycls[slot];
    } else {
      for (int t = 0; t < nT; ++t) {
        CK(cudaMemcpyAsync(pre_y[slot] + (size_t)t * ldB, ybh[t], B * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
        yp[t] = pre_y[slot] + (size_t)t * ldB;
      }
    }
    CK(cudaEventRecord(ev_h2d[slot], copy_stream));
    // main stream: device -> staging copies + the step, in order behind the previous step
    CK(cudaStreamWaitEvent(ctx->stream, ev_h2d[slot], 0));
    CKS(ns_begin_step());
    int rc = step_batch_impl(pre_x[slot], x_dtype, x_layout, yp.data(), y_kind, B, rho, cudaMemcpyDeviceToDevice);
    if (rc != AGP_OK) { ns_tail_now = false; return rc; }
    ns_end_step();
    CK(cudaEventRecord(ev_free[slot], ctx->stream));
    // result slot: batch_compute left the canonical mean of latent 0 behind mu0 (also copied to h_mu by the step itself)
    CK(cudaMemcpyAsync(h_res[slot], lat[0].mu0 + mp, m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_stat + slot, status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ev_done[slot], ctx->stream));
    *ticket = n_tickets++;
    return AGP_OK;
  }
  int result_wait(int64_t ticket, double* mu) override {
    if (!mu || ticket < 0 || ticket >= n_tickets) BAD("unknown ticket");
    if (ticket < n_tickets - 2) BAD("ticket expired: only the results of the last two asynchronous steps are kept");
    const int slot = (int)(ticket & 1);
    CK(cudaEventSynchronize(ev_done[slot]));
    const int s = h_stat[slot];
    if (s) {
      CK(cudaMemsetAsync(status, 0, sizeof(int), ctx->stream));
      if (s & ST_TAIL_TIMEOUT) { ctx->err = "persistent m x m tail timed out waiting for a tile"; return AGP_ERR_STATE; }
      if (s & ST_NS_NOCONV) { ctx->err = "experimental Newton-Schulz tail (AGP_TAIL_NS) did not converge"; return AGP_ERR_STATE; }
      if (s & ST_NOT_POSDEF) { ctx->err = "PosDefException: matrix is not positive definite; Cholesky factorization failed."; return AGP_ERR_NOT_POSDEF; }
      ctx->err = "K̃ has negative values";
      return AGP_ERR_KTILDE_NONPOS;
    }
    memcpy(mu, h_res[slot], m * sizeof(double));
    return AGP_OK;
  }

  int* h_status = nullptr;   // pinned: one stream synchronisation reads the sticky device status
  int sync_status() override {
    if (!h_status) CK(cudaMallocHost((void**)&h_status, sizeof(int)));
    CK(cudaMemcpyAsync(h_status, status, sizeof(int), cudaMemcpyDeviceToHost, st()));
    CK(cudaStreamSynchronize(st()));
    int s = *h_status;
    if (s) {
      CK(cudaMemset(status, 0, sizeof(int)));
      if (s & ST_PEER_TIMEOUT) { ctx->err = "peer exchange timed out (a rank of the latent-sharded group did not publish its moments)"; return AGP_ERR_STATE; }
      if (s & ST_TAIL_TIMEOUT) { ctx->err = "persistent m x m tail timed out waiting for a tile (CTAs of one launch not co-resident?); AGP_TAIL_VARIANT=2 selects the multi-launch tail"; return AGP_ERR_STATE; }
      if (s & ST_NS_NOCONV) { ctx->err = "experimental Newton-Schulz tail (AGP_TAIL_NS) did not converge: raise AGP_TAIL_NS_AFTER or the iteration count"; return AGP_ERR_STATE; }
      if (s & ST_NOT_POSDEF) { ctx->err = "PosDefException: matrix is not positive definite; Cholesky factorization failed."; return AGP_ERR_NOT_POSDEF; }
      ctx->err = "K̃ has negative values";
      return AGP_ERR_KTILDE_NONPOS;
    }
    return AGP_OK;
  }

  void* moments_ptr(int which, int64_t* ld) override {
    if (ld) *ld = ldB;
    return which == 0 ? (void*)mean_f : (void*)var_f;
  }

  // moments of the last minibatch under the UPDATED posterior (ELBO uses the post-update mu, Sigma)
  int elbo_moments() override {
    if (curB < 1) { ctx->err = "no minibatch to evaluate the ELBO on"; return AGP_ERR_STATE; }
    CKS(ensure_factors());
    const bool was_stale = kernel_matrices_stale;  // a prefetch / predict_f overwrote Knm / V: rebuild them first
    CKS(ensure_current_kernel_matrices());
    return moments_impl(cur_from_batch, curB, was_stale, 2);
  }

  int elbo(double rho, double* out3) override {
    if (!out3) BAD("null output");
    if (curB < 1) { ctx->err = "no minibatch to evaluate the ELBO on"; return AGP_ERR_STATE; }
    CKS(ensure_factors());
    if (Ql == Qg) CKS(elbo_moments());
    const int B = curB;
    CK(cudaMemsetAsync(d_out, 0, 8 * sizeof(double), st()));
    LikParams lp = lik_params(B, true /* labels already gathered into yb */, 0);
    if (model_kind == AGP_MODEL_MOSVGP) launch_lik(lp, false);
    elbo_lik_kernel<<<(B + 255) / 256, 256, 0, st()>>>(lp, d_out);
    ++launches;
    std::vector<double> ld(Ql), h(8);
    double kl = 0.0;
    for (int q = 0; q < Ql; ++q) {
      Latent& L = lat[q];
      CK(cudaMemsetAsync(d_out + 4, 0, 2 * sizeof(double), st()));
      ensure_muv(L);
      gauss_kl_x_kernel<<<(m + 7) / 8, 256, 0, st()>>>(L.Xv, mp, m, L.muv, L.mu0v, d_out + 4);
      ++launches;
      double t2[2], ldp;
      CK(cudaMemcpyAsync(t2, d_out + 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, st()));
      CK(cudaMemcpyAsync(&ldp, L.logdetP, sizeof(double), cudaMemcpyDeviceToHost, st()));
      CK(cudaStreamSynchronize(st()));
      // KLdivergences.jl:17 with logdet K - logdet Sigma = logdet P_v, tr(K \\ Sigma) = tr(Sigma_v), invquad = |mu_v - mu0_v|^2
      kl += 0.5 * (ldp + t2[0] + t2[1] - (double)m);
    }
    CK(cudaMemcpyAsync(h.data(), d_out, 4 * sizeof(double), cudaMemcpyDeviceToHost, st()));
    CK(cudaStreamSynchronize(st()));
    out3[0] = rho * h[0]; out3[1] = kl; out3[2] = rho * h[2];
    return AGP_OK;
  }

  // ---- accessors -------------------------------------------------------------------------------------
  int copy_mat(const double* d, double* h) {
    CK(cudaMemcpy2D(h, (size_t)m * 8, d, (size_t)mp * 8, (size_t)m * 8, m, cudaMemcpyDeviceToHost));
    return AGP_OK;
  }
  int get_posterior(int ql, double* mu, double* Sigma, double* eta1, double* eta2) override {
    if (ql < 0 || ql >= Ql) BAD("latent index out of range");
    if (ql == 0 && mu && !Sigma && !eta1 && !eta2 && h_mu_valid) {   // mean already on the host (host-batch step)
      CK(cudaStreamSynchronize(st()));
      memcpy(mu, h_mu, m * sizeof(double));
      return AGP_OK;
    }
    Latent& L = lat[ql];
    if (!L.white_valid) {  // nothing has run yet: the canonical parameters are the truth (Sigma = inv(-2 eta2) needs K only formally)
      if (!have_K) CKS(refresh_K());
    }
    if (Sigma || eta1 || eta2) CKS(ensure_factor(L));   // (the mean alone is served from the refined covariance, see ensure_muv)
    // canonical from whitened, fp64:  mu = L mu_v,  Sigma = L Sigma_v L^T,  eta1 = L^-T eta1_v,  eta2 = L^-T eta2_v L^-1
    if (eta1 || eta2) canonicalize(L);
    if (mu) { ensure_muv(L); symv_kernel<<<(m * 32 + 255) / 256, 256, 0, st()>>>(L.Lc, mp, m, L.muv, L.mu0 + mp); ++launches; }  // scratch behind mu0
    if (Sigma) {
      dgemm(true, true, L.Xv, L.Xv, L.X, 1.0, 0.0);        // Sigma_v = X^T X
      dgemm(false, false, L.X, L.Lc, L.W, 1.0, 0.0);       // W = Sigma_v L^T   (NT)
      dgemm(false, true, L.Lc, L.W, L.X, 1.0, 0.0);        // Sigma = L W
    }
    CK(cudaStreamSynchronize(st()));
    if (mu) CK(cudaMemcpy(mu, L.mu0 + mp, m * 8, cudaMemcpyDeviceToHost));
    if (eta1) CK(cudaMemcpy(eta1, L.eta1c, m * 8, cudaMemcpyDeviceToHost));
    if (Sigma) CKS(copy_mat(L.X, Sigma));
    if (eta2) CKS(copy_mat(L.eta2c, eta2));
    return AGP_OK;
  }
  int set_posterior(int ql, const double* eta1, const double* eta2) override {
    h_mu_valid = false; h_steps = 0;
    if (ql < 0 || ql >= Ql || !eta1 || !eta2) BAD("bad posterior arguments");
    Latent& L = lat[ql];
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemset(L.eta1c, 0, mp * 8));
    CK(cudaMemset(L.eta2c, 0, (size_t)mp * mp * 8));
    CK(cudaMemcpy(L.eta1c, eta1, m * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy2D(L.eta2c, (size_t)mp * 8, eta2, (size_t)m * 8, (size_t)m * 8, m, cudaMemcpyHostToDevice));
    L.white_valid = false;
    if (have_K) { CKS(whiten(L)); return sync_status(); }
    return AGP_OK;
  }
  int get_counters(int64_t* t, int64_t* cur) override {
    int64_t c[2];
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(c, counters, 16, cudaMemcpyDeviceToHost));
    if (t) *t = c[0];
    if (cur) *cur = c[1];
    return AGP_OK;
  }
  int set_counters(int64_t t, int64_t cur) override {
    if (t < 1 || cur < 0) BAD("bad counters");
    int64_t c[2] = {t, cur};
    prefetched = false; stats_early = false; drop_graph();
    h_steps = t - 1;
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(counters, c, 16, cudaMemcpyHostToDevice));
    CKS(upload_lr(t));
    return AGP_OK;
  }
  int get_local(const char* name, int row, double* out, int B) override {
    if (!name || !out || B < 1 || B > Bcap || row < 0) BAD("bad local-variable query");
    std::string s(name);
    const double* base = nullptr; int rows = 1;
    if (s == "c") { base = lc; rows = R; }
    else if (s == "theta") { base = ltheta; rows = R; }
    else if (s == "gamma") { base = lgamma_; rows = R; }
    else if (s == "b") { base = lc; rows = R; }                                   // Laplace: b lives in the c array
    else if (s == "phi") { base = lc + ldB; rows = 1; if (!is_het) BAD("unknown local variable"); }
    else if (s == "sigma_g") { base = lgamma_ + ldB; rows = 1; if (!is_het) BAD("unknown local variable"); }
    else if (s == "alpha") { base = lalpha; rows = 1; }
    else if (s == "mean_f" || s == "var_f") {
      int64_t xe = 0;
      if (peer) { CK(cudaStreamSynchronize(st())); CK(cudaMemcpy(&xe, d_xepoch, 8, cudaMemcpyDeviceToHost)); }
      base = (s == "mean_f" ? mean_f : var_f) + (xe & 1) * par_stride; rows = Qg;
    }
    else if (s == "grad_mu") { base = gmu; rows = Ql; }
    else if (s == "grad_Sigma") { base = gS; rows = Ql; }
    else if (s == "y") { base = yb; rows = nT; }
    else if (s == "Ktilde") { if (row >= Ql) BAD("row out of range"); base = lat[row].Ktilde; rows = row + 1; row = 0; CK(cudaStreamSynchronize(st())); CK(cudaMemcpy(out, base, B * 8, cudaMemcpyDeviceToHost)); return AGP_OK; }
    else BAD("unknown local variable");
    if (row >= rows) BAD("row out of range");
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(out, base + (size_t)row * ldB, B * 8, cudaMemcpyDeviceToHost));
    return AGP_OK;
  }
  int get_kernel_matrices(int ql, double* Knm, double* kappa, int B) override {
    if (ql < 0 || ql >= Ql || B < 1 || B > Bcap) BAD("bad kernel-matrix query");
    if (curB > 0) CKS(ensure_current_kernel_matrices());
    Latent& L = lat[ql];
    if (!pKS) CKS(dalloc(&pKS, (size_t)Bcap * ldm));
    if (kappa) {  // kappa = Knm / K = V L^-1   (NN product with the T shadow of L^-1)
      GemmParams<T> g{};
      g.A = L.V; g.lda = ldm; g.B = L.Linv_T; g.ldb = ldm; g.C = pKS; g.ldc = ldm; g.M = B; g.N = m; g.K = m; g.alpha = 1.0;
      gemm_simt_launch<T, false, true, EPI_PLAIN>(g, 1, st());
      ++launches;
    }
    CK(cudaStreamSynchronize(st()));
    std::vector<T> tmp((size_t)B * ldm);
    for (int w = 0; w < 2; ++w) {
      double* dst = w == 0 ? Knm : kappa;
      if (!dst) continue;
      CK(cudaMemcpy(tmp.data(), w == 0 ? L.Knm : pKS, tmp.size() * sizeof(T), cudaMemcpyDeviceToHost));
      for (int b = 0; b < B; ++b) for (int j = 0; j < m; ++j) dst[(size_t)b * m + j] = (double)tmp[(size_t)b * ldm + j];
    }
    return AGP_OK;
  }
  int get_Kinv(int ql, double* Kinv, double* logdetK) override {
    if (ql < 0 || ql >= Ql) BAD("latent index out of range");
    if (!have_K) { ctx->err = "agp_refresh_K has not run"; return AGP_ERR_STATE; }
    CK(cudaStreamSynchronize(st()));
    if (Kinv) CKS(copy_mat(lat[ql].Kinv, Kinv));
    if (logdetK) *logdetK = lat[ql].logdetK;
    return AGP_OK;
  }

  // ---- prediction (training/predictions.jl:25-50) ----------------------------------------------------
  int predict_f(const void* Xt, int x_dtype, int x_layout, int64_t nt, int want_var, double* mu_out, double* var_out) override;
  int proba_logistic(const double* mu, const double* var, int64_t nn_, const double* nodes, const double* w, int nq, double* p,
                     double* pv) override {
    if (!mu || !var || !nodes || !w || !p || !pv || nn_ < 1 || nq < 1) BAD("bad proba arguments");
    double *dm, *dv, *dn, *dw, *dp, *dpv;
    CK(cudaMalloc(&dm, nn_ * 8)); CK(cudaMalloc(&dv, nn_ * 8)); CK(cudaMalloc(&dp, nn_ * 8)); CK(cudaMalloc(&dpv, nn_ * 8));
    CK(cudaMalloc(&dn, nq * 8)); CK(cudaMalloc(&dw, nq * 8));
    CK(cudaMemcpyAsync(dm, mu, nn_ * 8, cudaMemcpyHostToDevice, st())); CK(cudaMemcpyAsync(dv, var, nn_ * 8, cudaMemcpyHostToDevice, st()));
    CK(cudaMemcpyAsync(dn, nodes, nq * 8, cudaMemcpyHostToDevice, st())); CK(cudaMemcpyAsync(dw, w, nq * 8, cudaMemcpyHostToDevice, st()));
    proba_logistic_kernel<<<(int)((nn_ + 127) / 128), 128, 0, st()>>>(dm, dv, nn_, dn, dw, nq, dp, dpv);
    ++launches;
    CK(cudaMemcpyAsync(p, dp, nn_ * 8, cudaMemcpyDeviceToHost, st())); CK(cudaMemcpyAsync(pv, dpv, nn_ * 8, cudaMemcpyDeviceToHost, st()));
    CK(cudaStreamSynchronize(st()));
    cudaFree(dm); cudaFree(dv); cudaFree(dn); cudaFree(dw); cudaFree(dp); cudaFree(dpv);
    return AGP_OK;
  }

  int set_quadrature(const double* nodes, const double* w, int nn) override {
    if (!nodes || !w || nn < 1 || nn > 128) BAD("bad quadrature rule (1..128 nodes)");
    CK(cudaMemcpyAsync(d_qnodes, nodes, nn * 8, cudaMemcpyHostToDevice, st()));
    CK(cudaMemcpyAsync(d_qw, w, nn * 8, cudaMemcpyHostToDevice, st()));
    CK(cudaStreamSynchronize(st()));
    nq = nn;
    return AGP_OK;
  }
  int get_lik_param(int task, double* v) override {
    if (task < 0 || task >= nT || !v) BAD("bad likelihood-parameter query");
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(v, (h_lik_kind[task] == AGP_LIK_GAUSSIAN ? d_p0 : d_lam) + task, 8, cudaMemcpyDeviceToHost));   // Gaussian: the live sigma^2
    return AGP_OK;
  }
  int set_lik_param(int task, double v) override {
    if (task < 0 || task >= nT || !(v > 0)) BAD("bad likelihood parameter");
    CK(cudaStreamSynchronize(st()));
    CK(cudaMemcpy(d_lam + task, &v, 8, cudaMemcpyHostToDevice));
    return AGP_OK;
  }
  int proba_link(int link, double p0, const double* mu, const double* var, int64_t nn_, double* p, double* pv) override {
    if (!mu || !var || !p || !pv || nn_ < 1 || link < 0 || link > 3) BAD("bad proba arguments");
    if (nq < 1) { ctx->err = "agp_set_quadrature has not been called"; return AGP_ERR_STATE; }
    double *dm, *dv, *dp, *dpv;
    CK(cudaMalloc(&dm, nn_ * 8)); CK(cudaMalloc(&dv, nn_ * 8)); CK(cudaMalloc(&dp, nn_ * 8)); CK(cudaMalloc(&dpv, nn_ * 8));
    CK(cudaMemcpyAsync(dm, mu, nn_ * 8, cudaMemcpyHostToDevice, st())); CK(cudaMemcpyAsync(dv, var, nn_ * 8, cudaMemcpyHostToDevice, st()));
    proba_link_kernel<<<(int)((nn_ + 127) / 128), 128, 0, st()>>>(link, p0, dm, dv, nn_, d_qnodes, d_qw, nq, dp, dpv);
    ++launches;
    CK(cudaMemcpyAsync(p, dp, nn_ * 8, cudaMemcpyDeviceToHost, st())); CK(cudaMemcpyAsync(pv, dpv, nn_ * 8, cudaMemcpyDeviceToHost, st()));
    CK(cudaStreamSynchronize(st()));
    cudaFree(dm); cudaFree(dv); cudaFree(dp); cudaFree(dpv);
    return AGP_OK;
  }

  int profile_enable(int on) override {
    resolve_pending();
    prof = on != 0;
    for (int i = 0; i < PH_COUNT; ++i) { ph_ms[i] = 0; ph_launch[i] = 0; }
    return AGP_OK;
  }
  int profile_read(int maxp, const char** names, double* ms, int64_t* ls) override {
    CK(cudaStreamSynchronize(st()));
    resolve_pending();
    int k = std::min(maxp, (int)PH_COUNT);
    for (int i = 0; i < k; ++i) { if (names) names[i] = kPhaseNames[i]; if (ms) ms[i] = ph_ms[i]; if (ls) ls[i] = ph_launch[i]; }
    return PH_COUNT;
  }
  int64_t launch_count() override { return launches; }

  // bench hook: one hot kernel, `reps` launches back to back between two events (see include/agp_b200.h)
  int time_kernel(int which, int reps, double* ms_out) override {
    if (!ms_out || reps < 1 || which < 0 || which > 3) BAD("bad time_kernel arguments");
    if (!have_K || !have_data || !idx_pool || curB < 1 || pool_B != curB) { ctx->err = "time_kernel needs a completed resident-list step"; return AGP_ERR_STATE; }
    if (prec != AGP_PREC_TF32X3) { ctx->err = "time_kernel measures the tcgen05 kernels (precision tf32x3)"; return AGP_ERR_STATE; }
    const int B = curB;
    Latent& L = lat[0];
    CKS(ensure_factors());
    CKS(sync_status());
    int64_t c[2];
    CK(cudaMemcpy(c, counters, 16, cudaMemcpyDeviceToHost));
    T* xxr = nullptr;
    if (which == 0) {
      CKS(dalloc(&xxr, (size_t)reps * B));
      for (int r = 0; r < reps; ++r)
        xx_gather_kernel<T><<<(B + 255) / 256, 256, 0, st()>>>(idx_pool + ((c[1] + r) % n_lists) * B, B, xx, xxr + (size_t)r * B);
    }
    if (which == 3 && !L.um.gram_tn) CKS(umma_scale_transpose(ctx_err(), L.um, (const float*)(const void*)L.V, gS, 1.0, gmu, L.v1, B, mk, st()));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaStreamSynchronize(st()));
    CK(cudaEventRecord(e0, st()));
    int rc = AGP_OK;
    for (int r = 0; r < reps && rc == AGP_OK; ++r) {
      if (which == 0) {
        if (L.knm_tc)
          rc = umma_knm(ctx_err(), L.uk, (const float*)(const void*)X, Dp, Dp, idx_pool + ((c[1] + r) % n_lists) * B,
                        (const float*)(const void*)(xxr + (size_t)r * B), (const float*)(const void*)L.zz, B, L.kind, L.scale * L.scale, L.variance, st());
        else { ctx->err = "K_nm tensor-core kernel not active (D > 128 or D <= 16)"; rc = AGP_ERR_STATE; }
      } else if (which == 1 || which == 2) {
        UmmaEpilogue ep{};
        ep.mode = which == 1 ? UMMA_EPI_STORE_SUMSQ : UMMA_EPI_STATS_ONLY;
        ep.acc0 = L.racc + (which == 1 ? 0 : ldB); ep.acc1 = L.racc + 2 * ldB; ep.tvec = L.tvec;
        rc = umma_gemm_nt(ctx_err(), L.um, which == 1 ? UM_KNM : UM_V, which == 1 ? UM_LINV : UM_X, (float*)(void*)(which == 1 ? L.V : L.VS), B, mk, ep, st());
      } else {
        int ns = n_split;
        if (L.um.gram_tn) rc = umma_gram_tn(ctx_err(), L.um, (float*)(void*)L.Gpart, gS, 1.0, gmu, L.v1, B, mk, &ns, st());
        else rc = umma_gram(ctx_err(), L.um, (float*)(void*)L.Gpart, B, mk, &ns, st());
      }
      ++launches;
    }
    CK(cudaEventRecord(e1, st()));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(xxr);
    *ms_out = (double)ms / reps;
    // Knm / V / accumulators now hold scratch values: re-prime the pipeline before the next step / ELBO
    cudaMemsetAsync(L.v1, 0, m * sizeof(double), st());
    racc2_precleared = false;
    kernel_matrices_stale = true;
    if (prefetched) {  // fall back to the last consumed minibatch (same as predict_f)
      prefetched = false; stats_early = false;
      cudaMemcpyAsync(idx_cur, idx_prev, (size_t)curB * 8, cudaMemcpyDeviceToDevice, st());
      xx_gather_kernel<T><<<(curB + 255) / 256, 256, 0, st()>>>(idx_cur, curB, xx, xx_cur);
      ++launches;
    }
    drop_graph();
    return rc;
  }
  int use_graph(int on) override { want_graph = on != 0; if (!on) drop_graph(); return AGP_OK; }
};

// _predict_f (training/predictions.jl:25-50, diag = true): the same row-moment pipeline as the step, on test rows.
// NOTE: uses the step's Knm / V / VS buffers; the last minibatch's kernel matrices are saved and restored around it
// only logically (a later ELBO call recomputes them from the retained minibatch indices).
template <typename T>
int Engine<T>::predict_f(const void* Xt, int x_dtype, int x_layout, int64_t nt, int want_var, double* mu_out, double* var_out) {
  if (!Xt || !mu_out || nt < 1 || (want_var && !var_out)) BAD("bad predict arguments");
  if (!have_K) CKS(refresh_K());  // predictions.jl:28-29: compute_K when no state is passed
  CKS(ensure_factors());
  CKS(sync_status());             // surface errors of earlier asynchronous steps before the flag is reused
  double *dmu = nullptr, *dvar = nullptr;
  CK(cudaMalloc(&dmu, (size_t)Ql * ldB * 8)); CK(cudaMalloc(&dvar, (size_t)Ql * ldB * 8));
  if (!pXb) { CKS(dalloc(&pXb, (size_t)Bcap * Dp)); CKS(dalloc(&pxxb, Bcap)); }
  int rc = AGP_OK;
  for (int64_t r0 = 0; r0 < nt && rc == AGP_OK; r0 += Bcap) {
    int B = (int)std::min<int64_t>(Bcap, nt - r0);
    int Bk = B;
    if (prec == AGP_PREC_TF32X3) Bk = (int)rup(B, 128);  // tensor-core tiles: pad with (zeroed) rows
    if (Bk != B) cudaMemsetAsync(pXb, 0, (size_t)Bcap * Dp * sizeof(T), st());
    rc = upload_rows(Xt, x_dtype, x_layout, nt, r0, B, pXb, pxxb);
    if (rc != AGP_OK) break;
    int st_before = 0;
    rc = moments_rows(pXb, pxxb, nullptr, Bk, true, dmu, dvar, ldB, want_var != 0);
    (void)st_before;
    if (rc != AGP_OK) break;
    for (int q = 0; q < Ql; ++q) {
      if (cudaMemcpyAsync(mu_out + (size_t)q * nt + r0, dmu + (size_t)q * ldB, B * 8, cudaMemcpyDeviceToHost, st()) != cudaSuccess) rc = AGP_ERR_CUDA;
      if (want_var && cudaMemcpyAsync(var_out + (size_t)q * nt + r0, dvar + (size_t)q * ldB, B * 8, cudaMemcpyDeviceToHost, st()) != cudaSuccess) rc = AGP_ERR_CUDA;
    }
    if (cudaStreamSynchronize(st()) != cudaSuccess) rc = AGP_ERR_CUDA;
  }
  cudaFree(dmu); cudaFree(dvar);
  // the predictive variance of the reference is NOT checked for positivity (predictions.jl:38-44): drop a Ktilde flag
  cudaMemsetAsync(status, 0, sizeof(int), st());
  kernel_matrices_stale = true;
  if (prefetched) {  // the prefetched kernel matrices were overwritten: fall back to the last consumed minibatch
    prefetched = false; stats_early = false; drop_graph();
    cudaMemcpyAsync(idx_cur, idx_prev, (size_t)curB * 8, cudaMemcpyDeviceToDevice, st());
    xx_gather_kernel<T><<<(curB + 255) / 256, 256, 0, st()>>>(idx_cur, curB, xx, xx_cur);
    ++launches;
  }
  if (rc == AGP_ERR_CUDA && ctx->err.empty()) ctx->err = "CUDA failure in predict_f";
  return rc;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int agp_abi_version(void) { return AGP_ABI_VERSION; }

int agp_ctx_create(int device, void* cuda_stream, agp_ctx** out) {
  if (!out) return AGP_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return AGP_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return AGP_ERR_CUDA;
  agp_ctx* c = new agp_ctx();
  c->device = device;
  if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return AGP_ERR_CUDA; }
    c->own_stream = true;
  }
  *out = c;
  return AGP_OK;
}
void agp_ctx_destroy(agp_ctx* ctx) {
  if (!ctx) return;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}
const char* agp_last_error(const agp_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int agp_model_create(agp_ctx* ctx, const agp_model_desc* desc, agp_model** out) {
  if (!ctx || !desc || !out) return AGP_ERR_BAD_ARG;
  *out = nullptr;
  EngineBase* e = nullptr;
  int rc;
  if (desc->precision == AGP_PREC_F64) { auto* p = new Engine<double>(); rc = p->init(ctx, desc); e = p; }
  else { auto* p = new Engine<float>(); rc = p->init(ctx, desc); e = p; }
  if (rc != AGP_OK) { delete e; return rc; }
  agp_model* mdl = new agp_model();
  mdl->eng = e;
  *out = mdl;
  return AGP_OK;
}
void agp_model_destroy(agp_model* model) {
  if (!model) return;
  delete model->eng;
  delete model;
}

#define ENG0(m) if (!(m) || !(m)->eng) return AGP_ERR_BAD_ARG; EngineBase* e = (m)->eng; cudaSetDevice(e->ctx->device)
#define ENG(m) ENG0(m); e->join_async()

int agp_data_upload(agp_model* model, const void* X, int x_dtype, int x_layout, int64_t n, const void* const* y, int y_kind) {
  ENG(model); return e->data_upload(X, x_dtype, x_layout, n, y, y_kind);
}
int agp_minibatches_upload(agp_model* model, const int64_t* idx, int64_t n_lists, int32_t B, int32_t base) {
  ENG(model); return e->minibatches_upload(idx, n_lists, B, base);
}
int agp_refresh_K(agp_model* model) { ENG(model); return e->refresh_K(); }
int agp_state_reset(agp_model* model) { ENG(model); return e->state_reset(); }
int agp_set_kernel(agp_model* model, int32_t ql, int32_t kind, double scale, double variance) {
  ENG(model); return e->set_kernel(ql, kind, scale, variance);
}
int agp_step(agp_model* model, const int64_t* idx, int32_t B, int32_t base, double rho) {
  ENG(model);
  int s = e->step_full(idx, B, base, rho);
  if (s != AGP_OK) return s;
  return e->sync_status();
}
int agp_step_async(agp_model* model, const int64_t* idx, int32_t B, int32_t base, double rho) {
  ENG(model); return e->step_full(idx, B, base, rho);
}
int agp_step_batch(agp_model* model, const void* xb, int x_dtype, int x_layout, const void* const* yb, int y_kind, int32_t B,
                   double rho) {
  ENG(model);
  int s = e->step_batch(xb, x_dtype, x_layout, yb, y_kind, B, rho);
  if (s != AGP_OK) return s;
  return e->sync_status();
}
int agp_sync(agp_model* model) { ENG(model); return e->sync_status(); }
int agp_step_batch_async(agp_model* model, const void* xb, int x_dtype, int x_layout, const void* const* yb, int y_kind, int32_t B,
                         double rho, int64_t* ticket) {
  ENG0(model); return e->step_batch_async(xb, x_dtype, x_layout, yb, y_kind, B, rho, ticket);
}
int agp_result_wait(agp_model* model, int64_t ticket, double* mu) { ENG0(model); return e->result_wait(ticket, mu); }
int agp_step_moments_async(agp_model* model, const int64_t* idx, int32_t B, int32_t base) {
  ENG(model); return e->step_moments(idx, B, base, false);
}
int agp_step_update_async(agp_model* model, double rho) { ENG(model); return e->step_update(rho); }
void* agp_moments_devptr(agp_model* model, int32_t which, int64_t* ld_out) {
  if (!model || !model->eng) return nullptr;
  return model->eng->moments_ptr(which, ld_out);
}
int agp_elbo_moments_async(agp_model* model) { ENG(model); return e->elbo_moments(); }
int agp_elbo(agp_model* model, double rho, double* out3) { ENG(model); return e->elbo(rho, out3); }
int agp_get_posterior(agp_model* model, int32_t ql, double* mu, double* Sigma, double* eta1, double* eta2) {
  ENG(model); return e->get_posterior(ql, mu, Sigma, eta1, eta2);
}
int agp_set_posterior(agp_model* model, int32_t ql, const double* eta1, const double* eta2) {
  ENG(model); return e->set_posterior(ql, eta1, eta2);
}
int agp_get_counters(agp_model* model, int64_t* t, int64_t* cur) { ENG(model); return e->get_counters(t, cur); }
int agp_set_counters(agp_model* model, int64_t t, int64_t cur) { ENG(model); return e->set_counters(t, cur); }
int agp_get_local(agp_model* model, const char* name, int32_t row, double* out, int32_t B) {
  ENG(model); return e->get_local(name, row, out, B);
}
int agp_get_kernel_matrices(agp_model* model, int32_t ql, double* Knm, double* kappa, int32_t B) {
  ENG(model); return e->get_kernel_matrices(ql, Knm, kappa, B);
}
int agp_get_Kinv(agp_model* model, int32_t ql, double* Kinv, double* logdetK) { ENG(model); return e->get_Kinv(ql, Kinv, logdetK); }
int agp_predict_f(agp_model* model, const void* Xt, int x_dtype, int x_layout, int64_t nt, int want_var, double* mu, double* var) {
  ENG(model); return e->predict_f(Xt, x_dtype, x_layout, nt, want_var, mu, var);
}
int agp_set_step_size(agp_model* model, double eta) { ENG(model); return e->set_step_size(eta); }
int agp_set_noise_optimiser(agp_model* model, int32_t task, int32_t kind, double eta, double beta1, double beta2, double eps) {
  ENG(model); return e->set_noise_optimiser(task, kind, eta, beta1, beta2, eps);
}
int agp_hyper_grads(agp_model* model, double rho, double* d_scale, double* d_variance, double* dZ) {
  ENG(model); return e->hyper_grads(rho, d_scale, d_variance, dZ);
}
int agp_set_Z(agp_model* model, int32_t latent_local, const double* Z) { ENG(model); return e->set_Z(latent_local, Z); }
int agp_keep_stale_K(agp_model* model, int32_t on) { ENG(model); return e->keep_stale_K(on); }
int agp_set_A_optimiser(agp_model* model, int32_t kind, double eta, double beta1, double beta2, double eps) {
  ENG(model); return e->set_A_optimiser(kind, eta, beta1, beta2, eps);
}
int agp_get_A(agp_model* model, double* A) { ENG(model); return e->get_A(A); }
int agp_peer_export(agp_model* model, void* handle64) { ENG(model); return e->peer_export(handle64); }
int agp_peer_attach(agp_model* model, int32_t world, int32_t rank, const void* handles) { ENG(model); return e->peer_attach(world, rank, handles); }
int agp_peer_detach(agp_model* model) { ENG(model); return e->peer_detach(); }
int agp_set_quadrature(agp_model* model, const double* nodes, const double* weights, int32_t n_nodes) {
  ENG(model); return e->set_quadrature(nodes, weights, n_nodes);
}
int agp_get_lik_param(agp_model* model, int32_t task, double* value) { ENG(model); return e->get_lik_param(task, value); }
int agp_set_lik_param(agp_model* model, int32_t task, double value) { ENG(model); return e->set_lik_param(task, value); }
int agp_proba_link(agp_model* model, int32_t link, double p0, const double* mu, const double* var, int64_t n, double* pred, double* pred_var) {
  ENG(model); return e->proba_link(link, p0, mu, var, n, pred, pred_var);
}
int agp_proba_logistic(agp_model* model, const double* mu, const double* var, int64_t n, const double* nodes, const double* weights,
                       int32_t n_nodes, double* p, double* p_var) {
  ENG(model); return e->proba_logistic(mu, var, n, nodes, weights, n_nodes, p, p_var);
}
int agp_profile_enable(agp_model* model, int on) { ENG(model); return e->profile_enable(on); }
int agp_profile_read(agp_model* model, int32_t maxp, const char** names, double* ms, int64_t* launches) {
  ENG(model); return e->profile_read(maxp, names, ms, launches);
}
int agp_time_kernel(agp_model* model, int32_t which, int32_t reps, double* ms) { ENG(model); return e->time_kernel(which, reps, ms); }
int64_t agp_launch_count(agp_model* model) { return (model && model->eng) ? model->eng->launch_count() : 0; }
int agp_use_graph(agp_model* model, int on) { ENG(model); return e->use_graph(on); }
int agp_online_carry(agp_model* model, int32_t latent_local, const double* Za, int32_t ma, const double* invDa, const double* prev_eta1, double prev_L) {
  ENG(model); return e->online_carry(latent_local, Za, ma, invDa, prev_eta1, prev_L);
}
int agp_online_extra_kl(agp_model* model, double* out) { ENG(model); return e->online_extra_kl(out); }
int agp_predict_f_cov(agp_model* model, const double* Xt, int64_t nt, double* mu, double* cov) { ENG(model); return e->predict_f_cov(Xt, nt, mu, cov); }
int agp_local_updates_async(agp_model* model) { ENG(model); return e->local_updates_only(); }
int agp_step_with_gradients(agp_model* model, const int64_t* idx, int32_t B, int32_t base, const double* grad_mu, const double* grad_Sigma) {
  ENG(model); return e->step_with_gradients(idx, B, base, grad_mu, grad_Sigma);
}

// EXPERIMENTAL (see include/agp_b200.h): self-contained, does not touch any model
int agp_experimental_ns_refine(agp_ctx* ctx, int32_t m, const double* P, double* Y, int32_t iters, int32_t mode, double* resid, double* ms) {
  if (!ctx || !P || !Y || m < 128 || m % 128 || iters < 0 || iters > 64) {
    if (ctx) ctx->err = "agp_experimental_ns_refine: needs m % 128 == 0, 0 <= iters <= 64 and non-null P, Y";
    return AGP_ERR_BAD_ARG;
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return AGP_ERR_CUDA;
  agp::UmmaNs ns;
  int rc = agp::umma_ns_alloc(&ctx->err, ns, m, ctx->stream);
  if (rc) { agp::umma_ns_free(ns); return rc; }
  const size_t nn = (size_t)m * m;
  std::vector<float> hp(nn), hy(nn);
  for (size_t i = 0; i < nn; ++i) { hp[i] = (float)P[i]; hy[i] = (float)Y[i]; }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float t_ms = 0.f;
  cudaError_t ce = cudaMemcpyAsync(ns.P(), hp.data(), nn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ns.Y(), hy.data(), nn * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(ns.P64, P, nn * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
  if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
  if (ce == cudaSuccess) ce = cudaEventRecord(e0, ctx->stream);
  if (ce == cudaSuccess) rc = agp::umma_ns_iterate(&ctx->err, ns, iters, mode, nullptr, 0, ctx->stream);
  if (ce == cudaSuccess && !rc) ce = cudaEventRecord(e1, ctx->stream);
  if (ce == cudaSuccess && !rc) ce = cudaMemcpyAsync(hy.data(), ns.Y(), nn * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  std::vector<double> hr(64, 0.0);
  if (ce == cudaSuccess && !rc) ce = cudaMemcpyAsync(hr.data(), ns.resid, 64 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess && !rc) ce = cudaStreamSynchronize(ctx->stream);
  if (ce == cudaSuccess && !rc) ce = cudaEventElapsedTime(&t_ms, e0, e1);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  agp::umma_ns_free(ns);
  if (rc) return rc;
  if (ce != cudaSuccess) { ctx->err = std::string("agp_experimental_ns_refine: ") + cudaGetErrorString(ce); return AGP_ERR_CUDA; }
  for (size_t i = 0; i < nn; ++i) Y[i] = (double)hy[i];
  if (resid) for (int i = 0; i < iters; ++i) resid[i] = std::sqrt(hr[i]);
  if (ms) *ms = (double)t_ms;
  return AGP_OK;
}

}  // extern "C"
