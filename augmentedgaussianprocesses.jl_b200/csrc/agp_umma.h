// Interface of the tcgen05 (5th-gen tensor core) path for the B x m contractions: 3xTF32 split GEMMs with
// TMA-staged operands and TMEM accumulators.  Implemented in agp_umma.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "agp_lik.cuh"

namespace agp {

enum UmmaMat : int { UM_KNM = 0, UM_V = 1, UM_LINV = 2, UM_X = 3, UM_COUNT = 4 };

// fused epilogues of the tensor-core GEMM
enum UmmaEpiMode : int {
  UMMA_EPI_STORE = 0,        // C = acc
  UMMA_EPI_STORE_SUMSQ = 1,  // C = acc ; acc0[row] += sum_col acc^2                       (V = Knm L^-T -> Ktilde)
  UMMA_EPI_STATS_ONLY = 2,   // no store ; acc0[row] += sum acc^2 ; acc1[row] += sum acc*tvec[col]   (V X^T -> var_f, mean_f)
  UMMA_EPI_STORE_MIRROR = 3, // C = acc and C^T = acc^T for off-diagonal tiles              (symmetric Gram product)
  // second-generation kernel only (experimental Newton-Schulz tail, see UmmaNs):
  UMMA_EPI_EYE_MINUS = 4,    // C = I - acc ; acc0[0] += |C|_F^2
  UMMA_EPI_ADD = 5,          // C = cin + acc   (cin: same shape / leading dimension as C)
  UMMA_EPI_STATS_SIGMA = 6   // no store ; acc0[row] += sum acc*cin[row][col] ; acc1[row] += sum acc*tvec[col]
                             //   (V Sigma_v -> var_f = rowsum((V Sigma_v) o V), mean_f = (V Sigma_v) eta1_v: statistics against the full covariance)
};
// Row finish fused into the UMMA_EPI_STATS_ONLY epilogue of umma_gemm_ps_kernel (single-latent SVGP steps): the thread that adds the LAST
// N-tile contribution of a sample (per-row arrival counter) forms Ktilde, mean_f and var_f and runs the sample's local update right there
// -- what rowfinish_lik_kernel does in a launch of its own (5.7 us on the critical chain of the C2 step).  cnt: [rows] zero-initialised,
// reset by the finishing thread.
struct UmmaRowFinish {
  unsigned* cnt = nullptr;
  int B = 0;                         // samples (rows >= B are padding)
  const double* sumsq_v = nullptr;   // row sums of squares of V (K_nm L^-T product)
  double kdiag_jit = 0.0; double* Ktilde = nullptr; int* status = nullptr;
  LikParams lp;
};
struct UmmaEpilogue {
  int mode = 0;
  double* acc0 = nullptr; double* acc1 = nullptr; const double* tvec = nullptr;
  const float* cin = nullptr;
  const UmmaRowFinish* fin = nullptr;   // host pointer, copied into the launch (pre-split kernel + UMMA_EPI_STATS_ONLY only)
};

struct UmmaLatent {
  int m = 0, ldm = 0, Bcap = 0;
  float* UT = nullptr;   // U^T = (diag(sqrt(rho w)) V)^T, [m][Bcap], operand of the Gram product
  void* tmaps = nullptr; // host array of CUtensorMap (raw fp32 operands: Knm, V, L^-1, X, U^T)
  // second-generation GEMM kernel (opt-in through the environment variable AGP_UMMA_V2; see agp_umma.cu): bit 0 = on,
  // bit 1 = keep splitting the m x m operand inside the kernel.  Bsplit = TF32 hi / lo copies of L^-1 and X.
  int v2 = 0;
  float* Bsplit = nullptr;
  // pre-split right operand with the first-generation main loop (umma_gemm_ps_kernel; default ON, AGP_UMMA_PS=0 disables): Bsplit holds
  // the planes, the worker groups only split the A operand
  int ps = 0;
  // Gram product straight from V (umma_gram_tn_kernel; default ON, AGP_GRAM_TN=0 disables): no scale-transpose pass, U^T is not used
  int gram_tn = 0;
};

bool umma_shape_ok(int m, int Bcap);
// tensor maps are built once on the engine's fp32 operand buffers (all [rows][ldm], rows = Bcap or m)
int umma_latent_alloc(std::string* err, UmmaLatent& u, int m, int ldm, int Bcap, const float* Knm, const float* V, const float* Linv,
                      const float* X, cudaStream_t st);
void umma_latent_free(UmmaLatent& u);
// C[M x N] = A[M x K] * B[N x K]^T with K = m, 3xTF32 (operands split to hi/lo inside the kernel)
int umma_gemm_nt(std::string* err, UmmaLatent& u, int a_which, int b_which, float* C, int M, int N, const UmmaEpilogue& ep,
                 cudaStream_t st);
// v2 only (no-ops / null otherwise): refresh the pre-split copy of an m x m right operand (which = UM_LINV or UM_X) from its fp32
// matrix [m][ldm]; pointer to the hi (lo = 0) or lo (lo = 1) copy, for producers that write the split themselves
int umma_presplit(std::string* err, UmmaLatent& u, int which, const float* src, cudaStream_t st);
float* umma_split_ptr(const UmmaLatent& u, int which, int lo);
// U^T = (diag(sqrt(rho w)) V)^T into u.UT, fused with v1 += V^T g
int umma_scale_transpose(std::string* err, UmmaLatent& u, const float* V, const double* w, double rho, const double* g, double* v1,
                         int B, int m, cudaStream_t st);
// Gpart[s] = sum_{b in split s} U_b U_b^T (full symmetric tiles) ; *n_split in: capacity, out: used
// per-step chain launches carry the programmatic-dependent-launch attribute when on (set per call site by the engine)
void umma_set_pdl(bool on);
// cap of the persistent grid of umma_gemm_nt (0 = every SM): set around launches that run beside the persistent m x m tail
void umma_set_grid_cap(int n);
// restrict the NEXT umma_gemm_nt launches with a lower-triangular right operand to the N tiles [lo, hi] (lo < 0: all tiles): the step's
// V X^T statistics are issued N tile by N tile as the rows of X leave the m x m tail
void umma_set_tile_range(int lo, int hi);
int umma_gram(std::string* err, UmmaLatent& u, float* Gpart, int B, int m, int* n_split, cudaStream_t st);
// the same partials straight from V = u's UM_V matrix, scaled by sqrt(rho w) along the samples inside the kernel; v1 += V^T g
int umma_gram_tn(std::string* err, UmmaLatent& u, float* Gpart, const double* w, double rho, const double* g, double* v1, int B, int m,
                 int* n_split, cudaStream_t st);
// the Gram product in two launches: part 0 = tile (0, 0) over S slices, part 1 = the other upper tiles over S slices (<= grid_max CTAs)
int umma_gram_part(std::string* err, UmmaLatent& u, float* Gpart, int B, int m, int part, int S, int grid_max, cudaStream_t st);
int umma_gram_splits(int B, int cap);   // largest usable split count <= cap

// ---- grouped launches: the same-shaped product of several latent GPs in ONE persistent launch (multi-latent models: the per-launch
// fixed cost of a C2 / C4 sized product, ~8 us, is paid once; the persistent CTAs stay balanced over n x tiles work units) ----
struct UmmaGroups { void* dev = nullptr; int n = 0; int ps = 0; };   // device array, one entry (tensor maps + epilogue targets) per latent
// a_which / b_which: UmmaMat of the operands (-1 = the latent's U^T buffer: Gram product); C / acc0 / acc1 / tvec: per-latent epilogue
// targets (null arrays allowed).  Call again after a latent's buffers or tensor maps change.
int umma_groups_build(std::string* err, UmmaGroups& gs, UmmaLatent* const* lats, int n, int a_which, int b_which, float* const* C,
                      double* const* acc0, double* const* acc1, const double* const* tvec, cudaStream_t st);
void umma_groups_free(UmmaGroups& gs);
// C_q[M x N] = A_q B_q^T for every group q (b_tri: B_q lower triangular, k-blocks above the diagonal skipped); epi_mode: UmmaEpiMode
int umma_gemm_nt_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& shape, int b_tri, int M, int N, int epi_mode, cudaStream_t st);
// Gpart_q[s] = split-K partials of U_q^T U_q for every group; *n_split in: capacity of the partial buffers, out: slices used
int umma_gram_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& shape, int B, int m, int* n_split, cudaStream_t st);

// the same partials straight from every group's V (umma_gram_tn_kernel<true>; no scale-transpose pass, v1_q += V_q^T g_q inside): a
// separate device array (entries: V's tensor map, Gpart, the sample weights w_q, g_q, v1_q); rho multiplies every weight
int umma_gram_tn_groups_build(std::string* err, UmmaGroups& gs, UmmaLatent* const* lats, int n, float* const* Gpart, const double* const* w,
                              const double* const* g, double* const* v1, cudaStream_t st);
int umma_gram_tn_grouped(std::string* err, const UmmaGroups& gs, const UmmaLatent& shape, double rho, int B, int m, int* n_split, cudaStream_t st);

// ---- EXPERIMENTAL, not on the product path (never run on a GPU yet): Newton-Schulz refinement of an m x m inverse ----
// Y <- Y + Y (I - P Y) as two 3xTF32 tensor-core products per iteration (T = I - Y P, then Y' = Y + Y T^T), the candidate
// replacement of the fp64 Cholesky tail once the Robbins-Monro step is small (profiles/r1/studies/newton_schulz_*.txt:
// spectral radius of I - P_new Sigma_old ~ 0.1-0.3 after the first iterations of C2, 3 iterations reach the fp32 floor).
struct UmmaNs {
  int m = 0, ldm = 0;
  float* buf = nullptr;              // [P | Y | W0 | W1 | T], each [m][ldm] fp32
  double* resid = nullptr;           // [64] |I - Y P|_F^2 at the start of each iteration of the last umma_ns_iterate call
  double* P64 = nullptr;             // [m][m] fp64 P for the DMMA residual (the caller may pass its own matrix instead)
  void* maps = nullptr;
  float* P() const { return buf; }                                   // fp32 P, only read by the 3xTF32 residual (mode bit 0 clear)
  float* Y() const { return buf + (size_t)1 * m * ldm; }             // the iterate: read at entry, rewritten at exit (fixed address,
  float* W(int i) const { return buf + (size_t)(2 + i) * m * ldm; }  //   so a captured CUDA graph stays valid); W0 / W1: ping-pong work buffers
  float* T() const { return buf + (size_t)4 * m * ldm; }
};
int umma_ns_alloc(std::string* err, UmmaNs& ns, int m, cudaStream_t st);
// statistics of A (a_which, [M][m]) against the full symmetric ns.Y() on the v2 kernel: ep.mode = UMMA_EPI_STATS_SIGMA, ep.cin = A's matrix
int umma_gemm_sigma(std::string* err, UmmaLatent& u, UmmaNs& ns, int a_which, int M, const UmmaEpilogue& ep, cudaStream_t st);
void umma_ns_free(UmmaNs& ns);
// `iters` refinements of ns.Y() in place.  mode bit 0: residual T = I - Y P in fp64 on DMMA (ns_resid_f64_kernel, reads P64 with
// leading dimension ldp; P64 == nullptr -> ns.P64, ldp = m) instead of 3xTF32 (reads ns.P()); bit 1: symmetrise the result
int umma_ns_iterate(std::string* err, UmmaNs& ns, int iters, int mode, const double* P64, int64_t ldp, cudaStream_t st);

// ---- K_nm construction on the tensor core (agp_knm.cu) ----
struct UmmaKnm {
  int m = 0, D = 0, Kp = 0;                   // Kp = D rounded up to a multiple of 32
  float *Zhi = nullptr, *Zlo = nullptr;       // TF32 hi / lo split of the inducing points [m][Kp]
  void* maps = nullptr;                       // CUtensorMap: Zhi, Zlo (loads), Knm (store)
  int force_groups = 0;                       // > 0: override the column split (tuning / microbenchmarks)
};
bool umma_knm_shape_ok(int m, int Bcap, int D);
// (re)split Z and build the tensor maps (Knm: the engine's [Bcap][ldk] output buffer); call again whenever Z changes
int umma_knm_setup(std::string* err, UmmaKnm& k, const float* Z, int64_t ldz, int m, int D, float* Knm, int64_t ldk, int Bcap,
                   cudaStream_t st);
void umma_knm_free(UmmaKnm& k);
// Knm[b][j] = variance * base(scale2 * |x_b - z_j|^2) for B minibatch rows (gathered through `gather` when given)
int umma_knm(std::string* err, UmmaKnm& k, const float* X, int64_t ldx, int Dp, const int64_t* gather, const float* xx, const float* zz,
             int B, int kind, double scale2, double variance, cudaStream_t st);

}  // namespace agp
