// Interface of the tcgen05 (5th-gen tensor core) path for the B x m contractions: 3xTF32 split GEMMs with
// TMA-staged operands and TMEM accumulators.  Implemented in agp_umma.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace agp {

enum UmmaMat : int { UM_KNM = 0, UM_V = 1, UM_LINV = 2, UM_X = 3, UM_COUNT = 4 };

struct UmmaLatent {
  int m = 0, ldm = 0, Bcap = 0;
  float* hi[UM_COUNT] = {nullptr, nullptr, nullptr, nullptr};  // TF32-rounded high parts
  float* lo[UM_COUNT] = {nullptr, nullptr, nullptr, nullptr};  // residuals
  float* kT_hi = nullptr; float* kT_lo = nullptr;              // V^T (m x B) and diag(w)-scaled copy for the Gram product
  float* kTw_hi = nullptr; float* kTw_lo = nullptr;
  void* tmaps = nullptr;                                        // host array of CUtensorMap
};

bool umma_shape_ok(int m, int Bcap);
int umma_latent_alloc(std::string* err, UmmaLatent& u, int m, int ldm, int Bcap, cudaStream_t st);
void umma_latent_free(UmmaLatent& u);
// split src (rows x m, ld = ldm) into hi/lo of matrix `which`
int umma_split_matrix(std::string* err, UmmaLatent& u, int which, const float* src, int rows, cudaStream_t st);
// C[M x N] = A[M x K] * B[N x K]^T with K = m, 3xTF32
int umma_gemm_nt(std::string* err, UmmaLatent& u, int a_which, int b_which, float* C, int M, int N, cudaStream_t st);
// Gpart[s] = sum_{b in split s} rho*w_b kappa_b kappa_b^T ; *n_split in: capacity, out: used
int umma_gram(std::string* err, UmmaLatent& u, const float* kappa, const double* w, double rho, float* Gpart, int B, int m,
              int* n_split, cudaStream_t st);

}  // namespace agp
