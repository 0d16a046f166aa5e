// placeholder until the tcgen05 kernels land (next commit): AGP_PREC_TF32X3 fails loudly.
#include "agp_umma.h"
namespace agp {
bool umma_shape_ok(int, int) { return false; }
int umma_latent_alloc(std::string* err, UmmaLatent&, int, int, int, cudaStream_t) { *err = "tcgen05 path not built"; return 5; }
void umma_latent_free(UmmaLatent&) {}
int umma_split_matrix(std::string* err, UmmaLatent&, int, const float*, int, cudaStream_t) { *err = "tcgen05 path not built"; return 5; }
int umma_gemm_nt(std::string* err, UmmaLatent&, int, int, float*, int, int, cudaStream_t) { *err = "tcgen05 path not built"; return 5; }
int umma_gram(std::string* err, UmmaLatent&, const float*, const double*, double, float*, int, int, int*, cudaStream_t) { *err = "tcgen05 path not built"; return 5; }
}
