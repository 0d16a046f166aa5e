// per-role timeline of the persistent tcgen05 3xTF32 GEMM (agp_umma.cu, umma_gemm_nt_kernel) at C2 sizes
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAGP_UMMA_TRACE -o gemm_trace gemm_trace.cu -lcuda
// prints, for a few CTAs, the merged event list (role, tag, cycles since the CTA's first event); tags: see agp_umma.cu (UTT)
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_umma.cu"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace agp;
static void dump(const char* name, int grid) {
  std::vector<unsigned long long> tr((size_t)160 * 4 * UT_SLOTS * 2);
  std::vector<int> nn(160 * 4);
  cudaMemcpyFromSymbol(tr.data(), agp_ut_trace, tr.size() * 8);
  cudaMemcpyFromSymbol(nn.data(), agp_ut_n, nn.size() * 4);
  printf("=== %s (grid %d)\n", name, grid);
  const int ctas[] = {0, 1, grid / 2, grid - 1};
  for (int ci = 0; ci < 4; ++ci) {
    const int c = ctas[ci];
    struct Ev { unsigned long long t; int role, tag; };
    std::vector<Ev> ev;
    for (int r = 0; r < 4; ++r)
      for (int i = 0; i < std::min(nn[c * 4 + r], UT_SLOTS); ++i) {
        const unsigned long long* p = tr.data() + (((size_t)c * 4 + r) * UT_SLOTS + i) * 2;
        ev.push_back({p[1], r, (int)p[0]});
      }
    std::sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.t < b.t; });
    if (ev.empty()) continue;
    printf("cta %d:", c);
    for (auto& e : ev) printf(" [r%d %d @%llu]", e.role, e.tag, e.t - ev[0].t);
    printf("\n");
  }
  std::vector<int> z(160 * 4, 0);
  cudaMemcpyToSymbol(agp_ut_n, z.data(), z.size() * 4);
}
int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 8192, m = argc > 2 ? atoi(argv[2]) : 512;
  std::mt19937 rng(3);
  std::normal_distribution<float> nd;
  std::vector<float> A((size_t)B * m), L((size_t)m * m, 0.f), X((size_t)m * m, 0.f);
  for (auto& v : A) v = nd(rng);
  for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) { L[(size_t)i * m + j] = nd(rng) / sqrtf((float)m); X[(size_t)i * m + j] = nd(rng) / sqrtf((float)m); }
  float *dA, *dV, *dL, *dX, *dVS, *dG; double *acc, *tvec;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dV, A.size() * 4); cudaMalloc(&dVS, A.size() * 4); cudaMalloc(&dL, L.size() * 4); cudaMalloc(&dX, X.size() * 4);
  cudaMalloc(&dG, (size_t)32 * m * m * 4); cudaMalloc(&acc, 3 * B * 8); cudaMalloc(&tvec, m * 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dL, L.data(), L.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice); cudaMemset(acc, 0, 3 * B * 8); cudaMemset(tvec, 0, m * 8);
  std::string err; UmmaLatent u;
  if (umma_latent_alloc(&err, u, m, m, B, dA, dV, dL, dX, 0)) { printf("alloc: %s\n", err.c_str()); return 1; }
  umma_presplit(&err, u, UM_LINV, dL, 0); umma_presplit(&err, u, UM_X, dX, 0);   // pre-split right operands (no-op unless AGP_UMMA_PS / V2)
  printf("pre-split right operand: %d\n", u.ps);
  std::vector<int> z(160 * 4, 0);
  UmmaEpilogue ep1{}; ep1.mode = UMMA_EPI_STORE_SUMSQ; ep1.acc0 = acc;
  UmmaEpilogue ep2{}; ep2.mode = UMMA_EPI_STATS_ONLY; ep2.acc0 = acc + B; ep2.acc1 = acc + 2 * B; ep2.tvec = tvec;
  for (int w = 0; w < 3; ++w) { umma_gemm_nt(&err, u, UM_KNM, UM_LINV, dV, B, m, ep1, 0); umma_gemm_nt(&err, u, UM_V, UM_X, dVS, B, m, ep2, 0); }
  cudaDeviceSynchronize();
  cudaMemcpyToSymbol(agp_ut_n, z.data(), z.size() * 4);
  umma_gemm_nt(&err, u, UM_KNM, UM_LINV, dV, B, m, ep1, 0);
  cudaDeviceSynchronize();
  dump("V = A L^T (triangular B, store + sumsq)", 148);
  umma_gemm_nt(&err, u, UM_V, UM_X, dVS, B, m, ep2, 0);
  cudaDeviceSynchronize();
  dump("V X^T (triangular B, statistics only)", 148);
  std::vector<double> w(B, 1.0), gz(B, 0.0); double *dw, *dg, *dv1;
  cudaMalloc(&dw, B * 8); cudaMalloc(&dg, B * 8); cudaMalloc(&dv1, m * 8);
  cudaMemcpy(dw, w.data(), B * 8, cudaMemcpyHostToDevice); cudaMemcpy(dg, gz.data(), B * 8, cudaMemcpyHostToDevice);
  umma_scale_transpose(&err, u, dV, dw, 1.0, dg, dv1, B, m, 0);
  int ns = 32;
  umma_gram(&err, u, dG, B, m, &ns, 0);
  cudaDeviceSynchronize();
  cudaMemcpyToSymbol(agp_ut_n, z.data(), z.size() * 4);
  ns = 32; umma_gram(&err, u, dG, B, m, &ns, 0);
  cudaDeviceSynchronize();
  dump("Gram U^T U (split-K, mirror)", 148);
  if (u.gram_tn) {
    ns = 32; umma_gram_tn(&err, u, dG, dw, 1.0, dg, dv1, B, m, &ns, 0);
    cudaDeviceSynchronize();
    cudaMemcpyToSymbol(agp_ut_n, z.data(), z.size() * 4);
    ns = 32; umma_gram_tn(&err, u, dG, dw, 1.0, dg, dv1, B, m, &ns, 0);
    cudaDeviceSynchronize();
    dump("Gram straight from V (umma_gram_tn_kernel)", 148);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
