// standalone correctness + timing harness for the persistent tcgen05 3xTF32 GEMM (agp_umma.cu)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gemm_bench gemm_bench.cu -lcuda
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_umma.cu"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace agp;
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
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto fn, double flops) {
    for (int w = 0; w < 3; ++w) fn();
    cudaEventRecord(e0);
    const int reps = 20;
    for (int r = 0; r < reps; ++r) fn();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-10s %.2f us per launch, %.1f useful TFLOP/s   (%s)\n", name, ms * 1000 / reps, flops / (ms * 1e-3 / reps) / 1e12, cudaGetErrorString(cudaGetLastError()));
  };
  UmmaEpilogue ep1{}; ep1.mode = UMMA_EPI_STORE_SUMSQ; ep1.acc0 = acc;
  timeit("V=A*L^T", [&] { umma_gemm_nt(&err, u, UM_KNM, UM_LINV, dV, B, m, ep1, 0); }, 1.0 * B * m * m);
  UmmaEpilogue ep2{}; ep2.mode = UMMA_EPI_STATS_ONLY; ep2.acc0 = acc + B; ep2.acc1 = acc + 2 * B; ep2.tvec = tvec;
  timeit("V*X^T", [&] { umma_gemm_nt(&err, u, UM_V, UM_X, dVS, B, m, ep2, 0); }, 1.0 * B * m * m);
  std::vector<double> w(B, 1.0), gz(B, 0.0); double *dw, *dg, *dv1;
  cudaMalloc(&dw, B * 8); cudaMalloc(&dg, B * 8); cudaMalloc(&dv1, m * 8);
  cudaMemcpy(dw, w.data(), B * 8, cudaMemcpyHostToDevice); cudaMemcpy(dg, gz.data(), B * 8, cudaMemcpyHostToDevice);
  umma_scale_transpose(&err, u, dV, dw, 1.0, dg, dv1, B, m, 0);
  int ns = 32;
  timeit("gram", [&] { ns = 32; umma_gram(&err, u, dG, B, m, &ns, 0); }, 2.0 * B * m * m);
  // correctness: V against fp64 on sampled entries; gram against V^T V
  std::vector<float> V((size_t)B * m); cudaMemcpy(V.data(), dV, V.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int t = 0; t < 4000; ++t) {
    int b = rng() % B, j = rng() % m; double s = 0;
    for (int k = 0; k <= j; ++k) s += (double)A[(size_t)b * m + k] * L[(size_t)j * m + k];
    maxerr = std::max(maxerr, fabs(s - V[(size_t)b * m + j])); maxref = std::max(maxref, fabs(s));
  }
  printf("V: max abs err %.3e (max |ref| %.2f)\n", maxerr, maxref);
  std::vector<float> Gp((size_t)ns * m * m); cudaMemcpy(Gp.data(), dG, Gp.size() * 4, cudaMemcpyDeviceToHost);
  double gerr = 0, gref = 0;
  for (int t = 0; t < 300; ++t) {
    int i = rng() % m, j = rng() % m; double s = 0, gsum = 0;
    for (int b = 0; b < B; ++b) s += (double)V[(size_t)b * m + i] * V[(size_t)b * m + j];
    for (int sp = 0; sp < ns; ++sp) gsum += Gp[(size_t)sp * m * m + (size_t)i * m + j];
    gerr = std::max(gerr, fabs(s - gsum)); gref = std::max(gref, fabs(s));
  }
  printf("gram (%d splits): max abs err %.3e (max |ref| %.1f)\n", ns, gerr, gref);
  // ---- Gram product straight from V (umma_gram_tn_kernel): random weights / gradient vector, against fp64 ----
  if (u.gram_tn) {
    std::uniform_real_distribution<double> ud(0.0, 0.5);
    const double rho = 3.7;
    for (int b = 0; b < B; ++b) { w[b] = ud(rng); gz[b] = ud(rng) - 0.25; }
    w[5] = 0.0; w[B - 1] = -1e-3;   // clamped at zero
    cudaMemcpy(dw, w.data(), B * 8, cudaMemcpyHostToDevice); cudaMemcpy(dg, gz.data(), B * 8, cudaMemcpyHostToDevice);
    cudaMemset(dv1, 0, m * 8);
    int ns2 = 32;
    umma_gram_tn(&err, u, dG, dw, rho, dg, dv1, B, m, &ns2, 0);
    cudaDeviceSynchronize();
    printf("gram_tn launch: %s %s\n", cudaGetErrorString(cudaGetLastError()), err.c_str());
    std::vector<float> Gq((size_t)ns2 * m * m); cudaMemcpy(Gq.data(), dG, Gq.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<double> v1(m); cudaMemcpy(v1.data(), dv1, m * 8, cudaMemcpyDeviceToHost);
    double e2 = 0, r2 = 0, asym = 0;
    for (int t = 0; t < 400; ++t) {
      int i = rng() % m, j = rng() % m; double s = 0, gsum = 0, gsumT = 0;
      if (t < 8) { i = (t * 64 + 3) % m; j = i; }
      for (int b = 0; b < B; ++b) s += (double)V[(size_t)b * m + i] * V[(size_t)b * m + j] * std::max(rho * w[b], 0.0);
      for (int sp = 0; sp < ns2; ++sp) { gsum += Gq[(size_t)sp * m * m + (size_t)i * m + j]; gsumT += Gq[(size_t)sp * m * m + (size_t)j * m + i]; }
      e2 = std::max(e2, fabs(s - gsum)); r2 = std::max(r2, fabs(s)); asym = std::max(asym, fabs(gsum - gsumT));
    }
    double ev = 0, rv = 0;
    for (int j = 0; j < m; ++j) {
      double s = 0;
      for (int b = 0; b < B; ++b) s += (double)V[(size_t)b * m + j] * gz[b];
      ev = std::max(ev, fabs(s - v1[j])); rv = std::max(rv, fabs(s));
    }
    printf("gram_tn (%d splits): max abs err %.3e (max |ref| %.1f), asymmetry %.3e; V^T g: max abs err %.3e (max |ref| %.2f)\n", ns2, e2, r2, asym, ev, rv);
    timeit("gram_tn", [&] { ns2 = 32; umma_gram_tn(&err, u, dG, dw, rho, dg, dv1, B, m, &ns2, 0); }, 2.0 * B * m * m);
    timeit("scale_T", [&] { umma_scale_transpose(&err, u, dV, dw, rho, dg, dv1, B, m, 0); }, 0.0);
  }
  return 0;
}
