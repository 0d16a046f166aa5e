// standalone timing harness for the tcgen05 K_nm kernel (agp_knm.cu): per-CTA phase timestamps (globaltimer) + launch time
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAGP_KNM_TIMING -o knm_bench knm_bench.cu -lcuda
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_knm.cu"
namespace agp { bool umma_shape_ok(int m, int Bcap) { return m % 128 == 0 && Bcap % 128 == 0; } }
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace agp;
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000, D = argc > 2 ? atoi(argv[2]) : 32, m = argc > 3 ? atoi(argv[3]) : 512, B = argc > 4 ? atoi(argv[4]) : 8192;
  const int Dp = (D + 3) / 4 * 4;
  std::mt19937 rng(1);
  std::normal_distribution<float> nd;
  std::vector<float> X((size_t)n * Dp, 0.f), xx(n), Z((size_t)m * Dp, 0.f), zz(m);
  for (int i = 0; i < n; ++i) { float s = 0; for (int k = 0; k < D; ++k) { float v = nd(rng); X[(size_t)i * Dp + k] = v; s += v * v; } xx[i] = s; }
  std::vector<int64_t> idx(B);
  for (int b = 0; b < B; ++b) idx[b] = rng() % n;
  for (int j = 0; j < m; ++j) { float s = 0; for (int k = 0; k < D; ++k) { float v = X[(size_t)idx[j] * Dp + k] + 0.1f * nd(rng); Z[(size_t)j * Dp + k] = v; s += v * v; } zz[j] = s; }
  std::vector<float> xxb(B);
  for (int b = 0; b < B; ++b) xxb[b] = xx[idx[b]];
  float *dX, *dxx, *dZ, *dzz, *dK; int64_t* didx;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dxx, B * 4); cudaMalloc(&dZ, Z.size() * 4); cudaMalloc(&dzz, m * 4); cudaMalloc(&dK, (size_t)B * m * 4); cudaMalloc(&didx, B * 8);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dxx, xxb.data(), B * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dzz, zz.data(), m * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(didx, idx.data(), B * 8, cudaMemcpyHostToDevice);
  std::string err; UmmaKnm k;
  if (umma_knm_setup(&err, k, dZ, Dp, m, D, dK, m, B, 0)) { printf("setup: %s\n", err.c_str()); return 1; }
  const double s2 = 1.0 / D;
  k.force_groups = argc > 5 ? atoi(argv[5]) : 0;
  const int kind = argc > 6 ? atoi(argv[6]) : 0;
  // L2 flush buffer
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float tot = 0; const int reps = 20;
  for (int r = 0; r < reps + 3; ++r) {
    cudaMemsetAsync(flush, r, 256 << 20);
    cudaEventRecord(e0);
    if (umma_knm(&err, k, dX, Dp, Dp, didx, dxx, dzz, B, kind, s2, 1.0, 0)) { printf("launch: %s\n", err.c_str()); return 1; }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r >= 3) tot += ms;
  }
  printf("knm B=%d m=%d D=%d: %.2f us per launch (event pair, L2 flushed) err=%s\n", B, m, D, tot * 1000 / reps, cudaGetErrorString(cudaGetLastError()));
#ifdef AGP_KNM_TIMING
  const int nct = std::min(1024, (B / 128) * (k.force_groups > 0 ? k.force_groups : 1));
  std::vector<unsigned long long> t(1024 * 8);
  cudaMemcpyFromSymbol(t.data(), agp_knm_t, sizeof(unsigned long long) * 1024 * 8);
  unsigned long long t0 = ~0ull;
  for (int c = 0; c < nct; ++c) t0 = std::min(t0, t[c * 8]);
  double avg[6] = {0}, mx[6] = {0};
  for (int c = 0; c < nct; ++c) for (int s = 0; s < 6; ++s) { double v = (double)(t[c * 8 + s] - t0); avg[s] += v / nct; mx[s] = std::max(mx[s], v); }
  const char* nm[6] = {"entry", "setup done (tmem alloc, sync)", "A gathered+split+stored", "tmem_full (MMA done)", "epilogue issued", "stores drained"};
  for (int s = 0; s < 6; ++s) printf("  %-32s avg %7.0f ns   max %7.0f ns\n", nm[s], avg[s], mx[s]);
#endif
  // correctness spot check
  std::vector<float> K((size_t)B * m); cudaMemcpy(K.data(), dK, K.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int t2 = 0; t2 < 2000; ++t2) {
    int b = rng() % B, j = rng() % m; double d2 = 0;
    for (int kk = 0; kk < D; ++kk) { double d = (double)X[(size_t)idx[b] * Dp + kk] - Z[(size_t)j * Dp + kk]; d2 += d * d; }
    double dd = sqrt(s2 * d2), ref = kind == 0 ? exp(-0.5 * s2 * d2) : kind == 1 ? (1 + sqrt(3.0) * dd) * exp(-sqrt(3.0) * dd) : (1 + sqrt(5.0) * dd + 5.0 / 3.0 * s2 * d2) * exp(-sqrt(5.0) * dd);
    maxerr = std::max(maxerr, fabs(ref - K[(size_t)b * m + j]));
  }
  printf("max abs err vs fp64 (2000 samples): %.3e\n", maxerr);
  return 0;
}
