// correctness + timing + timeline of the persistent m x m tail (agp_tail3.cuh) against the multi-launch tail (agp_tail2.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAGP_T3_TRACE tail3_test.cu -o tail3_test ; ./tail3_test [m] [nlat] [G] [trace]
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_tail3.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
using namespace agp;
template <typename K>
static void launch2(K kern, int grid, TailStepParams tp, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TAIL_THREADS); cfg.dynamicSmemBytes = TAIL2_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, tp);
}
static void launch_seq(TailStepParams tp, cudaStream_t st) {
  launch2(tail2_potf2_first_kernel<0>, 1, tp, st);
  for (int k = 0; k < tp.nblk; ++k) {
    int r = tp.nblk - 1 - k, tiles = r * (r + 1) / 2 + r * (k + 1) + k;
    if (!tiles) continue;
    tp.k = k;
    launch2(tail2_step_kernel, tiles, tp, st);
  }
}
int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 512, nblk = m / 64;
  const int nlat = argc > 2 ? atoi(argv[2]) : 1;
  int G = argc > 3 ? atoi(argv[3]) : 0;
  const int want_trace = argc > 4 ? atoi(argv[4]) : 0;
  if (G <= 0) G = std::max(2, std::min(tail3_max_tasks(nblk) + 1, 148 / nlat));
  if (nblk == 1) G = 1;
  std::vector<double> A((size_t)m * m), Gm((size_t)m * 96);
  srand(1);
  for (auto& g : Gm) g = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0; for (int k = 0; k < 96; ++k) s += Gm[(size_t)i * 96 + k] * Gm[(size_t)j * 96 + k];
    A[(size_t)i * m + j] = A[(size_t)j * m + i] = 40.0 * s + (i == j ? 1.0 : 0.0);
  }
  size_t bytes = (size_t)m * m * 8;
  double *dA; cudaMalloc(&dA, bytes); cudaMemcpy(dA, A.data(), bytes, cudaMemcpyHostToDevice);
  int* ds; cudaMalloc(&ds, 4); cudaMemset(ds, 0, 4);
  std::vector<Tail3Lat> h(nlat);
  std::vector<double*> dP(nlat), dW(nlat), dX(nlat), dD(nlat), dl(nlat);
  const size_t fw = tail3_flag_words(nblk);
  u64* dflags; cudaMalloc(&dflags, fw * nlat * 8); cudaMemset(dflags, 0, fw * nlat * 8);
  for (int q = 0; q < nlat; ++q) {
    cudaMalloc(&dP[q], bytes); cudaMalloc(&dW[q], bytes); cudaMalloc(&dX[q], bytes); cudaMalloc(&dD[q], (size_t)m * 64 * 8); cudaMalloc(&dl[q], 8);
    cudaMemset(dX[q], 0, bytes); cudaMemset(dW[q], 0, bytes); cudaMemset(dl[q], 0, 8);
    h[q] = Tail3Lat{dP[q], dW[q], dX[q], dD[q], dl[q], dflags + fw * q};
  }
  Tail3Lat* dlat; cudaMalloc(&dlat, nlat * sizeof(Tail3Lat)); cudaMemcpy(dlat, h.data(), nlat * sizeof(Tail3Lat), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(tail2_potf2_first_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaFuncSetAttribute(tail2_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaFuncSetAttribute(tail3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL3_SMEM);
  cudaStream_t st; cudaStreamCreate(&st);
  // reference: tail2 on latent 0's buffers
  double* dXref; cudaMalloc(&dXref, bytes); cudaMemset(dXref, 0, bytes);
  {
    TailStepParams tp{}; tp.P = dP[0]; tp.W = dW[0]; tp.Xout = dXref; tp.Dinv = dD[0]; tp.ld = m; tp.nblk = nblk; tp.logdet = dl[0]; tp.status = ds;
    cudaMemcpyAsync(dP[0], dA, bytes, cudaMemcpyDeviceToDevice, st);
    launch_seq(tp, st);
    cudaStreamSynchronize(st);
    cudaMemset(dl[0], 0, 8);
  }
  std::vector<double> Xr((size_t)m * m), X((size_t)m * m);
  cudaMemcpy(Xr.data(), dXref, bytes, cudaMemcpyDeviceToHost);
  Tail3Params p3{}; p3.lat = dlat; p3.nlat = nlat; p3.G = G; p3.nblk = nblk; p3.ld = m; p3.status = ds;
  auto run3 = [&]() {
    for (int q = 0; q < nlat; ++q) cudaMemcpyAsync(dP[q], dA, bytes, cudaMemcpyDeviceToDevice, st);
    tail3_kernel<<<nlat * G, TAIL_THREADS, TAIL3_SMEM, st>>>(p3);
  };
  for (int rep = 0; rep < 3; ++rep) run3();      // several epochs
  cudaStreamSynchronize(st);
  printf("tail3 m=%d nlat=%d G=%d grid=%d: %s\n", m, nlat, G, nlat * G, cudaGetErrorString(cudaGetLastError()));
  int stt; cudaMemcpy(&stt, ds, 4, cudaMemcpyDeviceToHost);
  for (int q = 0; q < nlat; q += std::max(1, nlat - 1)) {
    cudaMemcpy(X.data(), dX[q], bytes, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, up = 0;
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) {
      double d = X[(size_t)i * m + j] - Xr[(size_t)i * m + j]; num += d * d; den += Xr[(size_t)i * m + j] * Xr[(size_t)i * m + j];
      if (j > i) up = fmax(up, fabs(X[(size_t)i * m + j]));
    }
    printf("  latent %d: rel-Fro |X3 - X2| = %.3e  max|upper| = %.1e  status %d\n", q, sqrt(num / den), up, stt);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  // timing: kernel only (P restored by copies outside the event pair is not possible in a loop: time copy+kernel and copy alone)
  cudaEventRecord(e0, st); for (int w = 0; w < 20; ++w) run3(); cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  float ms_c;
  cudaEventRecord(e0, st); for (int w = 0; w < 20; ++w) for (int q = 0; q < nlat; ++q) cudaMemcpyAsync(dP[q], dA, bytes, cudaMemcpyDeviceToDevice, st);
  cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_c, e0, e1);
  printf("  tail3: %.1f us per launch (copy+kernel %.1f, copies %.1f)\n", (ms - ms_c) * 1000 / 20, ms * 1000 / 20, ms_c * 1000 / 20);
  {
    TailStepParams tp{}; tp.P = dP[0]; tp.W = dW[0]; tp.Xout = dXref; tp.Dinv = dD[0]; tp.ld = m; tp.nblk = nblk; tp.logdet = dl[0]; tp.status = ds;
    cudaEventRecord(e0, st);
    for (int w = 0; w < 20; ++w) { cudaMemcpyAsync(dP[0], dA, bytes, cudaMemcpyDeviceToDevice, st); launch_seq(tp, st); }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("  tail2 (PDL chain, 1 latent): %.1f us incl. copy\n", ms * 1000 / 20);
  }
#ifdef AGP_T3_TRACE
  if (want_trace) {
    int zero[160] = {0};
    cudaMemcpyToSymbol(agp_t3_trace_n, zero, sizeof(zero));
    run3(); cudaStreamSynchronize(st);
    std::vector<unsigned long long> tr((size_t)160 * T3_TRACE_SLOTS * 3); int n[160];
    cudaMemcpyFromSymbol(tr.data(), agp_t3_trace, tr.size() * 8); cudaMemcpyFromSymbol(n, agp_t3_trace_n, sizeof(n));
    unsigned long long g0 = ~0ull;
    for (int b = 0; b < std::min(nlat * G, 160); ++b) for (int e = 0; e < std::min(n[b], T3_TRACE_SLOTS); ++e) g0 = std::min(g0, tr[((size_t)b * T3_TRACE_SLOTS + e) * 3 + 2]);
    for (int b = 0; b < std::min(nlat * G, want_trace); ++b) {
      printf("cta %d:", b);
      unsigned long long c0 = n[b] ? tr[((size_t)b * T3_TRACE_SLOTS) * 3 + 1] : 0;
      for (int e = 0; e < std::min(n[b], T3_TRACE_SLOTS); ++e) {
        unsigned long long* r = &tr[((size_t)b * T3_TRACE_SLOTS + e) * 3];
        printf(" [%llu c%.2f g%.1f]", r[0], (double)((long long)(r[1] - c0)) / 1965.0, (double)(r[2] - g0) / 1000.0);
      }
      printf("\n");
    }
  }
#endif
  return 0;
}
