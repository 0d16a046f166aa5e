// standalone timing / probe harness for tile_potf2_inv (agp_tail.cuh)
#ifndef AGP_TAIL_DEBUG_T
#define AGP_TAIL_DEBUG_T 255
#endif
#define AGP_TAIL_DEBUG 1
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_tail.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace agp;
int main() {
  const int n = 64;
  std::vector<double> A(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 2.0 + i * 0.01 : 0.0) + 0.5 * cos(0.37 * (i - j)) / (1 + abs(i - j));
  double *dP, *dX, *dD, *dl; int* ds;
  cudaMalloc(&dP, n * n * 8); cudaMalloc(&dX, n * n * 8); cudaMalloc(&dD, n * n * 8); cudaMalloc(&dl, 8); cudaMalloc(&ds, 4);
  cudaMemcpy(dP, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(dl, 0, 8); cudaMemset(ds, 0, 4);
  cudaFuncSetAttribute(tail_potf2_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  TailStepParams p{}; p.P = dP; p.W = nullptr; p.Xout = dX; p.Dinv = dD; p.ld = n; p.nblk = 1; p.logdet = dl; p.status = ds;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) tail_potf2_first_kernel<<<1, POTF2_THREADS, TAIL_SMEM>>>(p);
  cudaEventRecord(e0);
  for (int w = 0; w < 20; ++w) tail_potf2_first_kernel<<<1, POTF2_THREADS, TAIL_SMEM>>>(p);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("potf2_inv 64x64: %.2f us per launch (20 back-to-back), err=%s\n", ms * 1000 / 20, cudaGetErrorString(cudaGetLastError()));
  long long dbg[256]; cudaMemcpyFromSymbol(dbg, agp_dbg, sizeof(dbg));
  for (int j = 0; j < 64; j += 1) {
    long long pre = dbg[j * 4 + 0], post = dbg[j * 4 + 1], end = dbg[j * 4 + 2];
    long long prev_end = j ? dbg[(j - 1) * 4 + 2] : pre; long long fdone = dbg[j * 4 + 3];
    if (0) printf("pivot %2d: publish %5lld  barrier-wait %5lld  lds+f %5lld fma %5lld   total %5lld\n", j, pre - prev_end, post - pre, fdone - post, end - fdone, end - prev_end);
  }
  long long own[256]; cudaMemcpyFromSymbol(own, agp_own, sizeof(own));
  for (int j = 1; j < 1; j += 7) printf("owner pivot %2d: since prev owner-done %5lld | newton %5lld | publish r %5lld\n", j, own[j*4+0]-own[(j-1)*4+2], own[j*4+1]-own[j*4+0], own[j*4+2]-own[j*4+1]);
  // check X * A * X^T = I
  std::vector<double> X(n * n); cudaMemcpy(X.data(), dX, n * n * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int a = 0; a < n; ++a) for (int b = 0; b < n; ++b) s += X[i * n + a] * A[a * n + b] * X[j * n + b];
    maxerr = fmax(maxerr, fabs(s - (i == j)));
  }
  printf("max |X A X^T - I| = %.3e\n", maxerr);
  return 0;
}
