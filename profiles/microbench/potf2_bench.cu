// standalone timing / correctness harness for the 64 x 64 Cholesky+inverse tile variants of agp_tail.cuh (run on the B200 box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 potf2_bench.cu -o potf2_bench
#include "potf2_variants.cuh"
#include <cstdio>
#include <vector>
#include <cmath>
using namespace agp;
template <int VAR>
void launch1(const TailStepParams& p) {
  if (VAR == 0) tail_potf2_first_kernel<0><<<1, POTF2_THREADS, TAIL_SMEM>>>(p);
  else if (VAR == 1) potf2_pipe_kernel<2><<<1, POTF2_THREADS, TAIL_SMEM>>>(p);
  else potf2_pipe_kernel<1><<<1, POTF2_THREADS, TAIL_SMEM>>>(p);
}
template <int VAR>
void run(const char* name, const TailStepParams& p, const std::vector<double>& A, int n) {
  cudaFuncSetAttribute(tail_potf2_first_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  cudaFuncSetAttribute(potf2_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  cudaFuncSetAttribute(potf2_pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) launch1<VAR>(p);
  cudaEventRecord(e0);
  for (int w = 0; w < 50; ++w) launch1<VAR>(p);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> X(n * n); cudaMemcpy(X.data(), p.Xout, n * n * 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  std::vector<double> T(n * n);
  for (int i = 0; i < n; ++i) for (int b = 0; b < n; ++b) { double s = 0; for (int a = 0; a < n; ++a) s += X[i * n + a] * A[a * n + b]; T[i * n + b] = s; }
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int b = 0; b < n; ++b) s += T[i * n + b] * X[j * n + b]; maxerr = fmax(maxerr, fabs(s - (i == j))); }
  printf("%-28s %.2f us per launch (50 back-to-back) = %.0f cycles/pivot @1.965GHz   max|X A X^T - I| = %.3e  err=%s\n", name, ms * 1000 / 50,
         ms * 1000 / 50 * 1965 / 64, maxerr, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int n = 64;
  std::vector<double> A(n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) A[i * n + j] = (i == j ? 2.0 + i * 0.01 : 0.0) + 0.5 * cos(0.37 * (i - j)) / (1 + abs(i - j));
  double *dP, *dX, *dD, *dl; int* ds;
  cudaMalloc(&dP, n * n * 8); cudaMalloc(&dX, n * n * 8); cudaMalloc(&dD, n * n * 8); cudaMalloc(&dl, 8); cudaMalloc(&ds, 4);
  cudaMemcpy(dP, A.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemset(dl, 0, 8); cudaMemset(ds, 0, 4);
  TailStepParams p{}; p.P = dP; p.W = nullptr; p.Xout = dX; p.Dinv = dD; p.ld = n; p.nblk = 1; p.logdet = dl; p.status = ds;
  run<0>("barrier-per-pivot (old)", p, A, n);
  cudaMemset(dX, 0, n * n * 8);
  run<1>("pipelined, 2 Newton", p, A, n);
  cudaMemset(dX, 0, n * n * 8);
  run<2>("pipelined, 1 Newton", p, A, n);
  int st; cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost); printf("status %d\n", st);
  return 0;
}
