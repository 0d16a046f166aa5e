// correctness + timing of the m x m tail generations (agp_tail.cuh vs agp_tail2.cuh) on a random SPD matrix, standalone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tail2_test.cu -o tail2_test ; ./tail2_test [m]
#include "../../augmentedgaussianprocesses.jl_b200/csrc/agp_tail2.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
using namespace agp;
static bool g_pdl = false;
static int g_chain = 1;   // argv[2]: pivot-chain implementation (0 = scalar, 1 = look-ahead)
template <typename K>
static void launch2(K kern, int grid, TailStepParams tp, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TAIL_THREADS); cfg.dynamicSmemBytes = TAIL2_SMEM; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, tp);
}
static void launch_seq(int gen, TailStepParams tp, cudaStream_t st) {
  if (gen == 0) tail_potf2_first_kernel<0><<<1, TAIL_THREADS, TAIL_SMEM, st>>>(tp);
  else if (g_chain == 0) launch2(tail2_potf2_first_kernel<0, 0>, 1, tp, st);
  else launch2(tail2_potf2_first_kernel<0, 1>, 1, tp, st);
  for (int k = 0; k < tp.nblk; ++k) {
    int r = tp.nblk - 1 - k, tiles = r * (r + 1) / 2 + r * (k + 1) + k;
    if (!tiles) continue;
    tp.k = k;
    if (gen == 0) tail_step_kernel<0><<<tiles, TAIL_THREADS, TAIL_SMEM, st>>>(tp);
    else if (g_chain == 0) launch2(tail2_step_kernel<0>, tiles, tp, st);
    else launch2(tail2_step_kernel<1>, tiles, tp, st);
  }
}
int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 512, nblk = m / 64;
  g_chain = argc > 2 ? atoi(argv[2]) : 1;
  printf("pivot chain variant %d\n", g_chain);
  std::vector<double> A((size_t)m * m), G((size_t)m * 96);
  srand(1);
  for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < m; ++i) for (int j = 0; j <= i; ++j) {
    double s = 0; for (int k = 0; k < 96; ++k) s += G[(size_t)i * 96 + k] * G[(size_t)j * 96 + k];
    A[(size_t)i * m + j] = A[(size_t)j * m + i] = 40.0 * s + (i == j ? 1.0 : 0.0);     // cond ~ 1e3: I + low-rank-ish PSD
  }
  // host reference: Cholesky + inverse of the factor
  std::vector<double> R(A), Xr((size_t)m * m, 0.0);
  for (int j = 0; j < m; ++j) {
    double d = R[(size_t)j * m + j]; for (int k = 0; k < j; ++k) d -= R[(size_t)j * m + k] * R[(size_t)j * m + k];
    d = sqrt(d); R[(size_t)j * m + j] = d;
    for (int i = j + 1; i < m; ++i) { double s = R[(size_t)i * m + j]; for (int k = 0; k < j; ++k) s -= R[(size_t)i * m + k] * R[(size_t)j * m + k]; R[(size_t)i * m + j] = s / d; }
  }
  for (int c = 0; c < m; ++c) for (int i = c; i < m; ++i) {
    double s = (i == c); for (int k = c; k < i; ++k) s -= R[(size_t)i * m + k] * Xr[(size_t)k * m + c];
    Xr[(size_t)i * m + c] = s / R[(size_t)i * m + i];
  }
  double ldref = 0; for (int j = 0; j < m; ++j) ldref += 2 * log(R[(size_t)j * m + j]);
  double *dA, *dP, *dW, *dX, *dD, *dl; int* ds;
  size_t bytes = (size_t)m * m * 8;
  cudaMalloc(&dA, bytes); cudaMalloc(&dP, bytes); cudaMalloc(&dW, bytes); cudaMalloc(&dX, bytes); cudaMalloc(&dD, (size_t)m * 64 * 8); cudaMalloc(&dl, 8); cudaMalloc(&ds, 4);
  cudaMemcpy(dA, A.data(), bytes, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(tail_potf2_first_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  cudaFuncSetAttribute(tail_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM);
  cudaFuncSetAttribute(tail2_potf2_first_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaFuncSetAttribute(tail2_potf2_first_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaFuncSetAttribute(tail2_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaFuncSetAttribute(tail2_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM);
  cudaStream_t st; cudaStreamCreate(&st);
  TailStepParams tp{}; tp.P = dP; tp.W = dW; tp.Xout = dX; tp.Dinv = dD; tp.ld = m; tp.nblk = nblk; tp.logdet = dl; tp.status = ds;
  for (int gen = 0; gen < 3; ++gen) {
    g_pdl = gen == 2;
    if (gen == 2) printf("(gen 2 = gen 1 launched with programmatic dependent launch)\n");
    cudaMemset(dX, 0, bytes); cudaMemset(dW, 0, bytes); cudaMemset(dl, 0, 8); cudaMemset(ds, 0, 4);
    cudaMemcpyAsync(dP, dA, bytes, cudaMemcpyDeviceToDevice, st);
    launch_seq(gen, tp, st);
    cudaStreamSynchronize(st);
    std::vector<double> X((size_t)m * m); double ld; int stt;
    cudaMemcpy(X.data(), dX, bytes, cudaMemcpyDeviceToHost); cudaMemcpy(&ld, dl, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&stt, ds, 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, up = 0;
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) {
      double d = X[(size_t)i * m + j] - Xr[(size_t)i * m + j]; num += d * d; den += Xr[(size_t)i * m + j] * Xr[(size_t)i * m + j];
      if (j > i) up = fmax(up, fabs(X[(size_t)i * m + j]));
    }
    // time: graph of (copy + sequence), 20 replays
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
    cudaMemcpyAsync(dP, dA, bytes, cudaMemcpyDeviceToDevice, st);
    launch_seq(gen, tp, st);
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(e0, st);
    for (int w = 0; w < 20; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // the first kernel alone
    cudaEventRecord(e0, st);
    for (int w = 0; w < 20; ++w) { if (gen == 0) tail_potf2_first_kernel<0><<<1, TAIL_THREADS, TAIL_SMEM, st>>>(tp); else tail2_potf2_first_kernel<0><<<1, TAIL_THREADS, TAIL2_SMEM, st>>>(tp); }
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms1; cudaEventElapsedTime(&ms1, e0, e1);
    printf("gen %d  m=%d: rel-Fro |X - Xref| = %.3e  max|upper| = %.1e  logdet err = %.3e  status %d | chol_inv %.1f us per graph replay (incl. %zu KB copy), 64x64 potf2+inv alone %.2f us  [%s]\n",
           gen, m, sqrt(num / den), up, fabs(ld - ldref) / fabs(ldref), stt, ms * 1000 / 20, bytes >> 10, ms1 * 1000 / 20, cudaGetErrorString(cudaGetLastError()));
  }
  // ablation: time of the 64 x 64 tile kernel with phases removed (20 back-to-back launches each; includes ~2 us launch gap)
  cudaEvent_t a0, a1; cudaEventCreate(&a0); cudaEventCreate(&a1);
#define ABLRUN(M, name) { cudaFuncSetAttribute(tail2_potf2_first_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL2_SMEM); \
    for (int w = 0; w < 3; ++w) tail2_potf2_first_kernel<M><<<1, TAIL_THREADS, TAIL2_SMEM, st>>>(tp); \
    cudaEventRecord(a0, st); for (int w = 0; w < 20; ++w) tail2_potf2_first_kernel<M><<<1, TAIL_THREADS, TAIL2_SMEM, st>>>(tp); cudaEventRecord(a1, st); cudaEventSynchronize(a1); \
    float ms_; cudaEventElapsedTime(&ms_, a0, a1); printf("  %-34s %.2f us\n", name, ms_ * 1000 / 20); }
  cudaMemcpy(dP, dA, bytes, cudaMemcpyDeviceToDevice);
  ABLRUN(0, "full");
  ABLRUN(1, "no pivot chain");
  ABLRUN(2, "no strip/row-block (b)");
  ABLRUN(4, "no trailing (c)");
  ABLRUN(8, "no result stores");
  ABLRUN(15, "nothing (load + init + barriers)");
  ABLRUN(14, "chain only");
  return 0;
}
