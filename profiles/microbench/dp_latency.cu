// fp64 pipe microbenchmark for the m x m tail design (run on the B200 box): dependent-chain latency, per-SM throughput,
// barrier cost, shared-memory round trip.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 dp_latency.cu -o dp_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_chain(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_chain_f(float* out, long long* cyc, float a, float b) {
  float x = a;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fmaf(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_tput(double* out, long long* cyc, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = a + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 128; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], b, a);
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_sync(long long* cyc) {
  long long t0 = clock64();
  for (int i = 0; i < 256; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* cyc, double a) {
  double x = a;
  long long t0 = clock64();
  for (int i = 0; i < 64; ++i) x = rsqrt(x) + 1.5;
  long long t1 = clock64();
  double y = a;
  for (int i = 0; i < 64; ++i) { double r = (double)rsqrtf((float)y); r = r * (1.5 - 0.5 * y * r * r); r = r * (1.5 - 0.5 * y * r * r); y = r + 1.5; }
  long long t2 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
__global__ void k_smem_rt(double* out, long long* cyc) {
  __shared__ double s[256];
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < 128; ++i) { s[threadIdx.x] = x; __syncthreads(); x = s[(threadIdx.x + 1) & 255] + 1.0; }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; float* outf;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&outf, 1 << 20); cudaMalloc(&cyc, 64);
  long long h[2];
  k_chain<<<1, 32>>>(out, cyc, 1.0, 0.999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.1f cyc\n", h[0] / 256.0);
  k_chain_f<<<1, 32>>>(outf, cyc, 1.0f, 0.999f); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("FFMA dependent latency: %.1f cyc\n", h[0] / 256.0);
  for (int nt : {32, 128, 256, 512, 1024}) {
    k_tput<<<1, nt>>>(out, cyc, 1.0, 0.999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA throughput %4d threads: %.2f DFMA/clk/SM\n", nt, 128.0 * 8 * nt / h[0]);
  }
  for (int nt : {64, 256, 1024}) { k_sync<<<1, nt>>>(cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("__syncthreads %4d threads: %.1f cyc\n", nt, h[0] / 256.0); }
  k_rsqrt<<<1, 32>>>(out, cyc, 2.0); cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost); printf("rsqrt(double)+add chain: %.1f cyc ; rsqrtf+2 Newton+add chain: %.1f cyc\n", h[0] / 64.0, h[1] / 64.0);
  k_smem_rt<<<1, 256>>>(out, cyc); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("STS -> bar -> LDS -> DADD round trip (256 thr): %.1f cyc\n", h[0] / 128.0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock %d kHz\n", clk);
  return 0;
}
