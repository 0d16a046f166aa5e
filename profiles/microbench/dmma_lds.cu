// microbenchmarks for the fp64 tile products of the m x m tail: DMMA (mma.sync.m8n8k4.f64) throughput / latency and
// shared-memory broadcast bandwidth.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 dmma_lds.cu -o dmma_lds
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NCH>
__global__ void k_dmma(double* out, long long* cyc, double a, double b) {
  double c[NCH][2];
  for (int i = 0; i < NCH; ++i) { c[i][0] = i; c[i][1] = -i; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it)
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma(c[i][0], c[i][1], a, b);
  __syncthreads();
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NCH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ __align__(16) double s[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = i;
  __syncthreads();
  double acc = 0;
  int lane = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) acc += s[(it + u * 8) & 1023];                       // LDS.64, all lanes the same address
      else if (MODE == 1) acc += s[((it + u * 32) & 1023) + lane];        // LDS.64, 32 distinct consecutive
      else if (MODE == 2) { double2 v = *reinterpret_cast<const double2*>(&s[((it * 2 + u * 16) & 1022)]); acc += v.x + v.y; }  // LDS.128 same address
      else { double2 v = *reinterpret_cast<const double2*>(&s[((it * 2 + u * 64) & 1023) + 2 * lane]); acc += v.x + v.y; }      // LDS.128 distinct
    }
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64);
  long long h;
  k_dmma<1><<<1, 32>>>(out, cyc, 1.0, 0.5); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("DMMA m8n8k4 dependent latency: %.1f cyc\n", h / 64.0);
  for (int nt : {32, 128, 256, 512}) {
    k_dmma<8><<<1, nt>>>(out, cyc, 1.0, 0.5); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA throughput %4d threads (8 chains/warp): %.1f FMA/clk/SM  (%.2f cyc per warp-DMMA per SMSP)\n", nt, 64.0 * 8 * (nt / 32) * 256 / h,
           (double)h / (64.0 * 8 * ((nt / 32 + 3) / 4)));
  }
  const char* names[4] = {"LDS.64 same address", "LDS.64 32 distinct", "LDS.128 same address", "LDS.128 32 distinct"};
  for (int nt : {32, 256}) {
    k_lds<0><<<1, nt>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %4d thr: %.2f cyc per warp-load per SM\n", names[0], nt, h / (64.0 * 16 * (nt / 32)));
    k_lds<1><<<1, nt>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %4d thr: %.2f cyc per warp-load per SM\n", names[1], nt, h / (64.0 * 16 * (nt / 32)));
    k_lds<2><<<1, nt>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %4d thr: %.2f cyc per warp-load per SM\n", names[2], nt, h / (64.0 * 16 * (nt / 32)));
    k_lds<3><<<1, nt>>>(out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-22s %4d thr: %.2f cyc per warp-load per SM\n", names[3], nt, h / (64.0 * 16 * (nt / 32)));
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
