// single-warp 16 x 16 pivot-chain variants (phase (a) of tile2_potf2_inv): where do the cycles per pivot go?
#include <cstdio>
#include <cuda_runtime.h>
template <int NEWTON>
__device__ __forceinline__ double rcpc(double d) {
  double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
  for (int i = 0; i < NEWTON; ++i) y = fma(y, fma(-d, y, 1.0), y);
  return y;
}
// MODE 0: as shipped (status check, 2 Newton)  1: no status check  2: no check, 1 Newton  3: d by shuffle, col by smem
//      4: everything by shuffle (no smem)  5: arithmetic only (no exchange at all; wrong numerics, timing floor)
//      6: rcp of fp32 seed (MUFU.RCP f32 + 2 Newton)   7: like 1 but col double-buffer replaced by 16 separate columns (no WAR)
template <int MODE>
__global__ void k_chain(const double* A, double* out, long long* cyc, int* status) {
  __shared__ __align__(16) double col[16 * 16];
  __shared__ double dv[16];
  const int lane = threadIdx.x, rr = lane & 15; const bool arow = lane < 16;
  double x[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) { int hi = rr > q ? rr : q, lo = rr > q ? q : rr; x[q] = arow ? A[hi * 16 + lo] : (q == rr ? 1.0 : 0.0); }
  __syncwarp();
  long long t0 = clock64();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    double* cj = col + ((MODE == 7) ? j * 16 : (j & 1) * 16);
    double d, cq[16];
    if (MODE <= 3 || MODE >= 6) { if (arow) cj[rr] = x[j]; __syncwarp(); }
    if (MODE == 3 || MODE == 4) d = __shfl_sync(0xffffffffu, x[j], j);
    else if (MODE == 5) d = x[j] + 2.0;
    else d = cj[j];
    if (MODE == 0) { if (!(d > 0.0)) { if (lane == 0) atomicOr(status, 2); d = 1.0; } }
    if (lane == j) dv[j] = d;
    double inv;
    if (MODE == 2) inv = rcpc<1>(d);
    else if (MODE == 6) { float f = __frcp_rn((float)d); double y = (double)f; y = fma(y, fma(-d, y, 1.0), y); inv = fma(y, fma(-d, y, 1.0), y); }
    else inv = rcpc<2>(d);
    double g = -x[j] * inv;
    if (arow && rr <= j) g = 0.0;
#pragma unroll
    for (int q = j + 1; q < 16; ++q) {
      if (MODE == 4) cq[q] = __shfl_sync(0xffffffffu, x[j], q);
      else if (MODE == 5) cq[q] = x[q] * 0.5;
      else cq[q] = cj[q];
    }
#pragma unroll
    for (int q = j + 1; q < 16; ++q) x[q] = fma(cq[q], g, x[q]);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += x[q];
  out[lane] = s + dv[rr];
  if (lane == 0) cyc[0] = t1 - t0;
}
int main() {
  double hA[256];
  for (int i = 0; i < 16; ++i) for (int j = 0; j < 16; ++j) hA[i * 16 + j] = (i == j ? 3.0 : 0.0) + 0.3 / (1 + abs(i - j));
  double *dA, *out; long long* cyc; int* st;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&out, 256); cudaMalloc(&cyc, 8); cudaMalloc(&st, 4); cudaMemset(st, 0, 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  long long h; double o[32];
#define RUN(M, name) for (int rep = 0; rep < 2; ++rep) k_chain<M><<<1, 32>>>(dA, out, cyc, st); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o, out, 256, cudaMemcpyDeviceToHost); \
  printf("%-58s %6.1f cycles per pivot   (check %.6f)\n", name, h / 16.0, o[17]);
  RUN(0, "0 shipped: smem col+d, status check, 2 Newton");
  RUN(1, "1 no status check");
  RUN(2, "2 no status check, 1 Newton");
  RUN(3, "3 d by SHFL, column by smem");
  RUN(4, "4 d and column by SHFL (no smem)");
  RUN(5, "5 arithmetic only (no exchange; floor)");
  RUN(6, "6 fp32 MUFU.RCP seed + 2 Newton");
  RUN(7, "7 no status check, one smem column per pivot (no reuse)");
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
