// Microbenchmark: latency of dependent mma.sync.m8n8k4.f64 chains on one warp / two warps per SM sub-partition, and
// of the LDS -> DMMA pattern of the panel kernel's tile loop.  (What bounds k_schur_panel's mma warps?)
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void chain(double* out, long long* clk, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
// LDS (address from a dependent integer op) -> DMMA on one accumulator, like one (tile, point) step
__global__ void lds_chain(double* out, long long* clk, int iters, int stride) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double c0 = 0, c1 = 0;
  const int lane = threadIdx.x & 31;
  const int foff = (lane & 3) * stride + (lane >> 2);
  unsigned pm = 0xfff;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    unsigned m = pm;
    while (m) {
      const int p = __ffs(m) - 1; m &= m - 1;
      const double fa = sm[(foff + p * 4 * stride + 16) & 4095], fb = sm[(foff + p * 4 * stride + 40) & 4095];
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(fa), "d"(fb));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, sizeof(double) * 148 * 1024); cudaMalloc(&c, 64);
  long long h;
  const int iters = 4096;
  for (int warps : {1, 4, 8, 16}) {
    chain<1><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d: 1 dependent chain  : %.1f clk per DMMA\n", warps, (double)h / iters);
    chain<2><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d: 2 independent chains: %.1f clk per DMMA\n", warps, (double)h / iters / 2);
    chain<4><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d: 4 independent chains: %.1f clk per DMMA\n", warps, (double)h / iters / 4);
    chain<8><<<1, 32 * warps>>>(d, c, iters); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d: 8 independent chains: %.1f clk per DMMA\n", warps, (double)h / iters / 8);
    for (int stride : {120, 124}) {
      lds_chain<<<1, 32 * warps>>>(d, c, 512, stride); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("warps %2d: ffs -> 2 LDS -> DMMA loop (stride %d): %.1f clk per step\n", warps, stride, (double)h / 512 / 12);
    }
  }
  return 0;
}
