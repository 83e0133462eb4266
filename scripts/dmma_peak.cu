// Microbenchmark: mma.sync.m8n8k4.f64 rate on one GPU (is the FP64 tensor path worth using on B200?)
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
  double c[NACC][2];
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// DMMA fed from shared memory: per k-step 3 A + 3 B fragment loads, 9 mma (24x24 warp tile)
__global__ void dmma_lds_kernel(double* out, int iters) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double c[9][2];
  for (int i = 0; i < 9; ++i) { c[i][0] = 0; c[i][1] = 0; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
    double a[3], b[3];
    const int base = ((it * 4 + (lane & 3)) * 129 + (lane >> 2) + wid * 24) & 4095;
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = sm[(base + 8 * i) & 4095]; b[i] = sm[(base + 8 * i + 2000) & 4095]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[3 * i + j][0]), "+d"(c[3 * i + j][1]) : "d"(a[i]), "d"(b[j]));
  }
  double s = 0;
  for (int i = 0; i < 9; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; printf("%s SMs %d\n", p.name, sms);
  double* d; cudaMalloc(&d, sizeof(double) * sms * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 12, 16, 32}) {
    int threads = warps * 32;
    float ms = timeit([&] { dmma_kernel<8><<<sms, threads>>>(d, iters); });
    double fma = 256.0 * 8 * iters * (double)warps * sms;
    printf("dmma m8n8k4 8 acc, warps/SM %2d: %.2f TFLOP/s (%.1f FMA/clk/SM at 1.965 GHz)\n", warps, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { dmma_kernel<2><<<sms, threads>>>(d, iters); });
    fma = 256.0 * 2 * iters * (double)warps * sms;
    printf("dmma m8n8k4 2 acc, warps/SM %2d: %.2f TFLOP/s\n", warps, 2 * fma / ms * 1e-9);
    ms = timeit([&] { dmma_lds_kernel<<<sms, threads>>>(d, iters); });
    fma = 256.0 * 9 * iters * (double)warps * sms;
    printf("dmma 24x24 tile from shared, warps/SM %2d: %.2f TFLOP/s\n", warps, 2 * fma / ms * 1e-9);
  }
  return 0;
}
