// Microbenchmark: sustained FP64 / FP32 FMA rate and shared-memory float64 load rate on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters) {
  T a[ILP], b = (T)1.0000001, c = (T)0.9999999;
  for (int i = 0; i < ILP; ++i) a[i] = (T)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = a[i] * b + c;
  }
  T s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void lds_kernel(double* out, int iters) {
  __shared__ double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  int base = (threadIdx.x * 19) & 4095;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) s += sm[(base + q + it) & 4095];
  }
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
  int sms = p.multiProcessorCount; printf("%s SMs %d clock %d kHz\n", p.name, sms, p.clockRate);
  double* d; cudaMalloc(&d, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;
    float ms = timeit([&] { fma_kernel<double, 8><<<sms, threads>>>(d, iters); });
    double fl = 2.0 * 8 * iters * (double)threads * sms;
    printf("fp64 ILP8 warps/SM %2d: %.2f TFLOP/s  (%.1f FMA lanes/clk/SM at 1.965 GHz)\n", warps, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / sms / 1.965e9);
    ms = timeit([&] { fma_kernel<float, 8><<<sms, threads>>>((float*)d, iters); });
    printf("fp32 ILP8 warps/SM %2d: %.2f TFLOP/s\n", warps, fl / ms * 1e-9);
  }
  for (int warps : {2, 8, 32}) {
    int threads = warps * 32;
    float ms = timeit([&] { lds_kernel<<<sms, threads>>>(d, 4000); });
    double bytes = 8.0 * 16 * 4000 * (double)threads * sms;
    printf("lds.64 warps/SM %2d: %.1f B/clk/SM\n", warps, bytes / (ms * 1e-3) / sms / 1.965e9);
  }
  return 0;
}
