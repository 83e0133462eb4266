// Microbenchmark: latency of one 6x6x6 FP64 block update (216 FMAs) out of shared memory per thread,
// as a function of the number of warps doing it at once, and of plain dependent DFMA chains.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void block_update(const double* __restrict__ Aik, const double* __restrict__ Ajk, double* __restrict__ Aij) {
  double bj[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) bj[q] = Ajk[q];
#pragma unroll
  for (int r = 0; r < 6; r += 2) {
    double a0[6], a1[6], c0[6], c1[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) { a0[q] = Aik[6 * r + q]; a1[q] = Aik[6 * r + 6 + q]; c0[q] = Aij[6 * r + q]; c1[q] = Aij[6 * r + 6 + q]; }
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
      for (int c = 0; c < 6; ++c) { c0[c] = fma(-a0[q], bj[6 * c + q], c0[c]); c1[c] = fma(-a1[q], bj[6 * c + q], c1[c]); }
#pragma unroll
    for (int q = 0; q < 6; ++q) { Aij[6 * r + q] = c0[q]; Aij[6 * r + 6 + q] = c1[q]; }
  }
}
__global__ void k_blk(long long* out, int reps) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  for (int i = tid; i < 37 * 3 * 256; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; ++it) {
    block_update(sm + 37 * (tid % 64), sm + 37 * (64 + tid), sm + 37 * (64 + 256 + tid));
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) out[0] = (t1 - t0) / reps;
}
template <int ILP>
__global__ void k_chain(long long* out, double* sink, int n) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
  double x = 1.0000001, y = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = (t1 - t0);
}
int main() {
  long long* d; cudaMalloc(&d, 8); double* sk; cudaMalloc(&sk, 8 * 1024);
  cudaFuncSetAttribute(k_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, 37 * 3 * 256 * 8);
  for (int threads : {32, 64, 128, 256}) {
    k_blk<<<1, threads, 37 * 3 * 256 * 8>>>(d, 50); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("block_update per thread, %3d threads: %lld clk per update (+barrier)  err=%s\n", threads, h, cudaGetErrorString(cudaGetLastError()));
  }
  const int n = 2000;
  long long h;
  k_chain<1><<<1, 32>>>(d, sk, n); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("dependent DFMA chain, 1 warp, ILP 1: %.1f clk per DFMA\n", (double)h / n);
  k_chain<2><<<1, 32>>>(d, sk, n); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("ILP 2: %.1f clk per DFMA\n", (double)h / n / 2);
  k_chain<6><<<1, 32>>>(d, sk, n); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("ILP 6: %.1f clk per DFMA\n", (double)h / n / 6);
  k_chain<12><<<1, 32>>>(d, sk, n); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("ILP 12: %.1f clk per DFMA\n", (double)h / n / 12);
  k_chain<6><<<1, 256>>>(d, sk, n); cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); printf("ILP 6, 8 warps: %.1f clk per DFMA per warp\n", (double)h / n / 6);
  return 0;
}
