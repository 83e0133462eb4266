// How fast can 148 CTAs add their partial sums into ONE small matrix with float64 REDs?
// Mimics the flush of the whole-window linearise kernel: 74 contributors x 7260 entries (+ 148 x 240),
// and measures what spreading the contributors over G copies of the matrix buys.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/red_bench scripts/red_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kN = 7260;

__global__ void k_flush(double* S, int copies, int stride, int contributors, int rotate) {
  const int c = blockIdx.x % contributors;           // contributor id
  if ((int)blockIdx.x >= contributors * 2) return;
  const int half = blockIdx.x / contributors;        // two CTAs share one contributor's entries
  const int lo = half * (kN / 2), cnt = kN / 2;
  const int rot = rotate ? (int)((long long)c * cnt / contributors) : 0;
  double* dst = S + (size_t)(c % copies) * stride;
  for (int off = threadIdx.x; off < cnt; off += blockDim.x) {
    int o = off + rot;
    if (o >= cnt) o -= cnt;
    atomicAdd(dst + lo + o, 1.0 + c);
  }
}

int main() {
  double* S;
  const int stride = 8192;
  cudaMalloc(&S, sizeof(double) * stride * 80);
  cudaMemset(S, 0, sizeof(double) * stride * 80);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rotate = 0; rotate < 2; ++rotate)
    for (int copies : {1, 2, 4, 8, 16, 37, 74}) {
      for (int i = 0; i < 3; ++i) k_flush<<<148, 640>>>(S, copies, stride, 74, rotate);
      cudaEventRecord(e0);
      for (int i = 0; i < 50; ++i) k_flush<<<148, 640>>>(S, copies, stride, 74, rotate);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("rotate %d copies %2d : %.2f us per launch (537k REDs)\n", rotate, copies, ms * 1000.f / 50);
    }
  // empty kernel for the launch floor
  cudaEventRecord(e0);
  for (int i = 0; i < 50; ++i) k_flush<<<148, 640>>>(S, 1, stride, 0 + 1, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("2 active CTAs only: %.2f us per launch\n", ms * 1000.f / 50);
  return 0;
}
