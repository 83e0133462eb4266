// Microbenchmark: what does a predicated-off DMMA cost?  One strip step per point (3 LDS + 2 DMMA, as ptxas if-converts
// the panel kernel's branch-over), 12 points per pass, with masks that switch 12 / 6 / 1 / 0 of them on.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kYS = 132;
template <int U>
__device__ __forceinline__ void step(unsigned ya, unsigned yb, unsigned pm, double (&acc)[2][2]) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t.reg .f64 a, b0, b1;\n\t"
      "and.b32 t, %6, %7;\n\t"
      "setp.eq.u32 p, t, 0;\n\t"
      "@p bra.uni SKIP_%=;\n\t"
      "ld.shared.f64 a, [%4+%8];\n\t"
      "ld.shared.f64 b0, [%5+%8];\n\t"
      "ld.shared.f64 b1, [%5+%9];\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {a}, {b0}, {%0,%1};\n\t"
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%2,%3}, {a}, {b1}, {%2,%3};\n\t"
      "SKIP_%=:\n\t}"
      : "+d"(acc[0][0]), "+d"(acc[0][1]), "+d"(acc[1][0]), "+d"(acc[1][1])
      : "r"(ya), "r"(yb), "r"(pm), "n"(1u << U), "n"(U * 4 * kYS * 8), "n"(U * 4 * kYS * 8 + 64));
}
__global__ void k(double* out, long long* clk, const unsigned* pms, int n) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 6400; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double acc[2][2] = {{0, 0}, {0, 0}};
  const int lane = threadIdx.x & 31;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + 8 * ((lane & 3) * kYS + (lane >> 2));
  const unsigned pm = pms[0];
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    unsigned ya = base + 64 * (it & 3), yb = base + 128 * (it & 1);
    step<0>(ya, yb, pm, acc); step<1>(ya, yb, pm, acc); step<2>(ya, yb, pm, acc); step<3>(ya, yb, pm, acc);
    step<4>(ya, yb, pm, acc); step<5>(ya, yb, pm, acc); step<6>(ya, yb, pm, acc); step<7>(ya, yb, pm, acc);
    step<8>(ya, yb, pm, acc); step<9>(ya, yb, pm, acc); step<10>(ya, yb, pm, acc); step<11>(ya, yb, pm, acc);
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc[0][0] + acc[0][1] + acc[1][0] + acc[1][1];
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* d; long long* c; unsigned* pm; cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 64); cudaMalloc(&pm, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64000);
  const int n = 2048;
  for (int warps : {4, 8}) {
    for (unsigned m : {0xfffu, 0x555u, 0x001u, 0x0u}) {
      cudaMemcpy(pm, &m, 4, cudaMemcpyHostToDevice);
      k<<<1, 32 * warps, 64000>>>(d, c, pm, n);
      long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("warps %d mask %03x (%2d of 12 points on): %.1f clk per pass of 12 points (24 DMMAs when all on)\n", warps, m, __builtin_popcount(m), (double)h / n);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
