// Large windows (more cameras than the in-shared-memory Cholesky holds): block-sparse reduced
// camera matrix and a block-Jacobi preconditioned conjugate-gradient solve.
//
//  * symbolic phase (once per solve): camera co-visibility bitmap -> block-CSR structure
//    (row_ptr / col_idx / per-word prefix popcounts for O(1) block lookup / transposed positions)
//  * k_rcs_pcg: one cooperative kernel per LM iteration.  All PCG iterations run inside it with two
//    grid barriers each; only blocks with col <= row are stored, the SpMV reads the transposed
//    position for the upper half; the search direction is rebuilt on the fly (p = z + beta p_old)
//    so that no third barrier is needed.  Ends with the camera retraction, like the Cholesky kernel.
#include <cooperative_groups.h>

#include "ba_kernels.cuh"

namespace cg = cooperative_groups;

namespace ba {

// ------------------------------------------------------------------------------------------
// symbolic structure
// ------------------------------------------------------------------------------------------
__global__ void k_sym_mark(Work w, unsigned* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warp_total = (gridDim.x * blockDim.x) >> 5;
  const int W = w.bsr_W;
  for (int j = warp_global; j < w.n_pt; j += warp_total) {
    const int o0 = __ldg(w.pt_off + j);
    const int k = __ldg(w.pt_off + j + 1) - o0;
    for (int ia = lane; ia < k; ia += 32) {
      const int a = __ldg(w.obs_cam + o0 + ia);
      if (w.fixed[a]) continue;
      for (int ib = 0; ib < k; ++ib) {
        const int b = __ldg(w.obs_cam + o0 + ib);
        if (w.fixed[b]) continue;
        const unsigned m = 1u << (b & 31);
        unsigned* p = bits + (size_t)a * W + (b >> 5);
        if ((__ldg(p) & m) == 0u) atomicOr(p, m);
      }
    }
  }
  // every camera owns its diagonal block (fixed ones become identity rows in the solver)
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < w.n_cam; a += gridDim.x * blockDim.x)
    atomicOr(bits + (size_t)a * W + (a >> 5), 1u << (a & 31));
}

// per row: prefix popcounts of its words and the row total (warp per row)
__global__ void k_sym_rowcount(int n_cam, int W, const unsigned* __restrict__ bits,
                               int* __restrict__ wpre, int* __restrict__ row_nnz) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warp_total = (gridDim.x * blockDim.x) >> 5;
  for (int a = warp_global; a < n_cam; a += warp_total) {
    int run = 0;
    for (int base = 0; base < W; base += 32) {
      const int wi = base + lane;
      const int c = (wi < W) ? __popc(bits[(size_t)a * W + wi]) : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += v;
      }
      if (wi < W) wpre[(size_t)a * W + wi] = run + inc - c;
      run += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) row_nnz[a] = run;
  }
}

// exclusive scan of the row counts (single CTA) -> row_ptr, total in row_ptr[n_cam]
__global__ void k_sym_scan(int n_cam, const int* __restrict__ row_nnz, int* __restrict__ row_ptr) {
  __shared__ int s_warp[32];
  __shared__ int s_run;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n_cam; base += blockDim.x) {
    const int i = base + tid;
    const int c = (i < n_cam) ? row_nnz[i] : 0;
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int v = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
      int vi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(kFull, vi, o);
        if (lane >= o) vi += u;
      }
      s_warp[lane] = vi - v;
    }
    __syncthreads();
    const int run = s_run;
    if (i < n_cam) row_ptr[i] = run + s_warp[wid] + inc - c;
    __syncthreads();
    if (tid == blockDim.x - 1) s_run = run + s_warp[wid] + inc;
    __syncthreads();
  }
  if (tid == 0) row_ptr[n_cam] = s_run;
}

__global__ void k_sym_fill(int n_cam, int W, const unsigned* __restrict__ bits,
                           const int* __restrict__ wpre, const int* __restrict__ row_ptr,
                           int* __restrict__ col_idx) {
  const long long total = (long long)n_cam * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i / W), wi = (int)(i - (long long)a * W);
    unsigned m = bits[i];
    int at = row_ptr[a] + wpre[i];
    while (m) {
      const int bit = __ffs(m) - 1;
      col_idx[at++] = wi * 32 + bit;
      m &= m - 1;
    }
  }
}

__global__ void k_sym_tpos(Work w, int* __restrict__ tpos) {
  const int lane = threadIdx.x & 31;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int warp_total = (gridDim.x * blockDim.x) >> 5;
  for (int a = warp_global; a < w.n_cam; a += warp_total) {
    const int e0 = w.bsr_row_ptr[a], e1 = w.bsr_row_ptr[a + 1];
    for (int e = e0 + lane; e < e1; e += 32) tpos[e] = bsr_find(w, w.bsr_col[e], a);
  }
}

int launch_symbolic_mark(const Work& w, unsigned* bits, cudaStream_t s) {
  cudaMemsetAsync(bits, 0, sizeof(unsigned) * (size_t)w.n_cam * w.bsr_W, s);
  int blocks = (w.n_pt + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_sym_mark<<<blocks, 256, 0, s>>>(w, bits);
  return 1;
}

__global__ void k_sym_or(size_t words, int world, const unsigned* __restrict__ all, unsigned* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
    unsigned m = 0u;
    for (int r = 0; r < world; ++r) m |= all[(size_t)r * words + i];
    out[i] = m;
  }
}

int launch_symbolic_or(const Work& w, const unsigned* all, int world, unsigned* bits, cudaStream_t s) {
  const size_t words = (size_t)w.n_cam * w.bsr_W;
  int blocks = (int)((words + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_sym_or<<<blocks, 256, 0, s>>>(words, world, all, bits);
  return 1;
}

int launch_symbolic_count(const Work& w, unsigned* bits, int* wpre, int* row_nnz, int* row_ptr, cudaStream_t s) {
  int rb = (w.n_cam + 7) / 8;
  if (rb > 148 * 8) rb = 148 * 8;
  k_sym_rowcount<<<rb, 256, 0, s>>>(w.n_cam, w.bsr_W, bits, wpre, row_nnz);
  k_sym_scan<<<1, 1024, 0, s>>>(w.n_cam, row_nnz, row_ptr);
  return 2;
}

int launch_symbolic(const Work& w, unsigned* bits, int* wpre, int* row_nnz, int* row_ptr,
                    cudaStream_t s) {
  launch_symbolic_mark(w, bits, s);
  launch_symbolic_count(w, bits, wpre, row_nnz, row_ptr, s);
  return 3;
}

int launch_symbolic_fill(const Work& w, int* col_idx, int* tpos, cudaStream_t s) {
  long long total = (long long)w.n_cam * w.bsr_W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_sym_fill<<<blocks, 256, 0, s>>>(w.n_cam, w.bsr_W, w.bsr_bits, w.bsr_wpre, w.bsr_row_ptr, col_idx);
  int rb = (w.n_cam + 7) / 8;
  if (rb > 148 * 8) rb = 148 * 8;
  k_sym_tpos<<<rb, 256, 0, s>>>(w, tpos);
  return 2;
}

// ------------------------------------------------------------------------------------------
// block-Jacobi PCG on the block-sparse reduced camera system
// ------------------------------------------------------------------------------------------
static constexpr int kPcgThreads = 256;
static constexpr int kPcgWarps = kPcgThreads / 32;

__device__ __forceinline__ double fast_rsqrt_pcg(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a * r, r, 1.0);
    r = fma(0.5 * r, e, r);
  }
  return r;
}

// inverse of a 6x6 SPD block (lower triangle given, row-major) into Mi (full symmetric)
__device__ bool spd6_inverse(const double* __restrict__ D, double Mi[36]) {
  double L[36], inv[6];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) L[6 * i + j] = (j <= i) ? D[6 * i + j] : 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double akk = L[7 * k];
    ok = ok && (akk > 0.0) && (akk < 1e300);
    const double r = fast_rsqrt_pcg(akk);
    inv[k] = r;
    L[7 * k] = akk * r;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) L[6 * i + k] *= r;
#pragma unroll
    for (int i = k + 1; i < 6; ++i)
#pragma unroll
      for (int j = k + 1; j <= i; ++j) L[6 * i + j] = fma(-L[6 * i + k], L[6 * j + k], L[6 * i + j]);
  }
  // Li = L^-1 (lower), column by column
  double Li[36];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i < j) { Li[6 * i + j] = 0.0; continue; }
      double a = (i == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = j; q < i; ++q) a = fma(-L[6 * i + q], Li[6 * q + j], a);
      Li[6 * i + j] = a * inv[i];
    }
  }
  // Mi = Li^T Li
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      double a = 0.0;
#pragma unroll
      for (int q = r; q < 6; ++q) a = fma(Li[6 * q + r], Li[6 * q + c], a);
      Mi[6 * r + c] = a;
      Mi[6 * c + r] = a;
    }
  return ok;
}

__device__ __forceinline__ double block_sum_pcg(double v, double* s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s_red[wid] = v;
  __syncthreads();
  double a = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kPcgWarps; ++i) a += s_red[i];
  }
  return a;   // valid in thread 0
}

__global__ void __launch_bounds__(kPcgThreads)
k_rcs_pcg(Work w) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_red[kPcgWarps];
  LmCtl* ctl = w.ctl;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int gtid = blockIdx.x * blockDim.x + tid, gthreads = gridDim.x * blockDim.x;
  const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  const int nc = w.n_cam;
  double* __restrict__ x = w.dp;
  double* __restrict__ r = w.pcg_r;
  double* __restrict__ z = w.pcg_z;
  double* __restrict__ Ap = w.pcg_Ap;
  double* __restrict__ Minv = w.pcg_Minv;
  volatile double* acc = w.pcg_acc;
  const double tol2 = w.pcg_tol * w.pcg_tol;

  if (gtid == 0) {
    if (ctl->iter == 0) {
      double c0 = *w.cost_lin;
      ctl->cost_cur = c0;
      ctl->cost_hist[0] = c0;
    }
  }
  for (int i = gtid; i < 3 * (w.pcg_max_iter + 2); i += gthreads) w.pcg_acc[i] = 0.0;
  grid.sync();

  // ---- setup: block-Jacobi inverse, r = b, z = M^-1 r, x = 0 --------------------------------
  int fail_local = 0;
  double rz_p = 0.0, bb_p = 0.0;
  for (int a = gtid; a < nc; a += gthreads) {
    double Mi[36];
    double ra[6], za[6];
    if (w.fixed[a]) {
#pragma unroll
      for (int q = 0; q < 36; ++q) Mi[q] = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) { Mi[7 * q] = 1.0; ra[q] = 0.0; }
    } else {
      double D[36];
      const double* S = w.S + (size_t)bsr_find(w, a, a) * 36;
#pragma unroll
      for (int q = 0; q < 36; ++q) D[q] = S[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) { D[7 * q] += damp(w.diagU[6 * (size_t)a + q], lambda); ra[q] = w.b[6 * (size_t)a + q]; }
      if (!spd6_inverse(D, Mi)) fail_local = 1;
    }
#pragma unroll
    for (int q = 0; q < 36; ++q) Minv[36 * (size_t)a + q] = Mi[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s = fma(Mi[6 * q + c], ra[c], s);
      za[q] = s;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const size_t i = 6 * (size_t)a + q;
      x[i] = 0.0; r[i] = ra[q]; z[i] = za[q];
      w.pcg_p[0][i] = 0.0; w.pcg_p[1][i] = 0.0;
      rz_p += ra[q] * za[q];
      bb_p += ra[q] * ra[q];
    }
  }
  {
    double v = block_sum_pcg(rz_p, s_red);
    if (tid == 0 && v != 0.0) atomicAdd(w.pcg_acc + 0, v);
    v = block_sum_pcg(bb_p, s_red);
    if (tid == 0 && v != 0.0) atomicAdd(w.pcg_acc + 1, v);
    if (fail_local) atomicAdd(w.pcg_acc + 2, 1.0);
  }
  grid.sync();
  double rz = acc[0];
  const double bb = acc[1];
  int fail = (acc[2] != 0.0 || ctl->solve_fail != 0) ? 1 : 0;
  double beta = 0.0;
  int it = 0;
  int pi = 0;   // pcg_p[pi] = previous direction, pcg_p[pi ^ 1] = the one being written
  if (!fail && bb > 0.0 && isfinite(bb)) {
    for (; it < w.pcg_max_iter; ++it) {
      const double* __restrict__ pold = w.pcg_p[pi];
      double* __restrict__ pnew = w.pcg_p[pi ^ 1];
      // ---- SpMV: Ap = (S + lambda diag U) p with p = z + beta p_old rebuilt on the fly --------
      double pAp_p = 0.0;
      for (int a = gwarp; a < nc; a += gwarps) {
        double y[6] = {0, 0, 0, 0, 0, 0};
        const bool fa = w.fixed[a] != 0;
        if (!fa) {
          const int e0 = __ldg(w.bsr_row_ptr + a), e1 = __ldg(w.bsr_row_ptr + a + 1);
          for (int e = e0 + lane; e < e1; e += 32) {
            const int b = __ldg(w.bsr_col + e);
            if (w.fixed[b]) continue;
            double pb[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) pb[q] = fma(beta, pold[6 * (size_t)b + q], z[6 * (size_t)b + q]);
            if (b < a) {
              const double* M = w.S + (size_t)e * 36;
#pragma unroll
              for (int rr = 0; rr < 6; ++rr)
#pragma unroll
                for (int c = 0; c < 6; ++c) y[rr] = fma(M[6 * rr + c], pb[c], y[rr]);
            } else if (b > a) {
              const double* M = w.S + (size_t)__ldg(w.bsr_tpos + e) * 36;
#pragma unroll
              for (int rr = 0; rr < 6; ++rr)
#pragma unroll
                for (int c = 0; c < 6; ++c) y[rr] = fma(M[6 * c + rr], pb[c], y[rr]);
            } else {
              const double* M = w.S + (size_t)e * 36;
#pragma unroll
              for (int rr = 0; rr < 6; ++rr) {
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                  const double m = (c <= rr) ? M[6 * rr + c] : M[6 * c + rr];
                  y[rr] = fma(m, pb[c], y[rr]);
                }
                y[rr] = fma(damp(w.diagU[6 * (size_t)a + rr], lambda), pb[rr], y[rr]);
              }
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) y[q] = warp_sum(y[q]);
        if (lane < 6) {
          const size_t i = 6 * (size_t)a + lane;
          const double pa = fa ? 0.0 : fma(beta, pold[i], z[i]);
          double yv = y[0];
#pragma unroll
          for (int q = 1; q < 6; ++q) yv = (lane == q) ? y[q] : yv;
          pnew[i] = pa;
          Ap[i] = yv;
          pAp_p += pa * yv;
        }
      }
      {
        double v = block_sum_pcg(pAp_p, s_red);
        if (tid == 0 && v != 0.0) atomicAdd(w.pcg_acc + 3 + 3 * it, v);
      }
      grid.sync();
      const double pAp = acc[3 + 3 * it];
      if (!(pAp > 0.0) || !isfinite(pAp)) { fail = 1; break; }     // uniform
      const double alpha = rz / pAp;
      // ---- x, r, z updates ------------------------------------------------------------------
      double rz_n = 0.0, rr_n = 0.0;
      for (int a = gtid; a < nc; a += gthreads) {
        double ra[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const size_t i = 6 * (size_t)a + q;
          x[i] = fma(alpha, pnew[i], x[i]);
          ra[q] = fma(-alpha, Ap[i], r[i]);
          r[i] = ra[q];
        }
        const double* Mi = Minv + 36 * (size_t)a;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < 6; ++c) s = fma(Mi[6 * q + c], ra[c], s);
          z[6 * (size_t)a + q] = s;
          rz_n += ra[q] * s;
          rr_n += ra[q] * ra[q];
        }
      }
      {
        double v = block_sum_pcg(rz_n, s_red);
        if (tid == 0 && v != 0.0) atomicAdd(w.pcg_acc + 4 + 3 * it, v);
        v = block_sum_pcg(rr_n, s_red);
        if (tid == 0 && v != 0.0) atomicAdd(w.pcg_acc + 5 + 3 * it, v);
      }
      grid.sync();
      const double rz_new = acc[4 + 3 * it], rr = acc[5 + 3 * it];
      pi ^= 1;
      if (rr <= tol2 * bb) { ++it; break; }                         // uniform
      beta = rz_new / rz;
      rz = rz_new;
    }
  }
  // every thread holds the same `fail` / `it`
  if (fail) {
    for (int i = gtid; i < w.n; i += gthreads) x[i] = 0.0;
  }
  grid.sync();
  if (gtid == 0) {
    ctl->solve_fail = fail;
    ctl->pcg_iters_last = it;
    *w.cost_lin = 0.0;
  }
  // candidate camera poses
  const double* P = w.pose[cur];
  double* Pn = w.pose[cur ^ 1];
  for (int c = gtid; c < nc; c += gthreads) {
    double R[9], Rn[9], dw[3], dt[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = P[12 * (size_t)c + q];
    const bool upd = (w.fixed[c] == 0) && !fail;
    bool fin = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = x[6 * (size_t)c + q], b2 = x[6 * (size_t)c + 3 + q];
      fin = fin && isfinite(a) && isfinite(b2);
      dw[q] = upd ? a : 0.0;
      dt[q] = upd ? b2 : 0.0;
    }
    if (!fin) {
#pragma unroll
      for (int q = 0; q < 3; ++q) { dw[q] = 0.0; dt[q] = 0.0; x[6 * (size_t)c + q] = 0.0; x[6 * (size_t)c + 3 + q] = 0.0; }
    }
    if (upd && fin) so3_exp_mul(dw, R, Rn);
#pragma unroll
    for (int q = 0; q < 9; ++q) Pn[12 * (size_t)c + q] = (upd && fin) ? Rn[q] : R[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Pn[12 * (size_t)c + 9 + q] = P[12 * (size_t)c + 9 + q] + dt[q];
  }
  // clear the accumulation buffers for the next LM iteration
  for (size_t i = gtid; i < w.s_len; i += gthreads) w.S[i] = 0.0;
  for (int i = gtid; i < w.n; i += gthreads) { w.b[i] = 0.0; w.diagU[i] = 0.0; }
}

int pcg_grid_size(int n_cam, int sm_count) {
  int want = (n_cam + kPcgWarps - 1) / kPcgWarps;
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rcs_pcg, kPcgThreads, 0);
  if (occ < 1) occ = 1;
  int cap = sm_count * (occ > 2 ? 2 : occ);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return want;
}

int launch_rcs_pcg(const Work& w, int grid, cudaStream_t s) {
  Work wc = w;
  void* args[] = {&wc};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_rcs_pcg, dim3(grid), dim3(kPcgThreads), args, 0, s);
  return e == cudaSuccess ? 1 : -1;
}

}  // namespace ba
