// Large sequential windows: block-BANDED reduced camera matrix and a direct banded Cholesky.
//
// A map point is seen by keyframes that are close in time, so for a trajectory without loop
// closures the reduced camera matrix S is block-banded: S(a,b) != 0 only for |a-b| <= w, with
// w = max over points of (largest - smallest camera index of the point).  Cholesky fill stays
// inside the band, so the factorisation is exact at n_cam * w^2 block operations -- for the
// BAL-scale config (1 000 cameras, w = 15) three orders of magnitude less work than the
// thousands of block-Jacobi PCG iterations the ill-conditioned chain needs (measured: 35.7 ms
// of PCG per LM iteration on C4, profiles/r1_v3_c4_single_gpu_kernel_times.log).
//
//  * layout: S_band[I][d] = block (I, I-d), d = 0..w, 36 doubles row-major; the linearise kernels
//    reach a block with pure index arithmetic (s_block_ptr), and the multi-GPU all-reduce moves
//    n_cam (w+1) 288 bytes instead of the full block-CSR.
//  * k_rcs_band (one CTA): right-looking blocked Cholesky over a SLIDING WINDOW held in shared
//    memory.  At step K only rows K..K+w are live; they sit in a ring of w+2 block rows whose
//    (row, column) residues are folded into one packed triangle (for two residues only one
//    ordering is ever live), (w+2)(w+3)/2 blocks = 176 KB at w = 32.  The row entering the
//    window is prefetched into registers at the top of the step and stored into the cells the
//    previous step freed, so its L2 latency hides behind the panel and trailing phases.  The
//    next diagonal block is factored by one thread right after its own update (look-ahead).
//    Forward substitution rides along as one more row.  Columns of L go to global memory and
//    come back in double-buffered batches for the warp-synchronous backward substitution.
#include "ba_kernels.cuh"

namespace ba {

static constexpr int kBandThreads = 256;
static constexpr int kBandMaxW = 32;
static constexpr int kCS = 37;            // cell stride (doubles): odd -> conflict-free block rows
static constexpr int kPF = 5;             // prefetch registers per thread: (kBandMaxW+1)*36 <= kPF*kBandThreads
static constexpr int kBatch = 8;          // columns per backward-substitution batch

int band_max_w() { return kBandMaxW; }

size_t band_smem_bytes(int bw) {
  const size_t W2 = (size_t)bw + 2;
  const size_t cells = W2 * (W2 + 1) / 2 * kCS;
  const size_t back = 2 * (size_t)kBatch * ((size_t)(bw + 1) * 36 + 36 + 6) + 64 * 6;
  const size_t a = cells > back ? cells : back;
  return sizeof(double) * (a + 2 * W2 * 6 + 8);
}

// ------------------------------------------------------------------------------------------
// bandwidth of the co-visibility structure: max over points of (max camera - min camera)
// ------------------------------------------------------------------------------------------
__global__ void k_band_width(Work w, int* __restrict__ out) {
  int span = 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < w.n_pt; j += gridDim.x * blockDim.x) {
    const int o0 = __ldg(w.pt_off + j), o1 = __ldg(w.pt_off + j + 1);
    int lo = 0x7fffffff, hi = -1;
    for (int o = o0; o < o1; ++o) {
      const int c = __ldg(w.obs_cam + o);
      lo = min(lo, c);
      hi = max(hi, c);
    }
    if (hi >= lo) span = max(span, hi - lo);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) span = max(span, __shfl_xor_sync(kFull, span, o));
  if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(out, span);
}

int launch_band_width(const Work& w, int* out, cudaStream_t s) {
  cudaMemsetAsync(out, 0, sizeof(int), s);
  int blocks = (w.n_pt + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_band_width<<<blocks, 256, 0, s>>>(w, out);
  return 1;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double band_rsqrt(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a * r, r, 1.0);
    r = fma(0.5 * r, e, r);
  }
  return r;
}

__device__ __forceinline__ int tri_cell(int i, int j) {
  const int mx = max(i, j), mn = min(i, j);
  return (mx * (mx + 1) / 2 + mn) * kCS;
}

__device__ __forceinline__ int wrap(int r, int W2) { return r >= W2 ? r - W2 : r; }

__global__ void __launch_bounds__(kBandThreads, 1)
k_rcs_band(Work w) {
  extern __shared__ double sm[];
  const int nb = w.n_cam, bw = w.band_w, W2 = bw + 2, rowlen = (bw + 1) * 36;
  const int ncell = W2 * (W2 + 1) / 2;
  const int back_doubles = 2 * kBatch * ((bw + 1) * 36 + 36 + 6) + 64 * 6;
  const int a_doubles = max(ncell * kCS, back_doubles);
  double* __restrict__ A = sm;                 // window cells (factorisation) / column batches (backward)
  double* __restrict__ yr = A + a_doubles;     // ring: right-hand side rows of the window
  double* __restrict__ dinv = yr + W2 * 6;     // ring: 1 / diag(L_KK)
  __shared__ int s_fail;
  LmCtl* ctl = w.ctl;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;
  double* __restrict__ Sg = w.S;
  double* __restrict__ Lg = w.band_L;

  if (tid == 0) {
    s_fail = (ctl->solve_fail != 0) ? 1 : 0;
    if (ctl->iter == 0) {
      const double c0 = *w.cost_lin;
      ctl->cost_cur = c0;
      ctl->cost_hist[0] = c0;
    }
  }
  // value of element e of band row I as the factorisation wants it (damped, gauge rows = identity);
  // the accumulation buffers are cleared behind the read for the next LM iteration
  auto fetch = [&](int I, int e) -> double {
    const int d = e / 36, q = e - 36 * d;
    const size_t at = (size_t)I * rowlen + e;
    double v = Sg[at];
    Sg[at] = 0.0;
    if (d == 0) {
      const int r = q / 6, c = q - 6 * r;
      if (w.fixed[I] != 0) v = (r == c) ? 1.0 : 0.0;
      else if (r == c) v += damp(w.diagU[6 * (size_t)I + r], lambda);
    } else if (I - d < 0) {
      v = 0.0;
    }
    return v;
  };
  auto place = [&](int I, int ri, int e, double v) {
    const int d = e / 36, q = e - 36 * d;
    if (I - d < 0) return;
    int rj = ri - d;
    if (rj < 0) rj += W2;
    A[tri_cell(ri, rj) + q] = v;
  };

  // ---- initial window: rows 0..min(bw, nb-1) -------------------------------------------------
  {
    const int rows0 = min(bw + 1, nb);
    for (int I = 0; I < rows0; ++I)
      for (int e = tid; e < rowlen; e += nt) place(I, I, e, fetch(I, e));
    for (int i = tid; i < 6 * rows0; i += nt) {
      const int I = i / 6;
      yr[i] = (w.fixed[I] != 0) ? 0.0 : w.b[i];
    }
  }
  __syncthreads();

  auto factor_diag = [&](int rk) {
    double L[36], inv[6];
    double* D = A + tri_cell(rk, rk);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) L[6 * i + j] = (j <= i) ? D[6 * i + j] : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double akk = L[7 * k];
      ok = ok && (akk > 0.0) && (akk < 1e300);
      const double r = band_rsqrt(akk);
      inv[k] = r;
      L[7 * k] = akk * r;
#pragma unroll
      for (int i = k + 1; i < 6; ++i) L[6 * i + k] *= r;
#pragma unroll
      for (int i = k + 1; i < 6; ++i)
#pragma unroll
        for (int j = k + 1; j <= i; ++j) L[6 * i + j] = fma(-L[6 * i + k], L[6 * j + k], L[6 * i + j]);
    }
    if (!ok) s_fail = 1;
#pragma unroll
    for (int q = 0; q < 36; ++q) D[q] = L[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) dinv[rk * 6 + q] = inv[q];
  };
  if (tid == 0) factor_diag(0);
  __syncthreads();

  // per-thread constants of the row prefetch: block offset d and element q of its kPF band-row elements,
  // and the diag(U) entry it needs when it holds a diagonal element of the diagonal block
  int pd[kPF], pq[kPF];
#pragma unroll
  for (int u = 0; u < kPF; ++u) {
    const int e = tid + u * nt;
    pd[u] = e / 36;
    pq[u] = e - 36 * pd[u];
  }
  const int my_diag = (tid < 36 && tid / 6 == tid % 6) ? tid / 6 : -1;

  // ---- factorisation + forward substitution --------------------------------------------------
  int rK = 0;
  for (int K = 0; K < nb; ++K) {
    if (s_fail) break;                            // uniform: written before the last barrier
    const int m = min(bw, nb - 1 - K);
    const bool tr = (K >= 100 && K < 103);
    if (tr) dbg_mark(w, 40 + 5 * (K - 100), 500);
    // prefetch the row entering the window
    const int In = K + bw + 1;
    const int rn = (rK == 0) ? W2 - 1 : rK - 1;   // residue of In == residue of K-1
    // raw loads only (no value-dependent branch, so nothing waits for them before the row is placed):
    // the band row, the gauge flag of the camera and this thread's diag(U) entry
    double pre[kPF];
    double pre_du = 0.0, pre_b = 0.0;
    int pre_fx = 0;
    if (In < nb) {
      const double* __restrict__ src = Sg + (size_t)In * rowlen;
#pragma unroll
      for (int u = 0; u < kPF; ++u) {
        const int e = tid + u * nt;
        pre[u] = (e < rowlen) ? __ldcg(src + e) : 0.0;
      }
      pre_fx = w.fixed[In];
      if (my_diag >= 0) pre_du = __ldcg(w.diagU + 6 * (size_t)In + my_diag);
      if (tid < 6) pre_b = __ldcg(w.b + 6 * (size_t)In + tid);
#pragma unroll
      for (int u = 0; u < kPF; ++u) {
        const int e = tid + u * nt;
        if (e < rowlen) Sg[(size_t)In * rowlen + e] = 0.0;      // cleared behind the read for the next LM iteration
      }
    }
    // (2) panel rows and the rhs row:  v <- v * L_KK^-T
    {
      const double* __restrict__ Lk = A + tri_cell(rK, rK);
      const double* __restrict__ di = dinv + rK * 6;
      for (int row = tid; row < 6 * m + 1; row += nt) {
        double* v;
        if (row < 6 * m) {
          const int i = row / 6;
          v = A + tri_cell(wrap(rK + 1 + i, W2), rK) + 6 * (row - 6 * i);
        } else {
          v = yr + rK * 6;
        }
        double o[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) o[q] = v[q];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          o[q] *= di[q];
#pragma unroll
          for (int c = q + 1; c < 6; ++c) o[c] = fma(-o[q], Lk[6 * c + q], o[c]);
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) v[q] = o[q];
      }
    }
    if (tr) dbg_mark(w, 41 + 5 * (K - 100), 501);
    __syncthreads();
    if (tr) dbg_mark(w, 42 + 5 * (K - 100), 502);
    // (3) trailing update; warp 0 owns the next diagonal block and factors it (look-ahead)
    const int nblk_t = m * (m + 1) / 2;
    if (wid == 0) {
      if (m > 0) {
        if (lane < 6) {
          const int r1 = wrap(rK + 1, W2);
          const double* __restrict__ Aik = A + tri_cell(r1, rK) + 6 * lane;
          const double* __restrict__ Ajk = A + tri_cell(r1, rK);
          double* __restrict__ Aij = A + tri_cell(r1, r1) + 6 * lane;
          double a[6], acc[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) { a[q] = Aik[q]; acc[q] = Aij[q]; }
#pragma unroll
          for (int q = 0; q < 6; ++q)
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[c] = fma(-a[q], Ajk[6 * c + q], acc[c]);
#pragma unroll
          for (int q = 0; q < 6; ++q) Aij[q] = acc[q];
        }
        __syncwarp();
        if (lane == 0) factor_diag(wrap(rK + 1, W2));
      }
    } else {
      const int wt = nt - 32, t0 = tid - 32;
      for (int t = t0 + 1; t < nblk_t + m; t += wt) {
        if (t < nblk_t) {
          int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
          while (I * (I + 1) / 2 > t) --I;
          while ((I + 1) * (I + 2) / 2 <= t) ++I;
          const int J = t - I * (I + 1) / 2;
          const int ri = wrap(rK + 1 + I, W2), rj = wrap(rK + 1 + J, W2);
          const double* __restrict__ Aik = A + tri_cell(ri, rK);
          const double* __restrict__ Ajk = A + tri_cell(rj, rK);
          double* __restrict__ Aij = A + tri_cell(ri, rj);
          double bj[36];
#pragma unroll
          for (int q = 0; q < 36; ++q) bj[q] = Ajk[q];
#pragma unroll
          for (int r = 0; r < 6; r += 2) {
            double a0[6], a1[6], c0[6], c1[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              a0[q] = Aik[6 * r + q];
              a1[q] = Aik[6 * r + 6 + q];
              c0[q] = Aij[6 * r + q];
              c1[q] = Aij[6 * r + 6 + q];
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
#pragma unroll
              for (int c = 0; c < 6; ++c) {
                c0[c] = fma(-a0[q], bj[6 * c + q], c0[c]);
                c1[c] = fma(-a1[q], bj[6 * c + q], c1[c]);
              }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
              Aij[6 * r + q] = c0[q];
              Aij[6 * r + 6 + q] = c1[q];
            }
          }
        } else {
          const int J = t - nblk_t;
          const int rj = wrap(rK + 1 + J, W2);
          const double* __restrict__ Ajk = A + tri_cell(rj, rK);
          const double* yk = yr + rK * 6;
          double* yj = yr + rj * 6;
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double acc = yj[c];
#pragma unroll
            for (int q = 0; q < 6; ++q) acc = fma(-yk[q], Ajk[6 * c + q], acc);
            yj[c] = acc;
          }
        }
      }
      // column K of L, y_K and 1/diag go to global memory for the backward substitution
      for (int e = t0; e < (m + 1) * 36; e += wt) {
        const int d = e / 36, q = e - 36 * d;
        Lg[(size_t)K * rowlen + e] = A[tri_cell(wrap(rK + d, W2), rK) + q];
      }
      if (t0 < 6) {
        w.band_y[6 * (size_t)K + t0] = yr[rK * 6 + t0];
        w.band_dinv[6 * (size_t)K + t0] = dinv[rK * 6 + t0];
      }
    }
    if (tr) dbg_mark(w, 43 + 5 * (K - 100), 503);
    // the new row goes into the cells that step K-1 freed (residue of K-1): nobody reads them now
    if (In < nb) {
      // damping / gauge rows are applied here, to the diagonal block (d = 0: the first 36 elements)
      if (tid < 36) {
        const int r = tid / 6, c = tid - 6 * r;
        if (pre_fx != 0) pre[0] = (r == c) ? 1.0 : 0.0;
        else if (r == c) pre[0] += damp(pre_du, lambda);
      }
#pragma unroll
      for (int u = 0; u < kPF; ++u) {
        const int e = tid + u * nt;
        if (e < rowlen) {
          int rj = rn - pd[u];                       // In - d >= 1 inside the loop: every block of the row exists
          if (rj < 0) rj += W2;
          A[tri_cell(rn, rj) + pq[u]] = pre[u];
        }
      }
      if (tid < 6) yr[rn * 6 + tid] = (pre_fx != 0) ? 0.0 : pre_b;
    }
    if (tr) dbg_mark(w, 44 + 5 * (K - 100), 504);
    __syncthreads();
    rK = wrap(rK + 1, W2);
  }
  __syncthreads();
  dbg_mark(w, 70, 510);
  const bool fail = s_fail != 0;
  double* __restrict__ x = w.dp;

  if (!fail) {
    // ---- backward substitution L^T x = y in the column-oriented (axpy) form -------------------
    // Row K of L (blocks (K, K-d) = d-th block of column K-d) arrives in double-buffered batches.
    // Warp 0 walks the rows downwards: x_K = L_KK^-T y_K, then y_J -= L_KJ^T x_K for the rows J < K
    // inside the band.  Only the update of y_{K-1} is on the critical path of the next step (it is
    // done first), so a step costs one 6x6 product + one triangular product instead of a reduction
    // over the whole band column.
    const int rowsz = (bw + 1) * 36 + 36 + 6;         // blocks (K, K-d) d = 0..bw | Linv of L_KK | pad
    double* buf[2] = {A, A + kBatch * rowsz};
    double* yb = A + 2 * kBatch * rowsz;              // ring of 64 right-hand side rows
    const int nbatch = (nb + kBatch - 1) / kBatch;
    auto load_batch = [&](int bi, double* dst, int t0, int wt) {
      // batch bi covers rows K = nb-1-bi*kBatch-s, s = 0..kBatch-1; element e = (s, d, q).
      // Eight loads are in flight per thread before the first one is stored.
      const int per_row = (bw + 1) * 36, total = kBatch * per_row;
      for (int base = t0; base < total; base += 8 * wt) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = base + u * wt;
          v[u] = 0.0;
          if (e < total) {
            const int s = e / per_row, r = e - s * per_row;
            const int d = r / 36, q = r - 36 * d;
            const int K = nb - 1 - bi * kBatch - s;
            if (K - d >= 0) v[u] = __ldcg(Lg + (size_t)(K - d) * rowlen + d * 36 + q);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = base + u * wt;
          if (e < total) {
            const int s = e / per_row, r = e - s * per_row;
            dst[s * rowsz + r] = v[u];
          }
        }
      }
      // the rows that enter the band below this batch start from the forward-substitution result
      if (t0 < 6 * kBatch) {
        const int s = t0 / 6, c = t0 - 6 * s;
        const int J = nb - 1 - bi * kBatch - s - bw;
        if (J >= 0) yb[(J & 63) * 6 + c] = __ldcg(w.band_y + 6 * (size_t)J + c);
      }
    };
    auto invert_batch = [&](int bi, double* dst, int t0, int wt) {
      for (int t = t0; t < 6 * kBatch; t += wt) {
        const int s = t / 6, j = t - 6 * s;
        const int K = nb - 1 - bi * kBatch - s;
        if (K < 0) continue;
        const double* L = dst + s * rowsz;            // block d = 0 is L_KK
        double* Li = dst + s * rowsz + (bw + 1) * 36;
        double di[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) di[i] = __ldcg(w.band_dinv + 6 * (size_t)K + i);
        double col[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          double a = (i == j) ? 1.0 : 0.0;
#pragma unroll
          for (int q = 0; q < i; ++q) a = fma(-L[6 * i + q], col[q], a);
          col[i] = (i < j) ? 0.0 : a * di[i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) Li[6 * i + j] = col[i];
      }
    };
    // rows nb-1-bw .. nb-1 start from the forward-substitution result
    for (int i = tid; i < 6 * min(bw + 1, nb); i += nt) {
      const int J = nb - 1 - i / 6;
      yb[(J & 63) * 6 + (i % 6)] = w.band_y[6 * (size_t)J + (i % 6)];
    }
    __syncthreads();
    load_batch(0, buf[0], tid, nt);
    __syncthreads();
    invert_batch(0, buf[0], tid, nt);
    __syncthreads();
    // per-lane constants of the update rounds: pair p = lane + 32 u  ->  (d - 1, column)
    constexpr int kRounds = (6 * kBandMaxW + 31) / 32;
    int p_off[kRounds], p_d[kRounds];
#pragma unroll
    for (int u = 0; u < kRounds; ++u) {
      const int pp = lane + 32 * u;
      p_d[u] = pp / 6;
      p_off[u] = (p_d[u] + 1) * 36 + (pp - 6 * p_d[u]);
    }
    for (int bi = 0; bi < nbatch; ++bi) {
      double* cb = buf[bi & 1];
      if (bi == 10) dbg_mark(w, 75, 520);
      if (wid == 0) {
        for (int s = 0; s < kBatch; ++s) {
          const int K = nb - 1 - bi * kBatch - s;
          if (K < 0) break;
          const int mp = min(bw, K);
          const double* row = cb + s * rowsz;
          const double* Li = row + (bw + 1) * 36;
          // x_c = sum_{r >= c} Linv[r][c] y_r   (lanes c < 6), then every lane gets the six x
          const int c = lane < 6 ? lane : 0;
          double xa = 0.0, xb = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            const double l = (r >= c) ? Li[6 * r + c] : 0.0;
            const double yv = yb[(K & 63) * 6 + r];
            if (r & 1) xb = fma(l, yv, xb); else xa = fma(l, yv, xa);
          }
          const double xv = xa + xb;
          if (lane < 6) x[6 * (size_t)K + lane] = xv;
          double xk[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) xk[r] = __shfl_sync(kFull, xv, r);
          // y_J -= L_KJ^T x_K; the first six pairs (round 0, lanes 0..5) are row K-1, the only ones the
          // next step waits for.  All rounds are independent and issued back to back.
#pragma unroll
          for (int u = 0; u < kRounds; ++u) {
            if (p_d[u] < mp) {
              const double* Lb = row + p_off[u];
              double* yj = yb + ((K - 1 - p_d[u]) & 63) * 6 + (p_off[u] % 36) % 6;
              double a0 = *yj, a1 = 0.0;
              a0 = fma(-Lb[0], xk[0], a0);
              a1 = fma(-Lb[6], xk[1], a1);
              a0 = fma(-Lb[12], xk[2], a0);
              a1 = fma(-Lb[18], xk[3], a1);
              a0 = fma(-Lb[24], xk[4], a0);
              a1 = fma(-Lb[30], xk[5], a1);
              *yj = a0 + a1;
            }
          }
          __syncwarp();
        }
      } else if (bi + 1 < nbatch) {
        load_batch(bi + 1, buf[(bi + 1) & 1], tid - 32, nt - 32);
        // the inverses need the loaded diagonal blocks: named barrier over warps 1..7
        asm volatile("bar.sync 1, %0;" ::"r"(nt - 32));
        invert_batch(bi + 1, buf[(bi + 1) & 1], tid - 32, nt - 32);
      }
      if (bi == 10) dbg_mark(w, 76, 521);
      __syncthreads();
      if (bi == 10) dbg_mark(w, 77, 522);
    }
  } else {
    for (int c = tid; c < w.n; c += nt) x[c] = 0.0;
    // the accumulation buffers must still be cleared for the next iteration
    for (size_t i = tid; i < w.s_len; i += nt) Sg[i] = 0.0;
    __syncthreads();
  }
  dbg_mark(w, 71, 511);
  // clear b / diag(U) / cost for the next iteration
  for (int c = tid; c < w.n; c += nt) { w.b[c] = 0.0; w.diagU[c] = 0.0; }
  if (tid == 0) *w.cost_lin = 0.0;
  __syncthreads();
  bool bad = false;
  for (int c = tid; c < w.n; c += nt) {
    const double v = x[c];
    if (!isfinite(v)) { bad = true; x[c] = 0.0; }
  }
  if (bad) s_fail = 1;
  __syncthreads();
  if (tid == 0) ctl->solve_fail = s_fail;
  // candidate camera poses: R <- exp(dw) R, t <- t + dt
  const double* P = w.pose[cur];
  double* Pn = w.pose[cur ^ 1];
  for (int c = tid; c < nb; c += nt) {
    double R[9], Rn[9], dw[3], dt[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = P[12 * (size_t)c + q];
    const bool upd = (w.fixed[c] == 0) && (s_fail == 0);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double a = x[6 * (size_t)c + q], b2 = x[6 * (size_t)c + 3 + q];
      dw[q] = (upd && isfinite(a)) ? a : 0.0;
      dt[q] = (upd && isfinite(b2)) ? b2 : 0.0;
    }
    if (upd) so3_exp_mul(dw, R, Rn);
#pragma unroll
    for (int q = 0; q < 9; ++q) Pn[12 * (size_t)c + q] = upd ? Rn[q] : R[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Pn[12 * (size_t)c + 9 + q] = P[12 * (size_t)c + 9 + q] + dt[q];
  }
}

int configure_band() {
  cudaError_t e = cudaFuncSetAttribute(k_rcs_band, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)band_smem_bytes(kBandMaxW));
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_rcs_band(const Work& w, cudaStream_t s) {
  k_rcs_band<<<1, kBandThreads, band_smem_bytes(w.band_w), s>>>(w);
  return 1;
}

}  // namespace ba
