// Local bundle adjustment: the Levenberg-Marquardt hot path on sm_100a.
//
// Replaces the Adam/autograd loop of the reference
// (visual_slam/optimization/torch_optimizer.py:148-180) with, per LM iteration,
//   k_linearize_*   residuals + Jacobians + Huber weights + U/V/W/g blocks + Schur complement
//   k_rcs_cholesky  damped reduced camera system, in-shared-memory Cholesky, camera retraction
//   k_backsub       point back-substitution, candidate state, candidate cost, accept/reject
// All LM control state lives in device memory (LmCtl), so the loop never syncs with the host.
#include <stdlib.h>

#include "ba_kernels.cuh"
#include "ba_scatter.cuh"

namespace ba {

#ifndef BA_BACKSUB_MIN_CTAS
#define BA_BACKSUB_MIN_CTAS 2                // resident CTAs per SM the back-substitution kernel is compiled for
#endif
static constexpr int kWarps = 8;             // warps per CTA in the per-point kernels
static constexpr int kThreads = kWarps * 32;

// ------------------------------------------------------------------------------------------
// state initialisation
// ------------------------------------------------------------------------------------------
__global__ void k_init(Work w, const double* __restrict__ R_in, const double* __restrict__ t_in,
                       const double* __restrict__ X_in, LmInit p) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int nth = gridDim.x * blockDim.x;
  for (int i = tid; i < w.n_cam * 12; i += nth) {
    int c = i / 12, k = i % 12;
    double v = k < 9 ? R_in[9 * c + k] : t_in[3 * c + (k - 9)];
    w.pose[0][i] = v;
    w.pose[1][i] = v;
  }
  if (w.perm == nullptr) {
    for (int i = tid; i < 3 * w.n_pt; i += nth) w.X[0][i] = X_in[i];
  } else {
    // class order of the panel path: position p holds the caller's point perm[p]
    for (int i = tid; i < 3 * w.n_pt; i += nth) {
      const int p = i / 3, c = i - 3 * p;
      w.X[0][i] = X_in[3 * (size_t)__ldg(w.perm + p) + c];
    }
  }
  long long nred = (long long)w.s_len + 2 * w.n + 1;   // S | b | diagU | cost_lin contiguous
  for (long long i = tid; i < nred; i += nth) w.S[i] = 0.0;
  for (int i = tid; i < w.n; i += nth) w.dp[i] = 0.0;
  // point ranges of the mma linearise kernel's CTAs (equal observation shares): the binary searches over
  // pt_off are a chain of dependent L2 round trips, paid here once per solve instead of once per launch
  if (w.part != nullptr) {
    for (int i = tid; i <= w.part_n; i += nth) {
      int lo = 0;
      if (i == w.part_n) {
        lo = w.n_pt;
      } else if (i > 0) {
        const int target = (int)((long long)w.n_obs * i / w.part_n);
        int hi = w.n_pt;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (w.pt_off[mid] < target) lo = mid + 1; else hi = mid;
        }
        lo &= ~3;          // ranges start at multiples of 4 points: the panel kernel's bulk copies need 16-byte aligned sources
      }
      w.part[i] = lo;
    }
  }
  if (tid == 0) {
    LmCtl* c = w.ctl;
    c->lambda = p.lambda0;
    c->huber_delta = p.huber_delta;
    c->chi2_thr = p.chi2_thr;
    c->lambda_up = p.lambda_up;
    c->lambda_down = p.lambda_down;
    c->lambda_min = p.lambda_min;
    c->lambda_max = p.lambda_max;
    c->accept_rtol = p.accept_rtol;
    c->epoch_base = p.epoch_base;
    c->cost_cur = 0.0;
    c->cost_cand = 0.0;
    c->cost_lin = 0.0;
    c->cur = 0;
    c->iter = 0;
    c->solve_fail = 0;
    c->n_accepted = 0;
    c->ticket = 0u;
    c->pcg_iters_last = 0;
    w.cost_cand[0] = 0.0;
    w.cost_cand[1] = 0.0;
  }
  // histories start clean: a report of a solve with fewer iterations than the last one must not show stale entries
  if (blockIdx.x == 0) {
    LmCtl* c = w.ctl;
    for (int i = threadIdx.x; i < 65; i += blockDim.x) c->cost_hist[i] = 0.0;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
      c->cand_hist[i] = 0.0; c->lambda_hist[i] = 0.0; c->solve_fail_hist[i] = 0; c->pcg_iters_hist[i] = 0; c->accepted[i] = 0;
    }
  }
}

// max_iter == 0: nothing linearises, so the initial cost of the report comes from a plain evaluation
__global__ void k_cost0(Work w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const double c = w.cost_cand[0];
    w.ctl->cost_hist[0] = c;
    w.ctl->cost_cur = c;
    w.cost_cand[0] = 0.0;
  }
}
int launch_cost0(const Work& w, cudaStream_t s) {
  k_cost0<<<1, 32, 0, s>>>(w);
  return 1;
}

int launch_init(const Work& w, const double* R_in, const double* t_in, const double* X_in,
                const LmInit& p, cudaStream_t s) {
  long long work = (long long)w.s_len + 3LL * w.n_pt;
  int blocks = (int)((work + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_init<<<blocks, 256, 0, s>>>(w, R_in, t_in, X_in, p);
  return 1;
}

// ------------------------------------------------------------------------------------------
// linearise + Schur, general scatter variant (warp per point, global float64 RED atomics)
//
// Works for any window size and any per-point observation count; used directly for
// windows the tiled kernel does not cover and as an independent cross-check of it.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
k_linearize_scatter(Work w) {
  __shared__ double sW[kWarps][32][19];   // per-warp staging of W_ij (6x3), padded
  __shared__ int sCam[kWarps][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double lambda = ctl->lambda, delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  double cost = 0.0;
  const int warp_global = blockIdx.x * kWarps + wid;
  const int warp_total = gridDim.x * kWarps;
  for (int j = warp_global; j < w.n_pt; j += warp_total)
    scatter_point(w, j, lane, lambda, delta, pose, X, sW[wid], sCam[wid], cost);
  cost = warp_sum(cost);
  if (lane == 0 && cost != 0.0) atomicAdd(w.cost_lin, cost);
}

int launch_linearize_scatter(const Work& w, int sm_count, cudaStream_t s) {
  int blocks = (w.n_pt + kWarps - 1) / kWarps;
  if (blocks > 2 * sm_count) blocks = 2 * sm_count;
  if (blocks < 1) blocks = 1;
  k_linearize_scatter<<<blocks, kThreads, 0, s>>>(w);
  return 1;
}

// ------------------------------------------------------------------------------------------
// reduced camera system: damping, blocked (6x6) Cholesky in shared memory, camera retraction
//
// One CTA.  The matrix sits in shared memory in the same block-packed lower layout the
// linearise kernels accumulate into, so the load is one coalesced copy.  The forward substitution
// rides along with the factorisation (the right-hand side is carried as one more row); the backward
// substitution uses the stored inverse diagonal factors.
// ------------------------------------------------------------------------------------------
static constexpr size_t kSolveSmemMax = 227 * 1024 - 1024;   // opt-in limit minus static shared

size_t rcs_cholesky_smem_bytes(int n) {
  const size_t nb = (size_t)n / 6;
  return sizeof(double) * (nb * (nb + 1) / 2 * 37 + nb * 36 + 2 * (size_t)n + 8);
}
int rcs_cholesky_max_n() {
  int n = 6;
  while (rcs_cholesky_smem_bytes(n + 6) <= kSolveSmemMax) n += 6;
  return n;
}

// 1/sqrt(a) for the pivots: hardware approximation (MUFU.RSQ64H, ~2^-22) + two Newton steps.
// No slow-path call, so nothing spills around it; a <= 0, inf or nan yield nan/inf and the
// caller's pivot test rejects the step.
__device__ __forceinline__ double fast_rsqrt(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a * r, r, 1.0);       // 1 - a r^2
    r = fma(0.5 * r, e, r);
  }
  return r;
}

// Right-looking Cholesky of a 6x6 SPD block held in registers (lower triangle, row-major 6x6).
// On return L holds the factor (upper part zero) and inv[k] = 1 / L[k][k].
__device__ __forceinline__ bool chol6(double L[36], double inv[6]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double akk = L[7 * k];
    ok = ok && (akk > 0.0) && (akk < 1e300);
    const double r = fast_rsqrt(akk);
    inv[k] = r;
    L[7 * k] = akk * r;
#pragma unroll
    for (int i = k + 1; i < 6; ++i) L[6 * i + k] *= r;
#pragma unroll
    for (int i = k + 1; i < 6; ++i)
#pragma unroll
      for (int j = k + 1; j <= i; ++j) L[6 * i + j] = fma(-L[6 * i + k], L[6 * j + k], L[6 * i + j]);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i + 1; j < 6; ++j) L[6 * i + j] = 0.0;
  return ok;
}

// inverse of the lower-triangular 6x6 factor, column j by lane j (forward substitution on e_j)
__device__ __forceinline__ void tri6_inv_col(const double* __restrict__ L, const double* __restrict__ dinv,
                                             int j, double* __restrict__ Li) {
  double col[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double a = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < i; ++q) a = fma(-L[6 * i + q], col[q], a);
    col[i] = (i < j) ? 0.0 : a * dinv[i];
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) Li[6 * i + j] = col[i];
}

// shared-memory copy uses a 37-double block stride: with 36 (288 B) every block starts on one of
// four bank groups and the per-block loads of the trailing update serialise 4- to 8-way
static constexpr int kBS = 37;
__device__ __forceinline__ int sblk(int I, int J) { return (I * (I + 1) / 2 + J) * kBS; }

// ------------------------------------------------------------------------------------------
// Right-looking blocked Cholesky with the next block column taken off the trailing update's critical path.
// 12 warps: a PANEL group (4 warps) and a TRAILING group (8 warps), one CTA-wide barrier per block column.
// While the trailing group applies column K to the blocks (I,J), J >= K+2, the panel group makes column K+1
// final: rank-6 update of its blocks and of the rhs block by column K (one thread per block row), Cholesky of
// the diagonal block by thread 0, forward substitution along the rows below.  The sequential chain per column
// is (row update -> 6x6 factor -> row solve); measured with the clock64 trace on the C2 window (20 keyframes):
// 630 + 1 200 + 440 clk plus 250 clk of barriers, against 4 900 clk per column for the earlier single-group
// kernel (row solve -> barrier -> whole trailing sweep with the factor at its end -> barrier).
// ------------------------------------------------------------------------------------------
static constexpr int kPanelThreads = 128;
static constexpr int kTrailThreads = 256;
static constexpr int kSolveThreads = kPanelThreads + kTrailThreads;
static constexpr int kBarPanel = 1;

__device__ __forceinline__ void panel_bar() {
  asm volatile("bar.sync %0, %1;" ::"n"(kBarPanel), "n"(kPanelThreads) : "memory");
}

// v (6 entries) <- v * L^-T by forward substitution along the row
__device__ __forceinline__ void row_solve6(double* __restrict__ v, const double* __restrict__ Lk,
                                           const double* __restrict__ di) {
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

// c (6 entries) -= a (6 entries) * B^T, B a 6x6 block
__device__ __forceinline__ void row_update6(double* __restrict__ c, const double* __restrict__ a,
                                            const double* __restrict__ B) {
  double av[6], acc[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) { av[q] = a[q]; acc[q] = c[q]; }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) acc[cc] = fma(-av[q], B[6 * cc + q], acc[cc]);
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) c[q] = acc[q];
}

__global__ void __launch_bounds__(kSolveThreads, 1)
k_rcs_cholesky(Work w) {
  extern __shared__ double sm[];
  const int n = w.n, nb = w.n_cam;
  const int nblk = nb * (nb + 1) / 2;
  double* __restrict__ A = sm;                        // block-packed lower
  double* __restrict__ Linv = A + (size_t)nblk * kBS; // inverse diagonal factors
  double* __restrict__ y = Linv + (size_t)nb * 36;    // right-hand side -> L^-1 b
  double* x = y + n;                                  // solution
  __shared__ int s_fail;
  LmCtl* ctl = w.ctl;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const double lambda = ctl->lambda;
  const int cur = ctl->cur;

  if (tid == 0) {
    s_fail = (ctl->solve_fail != 0) ? 1 : 0;      // a linearise kernel may already have flagged the step
    if (ctl->iter == 0) {
      double c0 = *w.cost_lin;
      ctl->cost_cur = c0;
      ctl->cost_hist[0] = c0;
    }
  }
  dbg_mark(w, 0, 100);
  // this thread's rhs entry / undamped diagonal: requested first so that they travel with the copy below
  double dU = 0.0, bv = 0.0;
  if (tid < n) {
    bv = (w.fixed[tid / 6] != 0) ? 0.0 : w.b[tid];
    dU = w.diagU[tid];
  }
  // coalesced copy of the accumulated blocks (loads batched 10 deep: two round trips for a 20-keyframe window); the global buffer is zeroed
  // behind the copy for the next iteration
  {
    double* __restrict__ S = w.S;
    const int total = nblk * 36;
    for (int base = 0; base < total; base += 10 * nt) {
      double v[10];
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        const int i = base + u * nt + tid;
        v[u] = (i < total) ? S[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        const int i = base + u * nt + tid;
        if (i < total) { A[(i / 36) * kBS + (i % 36)] = v[u]; S[i] = 0.0; }
      }
    }
  }
  if (tid < n) y[tid] = bv;
  for (int c = tid + nt; c < n; c += nt) y[c] = (w.fixed[c / 6] != 0) ? 0.0 : w.b[c];
  __syncthreads();
  // damping of the pose blocks; fixed cameras -> identity rows/columns
  for (int c = tid; c < n; c += nt) {
    const int cam = c / 6, r = c - 6 * cam;
    double* D = A + sblk(cam, cam);
    if (w.fixed[cam] != 0) {
#pragma unroll
      for (int q = 0; q < 6; ++q) D[6 * r + q] = (q == r) ? 1.0 : 0.0;
    } else {
      D[7 * r] += damp(c < nt ? dU : w.diagU[c], lambda);
    }
    w.diagU[c] = 0.0;
    w.b[c] = 0.0;
  }
  if (tid == 0) *w.cost_lin = 0.0;
  __syncthreads();
  dbg_mark(w, 1, 101);

  double* Dinv = x;                     // x is free until the back substitution: 1/diag(L) lives there
  auto factor_diag = [&](int K) {
    double L[36], inv[6];
    double* D = A + sblk(K, K);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) L[6 * i + j] = (j <= i) ? D[6 * i + j] : 0.0;
    if (!chol6(L, inv)) s_fail = 1;
    // only the lower triangle is read by the row solves and the back substitution (the strict upper part
    // keeps whatever the updates left there)
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) D[6 * i + j] = L[6 * i + j];
#pragma unroll
    for (int q = 0; q < 6; ++q) Dinv[6 * K + q] = inv[q];
  };
  // rows below the diagonal block of column K and the rhs block K:  v <- v * L_KK^-T
  auto solve_column = [&](int K, int t0, int nthr) {
    const int m = nb - K - 1;
    const double* __restrict__ Lk = A + sblk(K, K);
    const double* __restrict__ di = Dinv + 6 * K;
    for (int row = t0; row < 6 * m + 1; row += nthr) {
      double* v = (row < 6 * m) ? (A + sblk(K + 1 + row / 6, K) + 6 * (row % 6)) : (y + 6 * K);
      row_solve6(v, Lk, di);
    }
  };

  if (tid == 0) factor_diag(0);
  __syncthreads();
  solve_column(0, tid, nt);
  __syncthreads();

  for (int K = 0; K + 1 < nb; ++K) {
    if (K < 6) dbg_mark(w, 2 + 3 * K, 110);
    // no early exit on a failed pivot: thread 0 may set s_fail while slower threads are still entering this
    // step, so a break here would not be uniform; NaNs run through and the step is rejected at the end
    const int m = nb - K - 1;             // block rows K+1 .. nb-1
    if (tid < kPanelThreads) {
      // ---- column K+1 becomes final ----------------------------------------------------------------
      // items: rows of block (K+1+i, K+1), i < m (the diagonal block's rows first: lanes 0..5 of warp 0),
      // and the rhs block K+1, each updated by column K with B = L(K+1, K)
      const double* __restrict__ B = A + sblk(K + 1, K);
      for (int t = tid; t < 6 * m + 1; t += kPanelThreads) {
        if (t < 6 * m) {
          const int I = K + 1 + t / 6, r = t % 6;
          row_update6(A + sblk(I, K + 1) + 6 * r, A + sblk(I, K) + 6 * r, B);
        } else {
          row_update6(y + 6 * (K + 1), y + 6 * K, B);
        }
      }
      if (tid < 32) {
        __syncwarp();
        if (tid == 0) factor_diag(K + 1);
      }
      panel_bar();
      solve_column(K + 1, tid, kPanelThreads);
    } else {
      // ---- trailing blocks (I,J), K+2 <= J <= I, and rhs blocks J >= K+2, by column K -----------------
      const int tt = tid - kPanelThreads;
      const int m2 = m - 1, J0 = K + 2;
      const int nblk_t = m2 * (m2 + 1) / 2;
      if (nblk_t >= 48) {
        // one thread per 6x6 block: 72 shared loads for 216 FMAs
        for (int t = tt; t < nblk_t + m2; t += kTrailThreads) {
          if (t < nblk_t) {
            int I = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while (I * (I + 1) / 2 > t) --I;
            while ((I + 1) * (I + 2) / 2 <= t) ++I;
            const int J = t - I * (I + 1) / 2;
            const double* __restrict__ Aik = A + sblk(J0 + I, K);
            const double* __restrict__ Ajk = A + sblk(J0 + J, K);
            double* __restrict__ Aij = A + sblk(J0 + I, J0 + J);
            double bj[36];
#pragma unroll
            for (int q = 0; q < 36; ++q) bj[q] = Ajk[q];
            // two block rows at a time: 12 independent accumulators cover the FP64 latency
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
            const int J = J0 + t - nblk_t;
            row_update6(y + 6 * J, y + 6 * K, A + sblk(J, K));
          }
        }
      } else {
        // few blocks left: one thread per block row keeps more lanes busy
        for (int t = tt; t < 6 * nblk_t + m2; t += kTrailThreads) {
          if (t < 6 * nblk_t) {
            const int bi = t / 6, r = t - 6 * bi;
            int I = (int)((sqrtf(8.0f * (float)bi + 1.0f) - 1.0f) * 0.5f);
            while (I * (I + 1) / 2 > bi) --I;
            while ((I + 1) * (I + 2) / 2 <= bi) ++I;
            const int J = bi - I * (I + 1) / 2;
            row_update6(A + sblk(J0 + I, J0 + J) + 6 * r, A + sblk(J0 + I, K) + 6 * r, A + sblk(J0 + J, K));
          } else {
            const int J = J0 + t - 6 * nblk_t;
            row_update6(y + 6 * J, y + 6 * K, A + sblk(J, K));
          }
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  dbg_mark(w, 20, 120);
  const bool fail = s_fail != 0;
  if (!fail) {
    // inverse diagonal factors for the back substitution, all blocks in parallel
    for (int t = tid; t < 6 * nb; t += nt) {
      const int K = t / 6;
      tri6_inv_col(A + sblk(K, K), Dinv + 6 * K, t - 6 * K, Linv + K * 36);
    }
    __syncthreads();
    // backward substitution L^T x = y by the panel group, two group barriers per block:
    // x_K = L_KK^-T y_K by threads 0..5, then y_J -= L_KJ^T x_K for J < K, one entry per thread; the
    // entries of L and y a thread will need are loaded before x_K is known
    if (tid < kPanelThreads) {
      const int Jt = tid / 6, ct = tid - 6 * Jt;            // entry t = tid of y (first pass)
      for (int K = nb - 1; K >= 0; --K) {
        const bool mine = tid < 6 * K;
        double av[6], yv = 0.0;
        if (mine) {
          const double* __restrict__ Akj = A + sblk(K, Jt) + ct;
          yv = y[tid];
#pragma unroll
          for (int r = 0; r < 6; ++r) av[r] = Akj[6 * r];
        }
        if (tid < 6) {
          const double* __restrict__ Li = Linv + K * 36;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            const double t2 = (r >= tid) ? Li[6 * r + tid] * y[6 * K + r] : 0.0;
            if (r & 1) a1 += t2; else a0 += t2;
          }
          x[6 * K + tid] = a0 + a1;
        }
        panel_bar();
        double xk[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) xk[r] = x[6 * K + r];
        if (mine) {
          double a0 = yv, a1 = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            if (r & 1) a1 = fma(-av[r], xk[r], a1); else a0 = fma(-av[r], xk[r], a0);
          }
          y[tid] = a0 + a1;
        }
        for (int t = tid + kPanelThreads; t < 6 * K; t += kPanelThreads) {
          const int J = t / 6, c = t - 6 * J;
          const double* __restrict__ Akj = A + sblk(K, J) + c;
          double a0 = y[t], a1 = 0.0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
            if (r & 1) a1 = fma(-Akj[6 * r], xk[r], a1); else a0 = fma(-Akj[6 * r], xk[r], a0);
          }
          y[t] = a0 + a1;
        }
        panel_bar();
      }
    }
    __syncthreads();
  } else {
    for (int c = tid; c < n; c += nt) x[c] = 0.0;
    __syncthreads();
  }
  dbg_mark(w, 21, 130);
  bool bad = false;
  for (int c = tid; c < n; c += nt) {
    double v = x[c];
    if (!isfinite(v)) { bad = true; v = 0.0; }
    w.dp[c] = v;
  }
  if (bad) s_fail = 1;
  __syncthreads();
  if (tid == 0) ctl->solve_fail = s_fail;
  // candidate camera poses: R <- exp(dw) R, t <- t + dt   (torch_optimizer.py:159, t additive)
  const double* P = w.pose[cur];
  double* Pn = w.pose[cur ^ 1];
  for (int c = tid; c < w.n_cam; c += nt) {
    double R[9], Rn[9], dw[3], dt[3];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = P[12 * c + q];
    bool upd = (w.fixed[c] == 0) && (s_fail == 0);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double a = x[6 * c + q], b2 = x[6 * c + 3 + q];
      dw[q] = (upd && isfinite(a)) ? a : 0.0;
      dt[q] = (upd && isfinite(b2)) ? b2 : 0.0;
    }
    if (upd) so3_exp_mul(dw, R, Rn);
#pragma unroll
    for (int q = 0; q < 9; ++q) Pn[12 * c + q] = upd ? Rn[q] : R[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Pn[12 * c + 9 + q] = P[12 * c + 9 + q] + dt[q];
  }
  dbg_mark(w, 22, 140);
}

int configure_kernels() {
  cudaError_t e = cudaFuncSetAttribute(k_rcs_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kSolveSmemMax);
  if (e != cudaSuccess) return (int)e;
  int r = configure_band();
  if (r != 0) return r;
  r = configure_stream();
  if (r != 0) return r;
  r = configure_mma();
  if (r != 0) return r;
  r = configure_panel();
  if (r != 0) return r;
  r = configure_bcr();
  if (r != 0) return r;
  return configure_tiled();
}

int launch_rcs_cholesky(const Work& w, cudaStream_t s) {
  size_t smem = rcs_cholesky_smem_bytes(w.n);
  k_rcs_cholesky<<<1, kSolveThreads, smem, s>>>(w);
  return 1;
}

// ------------------------------------------------------------------------------------------
// back-substitution, candidate points, candidate cost, accept / reject
// ------------------------------------------------------------------------------------------
// kStaged: the window's current poses, candidate poses and camera steps sit in shared memory (windows of
// up to kBackCams keyframes), so a point's critical path has no dependent global gathers; the next
// point's observation range, camera indices and pixels are prefetched while the current one is solved.
static constexpr int kBackCams = 48;

// kG lanes serve one point (32 / kG points per warp and round): sparse co-visibility (C4: 8 observations per
// point) would leave three quarters of a warp idle with one point per warp.
template <int kG>
__device__ __forceinline__ void group_sum_6_3(double (&V)[6], double (&g)[3]) {
#pragma unroll
  for (int o = kG / 2; o > 0; o >>= 1) {
    double tv[6], tg[3];
#pragma unroll
    for (int q = 0; q < 6; ++q) tv[q] = __shfl_xor_sync(kFull, V[q], o);
#pragma unroll
    for (int q = 0; q < 3; ++q) tg[q] = __shfl_xor_sync(kFull, g[q], o);
#pragma unroll
    for (int q = 0; q < 6; ++q) V[q] += tv[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) g[q] += tg[q];
  }
}

template <bool kStaged, int kG>
__global__ void __launch_bounds__(kThreads, BA_BACKSUB_MIN_CTAS)
k_backsub(Work w) {
  __shared__ double s_cost[kWarps];
  __shared__ double s_pose[kStaged ? kBackCams * 13 : 1];
  __shared__ double s_posen[kStaged ? kBackCams * 13 : 1];
  __shared__ double s_dp[kStaged ? kBackCams * 7 : 1];
  constexpr int kPer = 32 / kG;                       // points per warp and round
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = lane & (kG - 1), sub = lane / kG;
  LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double lambda = ctl->lambda, delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ posen = w.pose[cur ^ 1];
  const double* __restrict__ X = w.X[cur];
  double* __restrict__ Xn = w.X[cur ^ 1];
  const double* __restrict__ dp = w.dp;
  double cost = 0.0;

  // every lane of a warp runs the same number of rounds (the butterflies are warp-wide instructions);
  // a group past the last point carries k = -1
  const int first = (blockIdx.x * kWarps + wid) * kPer;
  const int stride = gridDim.x * kWarps * kPer;
  // prefetch pipeline: range of the point after next, observation of the next point
  int j = first + sub;
  int o0 = 0, k = -1, o0n = 0, kn = -1;
  if (j < w.n_pt) { o0 = __ldg(w.pt_off + j); k = __ldg(w.pt_off + j + 1) - o0; }
  if (j + stride < w.n_pt) { o0n = __ldg(w.pt_off + j + stride); kn = __ldg(w.pt_off + j + stride + 1) - o0n; }
  ObsIn in;
  in.cam = 0; in.u = in.v = 0.0;
  double X0 = 0.0, X1 = 0.0, X2 = 0.0;
  if (k > 0 && gl < k) in = load_obs(w, o0 + gl);
  if (k >= 0) { X0 = X[3 * (size_t)j]; X1 = X[3 * (size_t)j + 1]; X2 = X[3 * (size_t)j + 2]; }
  if (kStaged) {
    for (int i = threadIdx.x; i < 12 * w.n_cam; i += kThreads) {
      const int c = i / 12, q = i - 12 * c;
      s_pose[c * 13 + q] = pose[i];
      s_posen[c * 13 + q] = posen[i];
    }
    for (int i = threadIdx.x; i < 6 * w.n_cam; i += kThreads) {
      const int c = i / 6, q = i - 6 * c;
      s_dp[c * 7 + q] = (w.fixed[c] == 0) ? dp[i] : 0.0;
    }
    __syncthreads();
  }

  for (int jb = first; jb < w.n_pt; jb += stride, j += stride) {
    // ---- issue the next point's loads before touching this one -----------------------------------
    const int jn = j + stride;
    ObsIn nin;
    nin.cam = 0; nin.u = nin.v = 0.0;
    double N0x = 0.0, N1x = 0.0, N2x = 0.0;
    if (kn > 0 && gl < kn) nin = load_obs(w, o0n + gl);
    if (kn >= 0) { N0x = X[3 * (size_t)jn]; N1x = X[3 * (size_t)jn + 1]; N2x = X[3 * (size_t)jn + 2]; }
    int o0nn = 0, knn = -1;
    if (jn + stride < w.n_pt) {
      o0nn = __ldg(w.pt_off + jn + stride);
      knn = __ldg(w.pt_off + jn + stride + 1) - o0nn;
    }

    double V[6] = {0, 0, 0, 0, 0, 0}, q[3] = {0, 0, 0};
    for (int base = 0; base < k; base += kG) {
      const int i = base + gl;
      if (i < k) {
        const ObsIn ob = (base == 0) ? in : load_obs(w, o0 + i);
        ObsLin ol;
        const double* P = kStaged ? (s_pose + 13 * ob.cam) : (pose + 12 * (size_t)ob.cam);
        linearize_obs<true>(w.K, P, P + 9, X0, X1, X2, ob.u, ob.v, delta, ol);
        double eu = ol.ru, ev = ol.rv;
        if (kStaged) {
          const double* d = s_dp + 7 * ob.cam;
#pragma unroll
          for (int r = 0; r < 6; ++r) { eu += ol.pu[r] * d[r]; ev += ol.pv[r] * d[r]; }
        } else if (w.fixed[ob.cam] == 0) {
          const double* d = dp + 6 * (size_t)ob.cam;
#pragma unroll
          for (int r = 0; r < 6; ++r) { eu += ol.pu[r] * d[r]; ev += ol.pv[r] * d[r]; }
        }
        V[0] += ol.wu * ol.lu[0] * ol.lu[0] + ol.wv * ol.lv[0] * ol.lv[0];
        V[1] += ol.wu * ol.lu[0] * ol.lu[1] + ol.wv * ol.lv[0] * ol.lv[1];
        V[2] += ol.wu * ol.lu[0] * ol.lu[2] + ol.wv * ol.lv[0] * ol.lv[2];
        V[3] += ol.wu * ol.lu[1] * ol.lu[1] + ol.wv * ol.lv[1] * ol.lv[1];
        V[4] += ol.wu * ol.lu[1] * ol.lu[2] + ol.wv * ol.lv[1] * ol.lv[2];
        V[5] += ol.wu * ol.lu[2] * ol.lu[2] + ol.wv * ol.lv[2] * ol.lv[2];
        const double au = ol.wu * eu, av = ol.wv * ev;
#pragma unroll
        for (int r = 0; r < 3; ++r) q[r] += au * ol.lu[r] + av * ol.lv[r];
      }
    }
    group_sum_6_3<kG>(V, q);
    if (k > 0) {
      V[0] += damp(V[0], lambda);
      V[3] += damp(V[3], lambda);
      V[5] += damp(V[5], lambda);
      double Vi[6];
      sym3_inverse(V, Vi);
      const double d0 = -(Vi[0] * q[0] + Vi[1] * q[1] + Vi[2] * q[2]);
      const double d1 = -(Vi[1] * q[0] + Vi[3] * q[1] + Vi[4] * q[2]);
      const double d2 = -(Vi[2] * q[0] + Vi[4] * q[1] + Vi[5] * q[2]);
      const double N0 = X0 + d0, N1 = X1 + d1, N2 = X2 + d2;
      if (gl == 0) { Xn[3 * (size_t)j] = N0; Xn[3 * (size_t)j + 1] = N1; Xn[3 * (size_t)j + 2] = N2; }
      for (int base = 0; base < k; base += kG) {
        const int i = base + gl;
        if (i < k) {
          const ObsIn ob = (base == 0) ? in : load_obs(w, o0 + i);
          const double* P = kStaged ? (s_posen + 13 * ob.cam) : (posen + 12 * (size_t)ob.cam);
          double ru, rv;
          residual_only(w.K, P, P + 9, N0, N1, N2, ob.u, ob.v, ru, rv);
          cost += huber_rho(ru, delta) + huber_rho(rv, delta);
        }
      }
    } else if (k == 0 && gl == 0) {
      Xn[3 * (size_t)j] = X0; Xn[3 * (size_t)j + 1] = X1; Xn[3 * (size_t)j + 2] = X2;
    }
    // rotate the pipeline
    in = nin; X0 = N0x; X1 = N1x; X2 = N2x;
    o0 = o0n; k = kn; o0n = o0nn; kn = knn;
  }
  cost = warp_sum(cost);
  if (lane == 0) s_cost[wid] = cost;
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) c += s_cost[i];
    atomicAdd(w.cost_cand, c);
    // sharded: a failed solve on ANY rank must reject the step on every rank, so the flag travels with the
    // candidate cost through the all-reduce (sum > 0 = failed somewhere)
    if (w.world > 1 && blockIdx.x == 0) w.cost_cand[1] = (ctl->solve_fail != 0) ? 1.0 : 0.0;
    // the LM state the decision needs does not change during this kernel: requested now, so that the last CTA
    // has it by the time its ticket comes back
    const LmIn in = lm_load(ctl);
    __threadfence();
    const unsigned t = atomicAdd(&ctl->ticket, 1u);
    if (t == gridDim.x - 1 && w.world == 1) {
      __threadfence();
      const double cand = *((volatile double*)w.cost_cand);
      *w.cost_cand = 0.0;
      lm_accept(ctl, in, cand);
    }
  }
}

// lanes per point (BA_BACKSUB_GROUP overrides)
static int backsub_group(const Work& w) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("BA_BACKSUB_GROUP");
    const int v = e ? atoi(e) : 0;
    forced = (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
  }
  if (forced) return forced;
  if (w.n_pt <= 0) return 32;
  // measured on B200 (C1 / C2 / C4, 10 / 14.6 / 8 observations per point): the kernel is bound by the latency of a
  // point's chain, so more points in flight win -- 4 lanes per point beat 8, 16 and 32 everywhere, and where the points
  // alone fill the machine fewer still: C4 (1 M points) 0.190 / 0.203 / 0.255 ms with 1 / 2 / 4 lanes, C2 (10 k points)
  // 24.9 / 18.8 / 14.8 us, C1 (2 k) 15.9 / 13.5 / 12.1 us
  const double mean = (double)w.n_obs / (double)w.n_pt;
  if (mean <= 24.0) return w.n_pt >= 100000 ? 1 : (w.n_pt >= 40000 ? 2 : 4);
  return mean <= 48.0 ? 8 : (mean <= 96.0 ? 16 : 32);
}

template <bool kStaged>
static void launch_backsub_g(const Work& w, int g, int sm_count, cudaStream_t s) {
  const int per = 32 / g;
  int blocks = (w.n_pt + kWarps * per - 1) / (kWarps * per);
  const int cap = sm_count * BA_BACKSUB_MIN_CTAS;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (g == 1) k_backsub<kStaged, 1><<<blocks, kThreads, 0, s>>>(w);
  else if (g == 2) k_backsub<kStaged, 2><<<blocks, kThreads, 0, s>>>(w);
  else if (g == 4) k_backsub<kStaged, 4><<<blocks, kThreads, 0, s>>>(w);
  else if (g == 8) k_backsub<kStaged, 8><<<blocks, kThreads, 0, s>>>(w);
  else if (g == 16) k_backsub<kStaged, 16><<<blocks, kThreads, 0, s>>>(w);
  else k_backsub<kStaged, 32><<<blocks, kThreads, 0, s>>>(w);
}

int launch_backsub(const Work& w, int sm_count, cudaStream_t s) {
  const int g = backsub_group(w);
  if (w.n_cam <= kBackCams) launch_backsub_g<true>(w, g, sm_count, s);
  else launch_backsub_g<false>(w, g, sm_count, s);
  return 1;
}

// point index of every observation from the CSR offsets (host-pointer API: 4 bytes per observation that need not
// cross PCIe); a warp per point, so long observation lists are written coalesced
__global__ void k_expand_obs_pt(const int* __restrict__ pt_off, int n_pt, int* __restrict__ obs_pt) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
  for (int j = warp; j < n_pt; j += nwarp) {
    const int o0 = __ldg(pt_off + j), o1 = __ldg(pt_off + j + 1);
    for (int o = o0 + lane; o < o1; o += 32) obs_pt[o] = j;
  }
}
int launch_expand_obs_pt(const int* pt_off, int n_pt, int* obs_pt, cudaStream_t s) {
  if (n_pt <= 0) return 0;
  int blocks = (n_pt + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_expand_obs_pt<<<blocks, 256, 0, s>>>(pt_off, n_pt, obs_pt);
  return 1;
}

// argument check of the host-pointer API on the device (the host loops over 8 M observations cost 2-3 ms of every C4
// call): offsets start at 0, never fall, end at n_obs; cameras in range.  bad: 1 offsets, 2 cameras.
__global__ void k_validate(int n_cam, int n_pt, int n_obs, const int* __restrict__ obs_cam, const int* __restrict__ pt_off,
                           int* __restrict__ bad) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  int b = 0;
  if (tid == 0 && n_pt > 0 && (pt_off[0] != 0 || pt_off[n_pt] != n_obs)) b |= 1;
  for (int j = tid; j < n_pt; j += nth)
    if (__ldg(pt_off + j + 1) < __ldg(pt_off + j)) b |= 1;
  for (int o = tid; o < n_obs; o += nth)
    if ((unsigned)__ldg(obs_cam + o) >= (unsigned)n_cam) b |= 2;
  if (b) atomicOr(bad, b);
}
int launch_validate(int n_cam, int n_pt, int n_obs, const int* obs_cam, const int* pt_off, int* bad, cudaStream_t s) {
  cudaMemsetAsync(bad, 0, sizeof(int), s);
  int blocks = (n_obs + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_validate<<<blocks, 256, 0, s>>>(n_cam, n_pt, n_obs, obs_cam, pt_off, bad);
  return 1;
}

// accept / reject as its own kernel (multi-GPU: runs after the candidate-cost all-reduce)
__global__ void k_accept(Work w) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    LmIn in = lm_load(w.ctl);
    const double cand = w.cost_cand[0];
    if (w.cost_cand[1] != 0.0) in.solve_fail = 1;          // some rank's solve failed: every rank rejects
    w.cost_cand[0] = 0.0;
    w.cost_cand[1] = 0.0;
    lm_accept(w.ctl, in, cand);
  }
}
int launch_accept(const Work& w, cudaStream_t s) {
  k_accept<<<1, 32, 0, s>>>(w);
  return 1;
}

// ------------------------------------------------------------------------------------------
// outlier flags at the final accepted iterate + copy-out
// ------------------------------------------------------------------------------------------
__global__ void k_finalize(Work w, double* __restrict__ R_out, double* __restrict__ t_out,
                           double* __restrict__ X_out, unsigned char* __restrict__ flags) {
  const int cur = w.ctl->cur;
  const double chi2 = w.ctl->chi2_thr;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int nth = gridDim.x * blockDim.x;
  if (flags) {
    for (int o = tid; o < w.n_obs; o += nth) {
      ObsIn in = load_obs(w, o);
      int j = __ldg(w.obs_pt + o);
      double ru, rv;
      residual_only(w.K, pose + 12 * in.cam, pose + 12 * in.cam + 9, X[3 * j], X[3 * j + 1],
                    X[3 * j + 2], in.u, in.v, ru, rv);
      const unsigned char f = (ru * ru + rv * rv) > chi2 ? 1 : 0;
      if (w.perm == nullptr) flags[o] = f;
      else flags[__ldg(w.pt_off_in + __ldg(w.perm + j)) + (o - __ldg(w.pt_off + j))] = f;   // the caller's observation order
    }
  }
  if (w.perm == nullptr) {
    for (int i = tid; i < 3 * w.n_pt; i += nth) X_out[i] = X[i];
  } else {
    for (int i = tid; i < 3 * w.n_pt; i += nth) {
      const int p = i / 3, c = i - 3 * p;
      X_out[3 * (size_t)__ldg(w.perm + p) + c] = X[i];
    }
  }
  for (int i = tid; i < 12 * w.n_cam; i += nth) {
    int c = i / 12, k = i % 12;
    if (k < 9) R_out[9 * c + k] = pose[i];
    else t_out[3 * c + (k - 9)] = pose[i];
  }
}

int launch_finalize(const Work& w, double* R_out, double* t_out, double* X_out,
                    unsigned char* flags, cudaStream_t s) {
  int work = w.n_obs > 3 * w.n_pt ? w.n_obs : 3 * w.n_pt;
  int blocks = (work + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_finalize<<<blocks, 256, 0, s>>>(w, R_out, t_out, X_out, flags);
  return 1;
}

// ------------------------------------------------------------------------------------------
// objective only
// ------------------------------------------------------------------------------------------
__global__ void k_eval_cost(Work w, const double* __restrict__ pose, const double* __restrict__ X,
                            double* out, double delta) {
  double cost = 0.0;
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int nth = gridDim.x * blockDim.x;
  for (int o = tid; o < w.n_obs; o += nth) {
    ObsIn in = load_obs(w, o);
    int j = __ldg(w.obs_pt + o);
    double ru, rv;
    residual_only(w.K, pose + 12 * in.cam, pose + 12 * in.cam + 9, (double)X[3 * j],
                  (double)X[3 * j + 1], (double)X[3 * j + 2], in.u, in.v, ru, rv);
    cost += huber_rho(ru, delta) + huber_rho(rv, delta);
  }
  cost = warp_sum(cost);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, cost);
}

int launch_eval_cost(const Work& w, const double* pose, const double* X, double* out,
                     double delta, cudaStream_t s) {
  int blocks = (w.n_obs + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaMemsetAsync(out, 0, sizeof(double), s);
  k_eval_cost<<<blocks, 256, 0, s>>>(w, pose, X, out, delta);
  return 1;
}

}  // namespace ba
