// Parallel direct solver for the block-BANDED reduced camera system: block cyclic reduction over
// super-blocks of 15 keyframes.
//
// The sequential banded Cholesky (k_rcs_band, ba_band.cu) walks the trajectory one block column at a
// time in ONE CTA: 3.5 ms for 1 000 keyframes, 35 ms for 10 000, replicated on every rank of the sharded
// path -- the Amdahl term of the multi-GPU configurations.  With bandwidth w <= 15 keyframes the
// matrix is block-TRIDIAGONAL in super-blocks of 15 keyframes (90 x 90 scalars):
//      D_i x_i + L_i x_(i-1) + L_(i+1)^T x_(i+1) = b_i .
// Cyclic reduction eliminates every odd super-block of a level at once, one CTA per super-block:
//   phase 1  partial Cholesky of the tall panel [D_i ; L_i^T ; L_(i+1) ; b_i^T] over its 15 block
//            columns (same right-looking 6x6-block algorithm, look-ahead and 37-double cell stride as the
//            dense solver): D_i = C C^T,  E1 = L_i^T C^-T,  E2 = L_(i+1) C^-T,  y = C^-1 b_i;
//   phase 2  Schur complement onto the even neighbours, 15 block products per output block accumulated
//            in registers, one float64 RED per entry:
//            D_(i-1) -= E1 E1^T,  D_(i+1) -= E2 E2^T,  L'_(i+1,i-1) -= E2 E1^T,  b_(i-1) -= E1 y,  b_(i+1) -= E2 y.
// The even super-blocks form the next level (half as many); log2(n_cam / 15) levels later one block is
// left.  Back substitution runs the levels in reverse:  x_i = C^-T (y - E1^T x_(i-1) - E2^T x_(i+1)).
// Every level is one launch whose CTAs are independent, so all SMs work; the results equal the
// sequential factorisation's up to rounding (both are exact direct solves of the same SPD system).
#include "ba_kernels.cuh"

namespace ba {

namespace {

constexpr int kSB = 15;                       // keyframes per super-block (>= band width)
constexpr int kDB = kSB * (kSB + 1) / 2;      // 120 lower blocks of a diagonal super-block
constexpr int kLB = kSB * kSB;                // 225 blocks of a coupling super-block
constexpr int kD = kDB * 36, kL = kLB * 36, kV = kSB * 6;
constexpr int kFac = kD + kV + 2 * kL + kV;   // C | 1/diag | E1 | E2 | y  per eliminated super-block
constexpr int kCS = 37;                       // shared-memory cell stride (doubles)
constexpr int kRows = 3 * kSB;                // block rows of the tall panel
constexpr int kCells = kDB + 2 * kLB;         // 570
constexpr int kET = 512;                      // threads of the elimination kernel
constexpr int kMaxLevels = 24;

struct BcrPlan {
  int n_levels;                               // levels 0 .. n_levels-1; the last one has a single super-block
  int N[kMaxLevels];                          // super-blocks per level
  size_t off_D[kMaxLevels], off_L[kMaxLevels], off_b[kMaxLevels], off_F[kMaxLevels];
  size_t off_x, off_fail, total;
};

__host__ __device__ inline int elim_count(int N) { return N == 1 ? 1 : N / 2; }

inline BcrPlan make_plan(int n_cam) {
  BcrPlan p;
  int N = (n_cam + kSB - 1) / kSB;
  size_t at = 0;
  p.n_levels = 0;
  for (;;) {
    const int l = p.n_levels++;
    p.N[l] = N;
    p.off_D[l] = at; at += (size_t)N * kD;
    p.off_L[l] = at; at += (size_t)N * kL;
    p.off_b[l] = at; at += (size_t)N * kV;
    p.off_F[l] = at; at += (size_t)elim_count(N) * kFac;
    if (N == 1) break;
    N = (N + 1) / 2;
  }
  p.off_x = at; at += (size_t)p.N[0] * kV;
  p.off_fail = at; at += 2;
  p.total = at;
  return p;
}

__device__ __forceinline__ double bcr_rsqrt(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a * r, r, 1.0);
    r = fma(0.5 * r, e, r);
  }
  return r;
}

// cell of block (I, J) of the tall panel: rows 0..14 the lower triangle of D, rows 15..44 full rows
__device__ __forceinline__ int cell(int I, int J) {
  return (I < kSB ? I * (I + 1) / 2 + J : kDB + (I - kSB) * kSB + J) * kCS;
}

// ------------------------------------------------------------------------------------------
// level 0 from the banded accumulation buffers (damping, gauge rows, clearing behind the read)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bcr_assemble(Work w, double* __restrict__ D, double* __restrict__ L, double* __restrict__ b, int* fail) {
  const int i = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int bw = w.band_w, rowlen = (bw + 1) * 36;
  LmCtl* ctl = w.ctl;
  const double lambda = ctl->lambda;
  if (i == 0 && tid == 0) {
    fail[0] = (ctl->solve_fail != 0) ? 1 : 0;
    fail[1] = 0;
    if (ctl->iter == 0) {
      const double c0 = *w.cost_lin;
      ctl->cost_cur = c0;
      ctl->cost_hist[0] = c0;
    }
  }
  double* __restrict__ Di = D + (size_t)i * kD;
  double* __restrict__ Li = L + (size_t)i * kL;
  // diagonal super-block, lower blocks (a, c), c <= a
  for (int e = tid; e < kD; e += nt) {
    const int blk = e / 36, q = e - 36 * blk;
    int a = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
    while (a * (a + 1) / 2 > blk) --a;
    while ((a + 1) * (a + 2) / 2 <= blk) ++a;
    const int c = blk - a * (a + 1) / 2;
    const int I = kSB * i + a, J = kSB * i + c, d = I - J;
    const int r = q / 6, cc = q - 6 * r;
    double v = 0.0;
    if (I < w.n_cam) {
      if (d <= bw) v = w.S[(size_t)I * rowlen + d * 36 + q];
      if (d == 0) {
        if (w.fixed[I] != 0) v = (r == cc) ? 1.0 : 0.0;
        else if (r == cc) v += damp(w.diagU[6 * (size_t)I + r], lambda);
      }
    } else if (d == 0) {
      v = (r == cc) ? 1.0 : 0.0;               // padding keyframes: identity
    }
    Di[e] = v;
  }
  // coupling to the previous super-block: block (a, c) = S(15 i + a, 15 (i-1) + c)
  for (int e = tid; e < kL; e += nt) {
    const int blk = e / 36, q = e - 36 * blk;
    const int a = blk / kSB, c = blk - kSB * a;
    const int I = kSB * i + a, J = kSB * (i - 1) + c, d = I - J;
    double v = 0.0;
    if (i > 0 && I < w.n_cam && d <= bw) v = w.S[(size_t)I * rowlen + d * 36 + q];
    Li[e] = v;
  }
  for (int e = tid; e < kV; e += nt) {
    const int I = kSB * i + e / 6;
    b[(size_t)i * kV + e] = (I < w.n_cam && w.fixed[I] == 0) ? w.b[6 * (size_t)kSB * i + e] : 0.0;
  }
  __syncthreads();
  // clear this super-block's rows of the accumulation buffers for the next LM iteration
  for (int a = 0; a < kSB; ++a) {
    const int I = kSB * i + a;
    if (I >= w.n_cam) break;
    for (int e = tid; e < rowlen; e += nt) w.S[(size_t)I * rowlen + e] = 0.0;
    if (tid < 6) { w.b[6 * (size_t)I + tid] = 0.0; w.diagU[6 * (size_t)I + tid] = 0.0; }
  }
  if (i == 0 && tid == 0) *w.cost_lin = 0.0;
}

// next level = the even super-blocks; their couplings start from zero
__global__ void __launch_bounds__(256)
k_bcr_next(const double* __restrict__ D, const double* __restrict__ b, double* __restrict__ D2,
           double* __restrict__ L2, double* __restrict__ b2) {
  const int j = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < kD; e += nt) D2[(size_t)j * kD + e] = D[(size_t)(2 * j) * kD + e];
  for (int e = tid; e < kL; e += nt) L2[(size_t)j * kL + e] = 0.0;
  for (int e = tid; e < kV; e += nt) b2[(size_t)j * kV + e] = b[(size_t)(2 * j) * kV + e];
}

// ------------------------------------------------------------------------------------------
// elimination of the odd super-blocks of one level (or of the last remaining super-block)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kET, 1)
k_bcr_elim(int N, const double* __restrict__ D, const double* __restrict__ L, const double* __restrict__ b,
           double* __restrict__ F, double* __restrict__ D2, double* __restrict__ L2, double* __restrict__ b2,
           int* fail) {
  extern __shared__ double sm[];
  double* __restrict__ A = sm;                      // kCells cells
  double* __restrict__ y = A + kCells * kCS;        // rhs row
  double* __restrict__ Dinv = y + kV;
  __shared__ int s_fail;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int e_idx = blockIdx.x;
  const int i = (N == 1) ? 0 : 2 * e_idx + 1;
  const bool has1 = (N > 1);                        // left neighbour i-1 (always there for an odd block)
  const bool has2 = (N > 1) && (i + 1 < N);

  if (tid == 0) s_fail = 0;
  // ---- load: D_i, E1 = L_i^T (rows = keyframes of i-1), E2 = L_(i+1) (rows = keyframes of i+1) ---------
  for (int e = tid; e < kD; e += nt) A[(e / 36) * kCS + (e % 36)] = D[(size_t)i * kD + e];
  for (int e = tid; e < kL; e += nt) {
    const int blk = e / 36, q = e - 36 * blk;
    const int a = blk / kSB, c = blk - kSB * a;     // E1 block (a, c), element q = (r, cc)
    const int r = q / 6, cc = q - 6 * r;
    // (L_i[c][a])^T : element (r, cc) = L_i block (c, a) element (cc, r)
    A[(kDB + blk) * kCS + q] = has1 ? L[(size_t)i * kL + (c * kSB + a) * 36 + cc * 6 + r] : 0.0;
    A[(kDB + kLB + blk) * kCS + q] = has2 ? L[(size_t)(i + 1) * kL + e] : 0.0;
  }
  for (int e = tid; e < kV; e += nt) y[e] = b[(size_t)i * kV + e];
  __syncthreads();

  auto factor_diag = [&](int K) {
    double Lb[36], inv[6];
    double* Dk = A + cell(K, K);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) Lb[6 * r + c] = (c <= r) ? Dk[6 * r + c] : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double akk = Lb[7 * k];
      ok = ok && (akk > 0.0) && (akk < 1e300);
      const double rr = bcr_rsqrt(akk);
      inv[k] = rr;
      Lb[7 * k] = akk * rr;
#pragma unroll
      for (int r = k + 1; r < 6; ++r) Lb[6 * r + k] *= rr;
#pragma unroll
      for (int r = k + 1; r < 6; ++r)
#pragma unroll
        for (int c = k + 1; c <= r; ++c) Lb[6 * r + c] = fma(-Lb[6 * r + k], Lb[6 * c + k], Lb[6 * r + c]);
    }
    if (!ok) s_fail = 1;
#pragma unroll
    for (int q = 0; q < 36; ++q) Dk[q] = Lb[q];
#pragma unroll
    for (int q = 0; q < 6; ++q) Dinv[6 * K + q] = inv[q];
  };
  if (tid == 0) factor_diag(0);
  __syncthreads();

  // ---- phase 1: right-looking partial Cholesky over the 15 block columns ---------------------------
  for (int K = 0; K < kSB; ++K) {
    if (s_fail) break;                             // uniform: written before the last barrier
    const int mD = kSB - 1 - K;                    // rows of D below the diagonal block
    {
      // panel: rows below in D, every row of E1 / E2, the rhs row:  v <- v * L_KK^-T
      const double* __restrict__ Lk = A + cell(K, K);
      const double* __restrict__ di = Dinv + 6 * K;
      const int nrow = 6 * (mD + 2 * kSB) + 1;
      for (int row = tid; row < nrow; row += nt) {
        double* v;
        if (row < nrow - 1) {
          const int I = K + 1 + row / 6;
          v = A + cell(I, K) + 6 * (row % 6);
        } else {
          v = y + 6 * K;
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
    __syncthreads();
    // trailing: blocks (I, J), K < J <= 14, J <= I <= 44; then the rhs rows
    const int nD = mD * (mD + 1) / 2, nE = 2 * kSB * mD;
    for (int t = tid; t < nD + nE + mD; t += nt) {
      if (t < nD + nE) {
        int I, J;
        if (t < nD) {
          int a = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
          while (a * (a + 1) / 2 > t) --a;
          while ((a + 1) * (a + 2) / 2 <= t) ++a;
          I = K + 1 + a;
          J = K + 1 + (t - a * (a + 1) / 2);
        } else {
          const int u = t - nD;
          I = kSB + u / mD;
          J = K + 1 + u % mD;
        }
        const double* __restrict__ Aik = A + cell(I, K);
        const double* __restrict__ Ajk = A + cell(J, K);
        double* __restrict__ Aij = A + cell(I, J);
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
        if (t == 0) factor_diag(K + 1);            // look-ahead: block (K+1,K+1) is final now
      } else {
        const int J = K + 1 + (t - nD - nE);
        const double* __restrict__ Ajk = A + cell(J, K);
        const double* yk = y + 6 * K;
        double* yj = y + 6 * J;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double acc = yj[c];
#pragma unroll
          for (int q = 0; q < 6; ++q) acc = fma(-yk[q], Ajk[6 * c + q], acc);
          yj[c] = acc;
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicOr(fail, 1);
    return;                                        // the finish kernel zeroes the step
  }

  // ---- factors to global memory for the back substitution -------------------------------------------
  double* __restrict__ Fe = F + (size_t)e_idx * kFac;
  for (int e = tid; e < kD; e += nt) Fe[e] = A[(e / 36) * kCS + (e % 36)];
  for (int e = tid; e < kV; e += nt) { Fe[kD + e] = Dinv[e]; Fe[kD + kV + 2 * kL + e] = y[e]; }
  for (int e = tid; e < 2 * kL; e += nt) Fe[kD + kV + e] = A[(kDB + e / 36) * kCS + (e % 36)];
  if (!has1) return;

  // ---- phase 2: Schur complement onto the even neighbours (next-level indices (i-1)/2, (i+1)/2) ------
  const int jl = (i - 1) / 2, jr = (i + 1) / 2;
  const int n_out = kDB + (has2 ? kDB + kLB : 0);
  for (int t = tid; t < n_out + (has2 ? 2 : 1) * kSB; t += nt) {
    if (t < n_out) {
      int rowbase, colbase, p, q2;
      double* dst;
      if (t < kDB || (has2 && t < 2 * kDB)) {
        const bool right = t >= kDB;
        const int blk = right ? t - kDB : t;
        p = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
        while (p * (p + 1) / 2 > blk) --p;
        while ((p + 1) * (p + 2) / 2 <= blk) ++p;
        q2 = blk - p * (p + 1) / 2;
        rowbase = colbase = right ? 2 * kSB : kSB;
        dst = D2 + (size_t)(right ? jr : jl) * kD + (size_t)blk * 36;
      } else {
        const int blk = t - 2 * kDB;
        p = blk / kSB;
        q2 = blk - kSB * p;
        rowbase = 2 * kSB;                          // rows: keyframes of i+1 (E2)
        colbase = kSB;                              // columns: keyframes of i-1 (E1)
        dst = L2 + (size_t)jr * kL + (size_t)blk * 36;
      }
      double acc[36];
#pragma unroll
      for (int q = 0; q < 36; ++q) acc[q] = 0.0;
      for (int k = 0; k < kSB; ++k) {
        const double* __restrict__ Pa = A + cell(rowbase + p, k);
        const double* __restrict__ Pb = A + cell(colbase + q2, k);
        double bj[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) bj[q] = Pb[q];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double a[6];
#pragma unroll
          for (int q = 0; q < 6; ++q) a[q] = Pa[6 * r + q];
#pragma unroll
          for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[6 * r + c] = fma(a[q], bj[6 * c + q], acc[6 * r + c]);
        }
      }
#pragma unroll
      for (int q = 0; q < 36; ++q) atomicAdd(dst + q, -acc[q]);
    } else {
      // rhs: b_left -= E1 y, b_right -= E2 y   (one keyframe row block per task)
      const int u = t - n_out;
      const bool right = u >= kSB;
      const int p = right ? u - kSB : u;
      const int rowbase = right ? 2 * kSB : kSB;
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < kSB; ++k) {
        const double* __restrict__ Pa = A + cell(rowbase + p, k);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
          for (int q = 0; q < 6; ++q) acc[r] = fma(Pa[6 * r + q], y[6 * k + q], acc[r]);
      }
      double* dst = b2 + (size_t)(right ? jr : jl) * kV + 6 * p;
#pragma unroll
      for (int r = 0; r < 6; ++r) atomicAdd(dst + r, -acc[r]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// back substitution of one level:  x_i = C^-T (y - E1^T x_(i-1) - E2^T x_(i+1))
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
k_bcr_back(int N, int shift, const double* __restrict__ F, double* __restrict__ x, const int* fail) {
  extern __shared__ double sm[];
  double* __restrict__ C = sm;                      // kDB cells
  double* __restrict__ Linv = C + kDB * kCS;        // inverse diagonal factors
  double* __restrict__ v = Linv + kSB * 36;
  double* __restrict__ xs = v + kV;
  if (*fail) return;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  const int e_idx = blockIdx.x;
  const int i = (N == 1) ? 0 : 2 * e_idx + 1;
  const bool has1 = (N > 1), has2 = (N > 1) && (i + 1 < N);
  const double* __restrict__ Fe = F + (size_t)e_idx * kFac;
  const double* __restrict__ dinv = Fe + kD;
  const double* __restrict__ E1 = Fe + kD + kV;
  const double* __restrict__ E2 = E1 + kL;
  const double* __restrict__ yv = E2 + kL;
  const double* __restrict__ xl = has1 ? x + ((size_t)(i - 1) << shift) * kV : x;
  const double* __restrict__ xr = has2 ? x + ((size_t)(i + 1) << shift) * kV : x;
  for (int e = tid; e < kD; e += nt) C[(e / 36) * kCS + (e % 36)] = Fe[e];
  // v_c = y_c - sum_a E1[a][c]^T x_left,a - sum_a E2[a][c]^T x_right,a     (one scalar per thread)
  for (int t = tid; t < kV; t += nt) {
    const int c = t / 6, cc = t - 6 * c;
    double a0 = yv[t], a1 = 0.0;
    if (has1) {
      for (int a = 0; a < kSB; ++a) {
        const double* __restrict__ B = E1 + (a * kSB + c) * 36 + cc;
#pragma unroll
        for (int r = 0; r < 6; ++r) a0 = fma(-B[6 * r], xl[6 * a + r], a0);
      }
    }
    if (has2) {
      for (int a = 0; a < kSB; ++a) {
        const double* __restrict__ B = E2 + (a * kSB + c) * 36 + cc;
#pragma unroll
        for (int r = 0; r < 6; ++r) a1 = fma(-B[6 * r], xr[6 * a + r], a1);
      }
    }
    v[t] = a0 + a1;
  }
  __syncthreads();
  for (int t = tid; t < 6 * kSB; t += nt) {
    // column j of the inverse of the lower-triangular diagonal factor of block K
    const int K = t / 6, j = t - 6 * K;
    const double* __restrict__ Lk = C + (K * (K + 1) / 2 + K) * kCS;
    double col[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double a = (r == j) ? 1.0 : 0.0;
#pragma unroll
      for (int q = 0; q < r; ++q) a = fma(-Lk[6 * r + q], col[q], a);
      col[r] = (r < j) ? 0.0 : a * dinv[6 * K + r];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) Linv[K * 36 + 6 * r + j] = col[r];
  }
  __syncthreads();
  if (tid < 32) {
    for (int K = kSB - 1; K >= 0; --K) {
      const double* __restrict__ Li = Linv + K * 36;
      double xv = 0.0;
      if (lane < 6) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double t2 = (r >= lane) ? Li[6 * r + lane] * v[6 * K + r] : 0.0;
          if (r & 1) a1 += t2; else a0 += t2;
        }
        xv = a0 + a1;
        xs[6 * K + lane] = xv;
      }
      double xk[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) xk[r] = __shfl_sync(kFull, xv, r);
      for (int t = lane; t < 6 * K; t += 32) {
        const int J = t / 6, c = t - 6 * J;
        const double* __restrict__ Ckj = C + (K * (K + 1) / 2 + J) * kCS;
        double a0 = v[t], a1 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          if (r & 1) a1 = fma(-Ckj[6 * r + c], xk[r], a1); else a0 = fma(-Ckj[6 * r + c], xk[r], a0);
        }
        v[t] = a0 + a1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  double* __restrict__ xo = x + ((size_t)i << shift) * kV;
  for (int t = tid; t < kV; t += nt) xo[t] = xs[t];
}

// camera step, failure handling, candidate poses
__global__ void __launch_bounds__(256)
k_bcr_finish(Work w, const double* __restrict__ x, const int* fail) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  bool bad = *fail != 0;
  // a non-finite component anywhere invalidates the step: checked per camera below and flagged
  const double* P = w.pose[cur];
  double* Pn = w.pose[cur ^ 1];
  for (int c = tid; c < w.n_cam; c += nth) {
    double R[9], Rn[9], dw[3], dt[3];
    bool fin = true;
#pragma unroll
    for (int q = 0; q < 6; ++q) fin = fin && isfinite(x[6 * (size_t)c + q]);
    if (!fin && !bad) atomicOr((int*)fail + 1, 1);
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = P[12 * (size_t)c + q];
    const bool upd = (w.fixed[c] == 0) && !bad && fin;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      dw[q] = upd ? x[6 * (size_t)c + q] : 0.0;
      dt[q] = upd ? x[6 * (size_t)c + 3 + q] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) w.dp[6 * (size_t)c + q] = upd ? x[6 * (size_t)c + q] : 0.0;
    if (upd) so3_exp_mul(dw, R, Rn);
#pragma unroll
    for (int q = 0; q < 9; ++q) Pn[12 * (size_t)c + q] = upd ? Rn[q] : R[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Pn[12 * (size_t)c + 9 + q] = P[12 * (size_t)c + 9 + q] + dt[q];
  }
}
__global__ void k_bcr_flag(Work w, int* fail) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    w.ctl->solve_fail = (fail[0] | fail[1]) ? 1 : 0;
    fail[0] = 0;
    fail[1] = 0;
  }
}

size_t elim_smem() { return sizeof(double) * (kCells * kCS + 2 * kV + 8); }
size_t back_smem() { return sizeof(double) * (kDB * kCS + kSB * 36 + 2 * kV + 8); }

}  // namespace

bool bcr_supported(int n_cam, int band_w) { return band_w >= 0 && band_w <= kSB && n_cam >= 4 * kSB; }

size_t bcr_workspace_doubles(int n_cam) { return make_plan(n_cam).total; }

int configure_bcr() {
  cudaError_t e = cudaFuncSetAttribute(k_bcr_elim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)elim_smem());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_smem());
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_rcs_bcr(const Work& w, double* ws, cudaStream_t s) {
  const BcrPlan p = make_plan(w.n_cam);
  int* fail = reinterpret_cast<int*>(ws + p.off_fail);
  double* x = ws + p.off_x;
  int launches = 0;
  k_bcr_assemble<<<p.N[0], 256, 0, s>>>(w, ws + p.off_D[0], ws + p.off_L[0], ws + p.off_b[0], fail);
  ++launches;
  for (int l = 0; l < p.n_levels; ++l) {
    const int N = p.N[l];
    double *D2 = nullptr, *L2 = nullptr, *b2 = nullptr;
    if (N > 1) {
      D2 = ws + p.off_D[l + 1]; L2 = ws + p.off_L[l + 1]; b2 = ws + p.off_b[l + 1];
      k_bcr_next<<<p.N[l + 1], 256, 0, s>>>(ws + p.off_D[l], ws + p.off_b[l], D2, L2, b2);
      ++launches;
    }
    k_bcr_elim<<<elim_count(N), kET, elim_smem(), s>>>(N, ws + p.off_D[l], ws + p.off_L[l], ws + p.off_b[l],
                                                      ws + p.off_F[l], D2, L2, b2, fail);
    ++launches;
  }
  for (int l = p.n_levels - 1; l >= 0; --l) {
    k_bcr_back<<<elim_count(p.N[l]), 256, back_smem(), s>>>(p.N[l], l, ws + p.off_F[l], x, fail);
    ++launches;
  }
  int blocks = (w.n_cam + 255) / 256;
  k_bcr_finish<<<blocks, 256, 0, s>>>(w, x, fail);
  k_bcr_flag<<<1, 32, 0, s>>>(w, fail);
  return launches + 2;
}

}  // namespace ba
