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
#include <string.h>

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
// Level 0 from the banded accumulation buffers.  Sharded runs (kPeers): the exchange of the reduced camera system IS
// this kernel -- a CTA stages its super-block's 15 band rows of S and its rows of b / diagU in shared memory as the SUM
// over the ranks' reduction buffers, read over NVLink peer memory (16-byte loads, four in flight per thread; a peer whose
// shard does not touch the super-block's cameras -- every rank posts its camera range with its flag -- is skipped: with
// contiguous point shards of a trajectory that leaves one or two ranks per row), then assembles from there.
constexpr int kFlagStride = 4;                // 64-bit words per (owner, poster) flag slot: epoch | first camera | last camera | -
constexpr int kAsmRowMax = (kSB + 1) * 36;    // doubles of one band row at the widest band
inline size_t assemble_smem() { return sizeof(double) * ((size_t)kSB * kAsmRowMax + 2 * kV + 2); }

template <bool kPeers>
__global__ void __launch_bounds__(256)
k_bcr_assemble(Work w, double* __restrict__ D, double* __restrict__ L, double* __restrict__ b, int* fail, PeerView pv, int it) {
  extern __shared__ __align__(16) double s_asm[];
  const int i = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int bw = w.band_w, rowlen = (bw + 1) * 36;
  LmCtl* ctl = w.ctl;
  const double lambda = ctl->lambda;
  const size_t off_b = w.s_len, off_dU = w.s_len + (size_t)w.n, off_cost = w.s_len + 2 * (size_t)w.n;
  double* s_S = s_asm;                                   // [15][rowlen]
  double* s_b = s_asm + (size_t)kSB * kAsmRowMax;        // [90]
  double* s_dU = s_b + kV;                               // [90]
  __shared__ int s_late;
  __shared__ int s_use[kMaxPeers];
  __shared__ double s_cost;
  const int I0 = kSB * i;
  const int n_rows = max(0, min(kSB, w.n_cam - I0));
  if (kPeers) {
    // this rank's linearisation is complete (stream order): tell every peer (epoch + my camera range), then wait for
    // theirs.  The wait is bounded: a peer that never arrives fails the step instead of hanging the GPU.
    const long long epoch = ctl->epoch_base + it + 1;
    if (tid == 0) { s_late = 0; s_cost = 0.0; }
    __syncthreads();
    if (tid < pv.world) {
      int use = 1;
      if (tid != pv.rank) {
        if (i == 0) {
          volatile long long* out = reinterpret_cast<volatile long long*>(pv.flag_remote[tid]);
          out[1] = pv.cam_range[0];
          out[2] = pv.cam_range[1];
          __threadfence_system();
          out[0] = epoch;
        }
        const volatile long long* f = reinterpret_cast<const volatile long long*>(pv.flag_local + kFlagStride * tid);
        const long long t0 = clock64();
        while (f[0] < epoch) {
          if (clock64() - t0 > 6000000000ll) { s_late = 1; break; }
        }
        __threadfence_system();
        const long long clo = f[1], chi = f[2];
        use = (chi >= I0 && clo <= I0 + kSB - 1) ? 1 : 0;
      }
      s_use[tid] = use;
    }
    __syncthreads();
    // the sum over the ranks, in rank order on every rank: the replicated solves then start from identical bits
    const int n2 = (n_rows * rowlen) >> 1;               // double2 entries (rowlen is a multiple of 36)
    bool first = true;
    for (int q = 0; q < pv.world; ++q) {
      if (!s_use[q]) continue;
      const double2* __restrict__ src = reinterpret_cast<const double2*>(pv.red[q] + (size_t)I0 * rowlen);
      double2* dst = reinterpret_cast<double2*>(s_S);
#pragma unroll 4
      for (int e = tid; e < n2; e += nt) {
        const double2 v = __ldcg(src + e);
        if (first) dst[e] = v;
        else { double2 o = dst[e]; o.x += v.x; o.y += v.y; dst[e] = o; }
      }
      if (tid < 6 * n_rows) {
        const double vb = __ldcg(pv.red[q] + off_b + 6 * (size_t)I0 + tid);
        const double vu = __ldcg(pv.red[q] + off_dU + 6 * (size_t)I0 + tid);
        if (first) { s_b[tid] = vb; s_dU[tid] = vu; } else { s_b[tid] += vb; s_dU[tid] += vu; }
      }
      first = false;
    }
    if (i == 0 && tid == 0) {
      double c = 0.0;
      for (int q = 0; q < pv.world; ++q) c += __ldcg(pv.red[q] + off_cost);
      s_cost = c;
    }
    __syncthreads();
  }
  auto S_at = [&](int I, int dq) -> double {             // entry dq of band row I of this super-block
    return kPeers ? s_S[(size_t)(I - I0) * rowlen + dq] : w.S[(size_t)I * rowlen + dq];
  };
  if (i == 0 && tid == 0) {
    fail[0] = (ctl->solve_fail != 0 || (kPeers && s_late)) ? 1 : 0;
    fail[1] = 0;
    if (ctl->iter == 0) {
      const double c0 = kPeers ? s_cost : *w.cost_lin;
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
    const int I = I0 + a, J = I0 + c, d = I - J;
    const int r = q / 6, cc = q - 6 * r;
    double v = 0.0;
    if (I < w.n_cam) {
      if (d <= bw) v = S_at(I, d * 36 + q);
      if (d == 0) {
        if (w.fixed[I] != 0) v = (r == cc) ? 1.0 : 0.0;
        else if (r == cc) v += damp(kPeers ? s_dU[6 * a + r] : w.diagU[6 * (size_t)I + r], lambda);
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
    const int I = I0 + a, J = kSB * (i - 1) + c, d = I - J;
    double v = 0.0;
    if (i > 0 && I < w.n_cam && d <= bw) v = S_at(I, d * 36 + q);
    Li[e] = v;
  }
  for (int e = tid; e < kV; e += nt) {
    const int I = I0 + e / 6;
    b[(size_t)i * kV + e] = (I < w.n_cam && w.fixed[I] == 0) ? (kPeers ? s_b[e] : w.b[6 * (size_t)I0 + e]) : 0.0;
  }
  __syncthreads();
  // clear this super-block's rows of the accumulation buffers for the next LM iteration (with peers reading them the
  // clearing waits for the candidate-cost all-reduce, which every rank reaches after this kernel: ba_api.cu)
  if (kPeers) return;
  for (int a = 0; a < kSB; ++a) {
    const int I = I0 + a;
    if (I >= w.n_cam) break;
    for (int e = tid; e < rowlen; e += nt) w.S[(size_t)I * rowlen + e] = 0.0;
    if (tid < 6) { w.b[6 * (size_t)I + tid] = 0.0; w.diagU[6 * (size_t)I + tid] = 0.0; }
  }
  if (i == 0 && tid == 0) *w.cost_lin = 0.0;
}

// block index of the lower triangle -> (row, column)
__device__ __forceinline__ void tri_decode(int blk, int& p, int& q) {
  p = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
  while (p * (p + 1) / 2 > blk) --p;
  while ((p + 1) * (p + 2) / 2 <= blk) ++p;
  q = blk - p * (p + 1) / 2;
}

// acc[2][6] += Pa[2 rows][6] * Pb[6][6]^T   (two rows of one 6x6 block product)
__device__ __forceinline__ void mac2(const double* __restrict__ Pa, const double* __restrict__ Pb, double* acc) {
  double a0[6], a1[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) { a0[q] = Pa[q]; a1[q] = Pa[6 + q]; }
#pragma unroll
  for (int c = 0; c < 6; ++c) {
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const double bq = Pb[6 * c + q];
      acc[c] = fma(a0[q], bq, acc[c]);
      acc[6 + c] = fma(a1[q], bq, acc[6 + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// elimination of the odd super-blocks of one level (or of the last remaining super-block), phase 1.
// R CTAs per super-block: the partial Cholesky of the tall panel needs no exchange between block ROWS of the
// two neighbour panels (row I of E only ever reads the D part of the same column), so every CTA factors the 15 x 15
// diagonal part redundantly and carries its own slice of the 30 neighbour rows -- the depth of a level is the
// 15 column steps of ONE small panel, not of 45 rows.  Tasks are two rows of a 6x6 block (72 FMAs), so that a column
// step is one or two rounds over the 512 threads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kET, 1)
k_bcr_elim(int N, int R, const double* __restrict__ D, const double* __restrict__ L, const double* __restrict__ b,
           double* __restrict__ F, int* fail) {
  extern __shared__ double sm[];
  double* __restrict__ A = sm;                      // kCells cells
  double* __restrict__ y = A + kCells * kCS;        // rhs row
  double* __restrict__ Dinv = y + kV;
  __shared__ int s_fail;
  if (*(volatile int*)fail) return;                 // an earlier level failed: the finish kernel zeroes the step
  const int tid = threadIdx.x, nt = blockDim.x;
  const int e_idx = blockIdx.x / R, rk = blockIdx.x - e_idx * R;
  const int i = (N == 1) ? 0 : 2 * e_idx + 1;
  const bool has1 = (N > 1);                        // left neighbour i-1 (always there for an odd block)
  const bool has2 = (N > 1) && (i + 1 < N);
  if (!has1 && rk != 0) return;
  // my slice [s0, s1) of the stacked neighbour rows: 0..14 = E1 (keyframes of i-1), 15..29 = E2 (keyframes of i+1)
  const int rp = (2 * kSB + R - 1) / R;
  const int slim = has2 ? 2 * kSB : (has1 ? kSB : 0);
  const int s0 = min(rk * rp, slim), s1 = min(s0 + rp, slim);
  const int nEr = s1 - s0;

  if (tid == 0) s_fail = 0;
  // ---- load: D_i, my rows of E1 = L_i^T and E2 = L_(i+1), the rhs --------------------------------------
  for (int e = tid; e < kD; e += nt) A[(e / 36) * kCS + (e % 36)] = D[(size_t)i * kD + e];
  for (int e = tid; e < nEr * kSB * 36; e += nt) {
    const int bl = e / 36, q = e - 36 * bl;
    const int s = s0 + bl / kSB, c = bl % kSB;
    double v;
    if (s < kSB) {
      // (L_i[c][s])^T : element (r, cc) = L_i block (c, s) element (cc, r)
      const int r = q / 6, cc = q - 6 * r;
      v = L[(size_t)i * kL + (c * kSB + s) * 36 + cc * 6 + r];
    } else {
      v = L[(size_t)(i + 1) * kL + ((s - kSB) * kSB + c) * 36 + q];
    }
    A[(kDB + s * kSB + c) * kCS + q] = v;
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

  // ---- right-looking partial Cholesky over the 15 block columns ------------------------------------
  for (int K = 0; K < kSB; ++K) {
    if (s_fail) break;                             // uniform: written before the last barrier
    const int mD = kSB - 1 - K;                    // rows of D below the diagonal block
    {
      // panel: rows below in D, my rows of E1 / E2, the rhs row:  v <- v * L_KK^-T
      const double* __restrict__ Lk = A + cell(K, K);
      const double* __restrict__ di = Dinv + 6 * K;
      const int nrow = 6 * (mD + nEr) + 1;
      for (int row = tid; row < nrow; row += nt) {
        double* v;
        if (row < nrow - 1) {
          const int br = row / 6;
          const int I = br < mD ? K + 1 + br : kSB + s0 + (br - mD);
          v = A + cell(I, K) + 6 * (row - 6 * br);
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
    // trailing: blocks (I, J), K < J <= 14, I >= J in D or one of my neighbour rows, two rows per task; the rhs rows
    const int nD = mD * (mD + 1) / 2, nE = nEr * mD;
    const int nfine = 3 * (nD + nE);
    for (int t = tid; t < nfine + mD; t += nt) {
      if (t < nfine) {
        const int blk = t / 3, part = t - 3 * blk;
        int I, J;
        if (blk < nD) {
          int a, c;
          tri_decode(blk, a, c);
          I = K + 1 + a;
          J = K + 1 + c;
        } else {
          const int u = blk - nD;
          I = kSB + s0 + u / mD;
          J = K + 1 + u % mD;
        }
        const double* __restrict__ Aik = A + cell(I, K) + 12 * part;
        const double* __restrict__ Ajk = A + cell(J, K);
        double* __restrict__ Aij = A + cell(I, J) + 12 * part;
        double acc[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) acc[q] = 0.0;
        mac2(Aik, Ajk, acc);
#pragma unroll
        for (int q = 0; q < 12; ++q) Aij[q] -= acc[q];
        if (t < 3) {                               // look-ahead: block (K+1,K+1) is final once its three parts are
          __syncwarp(0x7u);
          if (t == 0) factor_diag(K + 1);
        }
      } else {
        const int J = K + 1 + (t - nfine);
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
    return;
  }

  // ---- factors to global memory: my neighbour rows; the first CTA also C, 1/diag and y ------------------
  double* __restrict__ Fe = F + (size_t)e_idx * kFac;
  if (rk == 0) {
    for (int e = tid; e < kD; e += nt) Fe[e] = A[(e / 36) * kCS + (e % 36)];
    for (int e = tid; e < kV; e += nt) { Fe[kD + e] = Dinv[e]; Fe[kD + kV + 2 * kL + e] = y[e]; }
  }
  for (int e = tid; e < nEr * kSB * 36; e += nt) {
    const int bl = s0 * kSB + e / 36;
    Fe[kD + kV + (size_t)bl * 36 + (e % 36)] = A[(kDB + bl) * kCS + (e % 36)];
  }
}

// ------------------------------------------------------------------------------------------
// phase 2 of a level, gathered per super-block of the NEXT level (no REDs, no pre-copy, deterministic):
//   D'_j = D_2j - E2(2j-1) E2(2j-1)^T - E1(2j+1) E1(2j+1)^T        ("D" CTAs)
//   b'_j = b_2j - E2(2j-1) y(2j-1)    - E1(2j+1) y(2j+1)
//   L'_j =      - E2(2j-1) E1(2j-1)^T                              ("L" CTAs; coupling of j to j-1)
// P CTAs of each kind per super-block; a task is two rows of one output block, KS lanes share its 15 (30) block
// products and add up with shuffles.
// ------------------------------------------------------------------------------------------
constexpr int kUT = 512;

__global__ void __launch_bounds__(kUT, 1)
k_bcr_update(int N, int P, int ksD, int ksL, const double* __restrict__ D, const double* __restrict__ b,
             const double* __restrict__ F, double* __restrict__ D2, double* __restrict__ L2, double* __restrict__ b2,
             const int* fail) {
  extern __shared__ double sm[];
  double* __restrict__ Ea = sm;                     // kLB cells
  double* __restrict__ Eb = Ea + kLB * kCS;
  double* __restrict__ ya = Eb + kLB * kCS;
  double* __restrict__ yb = ya + kV;
  if (*fail) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int j = blockIdx.x / (2 * P), rem = blockIdx.x - j * 2 * P;
  const bool kindL = rem >= P;
  const int part = kindL ? rem - P : rem;
  const bool hasL = j >= 1, hasR = 2 * j + 1 < N;   // eliminated neighbours 2j-1 (index j-1) and 2j+1 (index j)
  if (kindL && !hasL) return;                       // L'_0 is never read
  const double* __restrict__ FL = F + (size_t)(hasL ? j - 1 : 0) * kFac;
  const double* __restrict__ FR = F + (size_t)j * kFac;
  // operands: Ea = E2 of the left elimination; Eb = E1 of the right one (D) / E1 of the left one (L)
  const double* __restrict__ srcA = FL + kD + kV + kL;
  const double* __restrict__ srcB = kindL ? FL + kD + kV : FR + kD + kV;
  const bool useA = hasL, useB = kindL ? true : hasR;
  for (int e = tid; e < kL; e += nt) {
    const int at = (e / 36) * kCS + (e % 36);
    if (useA) Ea[at] = srcA[e];
    if (useB) Eb[at] = srcB[e];
  }
  if (!kindL) {
    for (int e = tid; e < kV; e += nt) {
      ya[e] = hasL ? FL[kD + kV + 2 * kL + e] : 0.0;
      yb[e] = hasR ? FR[kD + kV + 2 * kL + e] : 0.0;
    }
  }
  __syncthreads();
  const int KS = kindL ? ksL : ksD;
  const int n_fine = 3 * (kindL ? kLB : kDB);
  const int per = (n_fine + P - 1) / P;
  const int t0 = part * per, t1 = min(n_fine, t0 + per);
  const int ks = tid % KS, slot = tid / KS, stride = nt / KS;
  for (int base = t0; base < t1; base += stride) {  // uniform trip count: the shuffles below need whole warps
    const int task = base + slot;
    const bool active = task < t1;
    const int blk = active ? task / 3 : 0, pr = active ? task - 3 * blk : 0;
    int p, q;
    if (kindL) { p = blk / kSB; q = blk - kSB * p; } else { tri_decode(blk, p, q); }
    double acc[12];
#pragma unroll
    for (int x = 0; x < 12; ++x) acc[x] = 0.0;
    if (active) {
      if (kindL) {
        for (int k = ks; k < kSB; k += KS) mac2(Ea + (p * kSB + k) * kCS + 12 * pr, Eb + (q * kSB + k) * kCS, acc);
      } else {
        if (hasL)
          for (int k = ks; k < kSB; k += KS) mac2(Ea + (p * kSB + k) * kCS + 12 * pr, Ea + (q * kSB + k) * kCS, acc);
        if (hasR)
          for (int k = ks; k < kSB; k += KS) mac2(Eb + (p * kSB + k) * kCS + 12 * pr, Eb + (q * kSB + k) * kCS, acc);
      }
    }
    for (int o = 1; o < KS; o <<= 1) {
#pragma unroll
      for (int x = 0; x < 12; ++x) acc[x] += __shfl_xor_sync(kFull, acc[x], o);
    }
    if (active && ks == 0) {
      const size_t at = (size_t)blk * 36 + 12 * pr;
      if (kindL) {
        double* __restrict__ dst = L2 + (size_t)j * kL + at;
#pragma unroll
        for (int x = 0; x < 12; ++x) dst[x] = -acc[x];
      } else {
        const double* __restrict__ src = D + (size_t)(2 * j) * kD + at;
        double* __restrict__ dst = D2 + (size_t)j * kD + at;
#pragma unroll
        for (int x = 0; x < 12; ++x) dst[x] = src[x] - acc[x];
      }
    }
  }
  // rhs of the next level: the last D CTA (it has the fewest block tasks)
  if (!kindL && part == P - 1) {
    for (int o = tid; o < kV; o += nt) {
      const int p = o / 6, r = o - 6 * p;
      double a0 = 0.0, a1 = 0.0;
      for (int k = 0; k < kSB; ++k) {
        const double* __restrict__ Pa = Ea + (p * kSB + k) * kCS + 6 * r;
        const double* __restrict__ Pb = Eb + (p * kSB + k) * kCS + 6 * r;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          if (hasL) a0 = fma(Pa[q], ya[6 * k + q], a0);
          if (hasR) a1 = fma(Pb[q], yb[6 * k + q], a1);
        }
      }
      b2[(size_t)j * kV + o] = b[(size_t)(2 * j) * kV + o] - (a0 + a1);
    }
  }
}

// ------------------------------------------------------------------------------------------
// back substitution of one level:  x_i = C^-T (y - E1^T x_(i-1) - E2^T x_(i+1))
// ------------------------------------------------------------------------------------------
constexpr int kBT = 512;
constexpr int kBSlices = 5;                        // the 15 neighbour keyframes in 5 slices of 3

__global__ void __launch_bounds__(kBT, 1)
k_bcr_back(int N, int shift, const double* __restrict__ F, double* __restrict__ x, const int* fail) {
  extern __shared__ double sm[];
  double* __restrict__ C = sm;                      // kDB cells
  double* __restrict__ Linv = C + kDB * kCS;        // inverse diagonal factors
  double* __restrict__ v = Linv + kSB * 36;
  double* __restrict__ xs = v + kV;
  double* __restrict__ ps = xs + kV;                // kBSlices x kV partial sums
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
  const double* xl = has1 ? x + ((size_t)(i - 1) << shift) * kV : x;
  const double* xr = has2 ? x + ((size_t)(i + 1) << shift) * kV : x;
  // partial sums of  E1[a][c]^T x_left,a + E2[a][c]^T x_right,a  over a slice of a   (one scalar per thread)
  if (tid < kBSlices * kV) {
    const int sl = tid / kV, o = tid - sl * kV;
    const int c = o / 6, cc = o - 6 * c;
    double a0 = 0.0, a1 = 0.0;
    if (has1) {
#pragma unroll
      for (int a = 3 * sl; a < 3 * sl + 3; ++a) {
        const double* __restrict__ B = E1 + (a * kSB + c) * 36 + cc;
#pragma unroll
        for (int r = 0; r < 6; ++r) a0 = fma(B[6 * r], xl[6 * a + r], a0);
      }
    }
    if (has2) {
#pragma unroll
      for (int a = 3 * sl; a < 3 * sl + 3; ++a) {
        const double* __restrict__ B = E2 + (a * kSB + c) * 36 + cc;
#pragma unroll
        for (int r = 0; r < 6; ++r) a1 = fma(B[6 * r], xr[6 * a + r], a1);
      }
    }
    ps[tid] = a0 + a1;
  }
  for (int e = tid; e < kD; e += nt) C[(e / 36) * kCS + (e % 36)] = Fe[e];
  __syncthreads();
  if (tid < kV) {
    double a = yv[tid];
#pragma unroll
    for (int sl = 0; sl < kBSlices; ++sl) a -= ps[sl * kV + tid];
    v[tid] = a;
  } else if (tid >= 96 && tid < 96 + 6 * kSB) {
    // column j of the inverse of the lower-triangular diagonal factor of block K
    const int t = tid - 96;
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
size_t update_smem() { return sizeof(double) * (2 * kLB * kCS + 2 * kV + 8); }
size_t back_smem() { return sizeof(double) * (kDB * kCS + kSB * 36 + (2 + kBSlices) * kV + 8); }

}  // namespace

bool bcr_supported(int n_cam, int band_w) { return band_w >= 0 && band_w <= kSB && n_cam >= 4 * kSB; }

size_t bcr_workspace_doubles(int n_cam) { return make_plan(n_cam).total; }

int configure_bcr() {
  cudaError_t e = cudaFuncSetAttribute(k_bcr_elim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)elim_smem());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_bcr_assemble<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)assemble_smem());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_bcr_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem());
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_smem());
  return e == cudaSuccess ? 0 : (int)e;
}

// lanes that share one task of the update kernel: as many (1, 2 or 4) as keep one round of tasks inside the CTA
static int pick_ks(int n_fine, int P) {
  const int per = (n_fine + P - 1) / P;
  return per * 4 <= kUT ? 4 : (per * 2 <= kUT ? 2 : 1);
}

int launch_rcs_bcr(const Work& w, double* ws, cudaStream_t s, int sm_count, const PeerView* pv, int it) {
  const BcrPlan p = make_plan(w.n_cam);
  int* fail = reinterpret_cast<int*>(ws + p.off_fail);
  double* x = ws + p.off_x;
  int launches = 0;
  if (sm_count <= 0) sm_count = 148;
  if (pv != nullptr && pv->world > 1) {
    k_bcr_assemble<true><<<p.N[0], 256, assemble_smem(), s>>>(w, ws + p.off_D[0], ws + p.off_L[0], ws + p.off_b[0], fail, *pv, it);
  } else {
    PeerView none;
    memset(&none, 0, sizeof(none));
    k_bcr_assemble<false><<<p.N[0], 256, 0, s>>>(w, ws + p.off_D[0], ws + p.off_L[0], ws + p.off_b[0], fail, none, it);
  }
  ++launches;
  for (int l = 0; l < p.n_levels; ++l) {
    const int N = p.N[l], ne = elim_count(N);
    // CTAs per eliminated super-block: the SMs the level leaves idle, up to 8 (4 neighbour rows each)
    const int R = N == 1 ? 1 : std::max(1, std::min(8, sm_count / ne));
    k_bcr_elim<<<ne * R, kET, elim_smem(), s>>>(N, R, ws + p.off_D[l], ws + p.off_L[l], ws + p.off_b[l],
                                                ws + p.off_F[l], fail);
    ++launches;
    if (N > 1) {
      const int N2 = p.N[l + 1];
      const int P = std::max(1, std::min(8, sm_count / (2 * N2)));
      k_bcr_update<<<N2 * 2 * P, kUT, update_smem(), s>>>(N, P, pick_ks(3 * kDB, P), pick_ks(3 * kLB, P),
                                                          ws + p.off_D[l], ws + p.off_b[l], ws + p.off_F[l],
                                                          ws + p.off_D[l + 1], ws + p.off_L[l + 1], ws + p.off_b[l + 1],
                                                          fail);
      ++launches;
    }
  }
  for (int l = p.n_levels - 1; l >= 0; --l) {
    k_bcr_back<<<elim_count(p.N[l]), kBT, back_smem(), s>>>(p.N[l], l, ws + p.off_F[l], x, fail);
    ++launches;
  }
  int blocks = (w.n_cam + 255) / 256;
  k_bcr_finish<<<blocks, 256, 0, s>>>(w, x, fail);
  k_bcr_flag<<<1, 32, 0, s>>>(w, fail);
  return launches + 2;
}

}  // namespace ba
