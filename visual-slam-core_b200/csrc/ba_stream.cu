// Streamed linearise + Schur kernel (the default for point-sorted input, any window size).
//
// Per chunk of points a CTA stages, in shared memory and SORTED BY CAMERA,
//   Y_ij  = W_ij C_j   with  C_j C_j^T = (V_j + lambda diag V_j)^-1        (6x3 per observation)
//   Jt_ij = sqrt(w) J_p , rt_ij = sqrt(w) r                               (pose rows, residual)
// and then builds, for every 6x6 block (a, b) of the reduced camera matrix that the chunk touches,
// the COMPACT LIST of (row of Y_aj, row of Y_bj) pairs (count -> scan -> fill with shared integer
// atomics).  The Schur products  S_ab -= Y_aj Y_bj^T  are then accumulated in registers by groups
// of Q lanes per block that walk the block's list with stride Q: every lane always has work
// (the previous tiled kernel looped every block over every point and tested a slot table, which
// left ~10-25 % of the lanes busy on sequential trajectories).  Because rows are sorted by camera,
// the Q lanes of a group read consecutive rows of Y: 128-bit shared loads without bank conflicts.
// The per-camera work (U_a, b_a, diag U_a) is one more task per camera over a contiguous row range.
//
//  * whole-window mode (n_cam <= 26): block -> thread assignment is static, the accumulators live in
//    registers across ALL chunks of the CTA and are flushed once at the end;
//  * chunk-window mode (any n_cam): the active blocks of each chunk are compacted, Q is chosen so
//    that all threads are busy, the Q partial sums are combined with a halving shuffle reduction
//    (37 shuffles for 36 values) and flushed with float64 REDs once per chunk and block.
// A chunk whose cameras span more than 48 keyframes (loop closure) or a point with more
// observations than a chunk holds goes through the warp-per-point scatter routine.
#include "ba_kernels.cuh"
#include "ba_scatter.cuh"

namespace ba {

namespace {

constexpr int kT = 384;                      // threads per CTA
constexpr int kNW = kT / 32;
constexpr int kMaxObs = 576;                 // observations per chunk
constexpr int kMaxPts = 96;                  // points per chunk
constexpr int kLmax = 48;                    // cameras in a chunk-local window
constexpr int kLdense = 26;                  // whole-window mode: 26*27/2 + 26 = 377 tasks <= kT
constexpr int kMaxTasks = kLmax * (kLmax + 1) / 2 + kLmax;
constexpr int kMaxPairs = 6656;              // (row, row) items per chunk
constexpr int kYS = 18;                      // doubles per Y row (144 B: 16-byte aligned rows)
constexpr int kJS = 14;                      // doubles per J row (112 B)

struct __align__(16) StreamSmem {
  double Y[kMaxObs * kYS];
  double J[kMaxObs * kJS];
  double Pt[kMaxPts * 9];       // C_j (c00 c01 c02 c11 c12 c22) and z_j = C^T g_l
  double red[kNW];
  unsigned pool[kMaxPairs];     // block lists: row(higher camera) | row(lower camera) << 16
  int toff[kMaxTasks + 1];      // list offsets per block
  int tcur[kMaxTasks];          // counters, then fill cursors
  int camA[kMaxObs];            // camera of each staged observation (observation order)
  int off[kMaxPts + 1];
  int camOff[kLmax + 1];        // rows of camera c: [camOff[c], camOff[c+1])
  int camCur[kLmax];
  int wsum[kNW];
  unsigned short act[kMaxTasks];  // compacted ids of the tasks with work
  short rowOf[kMaxObs];         // observation -> row (-1: fixed camera)
  short ptl[kMaxObs];           // observation -> chunk-local point
  short rowPt[kMaxObs];         // row -> chunk-local point
  unsigned char rowHead[kMaxObs];  // row is the first observation of its (point, camera)
  int npts, o_lo, cmin, cmax, nact;
};

__device__ __forceinline__ int lower_bound_off(const int* __restrict__ off, int n, int target) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// the rare fallback stays out of line so that its registers do not count against the hot loops
__device__ __noinline__ void scatter_point_cold(const Work& w, int j, int lane, double lambda, double delta,
                                                const double* pose, const double* X, double (*sW)[19],
                                                int* sCam, double* cost) {
  double c = 0.0;
  scatter_point(w, j, lane, lambda, delta, pose, X, sW, sCam, c);
  *cost += c;
}

// acc (6x6, row-major) += Ya (6x3) * Yb^T (3x6)
__device__ __forceinline__ void block_item(double (&acc)[36], const double* __restrict__ Ya,
                                           const double* __restrict__ Yb) {
  const double2* __restrict__ pb = reinterpret_cast<const double2*>(Yb);
  double yb[18];
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const double2 t = pb[q];
    yb[2 * q] = t.x;
    yb[2 * q + 1] = t.y;
  }
  const double2* __restrict__ pa = reinterpret_cast<const double2*>(Ya);
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const double2 t0 = pa[3 * h], t1 = pa[3 * h + 1], t2 = pa[3 * h + 2];
    const double a0 = t0.x, a1 = t0.y, a2 = t1.x, b0 = t1.y, b1 = t2.x, b2 = t2.y;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      acc[12 * h + c] = fma(a0, yb[3 * c], acc[12 * h + c]);
      acc[12 * h + 6 + c] = fma(b0, yb[3 * c], acc[12 * h + 6 + c]);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      acc[12 * h + c] = fma(a1, yb[3 * c + 1], acc[12 * h + c]);
      acc[12 * h + 6 + c] = fma(b1, yb[3 * c + 1], acc[12 * h + 6 + c]);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      acc[12 * h + c] = fma(a2, yb[3 * c + 2], acc[12 * h + c]);
      acc[12 * h + 6 + c] = fma(b2, yb[3 * c + 2], acc[12 * h + 6 + c]);
    }
  }
}

// per-camera work of one row: acc[0..20] += lower(Jt^T Jt), acc[21..26] += Y z - Jt^T rt
__device__ __forceinline__ void camera_item(double (&acc)[36], const StreamSmem& sm, int row) {
  const double2* __restrict__ pj = reinterpret_cast<const double2*>(sm.J + row * kJS);
  double p[14];
#pragma unroll
  for (int q = 0; q < 7; ++q) {
    const double2 t = pj[q];
    p[2 * q] = t.x;
    p[2 * q + 1] = t.y;
  }
  int e = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c <= r; ++c) {
      acc[e] = fma(p[r], p[c], fma(p[6 + r], p[6 + c], acc[e]));
      ++e;
    }
#pragma unroll
  for (int r = 0; r < 6; ++r) acc[21 + r] -= fma(p[r], p[12], p[6 + r] * p[13]);
  if (sm.rowHead[row]) {
    const double* __restrict__ y = sm.Y + row * kYS;
    const double* __restrict__ z = sm.Pt + sm.rowPt[row] * 9 + 6;
    const double z0 = z[0], z1 = z[1], z2 = z[2];
#pragma unroll
    for (int r = 0; r < 6; ++r) acc[21 + r] += y[3 * r] * z0 + y[3 * r + 1] * z1 + y[3 * r + 2] * z2;
  }
}

// One level of the halving reduction over the lanes of a group: the lane whose `m` bit is clear
// keeps the first `kHalf` of the kN live values, its partner the rest.
template <int kN, int kHalf>
__device__ __forceinline__ void halve(double (&acc)[36], int m, int q, int& base, int& cnt) {
  const bool hi_lane = (q & m) != 0;
#pragma unroll
  for (int i = 0; i < kHalf; ++i) {
    const double lo = acc[i];
    const double hi = (i + kHalf < kN) ? acc[i + kHalf] : 0.0;
    const double send = hi_lane ? lo : hi;
    const double keep = hi_lane ? hi : lo;
    acc[i] = keep + __shfl_xor_sync(kFull, send, m);
  }
  base += hi_lane ? kHalf : 0;
  cnt = hi_lane ? max(cnt - kHalf, 0) : min(cnt, kHalf);
}

// Combine the Q partial accumulators of a task and add them to the reduction buffers.
// B < ntri: 6x6 block (hi, lo) of the window, S -= acc;  B >= ntri: camera B - ntri.
__device__ __forceinline__ void flush_task(const Work& w, double (&acc)[36], int Q, int q, int B,
                                           int ntri, int cmin) {
  int base = 0, cnt = 36;
  // all lanes of the warp take the same path (Q is uniform per CTA)
  if (Q >= 2) halve<36, 18>(acc, Q >> 1, q, base, cnt);
  if (Q >= 4) halve<18, 9>(acc, Q >> 2, q, base, cnt);
  if (Q >= 8) halve<9, 5>(acc, Q >> 3, q, base, cnt);
  if (Q >= 16) halve<5, 3>(acc, Q >> 4, q, base, cnt);
  if (Q >= 32) halve<3, 2>(acc, Q >> 5, q, base, cnt);
  if (B < 0) return;
  if (B < ntri) {
    int hi = (int)((sqrtf(8.0f * (float)B + 1.0f) - 1.0f) * 0.5f);
    while (hi * (hi + 1) / 2 > B) --hi;
    while ((hi + 1) * (hi + 2) / 2 <= B) ++hi;
    const int lo = B - hi * (hi + 1) / 2;
    double* __restrict__ Sd = s_block_ptr(w, cmin + hi, cmin + lo);
    const bool diag = hi == lo;
#pragma unroll
    for (int i = 0; i < 36; ++i) {
      if (i < cnt) {
        const int e = base + i;
        const double v = acc[i];
        const int r = e / 6, c = e - 6 * r;
        if (v != 0.0 && !(diag && c > r)) atomicAdd(Sd + e, -v);
      }
    }
  } else {
    const int cam = cmin + (B - ntri);
    double* __restrict__ Sd = s_block_ptr(w, cam, cam);
#pragma unroll
    for (int i = 0; i < 27; ++i) {
      if (i < cnt) {
        const int e = base + i;
        const double v = acc[i];
        if (e < 21) {
          int r = (e >= 15) ? 5 : (e >= 10) ? 4 : (e >= 6) ? 3 : (e >= 3) ? 2 : (e >= 1) ? 1 : 0;
          const int c = e - r * (r + 1) / 2;
          if (v != 0.0) {
            atomicAdd(Sd + 6 * r + c, v);
            if (r == c) atomicAdd(w.diagU + 6 * (size_t)cam + r, v);
          }
        } else if (e < 27) {
          if (v != 0.0) atomicAdd(w.b + 6 * (size_t)cam + (e - 21), v);
        }
      }
    }
  }
}

template <bool kDense>
__global__ void __launch_bounds__(kT, 1)
k_linearize_stream(Work w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StreamSmem& sm = *reinterpret_cast<StreamSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double lambda = ctl->lambda, delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];

  // this CTA's points: equal observation share
  const long long nobs = w.n_obs;
  const int t_lo = (int)(nobs * blockIdx.x / gridDim.x);
  const int t_hi = (int)(nobs * (blockIdx.x + 1) / gridDim.x);
  const int p_lo = (blockIdx.x == 0) ? 0 : lower_bound_off(w.pt_off, w.n_pt, t_lo);
  const int p_hi = (blockIdx.x == gridDim.x - 1) ? w.n_pt : lower_bound_off(w.pt_off, w.n_pt, t_hi);

  // whole-window mode: static tasks, persistent accumulators
  const int Ld = w.n_cam;
  const int ntri_d = Ld * (Ld + 1) / 2, ntask_d = ntri_d + Ld;
  int Qd = 1;
  if (kDense) while (2 * Qd * ntask_d <= kT && Qd < 32) Qd *= 2;

  double acc[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) acc[q] = 0.0;
  double cost = 0.0;

  int p = p_lo;
  int chunk_no = -1;
#define STREAM_MARK(ph) do { if (chunk_no < 6) dbg_mark(w, 100 + 16 * chunk_no + (ph), 200 + (ph)); } while (0)
  while (p < p_hi) {
    ++chunk_no;
    STREAM_MARK(0);
    // ---- chunk bounds: whole points within kMaxPts points / kMaxObs observations / kMaxPairs pairs
    if (wid == 0) {
      const int q0 = p + 3 * lane;
      int o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (q0 + e <= p_hi) ? __ldg(w.pt_off + q0 + e) : 0x3fffffff;
      const int base = __shfl_sync(kFull, o[0], 0);
      int pr[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        const int k = min(max(o[e + 1] - o[e], 0), 4096);
        pr[e] = (q0 + e + 1 <= p_hi) ? k * (k + 1) / 2 : 0;
      }
      const int s3 = pr[0] + pr[1] + pr[2];
      const int excl = warp_incl_scan(s3, lane) - s3;
      int run = excl, cnt = 0;
      bool ok = true;
#pragma unroll
      for (int e = 0; e < 3; ++e) {
        run += pr[e];
        ok = ok && (q0 + e + 1 <= p_hi) && (o[e + 1] - base <= kMaxObs) && (run <= kMaxPairs);
        cnt += ok ? 1 : 0;
      }
      // a lane contributes only if every earlier lane was complete (prefix sums are monotone, so
      // this holds automatically; the ballot makes it explicit)
      const unsigned full = __ballot_sync(kFull, cnt == 3);
      const int first_partial = __ffs(~full) - 1;          // 32 -> -1 never: ~full==0 gives ffs 0 -> -1
      int total;
      if (full == kFull) total = 96;
      else total = 3 * first_partial + __shfl_sync(kFull, cnt, first_partial);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = 3 * lane + e;
        if ((e < 3 || idx == total) && idx <= total) sm.off[idx] = o[e] - base;
      }
      if (lane == 0) {
        sm.npts = total;
        sm.o_lo = base;
        sm.cmin = 0x7fffffff;
        sm.cmax = -1;
      }
    } else if (tid < 32 + kLmax) {
      sm.camCur[tid - 32] = 0;
    }
    __syncthreads();
    const int npts = sm.npts;
    const int o_lo = sm.o_lo;
    if (npts == 0) {
      // a single point with more observations than a chunk holds
      if (wid == 0)
        scatter_point_cold(w, p, lane, lambda, delta, pose, X, reinterpret_cast<double (*)[19]>(sm.Y), sm.camA, &cost);
      p += 1;
      __syncthreads();
      continue;
    }
    const int nobs_c = sm.off[npts];

    // ---- camera window of the chunk ------------------------------------------------------------
    if (!kDense) {
      int mn = 0x7fffffff, mx = -1;
      for (int i = tid; i < nobs_c; i += kT) {
        const int cam = __ldg(w.obs_cam + o_lo + i);
        sm.camA[i] = cam;
        mn = min(mn, cam);
        mx = max(mx, cam);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(kFull, mn, o));
        mx = max(mx, __shfl_xor_sync(kFull, mx, o));
      }
      if (lane == 0 && mx >= 0) { atomicMin(&sm.cmin, mn); atomicMax(&sm.cmax, mx); }
      __syncthreads();
    }
    const int cmin = kDense ? 0 : sm.cmin;
    const int L = kDense ? Ld : (sm.cmax - sm.cmin + 1);
    if (!kDense && L > kLmax) {
      // cameras too far apart for the local window (e.g. a loop closure): scatter this chunk
      double (*sW)[19] = reinterpret_cast<double (*)[19]>(sm.Y) + wid * 32;
      int* sCam = sm.camA + wid * 32;
      __syncthreads();
      for (int pj = wid; pj < npts; pj += kNW)
        scatter_point_cold(w, p + pj, lane, lambda, delta, pose, X, sW, sCam, &cost);
      p += npts;
      __syncthreads();
      continue;
    }
    const int ntri = L * (L + 1) / 2, ntask = ntri + L;
    STREAM_MARK(1);

    // ---- pass 1: per-observation V / g_l contributions, cost, rows per camera ------------------
    for (int i = tid; i < nobs_c; i += kT) {
      const int o = o_lo + i;
      const int cam = kDense ? __ldg(w.obs_cam + o) : sm.camA[i];
      const int pj = __ldg(w.obs_pt + o) - p;
      const float2 uv = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
      const double* Xj = X + 3 * (size_t)(p + pj);
      ObsLin ol;
      linearize_obs<false>(w.K, pose + 12 * (size_t)cam, pose + 12 * (size_t)cam + 9, Xj[0], Xj[1], Xj[2],
                           (double)uv.x, (double)uv.y, delta, ol);
      double* s = sm.Y + i * kYS;
      s[0] = ol.wu * ol.lu[0] * ol.lu[0] + ol.wv * ol.lv[0] * ol.lv[0];
      s[1] = ol.wu * ol.lu[0] * ol.lu[1] + ol.wv * ol.lv[0] * ol.lv[1];
      s[2] = ol.wu * ol.lu[0] * ol.lu[2] + ol.wv * ol.lv[0] * ol.lv[2];
      s[3] = ol.wu * ol.lu[1] * ol.lu[1] + ol.wv * ol.lv[1] * ol.lv[1];
      s[4] = ol.wu * ol.lu[1] * ol.lu[2] + ol.wv * ol.lv[1] * ol.lv[2];
      s[5] = ol.wu * ol.lu[2] * ol.lu[2] + ol.wv * ol.lv[2] * ol.lv[2];
      const double au = ol.wu * ol.ru, av = ol.wv * ol.rv;
      s[6] = au * ol.lu[0] + av * ol.lv[0];
      s[7] = au * ol.lu[1] + av * ol.lv[1];
      s[8] = au * ol.lu[2] + av * ol.lv[2];
      sm.ptl[i] = (short)pj;
      if (kDense) sm.camA[i] = cam;
      cost += huber_rho(ol.ru, delta) + huber_rho(ol.rv, delta);
      if (w.fixed[cam] == 0) atomicAdd(&sm.camCur[cam - cmin], 1);
    }
    __syncthreads();
    STREAM_MARK(2);

    // ---- per point: V, g_l -> C_j (C C^T = V^-1), z_j = C^T g_l; per camera: row ranges --------
    if (tid < npts) {
      double V[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0};
      for (int i = sm.off[tid]; i < sm.off[tid + 1]; ++i) {
        const double* s = sm.Y + i * kYS;
#pragma unroll
        for (int q = 0; q < 6; ++q) V[q] += s[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) gl[q] += s[6 + q];
      }
      V[0] += damp(V[0], lambda);
      V[3] += damp(V[3], lambda);
      V[5] += damp(V[5], lambda);
      // LDL^T of V; V^-1 = M^T D^-1 M with M = L^-1 (unit lower) -> C = M^T D^-1/2 (upper)
      const double a = V[0], b = V[1], c = V[2], d = V[3], e = V[4], f = V[5];
      const double i0 = 1.0 / a;
      const double l10 = b * i0, l20 = c * i0;
      const double d1 = d - l10 * b;
      const double e1 = e - l20 * b;
      const double i1 = 1.0 / d1;
      const double l21 = e1 * i1;
      const double d2 = f - l20 * c - l21 * e1;
      const double i2 = 1.0 / d2;
      const double m10 = -l10, m20 = l10 * l21 - l20, m21 = -l21;
      const bool ok = (a > 0.0) && (d1 > 0.0) && (d2 > 0.0) && isfinite(i0) && isfinite(i1) && isfinite(i2);
      const double s0 = ok ? sqrt(i0) : 0.0, s1 = ok ? sqrt(i1) : 0.0, s2 = ok ? sqrt(i2) : 0.0;
      double* P = sm.Pt + tid * 9;
      const double c00 = s0, c01 = m10 * s1, c02 = m20 * s2, c11 = s1, c12 = m21 * s2, c22 = s2;
      P[0] = c00; P[1] = c01; P[2] = c02; P[3] = c11; P[4] = c12; P[5] = c22;
      P[6] = c00 * gl[0];
      P[7] = c01 * gl[0] + c11 * gl[1];
      P[8] = c02 * gl[0] + c12 * gl[1] + c22 * gl[2];
    }
    if (wid == kNW - 1) {
      const int v0 = (lane < L) ? sm.camCur[lane] : 0;
      const int v1 = (lane + 32 < L) ? sm.camCur[lane + 32] : 0;
      const int i0 = warp_incl_scan(v0, lane);
      const int t0 = __shfl_sync(kFull, i0, 31);
      const int i1 = warp_incl_scan(v1, lane) + t0;
      __syncwarp();
      if (lane < L) { sm.camOff[lane] = i0 - v0; sm.camCur[lane] = i0 - v0; }
      if (lane + 32 < L) { sm.camOff[lane + 32] = i1 - v1; sm.camCur[lane + 32] = i1 - v1; }
      const int tot = __shfl_sync(kFull, i1, 31);
      if (lane == 0) sm.camOff[L] = tot;
    }
    for (int t = tid; t < ntask; t += kT) sm.tcur[t] = 0;
    __syncthreads();
    STREAM_MARK(3);

    // ---- pass 2: rows sorted by camera: Y = W C, sqrt-weighted pose rows ------------------------
    for (int i = tid; i < nobs_c; i += kT) {
      const int o = o_lo + i;
      const int cam = sm.camA[i];
      if (w.fixed[cam] != 0) { sm.rowOf[i] = -1; continue; }
      const int pj = sm.ptl[i];
      bool head = true;
      for (int i2 = sm.off[pj]; i2 < i; ++i2) head = head && (sm.camA[i2] != cam);
      const int row = atomicAdd(&sm.camCur[cam - cmin], 1);
      sm.rowOf[i] = (short)row;
      sm.rowPt[row] = (short)pj;
      sm.rowHead[row] = head ? 1 : 0;
      const float2 uv = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
      const double* Xj = X + 3 * (size_t)(p + pj);
      ObsLin ol;
      linearize_obs<true>(w.K, pose + 12 * (size_t)cam, pose + 12 * (size_t)cam + 9, Xj[0], Xj[1], Xj[2],
                          (double)uv.x, (double)uv.y, delta, ol);
      const double* P = sm.Pt + pj * 9;
      const double c00 = P[0], c01 = P[1], c02 = P[2], c11 = P[3], c12 = P[4], c22 = P[5];
      double* y = sm.Y + row * kYS;
      double* jt = sm.J + row * kJS;
      const double su = sqrt(ol.wu), sv = sqrt(ol.wv);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double pa = ol.wu * ol.pu[r], pb = ol.wv * ol.pv[r];
        const double w0 = pa * ol.lu[0] + pb * ol.lv[0];
        const double w1 = pa * ol.lu[1] + pb * ol.lv[1];
        const double w2 = pa * ol.lu[2] + pb * ol.lv[2];
        y[3 * r + 0] = w0 * c00;
        y[3 * r + 1] = w0 * c01 + w1 * c11;
        y[3 * r + 2] = w0 * c02 + w1 * c12 + w2 * c22;
        jt[r] = su * ol.pu[r];
        jt[6 + r] = sv * ol.pv[r];
      }
      jt[12] = su * ol.ru;
      jt[13] = sv * ol.rv;
    }
    __syncthreads();
    STREAM_MARK(4);

    // ---- repeated (point, camera) observations fold into the first one; count the block lists --
    for (int i = tid; i < nobs_c; i += kT) {
      const int row = sm.rowOf[i];
      if (row < 0) continue;
      const int cam = sm.camA[i];
      const int pj = sm.ptl[i];
      if (!sm.rowHead[row]) {
        int ih = sm.off[pj];
        while (sm.camA[ih] != cam) ++ih;
        double* yh = sm.Y + sm.rowOf[ih] * kYS;
        const double* ys = sm.Y + row * kYS;
#pragma unroll
        for (int q = 0; q < 18; ++q) atomicAdd(yh + q, ys[q]);
        continue;
      }
      const int cl = cam - cmin;
      const int i_end = sm.off[pj + 1];
      for (int i2 = i; i2 < i_end; ++i2) {
        const int r2 = sm.rowOf[i2];
        if (r2 < 0 || !sm.rowHead[r2]) continue;
        const int c2 = sm.camA[i2] - cmin;
        const int hi = max(cl, c2), lo = min(cl, c2);
        atomicAdd(&sm.tcur[hi * (hi + 1) / 2 + lo], 1);
      }
    }
    __syncthreads();
    STREAM_MARK(5);

    // ---- scan: list offsets and the compacted list of tasks with work --------------------------
    {
      int v[4], s = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = 4 * tid + e;
        int c = 0, a = 0;
        if (t < ntri) { c = sm.tcur[t]; a = c > 0; }
        else if (t < ntask) { a = sm.camOff[t - ntri + 1] > sm.camOff[t - ntri]; }
        v[e] = (c << 12) | a;
        s += v[e];
      }
      const int incl = warp_incl_scan(s, lane);
      if (lane == 31) sm.wsum[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        const int x = (lane < kNW) ? sm.wsum[lane] : 0;
        const int xi = warp_incl_scan(x, lane);
        if (lane < kNW) sm.wsum[lane] = xi - x;
        if (lane == kNW - 1) { sm.nact = xi & 0xfff; sm.toff[ntask] = xi >> 12; }
      }
      __syncthreads();
      int run = sm.wsum[wid] + incl - s;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int t = 4 * tid + e;
        if (t < ntask) {
          sm.toff[t] = run >> 12;
          sm.tcur[t] = run >> 12;
          if (v[e] & 1) sm.act[run & 0xfff] = (unsigned short)t;
        }
        run += v[e];
      }
    }
    __syncthreads();
    STREAM_MARK(6);

    // ---- fill the block lists -------------------------------------------------------------------
    for (int i = tid; i < nobs_c; i += kT) {
      const int row = sm.rowOf[i];
      if (row < 0 || !sm.rowHead[row]) continue;
      const int cl = sm.camA[i] - cmin;
      const int i_end = sm.off[sm.ptl[i] + 1];
      for (int i2 = i; i2 < i_end; ++i2) {
        const int r2 = sm.rowOf[i2];
        if (r2 < 0 || !sm.rowHead[r2]) continue;
        const int c2 = sm.camA[i2] - cmin;
        const int hi = max(cl, c2), lo = min(cl, c2);
        const int pos = atomicAdd(&sm.tcur[hi * (hi + 1) / 2 + lo], 1);
        const unsigned rhi = (cl >= c2) ? (unsigned)row : (unsigned)r2;
        const unsigned rlo = (cl >= c2) ? (unsigned)r2 : (unsigned)row;
        sm.pool[pos] = rhi | (rlo << 16);
      }
    }
    __syncthreads();
    STREAM_MARK(7);

    // ---- accumulate: Q lanes per task walk its list with stride Q ------------------------------
    {
      const int nrun = kDense ? ntask_d : sm.nact;
      int Q = Qd;
      if (!kDense) {
        Q = 1;
        while (2 * Q * nrun <= kT && Q < 32) Q *= 2;
      }
      const int G = kT / Q;
      const int grp = tid / Q, q = tid - grp * Q;
      for (int base = 0; base < nrun; base += G) {
        const int task = base + grp;
        int B = -1;
        if (task < nrun) B = kDense ? task : (int)sm.act[task];
        if (B >= 0) {
          if (B < ntri) {
            const int end = sm.toff[B + 1];
            for (int it = sm.toff[B] + q; it < end; it += Q) {
              const unsigned item = sm.pool[it];
              block_item(acc, sm.Y + (item & 0xffffu) * kYS, sm.Y + (item >> 16) * kYS);
            }
          } else {
            const int c = B - ntri;
            const int end = sm.camOff[c + 1];
            for (int row = sm.camOff[c] + q; row < end; row += Q) camera_item(acc, sm, row);
          }
        }
        if (!kDense) {
          flush_task(w, acc, Q, q, B, ntri, cmin);
#pragma unroll
          for (int i = 0; i < 36; ++i) acc[i] = 0.0;
        }
      }
    }
    STREAM_MARK(8);
    p += npts;
    __syncthreads();
    STREAM_MARK(9);
  }
  dbg_mark(w, 98, 298);

  if (kDense) {
    const int grp = tid / Qd, q = tid - grp * Qd;
    flush_task(w, acc, Qd, q, grp < ntask_d ? grp : -1, ntri_d, 0);
  }

  cost = warp_sum(cost);
  if (lane == 0) sm.red[wid] = cost;
  __syncthreads();
  if (tid == 0) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kNW; ++i) c += sm.red[i];
    if (c != 0.0) atomicAdd(w.cost_lin, c);
  }
  dbg_mark(w, 99, 299);
}

}  // namespace

int configure_stream() {
  const int bytes = (int)sizeof(StreamSmem);
  cudaError_t e = cudaFuncSetAttribute(k_linearize_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_linearize_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_linearize_stream(const Work& w, int sm_count, cudaStream_t s) {
  int blocks = sm_count;
  const int by_obs = (w.n_obs + 127) / 128;            // do not spread tiny windows over idle CTAs
  if (blocks > by_obs) blocks = by_obs;
  if (blocks < 1) blocks = 1;
  if (w.n_cam <= kLdense) k_linearize_stream<true><<<blocks, kT, sizeof(StreamSmem), s>>>(w);
  else k_linearize_stream<false><<<blocks, kT, sizeof(StreamSmem), s>>>(w);
  return 1;
}

}  // namespace ba
