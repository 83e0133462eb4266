// Linearise + Schur on the FP64 tensor path (mma.sync m8n8k4 f64), the default for local windows
// and for sequential (banded) co-visibility.  One CTA per SM, 640 threads, warp specialised:
//
//  * 12 STAGING warps (120 registers each, setmaxnreg).  A warp linearises one point at a time with
//    lane a <-> camera (base + a) of a window of kWin keyframes: the point block V_j / g_l is a
//    butterfly reduction over the lanes, every lane then holds C_j (C C^T = (V + lambda diag V)^-1),
//    forms Y_aj = W_aj C_j (6x3) in registers and
//      - writes it into a DENSE shared-memory panel  Yt[3 p + d][6 a + r]  (zeros where camera a does
//        not see point p),
//      - adds the camera's own sums U_a = sum J_p^T w J_p and b_a into a PRIVATE shared-memory column
//        (a lane always serves the same camera: no atomics, no pose gathers, no slot tables on the
//        per-observation path).
//    Offsets, camera indices, pixels and coordinates of the NEXT point are prefetched while the
//    current one is linearised.
//  * 8 MMA warps (56 registers each).  The Schur complement of a panel of 12 points is the symmetric
//    product  S -= Yt^T Yt : a warp owns two 24x24 macro tiles (4 x 4 cameras, 3 x 3 mma tiles) whose
//    accumulators live in a thread-private fragment layout in shared memory between panels (two
//    conflict-free 64-bit accesses per element and panel), so the register file goes to the
//    latency-bound staging warps.  Operand fragments are 64-bit shared loads, conflict-free because
//    the panel stride is 8 (mod 16) doubles.
//  * The panel is double buffered between the two groups with named barriers (full / empty), so
//    global-load latency, FP64 FMAs and DMMAs overlap without a CTA-wide barrier.
//
//  window "whole"   (n_cam <= kWin): base = 0, one flush of float64 REDs per CTA.
//  window "sliding" (any n_cam): the window follows the points (which must arrive roughly sorted
//    along the trajectory); accumulators are flushed only when the window has to move, and macro
//    tiles whose cameras do not appear in a panel are skipped.
// A point outside the window, with repeated (point, keyframe) observations or with more than 32
// observations goes through the warp-per-point scatter routine (ba_scatter.cuh) on the spot.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "ba_kernels.cuh"
#include "ba_scatter.cuh"

namespace ba {

namespace {

constexpr int kSW = 12;                 // staging warps (3 warpgroups)
constexpr int kXW = 8;                  // mma warps (2 warpgroups)
constexpr int kMT = (kSW + kXW) * 32;
constexpr int kST = kSW * 32;
constexpr int kWin = 20;                // cameras per window: 5 macro rows of 4 cameras
constexpr int kG = kWin / 4;
constexpr int kNMacro = kG * (kG + 1) / 2;   // 15 macro tiles of the lower triangle
constexpr int kYS = 6 * kWin;           // panel row stride in doubles: 120 = 8 (mod 16)
constexpr int kPts = kSW;               // points per panel (one per staging warp)
constexpr int kK = 3 * kPts;            // panel rows (36: nine mma k-steps)
constexpr int kPS = 13;                 // pose stride in shared memory (odd: conflict-free per-lane rows)
constexpr int kScratch = 32 * 19 + 16;  // scatter_point: sW[32][19] + sCam[32]
constexpr int kBarFull = 1, kBarEmpty = 3, kBarStage = 5;
constexpr int kCluster = 2;             // whole-window mode: CTAs of a cluster add their tiles through DSMEM before the REDs
constexpr int kNone = 0x7fffffff;

static_assert(kYS % 16 == 8, "panel stride must be 8 mod 16 doubles for conflict-free fragments");
static_assert(kNMacro <= 2 * kXW, "two macro tiles per mma warp");
static_assert(kK % 4 == 0, "whole k-steps");

struct PanelFlags {
  int rebase;          // a point did not fit the window
  int cmin;            // smallest camera among the points of the panel
};

struct __align__(16) MmaSmem {
  double Yt[2][kK * kYS];
  double Sacc[kNMacro][18][32];   // macro-tile accumulators in mma fragment order, private per thread
  double Uacc[kSW][27][kWin];     // per staging warp and camera lane: U lower (21) + b (6); private
  double scr[kScratch];           // scatter_point scratch (one warp at a time, see lock)
  double pose[32 * kPS];
  double red[kSW];
  // per staging warp: cp.async landing zones of the next point (no registers are spent on the prefetch)
  float2 puv[kSW][2][32];      // pixels of observations o0 + lane
  int pcam[kSW][2][32];        // their cameras
  double pX[kSW][2][4];        // the point
  int prange[kSW][3][2];       // pt_off[j], pt_off[j + 1] of the points n, n+1, n+2 of the warp (-1: no point)
  int slot[kSW][32];
  PanelFlags pf[3];            // panel n uses pf[n % 3]: written while staging n, read after it, cleared one panel later
  int pbase[2];                // window base the panel in buffer b was staged with
  unsigned pmask[2][kPts];     // cameras (window-relative bits) with a non-zero row block, per staged point
  int lock;
};

__device__ __forceinline__ int lower_bound_off(const int* __restrict__ off, int n, int target) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 1/sqrt(a): hardware approximation (MUFU.RSQ64H, ~2^-22) + two Newton steps
__device__ __forceinline__ double fast_rsqrt(double a) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a * r, r, 1.0);
    r = fma(0.5 * r, e, r);
  }
  return r;
}
template <int kId, int kN>
__device__ __forceinline__ void bar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(kId), "n"(kN) : "memory");
}
template <int kId, int kN>
__device__ __forceinline__ void bar_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(kId), "n"(kN) : "memory");
}

// the rare fallback stays out of line so that its registers do not count against the hot loops
__device__ __noinline__ void scatter_point_cold(const Work& w, int j, int lane, double lambda, double delta,
                                                const double* __restrict__ pose, const double* __restrict__ X,
                                                double* scratch, int* lock, double* cost) {
  if (lane == 0) {
    while (atomicCAS(lock, 0, 1) != 0) {}
  }
  __syncwarp();
  double (*sW)[19] = reinterpret_cast<double (*)[19]>(scratch);
  int* sCam = reinterpret_cast<int*>(scratch + 32 * 19);
  double c = 0.0;
  scatter_point(w, j, lane, lambda, delta, pose, X, sW, sCam, c);
  *cost += c;
  __syncwarp();
  if (lane == 0) { __threadfence_block(); atomicExch(lock, 0); }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// clock64 trace of one chosen thread of CTA 0 (compile with -DBA_MMA_TRACE, run with BA_DBG_TIMES=1)
__device__ __forceinline__ void mark(const Work& w, bool who, int slot, int tag) {
#ifdef BA_MMA_TRACE
  if (w.dbg && blockIdx.x == 0 && who && slot < 2000) {
    w.dbg[1 + 2 * slot] = tag;
    w.dbg[2 + 2 * slot] = clock64();
  }
#endif
}

__device__ __forceinline__ void tile_of(int t, int& I, int& J) {
  I = 0;
  while (t > I) { t -= I + 1; ++I; }
  J = t;
}

// one macro tile over one panel: accumulators come from and go back to their shared-memory slots
__device__ __forceinline__ void mma_macro_tile(const double* __restrict__ Y, int ksteps, int I, int J,
                                               double* __restrict__ acc /* [18][32] + lane */) {
  const bool diag = I == J;
  double c[9][2];
#pragma unroll
  for (int q = 0; q < 9; ++q) { c[q][0] = acc[(2 * q) * 32]; c[q][1] = acc[(2 * q + 1) * 32]; }
  const double* pa = Y + 24 * I;
  const double* pb = Y + 24 * J;
#pragma unroll 1
  for (int kk = 0; kk < ksteps; ++kk) {
    double a[3], b[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i] = pa[8 * i]; b[i] = pb[8 * i]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j <= i || !diag) dmma(c[3 * i + j][0], c[3 * i + j][1], a[i], b[j]);   // diagonal macro tile: lower part only
    pa += 4 * kYS;
    pb += 4 * kYS;
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) { acc[(2 * q) * 32] = c[q][0]; acc[(2 * q + 1) * 32] = c[q][1]; }
}

// macro tile -> float64 REDs; entries that are exactly zero belong to absent blocks and are skipped
__device__ __forceinline__ void flush_macro_tile(const Work& w, int base, int mI, int mJ, int lane,
                                                 double* __restrict__ acc) {
  // CTAs start at different entries so that they do not all hit the same addresses at the same time
  const int rot = blockIdx.x % 18;
#pragma unroll 1
  for (int e0 = 0; e0 < 18; ++e0) {
    const int e = (e0 + rot >= 18) ? e0 + rot - 18 : e0 + rot;
    const double v = acc[e * 32];
    if (v == 0.0) continue;
    acc[e * 32] = 0.0;
    const int t = e >> 1, h = e & 1;
    const int i = t / 3, j = t - 3 * i;
    const int m = 24 * mI + 8 * i + (lane >> 2);
    const int n2 = 24 * mJ + 8 * j + 2 * (lane & 3) + h;
    const int a = m / 6, r = m - 6 * a, b = n2 / 6, cc = n2 - 6 * b;
    if (a > b || (a == b && cc <= r)) atomicAdd(s_block_ptr(w, base + a, base + b) + 6 * r + cc, -v);
  }
}

// Whole-window mode, all 640 threads of every CTA of the cluster: the CTA pair's macro-tile accumulators
// (same window base 0) are summed through distributed shared memory, each CTA takes half of the
// entries and issues ONE float64 RED per entry for the whole cluster.
__device__ __forceinline__ void cluster_flush(const Work& w, MmaSmem& sm, int tid) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  cluster.sync();                                   // every CTA's accumulators are complete
  const unsigned rank = cluster.block_rank();
  constexpr int kTotal = kNMacro * 18 * 32;
  constexpr int kShare = (kTotal + kCluster - 1) / kCluster;
  const double* src[kCluster];
#pragma unroll
  for (int q = 0; q < kCluster; ++q) src[q] = cluster.map_shared_rank(&sm.Sacc[0][0][0], q);
  const int lo = rank * kShare, hi = min(kTotal, lo + kShare);
  // every cluster adds into the same entries of S: clusters start at different offsets of their share so that
  // the L2 does not see all of them on the same addresses at the same time (same-address float64 REDs serialise),
  // and a thread's remote / local loads are all in flight before its first RED
  constexpr int kIter = (kShare + kMT - 1) / kMT;
  const int n_clusters = (int)(gridDim.x / kCluster);
  const int rot = (int)(((long long)(blockIdx.x / kCluster) * kShare) / (n_clusters > 0 ? n_clusters : 1));
  double v[kIter];
  int at[kIter];
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    int off = tid + u * kMT;
    const bool ok = off < hi - lo;
    off += rot;
    if (off >= hi - lo) off -= hi - lo;
    at[u] = ok ? lo + off : -1;
    double a = 0.0;
    if (ok) {
#pragma unroll
      for (int q = 0; q < kCluster; ++q) a += src[q][lo + off];
    }
    v[u] = a;
  }
#pragma unroll
  for (int u = 0; u < kIter; ++u) {
    if (at[u] < 0 || v[u] == 0.0) continue;
    const int idx = at[u];
    const int t = idx / 576, e = (idx - 576 * t) >> 5, lane = idx & 31;
    int mI, mJ;
    tile_of(t, mI, mJ);
    const int tt = e >> 1, h = e & 1;
    const int i = tt / 3, j = tt - 3 * i;
    const int m = 24 * mI + 8 * i + (lane >> 2);
    const int n2 = 24 * mJ + 8 * j + 2 * (lane & 3) + h;
    const int a = m / 6, r = m - 6 * a, b = n2 / 6, cc = n2 - 6 * b;
    if (a > b || (a == b && cc <= r)) atomicAdd(s_block_ptr(w, a, b) + 6 * r + cc, -v[u]);
  }
  cluster.sync();                                   // nobody leaves while its shared memory is being read
}

}  // namespace

template <bool kSliding>
__global__ void __launch_bounds__(kMT, 1)
k_linearize_mma(const __grid_constant__ Work w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MmaSmem& sm = *reinterpret_cast<MmaSmem*>(smem_raw);
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
  int p_lo, p_hi;
  if (w.part != nullptr && w.part_n == (int)gridDim.x) {
    p_lo = __ldg(w.part + blockIdx.x);
    p_hi = __ldg(w.part + blockIdx.x + 1);
  } else {
    p_lo = (blockIdx.x == 0) ? 0 : lower_bound_off(w.pt_off, w.n_pt, t_lo);
    p_hi = (blockIdx.x == gridDim.x - 1) ? w.n_pt : lower_bound_off(w.pt_off, w.n_pt, t_hi);
  }
  const int n_panels = (p_hi - p_lo + kPts - 1) / kPts;

  if (wid >= kSW) {
    // =============================== MMA warps ===============================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int mw = wid - kSW;
    int I0, J0, I1, J1;
    tile_of(mw, I0, J0);
    const bool has1 = mw + kXW < kNMacro;
    tile_of(has1 ? mw + kXW : mw, I1, J1);
    double* acc0 = &sm.Sacc[mw][0][lane];
    double* acc1 = &sm.Sacc[has1 ? mw + kXW : mw][0][lane];
#pragma unroll
    for (int e = 0; e < 18; ++e) { acc0[e * 32] = 0.0; if (has1) acc1[e * 32] = 0.0; }
    const int foff = (lane & 3) * kYS + (lane >> 2);
    int acc_base = -1;
    const bool me = tid == kST;
    mark(w, me, 200, 400);
    for (int n = 0; n < n_panels; ++n) {
      const int b = n & 1;
      if (b == 0) bar_sync<kBarFull, kMT>(); else bar_sync<kBarFull + 1, kMT>();
      if (n < 8) mark(w, me, 201 + 2 * n, 410 + n);
      const int base_n = sm.pbase[b];
      if (base_n != acc_base) {
        if (acc_base >= 0) {
          flush_macro_tile(w, acc_base, I0, J0, lane, acc0);
          if (has1) flush_macro_tile(w, acc_base, I1, J1, lane, acc1);
        }
        acc_base = base_n;
      }
      const int npts = min(kPts, p_hi - (p_lo + n * kPts));
      const int ksteps = (3 * npts + 3) >> 2;
      bool do0 = true, do1 = has1;
      if (kSliding) {
        unsigned m = (lane < kPts) ? sm.pmask[b][lane] : 0u;
        m = __reduce_or_sync(kFull, m);
        unsigned rows = 0;                       // macro rows with at least one camera present
#pragma unroll
        for (int g = 0; g < kG; ++g) rows |= ((m >> (4 * g)) & 0xfu) ? (1u << g) : 0u;
        do0 = ((rows >> I0) & 1u) && ((rows >> J0) & 1u);
        do1 = has1 && ((rows >> I1) & 1u) && ((rows >> J1) & 1u);
      }
      const double* Y = sm.Yt[b] + foff;
      if (do0) mma_macro_tile(Y, ksteps, I0, J0, acc0);
      if (do1) mma_macro_tile(Y, ksteps, I1, J1, acc1);
      if (n + 2 < n_panels) {
        if (b == 0) bar_arrive<kBarEmpty, kMT>(); else bar_arrive<kBarEmpty + 1, kMT>();
      }
      if (n < 8) mark(w, me, 202 + 2 * n, 420 + n);
    }
    mark(w, me, 230, 430);
    if (kSliding) {
      if (acc_base >= 0) {
        const bool swap = has1 && ((blockIdx.x / 18) & 1);
        if (swap) flush_macro_tile(w, acc_base, I1, J1, lane, acc1);
        flush_macro_tile(w, acc_base, I0, J0, lane, acc0);
        if (has1 && !swap) flush_macro_tile(w, acc_base, I1, J1, lane, acc1);
      }
      mark(w, me, 231, 431);
      return;
    }
    cluster_flush(w, sm, tid);
    mark(w, me, 231, 431);
    return;
  }

  // ================================= staging warps ===========================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
  mark(w, tid == 0, 30, 300);
  if (tid < 3) { sm.pf[tid].rebase = 0; sm.pf[tid].cmin = kNone; }
  if (tid == 0) sm.lock = 0;
  sm.slot[wid][lane] = -1;
  int base = 0;
  if (kSliding) {
    // window start: the smallest camera of the first panel's observations
    bar_sync<kBarStage, kST>();
    if (p_lo < p_hi) {
      const int q_hi = min(p_hi, p_lo + kPts);
      const int o_a = __ldg(w.pt_off + p_lo), o_b = __ldg(w.pt_off + q_hi);
      int c = kNone;
      for (int o = o_a + tid; o < o_b && o < o_a + 8 * kST; o += kST) c = min(c, __ldg(w.obs_cam + o));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(kFull, c, o));
      if (lane == 0 && c != kNone) atomicMin(&sm.pf[2].cmin, c);
    }
    bar_sync<kBarStage, kST>();
    base = (sm.pf[2].cmin == kNone) ? 0 : sm.pf[2].cmin;
    bar_sync<kBarStage, kST>();
    if (tid == 0) sm.pf[2].cmin = kNone;
  }

  double* __restrict__ Ul = &sm.Uacc[wid][0][lane < kWin ? lane : 0];   // this lane's column of the private sums
  if (lane < kWin) {
#pragma unroll
    for (int q = 0; q < 27; ++q) Ul[q * kWin] = 0.0;
  }
  double cost = 0.0;
  bool cam_free = false;

  auto load_window = [&]() {
    // poses of the window into shared memory, one padded row per camera (staging group only)
    for (int i = tid; i < 12 * kWin; i += kST) {
      const int a = i / 12, q = i - 12 * a;
      const int cam = base + a;
      sm.pose[a * kPS + q] = (cam < w.n_cam) ? pose[12 * (size_t)cam + q] : 0.0;
    }
    const int cam = base + lane;
    cam_free = (lane < kWin) && (cam < w.n_cam) && (w.fixed[cam] == 0);
    bar_sync<kBarStage, kST>();
  };

  auto flush_cams = [&]() {
    // per-camera sums: across the staging warps, then one RED per entry
    bar_sync<kBarStage, kST>();
    for (int t0 = tid; t0 < 27 * kWin; t0 += kST) {
      const int t = (t0 + 67 * (int)blockIdx.x) % (27 * kWin);      // CTAs start at different entries
      const int q = t / kWin, a = t - q * kWin;
      const int cam = base + a;
      double s = 0.0;
#pragma unroll
      for (int ww = 0; ww < kSW; ++ww) { s += sm.Uacc[ww][q][a]; sm.Uacc[ww][q][a] = 0.0; }
      if (cam >= w.n_cam || s == 0.0) continue;
      if (q < 21) {
        int r = 0, e = q;
        while (e > r) { e -= r + 1; ++r; }
        atomicAdd(s_block_ptr(w, cam, cam) + 6 * r + e, s);
        if (e == r) atomicAdd(w.diagU + 6 * (size_t)cam + r, s);
      } else {
        atomicAdd(w.b + 6 * (size_t)cam + (q - 21), s);
      }
    }
    bar_sync<kBarStage, kST>();
  };

  load_window();
  mark(w, tid == 0, 31, 301);

  // software pipeline over this warp's points j = p_lo + wid, + kSW, ...: the observation range of the
  // point after next and the cameras / pixels / coordinates of the next point travel with cp.async
  // straight into shared memory while the current point is linearised
  auto issue_range = [&](int idx) {              // idx-th point of this warp -> prange[idx % 3]
    const int jj = p_lo + wid + idx * kSW;
    int* dst = sm.prange[wid][idx % 3];
    if (lane < 2) {
      if (jj < p_hi) cp_async4(dst + lane, w.pt_off + jj + lane);
      else dst[lane] = -1;
    }
  };
  auto issue_point = [&](int idx) {              // needs prange[idx % 3] to have landed
    const int jj = p_lo + wid + idx * kSW;
    const int o0 = sm.prange[wid][idx % 3][0], o1 = sm.prange[wid][idx % 3][1];
    if (o0 < 0) return;
    const int k = o1 - o0;
    if (lane < k) {
      cp_async4(&sm.pcam[wid][idx & 1][lane], w.obs_cam + o0 + lane);
      cp_async8(&sm.puv[wid][idx & 1][lane], reinterpret_cast<const float2*>(w.obs_uv) + o0 + lane);
    }
    if (lane < 3) cp_async8(&sm.pX[wid][idx & 1][lane], X + 3 * (size_t)jj + lane);
  };
  issue_range(0);
  issue_range(1);
  cp_async_commit();
  cp_async_wait_all();
  __syncwarp();
  issue_point(0);
  cp_async_commit();
  const double* __restrict__ P = sm.pose + lane * kPS;

#pragma unroll 1
  for (int n = 0; n < n_panels; ++n) {
    const int b = n & 1;
    const int fs = n % 3;
    if (n >= 2) { if (b == 0) bar_sync<kBarEmpty, kMT>(); else bar_sync<kBarEmpty + 1, kMT>(); }
    if (n == 3) mark(w, tid == 0, 70, 340);
    double* __restrict__ Y = sm.Yt[b];
    const int slot = wid;
    const int j = p_lo + n * kPts + slot;
    // ---- this point's data has landed; send for the next point and the range after it ------------
    cp_async_wait_all();
    __syncwarp();
    issue_point(n + 1);
    issue_range(n + 2);
    cp_async_commit();
    if (n == 3) mark(w, tid == 0, 71, 341);
    const int pb = n & 1;
    const int in_o0 = sm.prange[wid][n % 3][0];
    const int in_k = (in_o0 < 0) ? -1 : sm.prange[wid][n % 3][1] - in_o0;
    const int in_cam = (in_k > 0 && lane < min(in_k, 32)) ? sm.pcam[wid][pb][lane] : kNone;

    unsigned present = 0u;
    bool valid = in_k >= 0;
    int ia = -1;
    float2 uv = make_float2(0.f, 0.f);
    if (valid) {
      const int k = in_k;
      bool ok = k <= 32;
      int rel = -1;
      if (ok && lane < k) {
        rel = in_cam - base;
        if (rel < 0 || rel >= kWin) ok = false;
      }
      ok = __all_sync(kFull, ok);
      // which observation does camera lane `lane` serve?  Observations in strictly increasing keyframe order
      // (the packer's and the generator's order) need no table: it is the number of observed keyframes below.
      const int prev = __shfl_up_sync(kFull, rel, 1);
      const bool rising = __all_sync(kFull, lane == 0 || lane >= k || rel > prev);
      if (ok && rising) {
        const unsigned seen_mask = __reduce_or_sync(kFull, (lane < k) ? (1u << rel) : 0u);
        ia = ((seen_mask >> lane) & 1u) ? __popc(seen_mask & ((1u << lane) - 1u)) : -1;
        if (ia >= 0) uv = sm.puv[wid][pb][ia];
      } else if (ok) {
        // any other order: through a slot table, which also finds repeated (point, keyframe) observations
        __syncwarp();
        if (lane < k) sm.slot[wid][rel] = lane;
        __syncwarp();
        const bool dup = lane < k && sm.slot[wid][rel] != lane;
        ia = sm.slot[wid][lane];
        if (ia >= 0) uv = sm.puv[wid][pb][ia];
        ok = !__any_sync(kFull, dup);
        __syncwarp();
        sm.slot[wid][lane] = -1;
      }
      if (kSliding && k > 0 && k <= 32) {
        int c = in_cam;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(kFull, c, o));
        if (lane == 0) { atomicMin(&sm.pf[fs].cmin, c); if (!ok) sm.pf[fs].rebase = 1; }
      }
      if (!ok) {
        // out of the window / repeated / oversized: the scatter routine takes the point
        scatter_point_cold(w, j, lane, lambda, delta, pose, X, sm.scr, &sm.lock, &cost);
        ia = -1;
        valid = false;
      }
    }
    const bool seen = valid && ia >= 0;
    if (n == 3) mark(w, tid == 0, 72, 342);
    const bool act = seen && cam_free;
    double* dst = Y + (size_t)(3 * slot) * kYS + 6 * (lane < kWin ? lane : 0);
    if (valid) {
      // ---- phase 1: projection, weights, point block contributions ---------------------------
      double Y0 = 0, Y1 = 0, Y2 = 0, ja = 0, jb = 0, jc = 0, jd = 0, ru = 0, rv = 0, wu = 0, wv = 0;
      double V[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0};
      if (seen) {
        const double R0 = P[0], R1 = P[1], R2 = P[2], R3 = P[3], R4 = P[4], R5 = P[5], R6 = P[6], R7 = P[7], R8 = P[8];
        const double X0 = sm.pX[wid][pb][0], X1 = sm.pX[wid][pb][1], X2 = sm.pX[wid][pb][2];
        Y0 = R0 * X0 + R1 * X1 + R2 * X2;
        Y1 = R3 * X0 + R4 * X1 + R5 * X2;
        Y2 = R6 * X0 + R7 * X1 + R8 * X2;
        const double xc = Y0 + P[9], yc = Y1 + P[10], zc = Y2 + P[11];
        const double iz = fast_rcp(zc + kDepthEps);
        const double xz = xc * iz, yz = yc * iz;
        ru = w.K.fx * xz + w.K.cx - (double)uv.x;
        rv = w.K.fy * yz + w.K.cy - (double)uv.y;
        wu = huber_w(ru, delta);
        wv = huber_w(rv, delta);
        ja = w.K.fx * iz; jc = -w.K.fx * xz * iz;
        jb = w.K.fy * iz; jd = -w.K.fy * yz * iz;
        const double lu0 = ja * R0 + jc * R6, lu1 = ja * R1 + jc * R7, lu2 = ja * R2 + jc * R8;
        const double lv0 = jb * R3 + jd * R6, lv1 = jb * R4 + jd * R7, lv2 = jb * R5 + jd * R8;
        const double ul0 = wu * lu0, ul1 = wu * lu1, ul2 = wu * lu2;
        const double vl0 = wv * lv0, vl1 = wv * lv1, vl2 = wv * lv2;
        V[0] = ul0 * lu0 + vl0 * lv0;
        V[1] = ul0 * lu1 + vl0 * lv1;
        V[2] = ul0 * lu2 + vl0 * lv2;
        V[3] = ul1 * lu1 + vl1 * lv1;
        V[4] = ul1 * lu2 + vl1 * lv2;
        V[5] = ul2 * lu2 + vl2 * lv2;
        gl[0] = ul0 * ru + vl0 * rv;
        gl[1] = ul1 * ru + vl1 * rv;
        gl[2] = ul2 * ru + vl2 * rv;
        cost += huber_rho(ru, delta) + huber_rho(rv, delta);
      }
      warp_sum_6_3(V, gl);     // butterfly over the lanes, the nine sums side by side
      if (n == 3) mark(w, tid == 0, 73, 343);
      present = __ballot_sync(kFull, act);
      if (act) {
        // ---- phase 2: C_j with C C^T = (V + lambda diag V)^-1, z = C^T g_l ---------------------
        V[0] += damp(V[0], lambda);
        V[3] += damp(V[3], lambda);
        V[5] += damp(V[5], lambda);
        // LDL^T of V; V^-1 = M^T D^-1 M with M = L^-1 (unit lower) -> C = M^T D^-1/2 (upper);
        // pivots through 1/sqrt (hardware seed + two Newton steps, no slow-path call): s = d^-1/2, 1/d = s^2
        const double r0 = fast_rsqrt(V[0]);
        const double i0 = r0 * r0;
        const double l10 = V[1] * i0, l20 = V[2] * i0;
        const double d1 = V[3] - l10 * V[1];
        const double e1 = V[4] - l20 * V[1];
        const double r1 = fast_rsqrt(d1);
        const double i1 = r1 * r1;
        const double l21 = e1 * i1;
        const double d2 = V[5] - l20 * V[2] - l21 * e1;
        const double r2 = fast_rsqrt(d2);
        const bool pd = (V[0] > 0.0) && (d1 > 0.0) && (d2 > 0.0) && isfinite(r0) && isfinite(r1) && isfinite(r2);
        const double c00 = pd ? r0 : 0.0, c11 = pd ? r1 : 0.0, c22 = pd ? r2 : 0.0;
        const double c01 = -l10 * c11, c02 = (l10 * l21 - l20) * c22, c12 = -l21 * c22;
        const double z0 = c00 * gl[0];
        const double z1 = c01 * gl[0] + c11 * gl[1];
        const double z2 = c02 * gl[0] + c12 * gl[1] + c22 * gl[2];
        if (n == 3) mark(w, tid == 0, 74, 344);
        // ---- phase 3: pose rows, Y = W C = (w_u p_u) (l_u C) + (w_v p_v) (l_v C), camera sums ----------
        // (J_l C is formed once per observation: 2 FMAs per entry of Y instead of 5, and neither J_l nor C
        //  stays live through the row loop)
        double luC0, luC1, luC2, lvC0, lvC1, lvC2;
        {
          const double lu0 = ja * P[0] + jc * P[6], lu1 = ja * P[1] + jc * P[7], lu2 = ja * P[2] + jc * P[8];
          const double lv0 = jb * P[3] + jd * P[6], lv1 = jb * P[4] + jd * P[7], lv2 = jb * P[5] + jd * P[8];
          luC0 = lu0 * c00; luC1 = lu0 * c01 + lu1 * c11; luC2 = lu0 * c02 + lu1 * c12 + lu2 * c22;
          lvC0 = lv0 * c00; lvC1 = lv0 * c01 + lv1 * c11; lvC2 = lv0 * c02 + lv1 * c12 + lv2 * c22;
        }
        // b_a: sum_r  p_u[r] (w_u (l_u C).z - w_u r_u) + p_v[r] (w_v (l_v C).z - w_v r_v)
        const double gu = wu * (luC0 * z0 + luC1 * z1 + luC2 * z2 - ru);
        const double gv = wv * (lvC0 * z0 + lvC1 * z1 + lvC2 * z2 - rv);
        double pu[6], pv[6];
        pu[0] = jc * Y1;             pu[1] = ja * Y2 - jc * Y0;  pu[2] = -ja * Y1;
        pu[3] = ja;                  pu[4] = 0.0;                pu[5] = jc;
        pv[0] = -jb * Y2 + jd * Y1;  pv[1] = -jd * Y0;           pv[2] = jb * Y0;
        pv[3] = 0.0;                 pv[4] = jb;                 pv[5] = jd;
        int e2 = 0;
        double yp[3];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double pa = wu * pu[r], pb = wv * pv[r];
          const double y0 = pa * luC0 + pb * lvC0;
          const double y1 = pa * luC1 + pb * lvC1;
          const double y2 = pa * luC2 + pb * lvC2;
          if (r & 1) {
            // panel row d, columns 6 a + r - 1, 6 a + r
            *reinterpret_cast<double2*>(dst + r - 1) = make_double2(yp[0], y0);
            *reinterpret_cast<double2*>(dst + kYS + r - 1) = make_double2(yp[1], y1);
            *reinterpret_cast<double2*>(dst + 2 * kYS + r - 1) = make_double2(yp[2], y2);
          } else {
            yp[0] = y0; yp[1] = y1; yp[2] = y2;
          }
#pragma unroll
          for (int c2 = 0; c2 <= r; ++c2, ++e2) Ul[e2 * kWin] += pa * pu[c2] + pb * pv[c2];
          Ul[(21 + r) * kWin] += pu[r] * gu + pv[r] * gv;
        }
      }
    }
    if (!act && lane < kWin) {
      const double2 zz = make_double2(0.0, 0.0);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double2* p2 = reinterpret_cast<double2*>(dst + d * kYS);
        p2[0] = zz; p2[1] = zz; p2[2] = zz;
      }
    }
    if (kSliding && lane == 0) sm.pmask[b][slot] = present;
    if (n == 3) mark(w, tid == 0, 75, 345);

    if (wid == 0 && lane == 0) sm.pbase[b] = base;
    bool rebase = false;
    int cmin = kNone;
    if (kSliding) {
      bar_sync<kBarStage, kST>();
      cmin = sm.pf[fs].cmin;
      rebase = sm.pf[fs].rebase != 0 && cmin != kNone && cmin != base;
      const int fp = (fs == 0) ? 2 : fs - 1;            // flags of the previous panel: everyone has read them
      if (tid == 0) { sm.pf[fp].rebase = 0; sm.pf[fp].cmin = kNone; }
    }
    __threadfence_block();
    if (b == 0) bar_arrive<kBarFull, kMT>(); else bar_arrive<kBarFull + 1, kMT>();
    if (n < 8) mark(w, tid == 0, 32 + n, 310 + n);
    if (rebase) {
      // the window has to move: flush this group's sums and restart at the panel's first camera;
      // the mma warps flush their tiles when they see the new base on the next panel
      flush_cams();
      base = cmin;
      load_window();
    }
  }
  mark(w, tid == 0, 60, 330);
  flush_cams();
  mark(w, tid == 0, 61, 331);

  cost = warp_sum(cost);
  if (lane == 0) sm.red[wid] = cost;
  bar_sync<kBarStage, kST>();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kSW; ++i) s += sm.red[i];
    if (s != 0.0) atomicAdd(w.cost_lin, s);
  }
  if (!kSliding) cluster_flush(w, sm, tid);
}

int configure_mma() {
  const int bytes = (int)sizeof(MmaSmem);
  cudaError_t e = cudaFuncSetAttribute(k_linearize_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linearize_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

int mma_window() { return kWin; }

// how many clusters of the whole-window kernel can be resident at once (diagnostics)
int mma_max_active_clusters(int cluster) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148 / cluster * cluster);
  cfg.blockDim = dim3(kMT);
  cfg.dynamicSmemBytes = sizeof(MmaSmem);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, k_linearize_mma<false>, &cfg) != cudaSuccess) return -1;
  return n;
}

int mma_grid(int n_cam, int n_pt, int sm_count) {
  int blocks = sm_count;
  const int by_pts = (n_pt + kPts - 1) / kPts;
  if (blocks > by_pts) blocks = by_pts;
  if (blocks < 1) blocks = 1;
  if (n_cam <= kWin) blocks = (blocks + kCluster - 1) / kCluster * kCluster;   // CTAs without points only join the flush
  return blocks;
}

int launch_linearize_mma(const Work& w, int sm_count, cudaStream_t s) {
  const int blocks = mma_grid(w.n_cam, w.n_pt, sm_count);
  if (w.n_cam <= kWin) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(kMT);
    cfg.dynamicSmemBytes = sizeof(MmaSmem);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_linearize_mma<false>, w) != cudaSuccess) return -1;
  } else {
    k_linearize_mma<true><<<blocks, kMT, sizeof(MmaSmem), s>>>(w);
  }
  return 1;
}

}  // namespace ba
