// Linearise + Schur, "panel" path (round 2): three kernels per LM iteration instead of one monolith whose
// per-point dependency chain (projection -> butterfly -> 3x3 factor -> pose rows -> private sums) kept one
// point per warp in flight.
//
//   k_point_blocks   point-major, kG lanes per point.  V_j = sum J_l^T w J_l, g_l, the reference cost; then the
//                    3x3 factor C_j (C C^T = (V + lambda diag V)^-1) and q_j = V^-1 g_l.  One 80-byte record per
//                    point: C (6) | q (3) | cmin, cmax | flags.  Nothing per observation is written.
//   k_cam_blocks     CAMERA-major over a camera-sorted copy of the observations (built once per solve by
//                    k_cs_*): every lane of a warp serves the same camera, so U_a = sum J_p^T w J_p and
//                    b_a = sum J_p^T w (J_l q_j - r) accumulate in registers over 256 observations and leave
//                    through one butterfly + 27 float64 REDs.  All 32 lanes carry an observation.
//   k_schur_panel    point-major, warp specialised.  STAGING warps (2 groups x 4): lane <-> observation.
//                    The observations / points / records of a panel arrive by 1-D bulk copies
//                    (cp.async.bulk + mbarrier, double buffered); a lane projects its observation, forms
//                    Y_aj = W_aj C_j (6 x 3) and stores it into the panel  Yt[4 p + d][6 a + r]  (4th row of a
//                    point = zero pad, so that one mma k-step = one point).  No cross-lane reduction, no
//                    private sums, no slot tables: the chain per lane is projection -> Y.
//                    MMA warps (8): S -= Yt^T Yt with mma.sync.m8n8k4.f64 per 8 x 8 tile.  The tiles of the lower
//                    triangle of the 20-keyframe window are owned in strips (tile row x 16 columns, see the mma
//                    warps below), ALL accumulators stay in registers for the life of a window, and a k-step only
//                    touches the tiles inside the point's camera interval [cmin, cmax]: a point seen by 8
//                    keyframes costs ~30 DMMAs instead of the 120 of the dense window.
//                    (Tried and dropped: leaving the zero pads of a panel buffer in place when it is refilled with
//                    the camera interval it held before -- correct, 0.957 -> 0.950 ms: the staging warps are bound
//                    by their instruction stream per observation, not by those stores.)
//   Points whose cameras do not fit the window, or that see a keyframe twice / out of order, take a per-lane
//   fallback (every observation pair of the point -> float64 REDs); the window then moves to the panel's first
//   camera (accumulators flushed by REDs, entries that are exactly zero skipped).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "ba_kernels.cuh"
#include "ba_scatter.cuh"

namespace ba {

namespace {

constexpr int kWin = 20;                  // keyframes per window
constexpr int kCols = 6 * kWin;            // 120 window columns
constexpr int kPCols = 128;               // panel columns: 16 tiles of 8 (columns 120..127 stay zero)
constexpr int kYS = kPCols + 4;            // panel row stride in doubles: 132 = 4 (mod 16).  A fragment load reads
                                          // (lane & 3) * kYS + (lane >> 2): 4 a + b, a, b = 0..3, are sixteen different
                                          // 8-byte banks.  (120 = 8 mod 16, the round-1 stride, folds a = 0, 2 and 1, 3
                                          // onto each other: ncu counted 2 wavefronts per half-warp on every fragment load.)
constexpr int kPP = 12;                   // points per panel (one mma k-step each)
constexpr int kRows = 4 * kPP;            // panel rows
constexpr int kGW = 4;                    // warps per staging group
constexpr int kGL = 32 * kGW;             // lanes per staging group
constexpr int kNG = 2;                    // staging groups (take the CTA's panels in turn)
constexpr int kNB = 3;                    // panel buffers (panel n lives in buffer n mod 3): a group starts its next panel while
                                          // the mma warps still read its last one
constexpr int kXW = 8;                    // mma warps
constexpr int kThreadsP = 32 * (kNG * kGW + kXW);   // 512
constexpr int kOB = 512;                  // observation capacity of a panel's staging buffers
constexpr int kBoundsCap = 2048;          // panel boundaries staged per super-chunk
constexpr int kBarFull = 1, kBarEmpty = 1 + kNB, kBarGroup = 1 + 2 * kNB, kBarStageAll = 1 + 2 * kNB + kNG;
static_assert(kBarStageAll < 16, "named barriers");
constexpr int kRecD = 10;                 // doubles per point record
constexpr int kFlagFallback = 1, kFlagSkip = 2;
constexpr int kTrace0 = 200;              // first of the eight panels of CTA 0 the clock64 trace covers (-DBA_PANEL_TRACE)

static_assert(kYS % 16 == 12 || kYS % 16 == 4, "panel stride must be 4 or 12 mod 16 doubles for conflict-free fragment loads");
static_assert(kYS % 2 == 0, "rows must stay 16-byte aligned");
static_assert(kPP % 4 == 0, "CTA point ranges start at multiples of 4 so that bulk copies stay 16-byte aligned");

struct Rec {               // 80 bytes, 16-byte aligned
  double c[6];             // C upper triangle: c00 c01 c02 c11 c12 c22
  double q[3];             // V^-1 g_l
  int cmin;                // smallest camera of the point (INT_MAX: no observation)
  int cmaxf;               // largest camera | flags << 28
};
static_assert(sizeof(Rec) == 8 * kRecD, "record layout");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
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
template <int kN>
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kN) : "memory"); }
template <int kN>
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kN) : "memory"); }

// ---- mbarrier + 1-D bulk copy (TMA without a tensor map) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// k_point_blocks
// ------------------------------------------------------------------------------------------------------------
template <int kG>
__global__ void __launch_bounds__(256, 3)
k_point_blocks(const __grid_constant__ Work w) {
  __shared__ double s_cost[8];
  constexpr int kPer = 32 / kG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gl = lane & (kG - 1), sub = lane / kG;
  const LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double lambda = ctl->lambda, delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  Rec* __restrict__ rec = reinterpret_cast<Rec*>(w.rec);
  double cost = 0.0;
  const int first = (blockIdx.x * 8 + wid) * kPer;
  const int stride = gridDim.x * 8 * kPer;
  for (int jb = first; jb < w.n_pt; jb += stride) {
    const int j = jb + sub;
    int o0 = 0, k = -1;
    if (j < w.n_pt) { o0 = __ldg(w.pt_off + j); k = __ldg(w.pt_off + j + 1) - o0; }
    double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    int cmin = 0x7fffffff, cmax = -1, bad = 0;
    if (k > 0) {
      const double X0 = X[3 * (size_t)j], X1 = X[3 * (size_t)j + 1], X2 = X[3 * (size_t)j + 2];
      for (int i = gl; i < k; i += kG) {
        const ObsIn ob = load_obs(w, o0 + i);
        if (i > 0 && __ldg(w.obs_cam + o0 + i - 1) >= ob.cam) bad = 1;      // not strictly rising: repeated / unordered keyframes
        cmin = min(cmin, ob.cam);
        cmax = max(cmax, ob.cam);
        const double* P = pose + 12 * (size_t)ob.cam;
        ObsLin ol;
        linearize_obs<false>(w.K, P, P + 9, X0, X1, X2, ob.u, ob.v, delta, ol);
        const double ul0 = ol.wu * ol.lu[0], ul1 = ol.wu * ol.lu[1], ul2 = ol.wu * ol.lu[2];
        const double vl0 = ol.wv * ol.lv[0], vl1 = ol.wv * ol.lv[1], vl2 = ol.wv * ol.lv[2];
        V[0] += ul0 * ol.lu[0] + vl0 * ol.lv[0];
        V[1] += ul0 * ol.lu[1] + vl0 * ol.lv[1];
        V[2] += ul0 * ol.lu[2] + vl0 * ol.lv[2];
        V[3] += ul1 * ol.lu[1] + vl1 * ol.lv[1];
        V[4] += ul1 * ol.lu[2] + vl1 * ol.lv[2];
        V[5] += ul2 * ol.lu[2] + vl2 * ol.lv[2];
        g[0] += ul0 * ol.ru + vl0 * ol.rv;
        g[1] += ul1 * ol.ru + vl1 * ol.rv;
        g[2] += ul2 * ol.ru + vl2 * ol.rv;
        cost += huber_rho(ol.ru, delta) + huber_rho(ol.rv, delta);
      }
    }
    // sums over the kG lanes of the point (the butterflies are warp-wide instructions: every lane takes part)
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
      cmin = min(cmin, __shfl_xor_sync(kFull, cmin, o));
      cmax = max(cmax, __shfl_xor_sync(kFull, cmax, o));
      bad |= __shfl_xor_sync(kFull, bad, o);
    }
    if (gl == 0 && j < w.n_pt) {
      Rec r;
      int flags = bad ? kFlagFallback : 0;
      if (k > 0) {
        V[0] += damp(V[0], lambda);
        V[3] += damp(V[3], lambda);
        V[5] += damp(V[5], lambda);
        // LDL^T of V; V^-1 = M^T D^-1 M, M = L^-1 unit lower -> C = M^T D^-1/2 (upper), C C^T = V^-1
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
        const double z0 = c00 * g[0];
        const double z1 = c01 * g[0] + c11 * g[1];
        const double z2 = c02 * g[0] + c12 * g[1] + c22 * g[2];
        r.c[0] = c00; r.c[1] = pd ? c01 : 0.0; r.c[2] = pd ? c02 : 0.0; r.c[3] = c11; r.c[4] = pd ? c12 : 0.0; r.c[5] = c22;
        r.q[0] = pd ? (c00 * z0 + c01 * z1 + c02 * z2) : 0.0;
        r.q[1] = pd ? (c11 * z1 + c12 * z2) : 0.0;
        r.q[2] = pd ? (c22 * z2) : 0.0;
        if (!pd) flags |= kFlagSkip;
      } else {
#pragma unroll
        for (int q = 0; q < 6; ++q) r.c[q] = 0.0;
        r.q[0] = r.q[1] = r.q[2] = 0.0;
        flags |= kFlagSkip;
        cmax = 0;
      }
      r.cmin = cmin;
      r.cmaxf = cmax | (flags << 28);
      rec[j] = r;
    }
  }
  cost = warp_sum(cost);
  if (lane == 0) s_cost[wid] = cost;
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += s_cost[i];
    if (c != 0.0) atomicAdd(w.cost_lin, c);
  }
}

// ------------------------------------------------------------------------------------------------------------
// camera-sorted copy of the observations (once per solve): cameras start at multiples of 32 entries, pads have
// point -1, so that every 32-entry block belongs to one camera
// ------------------------------------------------------------------------------------------------------------
__global__ void k_cs_count(Work w) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int o = tid; o < w.n_obs; o += nth) {
    const int c = __ldg(w.obs_cam + o);
    // lanes of a warp with the same camera share one atomic
    const unsigned peers = __match_any_sync(__activemask(), c);
    if ((int)(__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(w.cs_cnt + c, __popc(peers));
  }
}
// one block: block offsets per camera (exclusive scan of ceil(count / 32)), block -> camera table, cursors reset
__global__ void k_cs_scan(Work w) {
  __shared__ int s_part[1024];
  __shared__ int s_carry;
  const int tid = threadIdx.x;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < w.n_cam; base += 1024) {
    const int a = base + tid;
    const int nb = (a < w.n_cam) ? (w.cs_cnt[a] + 31) >> 5 : 0;
    s_part[tid] = nb;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                 // Hillis-Steele inclusive scan
      const int v = (tid >= o) ? s_part[tid - o] : 0;
      __syncthreads();
      s_part[tid] += v;
      __syncthreads();
    }
    const int excl = s_carry + s_part[tid] - nb;
    if (a < w.n_cam) {
      w.cs_blk_off[a] = excl;
      w.cs_cnt[a] = 0;                                   // becomes the fill cursor
    }
    __syncthreads();
    if (tid == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
  if (tid == 0) { w.cs_blk_off[w.n_cam] = s_carry; *w.cs_nblk = s_carry; }
}
__global__ void k_cs_fill(Work w) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // block -> camera table: a binary search over the cameras' first blocks per block (the scan kernel used to write it,
  // one thread per camera and 250 entries each at C4: 133 us on one SM)
  const int nblk = __ldg(w.cs_blk_off + w.n_cam);
  for (int b = tid; b < nblk; b += nth) {
    int lo = 0, hi = w.n_cam - 1;                        // last camera whose first block is <= b
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(w.cs_blk_off + mid) <= b) lo = mid; else hi = mid - 1;
    }
    w.cs_blk_cam[b] = lo;
  }
  for (int o = tid; o < w.n_obs; o += nth) {
    const int c = __ldg(w.obs_cam + o);
    const unsigned peers = __match_any_sync(__activemask(), c);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(w.cs_cnt + c, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int slot = 32 * __ldg(w.cs_blk_off + c) + base + __popc(peers & ((1u << lane) - 1u));
    w.cs_pt[slot] = __ldg(w.obs_pt + o);
    w.cs_uv[slot] = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
  }
}

// ------------------------------------------------------------------------------------------------------------
// class order (once per solve): points sorted by (first camera, camera span), a counting sort over n_cam x 32 classes.
// A panel of 12 consecutive points then holds one or two classes, i.e. one or two camera intervals, and every strip of
// the mma warps sees its live points as one run (see strip_runs).  In the caller's order (points sorted by an anchor
// camera, 2..16 keyframes each) the live points of a strip were a scattered half of the panel.
// Counters are 64-bit (points << 32 | observations) so that one atomic hands out a consistent (position, first
// observation) pair; the order inside a class is whatever the atomics make it, the observations of a point keep theirs.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSpanClasses = 32;
__device__ __forceinline__ int point_class(const Work& w, int j, int& k, int& cmin, int& cmax) {
  const int o0 = __ldg(w.pt_off + j);
  k = __ldg(w.pt_off + j + 1) - o0;
  cmin = 0x7fffffff; cmax = -1;
  if (k <= 0) return w.n_cam * kSpanClasses;            // no observation: behind everything
  for (int i = 0; i < k; ++i) {
    const int c = __ldg(w.obs_cam + o0 + i);
    cmin = min(cmin, c);
    cmax = max(cmax, c);
  }
  return cmin * kSpanClasses + min(cmax - cmin, kSpanClasses - 1);
}
__global__ void k_ro_count(Work w, unsigned long long* __restrict__ cnt, int* __restrict__ range) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  int lo = 0x7fffffff, hi = -1;
  for (int j = tid; j < w.n_pt; j += nth) {
    int k, cmin, cmax;
    const int key = point_class(w, j, k, cmin, cmax);
    atomicAdd(cnt + key, (1ull << 32) | (unsigned long long)max(k, 0));
    lo = min(lo, cmin);
    hi = max(hi, cmax);
  }
  // the shard's camera range (the sharded solver skips peers whose range misses a row of the reduced camera system)
  lo = __reduce_min_sync(kFull, lo);
  hi = __reduce_max_sync(kFull, hi);
  if ((threadIdx.x & 31) == 0) {
    if (lo != 0x7fffffff) atomicMin(range, lo);
    if (hi >= 0) atomicMax(range + 1, hi);
  }
}
// one block: exclusive scan of the class counters in place (they become the scatter cursors)
__global__ void k_ro_scan(unsigned long long* __restrict__ cnt, int n_classes) {
  __shared__ unsigned long long s_part[1024];
  __shared__ unsigned long long s_carry;
  const int tid = threadIdx.x;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  for (int base = 0; base < n_classes; base += 1024) {
    const int a = base + tid;
    const unsigned long long v = (a < n_classes) ? cnt[a] : 0ull;
    s_part[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                 // Hillis-Steele inclusive scan
      const unsigned long long t = (tid >= o) ? s_part[tid - o] : 0ull;
      __syncthreads();
      s_part[tid] += t;
      __syncthreads();
    }
    if (a < n_classes) cnt[a] = s_carry + s_part[tid] - v;
    __syncthreads();
    if (tid == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
}
__global__ void k_ro_scatter(Work w, PanelOrder o) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  if (tid == 0) o.off[w.n_pt] = w.n_obs;
  for (int j = tid; j < w.n_pt; j += nth) {
    int k, cmin, cmax;
    const int key = point_class(w, j, k, cmin, cmax);
    const unsigned long long at = atomicAdd(o.cnt + key, (1ull << 32) | (unsigned long long)max(k, 0));
    const int pos = (int)(at >> 32), ob = (int)(at & 0xffffffffull);
    o.perm[pos] = j;
    o.off[pos] = ob;
    const int o0 = __ldg(w.pt_off + j);
    for (int i = 0; i < k; ++i) {
      o.cam[ob + i] = __ldg(w.obs_cam + o0 + i);
      o.pt[ob + i] = pos;
      reinterpret_cast<float2*>(o.uv)[ob + i] = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o0 + i);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_cam_blocks
// ------------------------------------------------------------------------------------------------------------
constexpr int kCamBlocksPerWarp = 16;     // 512 observations between two flushes (C4: 0.277 / 0.246 / 0.233 / 0.240 ms with 4 / 8 / 16 / 32)

__device__ __forceinline__ void cam_flush(const Work& w, int cam, int lane, double (&acc)[27]) {
  // butterfly over the lanes, 27 sums side by side; lane q then owns entry q
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double t[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) t[q] = __shfl_xor_sync(kFull, acc[q], o);
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] += t[q];
  }
  double v = 0.0;
#pragma unroll
  for (int q = 0; q < 27; ++q) if (lane == q) v = acc[q];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  if (lane < 27 && v != 0.0) {
    if (lane < 21) {
      int r = 0, e = lane;
      while (e > r) { e -= r + 1; ++r; }
      atomicAdd(s_block_ptr(w, cam, cam) + 6 * r + e, v);
      if (e == r) atomicAdd(w.diagU + 6 * (size_t)cam + r, v);
    } else {
      atomicAdd(w.b + 6 * (size_t)cam + (lane - 21), v);
    }
  }
}

__global__ void __launch_bounds__(256, 2)
k_cam_blocks(const __grid_constant__ Work w, const int per_warp) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nblk = *w.cs_nblk;
  const LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  const Rec* __restrict__ rec = reinterpret_cast<const Rec*>(w.rec);
  const int b0 = warp * per_warp;
  if (b0 >= nblk) return;
  const int b1 = min(nblk, b0 + per_warp);
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  int cam = -1;
  bool free_cam = false;
  double P[12];
  // the block's camera, point and pixel are fetched one block ahead: what is left of the chain table -> point -> gather
  // inside a trip is the gather of X_j and q_j
  int c_n = __ldg(w.cs_blk_cam + b0);
  int j_n = __ldg(w.cs_pt + 32 * (size_t)b0 + lane);
  float2 uv_n = __ldg(w.cs_uv + 32 * (size_t)b0 + lane);
  for (int b = b0; b < b1; ++b) {
    const int c = c_n, j = j_n;
    const float2 uv = uv_n;
    if (b + 1 < b1) {
      c_n = __ldg(w.cs_blk_cam + b + 1);
      j_n = __ldg(w.cs_pt + 32 * (size_t)(b + 1) + lane);
      uv_n = __ldg(w.cs_uv + 32 * (size_t)(b + 1) + lane);
    }
    if (c != cam) {
      if (cam >= 0 && free_cam) cam_flush(w, cam, lane, acc);
      cam = c;
      free_cam = w.fixed[c] == 0;
#pragma unroll
      for (int q = 0; q < 12; ++q) P[q] = __ldg(pose + 12 * (size_t)c + q);
    }
    if (!free_cam) continue;
    if (j < 0) continue;
    const double X0 = X[3 * (size_t)j], X1 = X[3 * (size_t)j + 1], X2 = X[3 * (size_t)j + 2];
    const double q0 = rec[j].q[0], q1 = rec[j].q[1], q2 = rec[j].q[2];
    ObsLin ol;
    linearize_obs<true>(w.K, P, P + 9, X0, X1, X2, (double)uv.x, (double)uv.y, delta, ol);
    // b_a += J_p^T w (J_l q - r): the Schur correction W V^-1 g_l and the gradient in one term
    const double gu = ol.wu * (ol.lu[0] * q0 + ol.lu[1] * q1 + ol.lu[2] * q2 - ol.ru);
    const double gv = ol.wv * (ol.lv[0] * q0 + ol.lv[1] * q1 + ol.lv[2] * q2 - ol.rv);
    int e = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double pa = ol.wu * ol.pu[r], pb = ol.wv * ol.pv[r];
#pragma unroll
      for (int c2 = 0; c2 <= r; ++c2, ++e) acc[e] += pa * ol.pu[c2] + pb * ol.pv[c2];
      acc[21 + r] += ol.pu[r] * gu + ol.pv[r] * gv;
    }
  }
  if (cam >= 0 && free_cam) cam_flush(w, cam, lane, acc);
}

// ------------------------------------------------------------------------------------------------------------
// k_schur_panel
// ------------------------------------------------------------------------------------------------------------
struct __align__(16) StageIn {          // one panel's inputs, landed by bulk copies
  int cam[kOB + 8];
  int pt[kOB + 8];
  float2 uv[kOB + 8];
  double X[3 * kPP + 2];
  double rec[kRecD * kPP];
};

struct __align__(16) PanelSmem {
  double Yt[kNB][kRows * kYS];
  StageIn in[kNG][2];
  uint64_t bar[kNG][2];                 // mbarriers of the input stages
  int bounds[kBoundsCap + 1];           // pt_off at the panel boundaries of this CTA (one super-chunk)
  unsigned char piv[kNB][16];           // per point slot: first | last << 4 tile (8 columns) of the point's camera interval (0xff: none)
  unsigned char puni[kNB][4];           // [0]: the one tile interval (first | last << 4) all twelve points of the panel share, 0xff: none
  int pbase[kNB];                       // window base the panel was staged with
  int pnpts[kNB];
  int gbase[kNG];                       // the group's snapshot of cur_base for the panel it is staging
  int cur_base;                         // window base for the panels that start from now on
  int n_rebase, n_fallback;             // diagnostics (-DBA_PANEL_TRACE)
};

// Where a window should start so that it holds a point AND the widest point the co-visibility band allows around the same
// place.  Caller's order (points follow an anchor camera): the centre of the point's cameras minus half the band, minus
// one.  (Starting at the panel's first camera minus a margin left the rare 16-keyframe point behind the window for a
// whole run of panels: ten such points made one CTA three times slower than the rest, and the kernel with it.)
// Class order (first cameras never fall): the point's first camera -- centring sent the window back and forth between
// the narrow and the wide classes of one first camera (9-12 moves and ~500 fallback observations per CTA).
__device__ __forceinline__ int window_start(int cmin, int cmax, int half_band, bool class_order) {
  return class_order ? cmin : ((cmin + cmax) >> 1) - half_band - 1;
}

// Y = W C of one observation: rows d = 0..2 of y[r][d], r = 0..5
struct YObs { double y[6][3]; };

__device__ __forceinline__ void obs_Y(const Intr& K, const double* __restrict__ P, double X0, double X1, double X2,
                                      float2 uv, double delta, const double* __restrict__ c, YObs& out) {
  // the 96-byte pose record in six 16-byte loads
  const double2* P2 = reinterpret_cast<const double2*>(P);
  const double2 q0 = __ldg(P2), q1 = __ldg(P2 + 1), q2 = __ldg(P2 + 2), q3 = __ldg(P2 + 3), q4 = __ldg(P2 + 4), q5 = __ldg(P2 + 5);
  const double R0 = q0.x, R1 = q0.y, R2 = q1.x, R3 = q1.y, R4 = q2.x, R5 = q2.y, R6 = q3.x, R7 = q3.y, R8 = q4.x;
  const double T0 = q4.y, T1 = q5.x, T2 = q5.y;
  const double Y0 = R0 * X0 + R1 * X1 + R2 * X2;
  const double Y1 = R3 * X0 + R4 * X1 + R5 * X2;
  const double Y2 = R6 * X0 + R7 * X1 + R8 * X2;
  const double xc = Y0 + T0, yc = Y1 + T1, zc = Y2 + T2;
  const double iz = fast_rcp(zc + kDepthEps);
  const double xz = xc * iz, yz = yc * iz;
  const double ru = K.fx * xz + K.cx - (double)uv.x;
  const double rv = K.fy * yz + K.cy - (double)uv.y;
  const double wu = huber_w(ru, delta), wv = huber_w(rv, delta);
  const double ja = K.fx * iz, jc = -K.fx * xz * iz;
  const double jb = K.fy * iz, jd = -K.fy * yz * iz;
  const double lu0 = ja * R0 + jc * R6, lu1 = ja * R1 + jc * R7, lu2 = ja * R2 + jc * R8;
  const double lv0 = jb * R3 + jd * R6, lv1 = jb * R4 + jd * R7, lv2 = jb * R5 + jd * R8;
  const double c00 = c[0], c01 = c[1], c02 = c[2], c11 = c[3], c12 = c[4], c22 = c[5];
  const double luC0 = lu0 * c00, luC1 = lu0 * c01 + lu1 * c11, luC2 = lu0 * c02 + lu1 * c12 + lu2 * c22;
  const double lvC0 = lv0 * c00, lvC1 = lv0 * c01 + lv1 * c11, lvC2 = lv0 * c02 + lv1 * c12 + lv2 * c22;
  double pu[6], pv[6];
  pu[0] = jc * Y1;             pu[1] = ja * Y2 - jc * Y0;  pu[2] = -ja * Y1;
  pu[3] = ja;                  pu[4] = 0.0;                pu[5] = jc;
  pv[0] = -jb * Y2 + jd * Y1;  pv[1] = -jd * Y0;           pv[2] = jb * Y0;
  pv[3] = 0.0;                 pv[4] = jb;                 pv[5] = jd;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const double pa = wu * pu[r], pb = wv * pv[r];
    out.y[r][0] = pa * luC0 + pb * lvC0;
    out.y[r][1] = pa * luC1 + pb * lvC1;
    out.y[r][2] = pa * luC2 + pb * lvC2;
  }
}

// the same, written straight into the panel: rows d = 0..2 of the point at `dst` (row stride kYS), two pose parameters at
// a time, so that at most six results are alive together (the staging warps sit at the 128-register cap: with all 18
// entries of Y alive the loop spilled, 17 local loads per observation)
__device__ __forceinline__ void obs_Y_store(const Intr& K, const double* __restrict__ P, double X0, double X1, double X2,
                                            float2 uv, double delta, const double* __restrict__ c, double* __restrict__ dst) {
  const double2* P2 = reinterpret_cast<const double2*>(P);
  const double2 q0 = __ldg(P2), q1 = __ldg(P2 + 1), q2 = __ldg(P2 + 2), q3 = __ldg(P2 + 3), q4 = __ldg(P2 + 4), q5 = __ldg(P2 + 5);
  const double R0 = q0.x, R1 = q0.y, R2 = q1.x, R3 = q1.y, R4 = q2.x, R5 = q2.y, R6 = q3.x, R7 = q3.y, R8 = q4.x;
  const double T0 = q4.y, T1 = q5.x, T2 = q5.y;
  const double Y0 = R0 * X0 + R1 * X1 + R2 * X2;
  const double Y1 = R3 * X0 + R4 * X1 + R5 * X2;
  const double Y2 = R6 * X0 + R7 * X1 + R8 * X2;
  const double xc = Y0 + T0, yc = Y1 + T1, zc = Y2 + T2;
  const double iz = fast_rcp(zc + kDepthEps);
  const double xz = xc * iz, yz = yc * iz;
  const double ru = K.fx * xz + K.cx - (double)uv.x;
  const double rv = K.fy * yz + K.cy - (double)uv.y;
  const double wu = huber_w(ru, delta), wv = huber_w(rv, delta);
  const double ja = K.fx * iz, jc = -K.fx * xz * iz;
  const double jb = K.fy * iz, jd = -K.fy * yz * iz;
  const double lu0 = ja * R0 + jc * R6, lu1 = ja * R1 + jc * R7, lu2 = ja * R2 + jc * R8;
  const double lv0 = jb * R3 + jd * R6, lv1 = jb * R4 + jd * R7, lv2 = jb * R5 + jd * R8;
  const double c00 = c[0], c01 = c[1], c02 = c[2], c11 = c[3], c12 = c[4], c22 = c[5];
  const double luC0 = lu0 * c00, luC1 = lu0 * c01 + lu1 * c11, luC2 = lu0 * c02 + lu1 * c12 + lu2 * c22;
  const double lvC0 = lv0 * c00, lvC1 = lv0 * c01 + lv1 * c11, lvC2 = lv0 * c02 + lv1 * c12 + lv2 * c22;
  const double wja = wu * ja, wjc = wu * jc, wjb = wv * jb, wjd = wv * jd;
  auto put = [&](int pair, double pa0, double pb0, double pa1, double pb1) {   // weighted pose Jacobian entries of two parameters
    *reinterpret_cast<double2*>(dst + 0 * kYS + 2 * pair) = make_double2(pa0 * luC0 + pb0 * lvC0, pa1 * luC0 + pb1 * lvC0);
    *reinterpret_cast<double2*>(dst + 1 * kYS + 2 * pair) = make_double2(pa0 * luC1 + pb0 * lvC1, pa1 * luC1 + pb1 * lvC1);
    *reinterpret_cast<double2*>(dst + 2 * kYS + 2 * pair) = make_double2(pa0 * luC2 + pb0 * lvC2, pa1 * luC2 + pb1 * lvC2);
  };
  // pu = (jc Y1, ja Y2 - jc Y0, -ja Y1, ja, 0, jc), pv = (-jb Y2 + jd Y1, -jd Y0, jb Y0, 0, jb, jd), times the weights
  put(0, wjc * Y1, wjd * Y1 - wjb * Y2, wja * Y2 - wjc * Y0, -wjd * Y0);
  put(1, -wja * Y1, wjb * Y0, wja, 0.0);
  put(2, 0.0, wjb, wjc, wjd);
}

// fallback for one observation of a point that cannot go through the panel: its Y against the Y of EVERY
// observation of the point (pairs in both orders are visited by the two lanes involved, each adds the lower part)
__device__ __noinline__ void pair_fallback(const Work& w, const double* __restrict__ pose, const StageIn& in, int i_self,
                                           int i0, int i1, double X0, double X1, double X2, const double* __restrict__ c,
                                           double delta) {
  const int cam_a = in.cam[i_self];
  if (w.fixed[cam_a] != 0) return;
  YObs ya;
  obs_Y(w.K, pose + 12 * (size_t)cam_a, X0, X1, X2, in.uv[i_self], delta, c, ya);
  for (int i = i0; i < i1; ++i) {
    const int cam_b = in.cam[i];
    if (cam_b > cam_a || w.fixed[cam_b] != 0) continue;
    YObs yb;
    obs_Y(w.K, pose + 12 * (size_t)cam_b, X0, X1, X2, in.uv[i], delta, c, yb);
    double* blk = s_block_ptr(w, cam_a, cam_b);
    const bool same = cam_a == cam_b;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        if (same && cc > r) continue;                 // diagonal block: lower part only
        const double v = ya.y[r][0] * yb.y[cc][0] + ya.y[r][1] * yb.y[cc][1] + ya.y[r][2] * yb.y[cc][2];
        atomicAdd(blk + 6 * r + cc, -v);
      }
  }
}

// clock64 trace of chosen threads of CTA 0 (BA_DBG_TIMES=1, -DBA_PANEL_TRACE): slot = event index, tag = what
__device__ __forceinline__ void pmark(const Work& w, bool who, int slot, int tag) {
#ifdef BA_PANEL_TRACE
  if (w.dbg && blockIdx.x == 0 && who && slot >= 0 && slot < 2000) {
    w.dbg[1 + 2 * slot] = tag;
    w.dbg[2 + 2 * slot] = clock64();
  }
#endif
}

// ---- mma warps ----------------------------------------------------------------------------------------------------
// S -= Yt^T Yt per panel with mma.sync.m8n8k4.f64.  What shaped this loop (all measured):
//  * eight per-warp code specialisations starved the instruction fetch (stall_no_instruction on top): ONE code sequence,
//    which tiles an accumulator stands for is data (c_own);
//  * a predicated-off DMMA still occupies the tensor pipe and ptxas turns `if (inside) dmma` into exactly that: trip
//    counts must be data, not predicates;
//  * the tensor pipe takes one DMMA per 16.2 clk and sub-partition and a dependent DMMA can follow after 26 clk
//    (scripts/dmma_latency.cu): two mma warps per sub-partition with one or two accumulators in flight each keep it full,
//    so the unit of work can be small.  Round 2 started with 16 x 16 super-tiles (four DMMAs per four fragment loads); the
//    clock64 trace showed the mma warps, not the staging, setting the pace (3.3 k of 4.2 k clk per panel): a point seen by
//    eight keyframes touches 48 of the 128 columns at a random offset, i.e. 10 super-tiles = 36 DMMAs, dealt to 8 warps in
//    lumps of 3-4 (the busiest warp had 8 per point against a mean of 4.5).
//  * Now the unit is a STRIP: tile row i (8 columns) x strip column sj (16 columns) = the 8 x 8 tiles (i, 2 sj) and
//    (i, 2 sj + 1), one A fragment and two B fragments for two DMMAs; the strip on the diagonal of an even row is the
//    single tile (i, i).  64 strips in the lower triangle of the 15 x 15 tiles; warp (i + 3 sj) mod 8 owns a pair strip,
//    warp 5 sj mod 8 the single one: 7 + 1 strips = 30 accumulator doubles per lane, and for any place of a point's camera
//    interval in the window the sub-partitions' shares stay within 1.15 of the mean (C4: 31.8 DMMAs per point instead of
//    35.3, the busiest sub-partition 99 per panel instead of 126).
// Loop nest: STRIP-major inside a panel.  For each of my 8 strips a ballot says which of the panel's points have it
// inside their tile interval; a loop over the set bits does the work.  All accumulators stay in registers for the life
// of a window.
constexpr int kNT = kCols / 8;            // 15 tile rows / columns of 8 in the window
constexpr int kSlots = 8;                 // strips per mma warp: slot 0 = the single diagonal tile, 1..7 = pair strips
__host__ __device__ constexpr bool strip_single(int i, int sj) { return 2 * sj == i; }
__host__ __device__ constexpr int strip_owner(int i, int sj) { return strip_single(i, sj) ? ((5 * sj) & 7) : ((i + 3 * sj) & 7); }

struct OwnTable { unsigned char ij[kXW][kSlots]; };   // i | sj << 4 of slot s
constexpr OwnTable make_own() {
  OwnTable t{};
  for (int w = 0; w < kXW; ++w) {
    int n = 1;
    for (int s = 0; s < kSlots; ++s) t.ij[w][s] = 0xff;
    for (int i = 0; i < kNT; ++i)
      for (int sj = 0; 2 * sj <= i; ++sj) {
        if (strip_owner(i, sj) != w) continue;
        if (strip_single(i, sj)) t.ij[w][0] = (unsigned char)(i | (sj << 4));
        else if (n < kSlots) t.ij[w][n++] = (unsigned char)(i | (sj << 4));
      }
  }
  return t;
}
constexpr bool own_complete() {           // every strip of the lower triangle has a slot
  const OwnTable t = make_own();
  int n = 0;
  for (int w = 0; w < kXW; ++w)
    for (int s = 0; s < kSlots; ++s)
      if (t.ij[w][s] != 0xff) ++n;
  int want = 0;
  for (int i = 0; i < kNT; ++i) want += i / 2 + 1;
  return n == want;
}
static_assert(own_complete(), "strip ownership must cover the lower triangle with 8 slots per warp");
__constant__ OwnTable c_own = make_own();

// one accumulator entry -> float64 RED (entries that are exactly zero belong to blocks nobody touched); out of line: rare
__device__ __noinline__ void flush_entry(const Work& w, int base, int lane, int ti, int tj, int h, double v) {
  const int m = 8 * ti + (lane >> 2);
  const int n2 = 8 * tj + 2 * (lane & 3) + h;
  if (m >= kCols || n2 >= kCols) return;
  const int a = m / 6, r = m - 6 * a, b = n2 / 6, cc = n2 - 6 * b;
  if (a > b || (a == b && cc <= r)) atomicAdd(s_block_ptr(w, base + a, base + b) + 6 * r + cc, -v);
}

// one strip x the points of `pm`: Ya = panel + 8 i (A fragment), Yb = panel + 16 sj (B fragments at +0, +8), both with
// the lane's fragment offset.  A point that is not in `pm` costs no DMMA.  How this loop got here (clock64 trace + ncu source page of the mma warps, C4):
//  * `while (pm) { p = ffs(pm); ... }` with the next point's fragments loaded before this point's DMMAs: ptxas renamed
//    the prefetched fragments through MOVs at the loop end, i.e. waited for the loads inside the iteration that issued
//    them (BREV -> FLO -> IMAD -> IMAD -> LDS -> MOV in front of every DMMA pair); half of the mma warps' samples were
//    fixed-latency `wait` stalls, 3 % math-pipe throttle: the warps were bound by their own instruction stream (640
//    instructions per panel and warp for 50 DMMAs), not by the tensor pipe.
//  * `if (bit) { loads; dmma; }` per point is if-converted by ptxas (also from a PTX `bra.uni`) into predicated DMMAs,
//    and a predicated-off DMMA holds the tensor pipe as long as a real one (scripts/dmma_pred.cu: 771 vs 802 clk per
//    24 DMMAs with none / all of them on).  A loop of opaque trip count 0 / 1 is kept as a branch but adds a counter.
//  * a switch over pairs of mask bits (compare tree over unpredicated straight-line bodies) halved the instruction
//    count and was slower still (3.8-5.3 k clk per panel against 3.2-4.4 k): taken branches and one exposed LDS latency
//    per pair.
//  * what the tensor pipe wants from a warp is a long unpredicated run of DMMAs with the fragment loads batched in
//    front: the live points of a strip are walked as RUNS of consecutive points (dense loops with data trip counts,
//    unrolled by four).  In class order (k_ro_*) a panel holds one or two camera intervals and a strip has one run.
template <bool kPair>
__device__ __forceinline__ void strip_runs(const double* __restrict__ Ya, const double* __restrict__ Yb, unsigned pm,
                                           double (&acc)[2][2]) {
#pragma unroll 1
  while (pm) {
    const int p0 = __ffs(pm) - 1;
    const int len = __ffs(~(pm >> p0)) - 1;                // pm has at most kPP bits: the complement has a set bit
    pm &= ~(((1u << len) - 1u) << p0);
    const double* __restrict__ ya = Ya + p0 * (4 * kYS);
    const double* __restrict__ yb = Yb + p0 * (4 * kYS);
#pragma unroll 4
    for (int t = 0; t < len; ++t) {
      const double a_ = ya[t * (4 * kYS)];
      if (kPair) {
        const double b0_ = yb[t * (4 * kYS)], b1_ = yb[t * (4 * kYS) + 8];
        dmma(acc[0][0], acc[0][1], a_, b0_);
        dmma(acc[1][0], acc[1][1], a_, b1_);
      } else {
        dmma(acc[0][0], acc[0][1], a_, a_);
      }
    }
  }
}

// a strip x ALL twelve points of a panel (class order: most panels hold one camera interval): straight-line code, no
// masks, no trip counts
template <bool kPair>
__device__ __forceinline__ void strip_full(const double* __restrict__ Ya, const double* __restrict__ Yb, double (&acc)[2][2]) {
#pragma unroll
  for (int t = 0; t < kPP; ++t) {
    const double a_ = Ya[t * (4 * kYS)];
    if (kPair) {
      const double b0_ = Yb[t * (4 * kYS)], b1_ = Yb[t * (4 * kYS) + 8];
      dmma(acc[0][0], acc[0][1], a_, b0_);
      dmma(acc[1][0], acc[1][1], a_, b1_);
    } else {
      dmma(acc[0][0], acc[0][1], a_, a_);
    }
  }
}

__device__ __forceinline__ void mma_warp(const Work& w, PanelSmem& sm, int mw, int lane, int n_panels) {
  double acc[kSlots][2][2];
#pragma unroll
  for (int s = 0; s < kSlots; ++s)
#pragma unroll
    for (int t = 0; t < 2; ++t) { acc[s][t][0] = 0.0; acc[s][t][1] = 0.0; }
  // my strips: tile row i and strip column sj per slot (15 | 15: none), kept in registers (opaque: the compiler otherwise
  // re-derives them from constant memory inside the panel loop)
  int si[kSlots], sj[kSlots];
#pragma unroll
  for (int s = 0; s < kSlots; ++s) {
    const unsigned ij = c_own.ij[mw][s];
    int a = (int)(ij & 15u), b = (int)(ij >> 4);
    asm volatile("mov.s32 %0, %0;" : "+r"(a));
    asm volatile("mov.s32 %0, %0;" : "+r"(b));
    si[s] = a;
    sj[s] = b;
  }
  auto flush = [&](int base) {
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double v = acc[s][t][h];
          acc[s][t][h] = 0.0;
          if (v != 0.0) flush_entry(w, base, lane, si[s], 2 * sj[s] + t, h, v);
        }
  };
  const int foff = (lane & 3) * kYS + (lane >> 2);
  int acc_base = -1;
  for (int n = 0, g = 0; n < n_panels; ++n, g = (g + 1 == kNB) ? 0 : g + 1) {      // g: the panel's buffer
    bar_sync<kGL + 32 * kXW>(kBarFull + g);
    pmark(w, mw == 0 && lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 100 + 4 * (n - kTrace0), 1000 + 10 * n + 1);
    pmark(w, lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 400 + 16 * (n - kTrace0) + 2 * mw, 20000 + 1000 * mw + 10 * (n - kTrace0) + 1);
    const int base_n = sm.pbase[g];
    if (base_n != acc_base) {
      if (acc_base >= 0) flush(acc_base);
      acc_base = base_n;
    }
    const int npts = sm.pnpts[g];
    const double* Yg = sm.Yt[g] + foff;
    const int pu = sm.puni[g][0];
    if (pu != 0xff) {
      // one interval for the whole panel: a strip is in or out as a whole
      const int lo = pu & 15, hi = pu >> 4, lo2 = lo >> 1;
      pmark(w, mw == 0 && lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 101 + 4 * (n - kTrace0), 1000 + 10 * n + 2);
      if (si[0] >= lo && si[0] <= hi) strip_full<false>(Yg + 8 * si[0], Yg, acc[0]);
#pragma unroll
      for (int s = 1; s < kSlots; ++s)
        if (si[s] >= lo && si[s] <= hi && sj[s] >= lo2) strip_full<true>(Yg + 8 * si[s], Yg + 16 * sj[s], acc[s]);
    } else {
    // lane p speaks for point p: its tile interval [lo, hi] (lo > hi: the point is not in the panel product)
    int lo = 15, hi = -1;
    if (lane < npts) {
      const int iv = sm.piv[g][lane];
      if (iv != 0xff) { lo = iv & 15; hi = iv >> 4; }
    }
    const int lo2 = lo >> 1;
    pmark(w, mw == 0 && lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 101 + 4 * (n - kTrace0), 1000 + 10 * n + 2);
    {
      const unsigned pm = __ballot_sync(kFull, si[0] >= lo && si[0] <= hi);
      strip_runs<false>(Yg + 8 * si[0], Yg, pm, acc[0]);
    }
#pragma unroll
    for (int s = 1; s < kSlots; ++s) {
      const unsigned pm = __ballot_sync(kFull, si[s] >= lo && si[s] <= hi && sj[s] >= lo2);
      strip_runs<true>(Yg + 8 * si[s], Yg + 16 * sj[s], pm, acc[s]);
    }
    }
    pmark(w, mw == 0 && lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 102 + 4 * (n - kTrace0), 1000 + 10 * n + 3);
    pmark(w, lane == 0 && n >= kTrace0 && n < kTrace0 + 8, 400 + 16 * (n - kTrace0) + 2 * mw + 1, 20000 + 1000 * mw + 10 * (n - kTrace0) + 3);
    if (n + kNB < n_panels) bar_arrive<kGL + 32 * kXW>(kBarEmpty + g);
  }
  if (acc_base >= 0) flush(acc_base);
}

__global__ void __launch_bounds__(kThreadsP, 1)
k_schur_panel(const __grid_constant__ Work w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PanelSmem& sm = *reinterpret_cast<PanelSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  const double* __restrict__ rec = w.rec;
  const bool sliding = w.n_cam > kWin;
  const bool class_order = w.perm != nullptr;
  const int half_band = (min(w.band_w >= 0 ? w.band_w : kWin - 1, kWin - 3) + 1) >> 1;   // half the widest camera span a point can have

  const int p_lo = __ldg(w.part + blockIdx.x), p_hi = __ldg(w.part + blockIdx.x + 1);
  const int n_panels = (p_hi - p_lo + kPP - 1) / kPP;
#ifdef BA_PANEL_TRACE
  const long long t_start = clock64();
#endif

  if (tid == 0) {
    for (int g = 0; g < kNG; ++g) { mbar_init(&sm.bar[g][0], 1); mbar_init(&sm.bar[g][1], 1); }
    for (int b = 0; b < kNB; ++b) { sm.pbase[b] = 0; sm.pnpts[b] = 0; }
    sm.cur_base = 0;
    sm.n_rebase = 0; sm.n_fallback = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (sliding && tid < 32) {
    // the window starts where this CTA's points start (three cameras before the first camera of its first 32 points);
    // starting every CTA at camera 0 sent the first two panels of each through the pair fallback: a fixed ~0.15-0.7 ms
    // per launch, the whole kernel time of a small shard
    int c = 0x7fffffff;
    if (p_lo + tid < p_hi) {
      const Rec& r = reinterpret_cast<const Rec*>(rec)[p_lo + tid];
      if (!((((unsigned)r.cmaxf) >> 28) & kFlagSkip)) c = window_start(r.cmin, r.cmaxf & 0x0fffffff, half_band, class_order);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(kFull, c, o));
    __syncwarp();
    if (tid == 0 && c != 0x7fffffff) sm.cur_base = max(0, c);
  }
  __syncthreads();

  if (wid >= kNG * kGW) {
    // ========================================= MMA warps ====================================================
    mma_warp(w, sm, wid - kNG * kGW, lane, n_panels);
    return;
  }

  // =========================================== staging warps ================================================
  const int g = wid / kGW;                       // staging group = panel buffer
  const int gt = tid - g * kGL;                  // thread index inside the group
  // service warp of the group = its LAST warp (lanes 96..127 carry an observation least often): it issues the bulk
  // points / records copies of the next panel and keeps the panel's books, off the other warps' critical path
  const int sl = gt - (kGL - 32);                // lane inside the service warp (negative: not in it)
  const bool producer = sl == 0;

  for (int sc0 = 0; sc0 < n_panels; sc0 += kBoundsCap) {
    // panel boundaries of this super-chunk (all staging threads), then the groups run their panels
    const int sc_n = min(kBoundsCap, n_panels - sc0);
    bar_sync<kNG * kGL>(kBarStageAll);
    for (int i = tid; i <= sc_n; i += kNG * kGL) {
      const int jj = min(p_hi, p_lo + (sc0 + i) * kPP);
      sm.bounds[i] = __ldg(w.pt_off + jj);
    }
    bar_sync<kNG * kGL>(kBarStageAll);

    // issue the bulk copies of panel `n` (this group's) into input stage `st`
    auto issue = [&](int n, int st, bool trm) {
      StageIn& si = sm.in[g][st];
      const int j0 = p_lo + n * kPP;
      const int npts = min(kPP, p_hi - j0);
      const int o0 = sm.bounds[n - sc0], o1 = sm.bounds[n - sc0 + 1];
      const int nob = min(o1 - o0, kOB);
      // observations: the 16-byte aligned superset [a0, a1) of the panel's range goes by bulk copy (the up to 3 + 3
      // extra entries belong to neighbouring points and are never looked at); entry i of the stage = observation a0 + i.
      // Only where that superset would run past the end of the arrays (the last panel of the last CTA) do the final
      // entries come by plain loads.  (A first version fetched the unaligned head / tail of EVERY panel with dependent
      // plain loads on the producer thread: ~5 DRAM round trips per panel on the critical path of the whole group.)
      const int a0 = o0 & ~3;
      int a1 = (o0 + nob + 3) & ~3;
      const int lim = w.n_obs & ~3;
      unsigned bytes = 0;
      if (a1 > lim) {
        a1 = max(lim, a0);
        for (int o = max(a1, o0); sl == 0 && o < o0 + nob; ++o) {
          si.cam[o - a0] = __ldg(w.obs_cam + o); si.pt[o - a0] = __ldg(w.obs_pt + o);
          si.uv[o - a0] = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
        }
      }
      const int a_lo = a0, a_hi = a1;
      // points and records: CTA ranges start at multiples of 4 points, so the sources are 16-byte aligned;
      // an odd point count leaves 8 bytes for plain loads
      const unsigned xb = (24u * npts) & ~15u, rb = 80u * npts;
      if (sl == 0 && ((24u * npts) & 15u)) {
        si.X[3 * npts - 1] = X[3 * (size_t)(j0 + npts) - 1];
      }
      if (a_hi > a_lo) bytes += 16u * (unsigned)(a_hi - a_lo);
      bytes += xb + rb;
      // one arrival (service lane 0, with the byte count) + five copies: lane 0 of the group's warps 0..2 sends the
      // observation arrays, two lanes of the service warp the points and records.  (Five lanes of the service warp cost
      // it 1.1-1.5 k clk per panel -- divergent, one UBLKCP after the other -- and the whole group waits for that warp at
      // the panel buffer's barrier.)  A copy that completes before the count is posted only drives the transaction count
      // negative for a moment, the phase still needs the arrival.
      const unsigned nb4 = (a_hi > a_lo) ? 4u * (unsigned)(a_hi - a_lo) : 0u;
      pmark(w, trm, 8 * (n - kNG - kTrace0) + 6, 10 * (n - kNG) + 7);
      if (sl == 0) mbar_expect_tx(&sm.bar[g][st], bytes);
      pmark(w, trm, 8 * (n - kNG - kTrace0) + 7, 10 * (n - kNG) + 8);
      if (gt == 0 && nb4) bulk_g2s(&si.cam[a_lo - a0], w.obs_cam + a_lo, nb4, &sm.bar[g][st]);
      if (gt == 32 && nb4) bulk_g2s(&si.pt[a_lo - a0], w.obs_pt + a_lo, nb4, &sm.bar[g][st]);
      if (gt == 64 && nb4) bulk_g2s(&si.uv[a_lo - a0], reinterpret_cast<const float2*>(w.obs_uv) + a_lo, 2 * nb4, &sm.bar[g][st]);
      if (sl == 0 && xb) bulk_g2s(si.X, X + 3 * (size_t)j0, xb, &sm.bar[g][st]);
      if (sl == 1) bulk_g2s(si.rec, rec + kRecD * (size_t)j0, rb, &sm.bar[g][st]);
    };
    const bool issuer = (lane == 0) || sl == 1;           // the lanes that take part in issue()

    const int n_first = sc0 + g;                          // this group's panels: sc0 + g, + kNG, ...
    const int sc_end = sc0 + sc_n;
    if (issuer && n_first < sc_end) issue(n_first, (n_first / kNG) & 1, false);

    for (int n = n_first; n < sc_end; n += kNG) {
      const int it = n / kNG;                             // this group's it-th panel overall
      const int st = it & 1;
      const int pb = n % kNB;                             // the panel's buffer
      double* __restrict__ Yb = sm.Yt[pb];
      // the group's view of the window base for this panel (the other group may move cur_base at any time)
      if (producer) sm.gbase[g] = sm.cur_base;
      // everyone is done with the other stage (panel n - kNG): its buffers may be refilled
      bar_sync<kGL>(kBarGroup + g);
      const bool tr = sl == 0 && n >= kTrace0 && n < kTrace0 + 8;
      pmark(w, tr, 8 * (n - kTrace0), 10 * n + 1);
      if (issuer && n + kNG < sc_end) issue(n + kNG, st ^ 1, tr);
      pmark(w, tr, 8 * (n - kTrace0) + 1, 10 * n + 2);
      mbar_wait(&sm.bar[g][st], (it >> 1) & 1);
      pmark(w, tr, 8 * (n - kTrace0) + 2, 10 * n + 3);
      const StageIn& si = sm.in[g][st];
      const int j0 = p_lo + n * kPP;
      const int npts = min(kPP, p_hi - j0);
      const int o0 = sm.bounds[n - sc0], o1 = sm.bounds[n - sc0 + 1];
      const int a0 = o0 & ~3;
      const int nob = o1 - o0;
      const bool overflow = nob > kOB;                    // more observations than the stage holds: fallback, from global memory
      // Window base of THIS panel, decided by every warp on its own from the same records (no barrier): the group's
      // snapshot, unless a point of the panel lies beyond the window's end or nothing of the panel fits -- then the window
      // moves now, to window_start() of the panel's points, and the mma warps flush their accumulators when they see the
      // new base.  (Moving it only for LATER panels sent the points that triggered the move through the pair fallback,
      // ~50-100 k clk per point: one such point made a CTA of a small shard a third slower than the others.)  A point
      // that starts before the window never moves it back: it takes the fallback.
      int base = 0;
      bool moved = false;
      Rec myrec;
      myrec.cmin = 0x7fffffff; myrec.cmaxf = kFlagSkip << 28;
      if (lane < npts) myrec = reinterpret_cast<const Rec*>(si.rec)[lane];
      const int my_flags = ((unsigned)myrec.cmaxf) >> 28, my_cmax = myrec.cmaxf & 0x0fffffff;
      if (sliding) {
        base = sm.gbase[g];
        const bool live = !(my_flags & kFlagSkip);
        const bool ahead = live && my_cmax >= base + kWin;
        const bool fits = live && myrec.cmin >= base && my_cmax < base + kWin;
        if (__any_sync(kFull, ahead) || !__any_sync(kFull, fits)) {
          int ws = live ? window_start(myrec.cmin, my_cmax, half_band, class_order) : 0x7fffffff;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ws = min(ws, __shfl_xor_sync(kFull, ws, o));
          if (ws != 0x7fffffff && max(0, ws) != base) { base = max(0, ws); moved = true; }
        }
      }

      // the panel buffer is free again once the mma warps have consumed the panel that was in it.  (Letting the three
      // observation warps pass without the service warp -- late, it has just sent the next panel's copies, ~1.5 k clk of
      // MIO queue -- changed nothing: the mma warps, 3.2 k clk per panel all eight alike, release the buffer later still.)
      if (n >= kNB) bar_sync<kGL + 32 * kXW>(kBarEmpty + pb);
      pmark(w, tr, 8 * (n - kTrace0) + 3, 10 * n + 4);

      // slot intervals + the books, by the group's service warp
      if (sl >= 0) {
        int iv = 0xff;                                    // no interval: skipped by the mma warps
        if (sl < npts && !(my_flags & kFlagSkip)) {
          const bool fits = myrec.cmin >= base && my_cmax < base + kWin;
          if (fits && !(my_flags & kFlagFallback) && !overflow) {
            const int lo = (6 * (myrec.cmin - base)) >> 3, hi = (6 * (my_cmax - base) + 5) >> 3;   // tiles of 8 columns
            iv = lo | (hi << 4);
          }
        }
        if (sl < npts) sm.piv[pb][sl] = (unsigned char)iv;
        // all twelve points in one interval (class order): the mma warps take their strips whole
        const int iv0 = __shfl_sync(kFull, iv, 0);
        const bool uni = __all_sync(kFull, sl >= npts || iv == iv0) && iv0 != 0xff && npts == kPP;
        if (sl == 0) {
          sm.puni[pb][0] = (unsigned char)(uni ? iv0 : 0xff);
          sm.pbase[pb] = base; sm.pnpts[pb] = npts;
          if (moved) { sm.cur_base = base; sm.n_rebase++; }
        }
      }

      pmark(w, tr, 8 * (n - kTrace0) + 4, 10 * n + 5);
      // ---- lane <-> observation ---------------------------------------------------------------------------
      if (!overflow) {
        for (int i = gt; i < nob; i += kGL) {
          const int e = (o0 - a0) + i;                    // entry in the stage
          const int cam = si.cam[e];
          const int slot = si.pt[e] - j0;
          const Rec& r = reinterpret_cast<const Rec*>(si.rec)[slot];
          const int flags = ((unsigned)r.cmaxf) >> 28, cmax = r.cmaxf & 0x0fffffff;
          if (flags & kFlagSkip) continue;
          const double X0 = si.X[3 * slot], X1 = si.X[3 * slot + 1], X2 = si.X[3 * slot + 2];
          const bool fits = r.cmin >= base && cmax < base + kWin;
          if (!fits || (flags & kFlagFallback)) {
            // the point's observations are entries [e - k_before, ...): find its range in the stage
            int i0 = e, i1 = e + 1;
            while (i0 > (o0 - a0) && si.pt[i0 - 1] - j0 == slot) --i0;
            while (i1 < (o0 - a0) + nob && si.pt[i1] - j0 == slot) ++i1;
#ifdef BA_PANEL_TRACE
            atomicAdd(&sm.n_fallback, 1);
#endif
            pair_fallback(w, pose, si, e, i0, i1, X0, X1, X2, r.c, delta);
            continue;
          }
          const int rel = cam - base;
          const bool first = (i == 0) || (si.pt[e - 1] - j0 != slot);
          const bool last = (i == nob - 1) || (si.pt[e + 1] - j0 != slot);
          // the mma warps read whole strips (16 columns): zeros out to the 16-column boundaries around the point's cameras
          const int lo_col = 16 * ((6 * (r.cmin - base)) >> 4);
          const int hi_col = 16 * (((6 * (cmax - base) + 5) >> 4) + 1);
          const int next_col = last ? hi_col : 6 * (si.cam[e + 1] - base);
          double* row0 = Yb + (size_t)(4 * slot) * kYS;
          // my six columns: Y (rows 0..2) and the zero pad row
          if (w.fixed[cam] == 0) {
            obs_Y_store(w.K, pose + 12 * (size_t)cam, X0, X1, X2, si.uv[e], delta, r.c, row0 + 6 * rel);
          } else {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              double2* dst = reinterpret_cast<double2*>(row0 + d * kYS + 6 * rel);
              dst[0] = dst[1] = dst[2] = make_double2(0.0, 0.0);
            }
          }
          {
            double2* dst = reinterpret_cast<double2*>(row0 + 3 * kYS + 6 * rel);
            dst[0] = dst[1] = dst[2] = make_double2(0.0, 0.0);
          }
          // zeros for the columns nobody owns: before the point's first camera, and between mine and the next
          if (first) {
            for (int cc = lo_col; cc < 6 * rel; cc += 2)
#pragma unroll
              for (int d = 0; d < 4; ++d) *reinterpret_cast<double2*>(row0 + d * kYS + cc) = make_double2(0.0, 0.0);
          }
          for (int cc = 6 * rel + 6; cc < next_col; cc += 2)
#pragma unroll
            for (int d = 0; d < 4; ++d) *reinterpret_cast<double2*>(row0 + d * kYS + cc) = make_double2(0.0, 0.0);
        }
      } else {
        // oversized panel (does not happen with kOB = 512 unless points average > 42 observations): every
        // observation through the pair fallback, straight from global memory, one point per lane
        for (int s = gt; s < npts; s += kGL) {
          const int j = j0 + s;
          const Rec r = reinterpret_cast<const Rec*>(rec)[j];
          if ((((unsigned)r.cmaxf) >> 28) & kFlagSkip) continue;
          const int q0 = __ldg(w.pt_off + j), q1 = __ldg(w.pt_off + j + 1);
          const double X0 = X[3 * (size_t)j], X1 = X[3 * (size_t)j + 1], X2 = X[3 * (size_t)j + 2];
          for (int oa = q0; oa < q1; ++oa) {
            const int cam_a = __ldg(w.obs_cam + oa);
            if (w.fixed[cam_a] != 0) continue;
            YObs ya;
            obs_Y(w.K, pose + 12 * (size_t)cam_a, X0, X1, X2, __ldg(reinterpret_cast<const float2*>(w.obs_uv) + oa), delta, r.c, ya);
            for (int ob = q0; ob < q1; ++ob) {
              const int cam_b = __ldg(w.obs_cam + ob);
              if (cam_b > cam_a || w.fixed[cam_b] != 0) continue;
              YObs yb;
              obs_Y(w.K, pose + 12 * (size_t)cam_b, X0, X1, X2, __ldg(reinterpret_cast<const float2*>(w.obs_uv) + ob), delta, r.c, yb);
              double* blk = s_block_ptr(w, cam_a, cam_b);
              for (int rr = 0; rr < 6; ++rr)
                for (int cc = 0; cc < 6; ++cc) {
                  if (cam_a == cam_b && cc > rr) continue;
                  atomicAdd(blk + 6 * rr + cc, -(ya.y[rr][0] * yb.y[cc][0] + ya.y[rr][1] * yb.y[cc][1] + ya.y[rr][2] * yb.y[cc][2]));
                }
            }
          }
        }
      }
      pmark(w, tr, 8 * (n - kTrace0) + 5, 10 * n + 6);
      __threadfence_block();
      bar_arrive<kGL + 32 * kXW>(kBarFull + pb);
    }
  }
#ifdef BA_PANEL_TRACE
  bar_sync<kNG * kGL>(kBarStageAll);
  if (tid == 0 && w.dbg) printf("cta %d: points %d panels %d rebases %d fallback observations %d clocks %lld\n", (int)blockIdx.x, p_hi - p_lo, n_panels, sm.n_rebase, sm.n_fallback, clock64() - t_start);
#endif
}

template <int kG>
void launch_pb(const Work& w, int sm_count, cudaStream_t s) {
  constexpr int kPer = 32 / kG;
  long long need = ((long long)w.n_pt + 8 * kPer - 1) / (8 * kPer);
  int blocks = (int)(need < 1 ? 1 : need);
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  k_point_blocks<kG><<<blocks, 256, 0, s>>>(w);
}

}  // namespace

int configure_panel() {
  cudaError_t e = cudaFuncSetAttribute(k_schur_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PanelSmem));
  return e == cudaSuccess ? 0 : (int)e;
}

int panel_window() { return kWin; }
int panel_points_alignment() { return 4; }
size_t panel_rec_doubles(int n_pt) { return (size_t)kRecD * ((size_t)n_pt + 1); }
int panel_cs_blocks_max(int n_cam, int n_obs) { return n_obs / 32 + n_cam + 1; }

// the panel path's bulk copies need 16-byte aligned observation arrays
bool panel_supported(const Work& w) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return w.part != nullptr && w.rec != nullptr && w.cs_pt != nullptr && al(w.obs_cam) && al(w.obs_pt) && al(w.obs_uv) &&
         al(w.X[0]) && al(w.X[1]) && al(w.rec);
}

size_t panel_classes(int n_cam) { return (size_t)n_cam * kSpanClasses + 1; }

int launch_panel_reorder(const Work& w, const PanelOrder& o, cudaStream_t s) {
  const int ncls = (int)panel_classes(w.n_cam);
  cudaMemsetAsync(o.cnt, 0, sizeof(unsigned long long) * (size_t)(ncls + 1), s);
  static const int kEmptyRange[2] = {0x7fffffff, -1};
  cudaMemcpyAsync(o.range, kEmptyRange, sizeof(kEmptyRange), cudaMemcpyHostToDevice, s);
  int blocks = (w.n_pt + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_ro_count<<<blocks, 256, 0, s>>>(w, o.cnt, o.range);
  k_ro_scan<<<1, 1024, 0, s>>>(o.cnt, ncls);
  k_ro_scatter<<<blocks, 256, 0, s>>>(w, o);
  return 3;
}

// once per solve: camera-sorted copy of the observations for k_cam_blocks
int launch_panel_setup(const Work& w, int sm_count, cudaStream_t s) {
  cudaMemsetAsync(w.cs_cnt, 0, sizeof(int) * (size_t)w.n_cam, s);
  cudaMemsetAsync(w.cs_pt, 0xff, sizeof(int) * 32 * (size_t)w.cs_nblk_max, s);
  int blocks = (w.n_obs + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > sm_count * 8) blocks = sm_count * 8;
  k_cs_count<<<blocks, 256, 0, s>>>(w);
  k_cs_scan<<<1, 1024, 0, s>>>(w);
  k_cs_fill<<<blocks, 256, 0, s>>>(w);
  return 3;
}

int launch_linearize_panel(const Work& w, int sm_count, cudaStream_t s, cudaEvent_t* marks) {
  // marks (profile mode, else null): events after the point-block pass and after the camera pass
  // lanes per point of the point-block pass: the mean observation count picks it
  // lanes per point of the point-block pass.  The pass is bound by the latency of a point's chain (offsets -> observation
  // -> pose gather -> 3x3 factor by one lane of the group), so what counts is points in flight: with enough points to
  // fill the machine ONE lane per point wins (C4, 1 M points, 8 observations each: 0.128 ms against 0.132 / 0.174 /
  // 0.276 / 0.489 ms with 2 / 4 / 8 / 16 lanes); small windows need the lanes for parallelism.  BA_PB_GROUP overrides.
  const double mean_k = w.n_pt > 0 ? (double)w.n_obs / (double)w.n_pt : 0.0;
  static int force_g = -1;
  if (force_g < 0) { const char* e = getenv("BA_PB_GROUP"); force_g = e ? atoi(e) : 0; }
  int g = force_g;
  if (g != 1 && g != 2 && g != 4 && g != 8 && g != 16 && g != 32) {
    if (mean_k <= 16.0 && w.n_pt >= 100000) g = 1;
    else if (mean_k <= 16.0 && w.n_pt >= 40000) g = 2;
    else g = mean_k <= 6.0 ? 4 : (mean_k <= 12.0 ? 8 : (mean_k <= 24.0 ? 16 : 32));
  }
  switch (g) {
    case 1: launch_pb<1>(w, sm_count, s); break;
    case 2: launch_pb<2>(w, sm_count, s); break;
    case 4: launch_pb<4>(w, sm_count, s); break;
    case 8: launch_pb<8>(w, sm_count, s); break;
    case 16: launch_pb<16>(w, sm_count, s); break;
    default: launch_pb<32>(w, sm_count, s); break;
  }
  if (marks) cudaEventRecord(marks[0], s);
  {
    static int per_warp = -1;          // diagnostics: BA_CAM_BLOCKS = 32-observation blocks between two flushes of a warp
    if (per_warp < 0) { const char* e = getenv("BA_CAM_BLOCKS"); per_warp = (e && atoi(e) > 0) ? atoi(e) : kCamBlocksPerWarp; }
    const int warps = (w.cs_nblk_max + per_warp - 1) / per_warp;
    const int blocks = (warps + 7) / 8;
    k_cam_blocks<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(w, per_warp);
  }
  if (marks) cudaEventRecord(marks[1], s);
  k_schur_panel<<<w.part_n, kThreadsP, sizeof(PanelSmem), s>>>(w);
  return 3;
}

}  // namespace ba
