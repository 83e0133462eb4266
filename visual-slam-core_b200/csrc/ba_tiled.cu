// Tiled linearise + Schur kernel for local windows (n_cam <= kNCmax).
//
// The scatter kernel (ba_local.cu) pays one float64 global atomic per entry of every
// W_ij V_j^-1 W_i'j^T block: 43 M atomics per LM iteration on the 20-keyframe window.  Here the
// Schur complement is accumulated in REGISTERS: a CTA stages a chunk of points in shared memory
//   Y_ij = W_ij C_j  with  C_j C_j^T = (V_j + lambda diag V_j)^-1      (6x3 per observation)
//   Jt_ij = sqrt(w) J_p, rt_ij = sqrt(w) r                               (pose rows, residual)
// and then every thread owns one 6x6 block (a, b) of the reduced camera matrix and walks the
// chunk's points, adding Y_aj Y_bj^T when both cameras see point j (slot table lookup).  Diagonal
// blocks, b and diag(U) are split over two threads per camera.  One flush of float64 REDs per
// chunk and block.  The inner loop has no atomics, no shuffles and conflict-free shared loads
// (Y stride 19 doubles, broadcast of Y_a across the threads that share camera a).
#include <stdlib.h>

#include "ba_kernels.cuh"
#include "ba_scatter.cuh"

namespace ba {

static constexpr int kTT = 256;     // threads per CTA
static constexpr int kTP = 64;      // max points per chunk
static constexpr int kTO = 640;     // max observations per chunk
static constexpr int kYS = 19;      // Y stride (doubles), odd -> conflict-free 64-bit loads
static constexpr int kJS = 15;      // Jt (12) + rt (2), padded to odd
static constexpr int kNCmax = 40;

// small windows use camera indices directly; larger ones a per-chunk camera window [cmin, cmin+L)
bool tiled_supported(int n_cam) { (void)n_cam; return true; }

struct TiledSmem {
  double Y[kTO * kYS];
  double J[kTO * kJS];
  double Pt[kTP * 9];          // C_j (upper triangular: c00 c01 c02 c11 c12 c22) and z_j = C^T g_l
  int slot[kTP * kNCmax];      // (point, camera) -> staged observation index or -1
  short next[kTO];             // chain of further observations of the same (point, camera)
  short ptl[kTO];              // chunk-local point of each staged observation
  int off[kTP + 1];
  double red[kTT / 32];
  int npts, o_lo, cmin, cmax;
};

__device__ __forceinline__ int lower_bound_off(const int* __restrict__ off, int n, int target) {
  // first index i in [0, n] with off[i] >= target
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int kDbg, bool kWindowed>
__global__ void __launch_bounds__(kTT, 1)
k_linearize_tiled(Work w) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TiledSmem& sm = *reinterpret_cast<TiledSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  LmCtl* ctl = w.ctl;
  const int cur = ctl->cur;
  const double lambda = ctl->lambda, delta = ctl->huber_delta;
  const double* __restrict__ pose = w.pose[cur];
  const double* __restrict__ X = w.X[cur];
  const int ncam = w.n_cam;

  // this CTA's points: equal observation share
  const long long nobs = w.n_obs;
  const int t_lo = (int)(nobs * blockIdx.x / gridDim.x);
  const int t_hi = (int)(nobs * (blockIdx.x + 1) / gridDim.x);
  int p_lo = (blockIdx.x == 0) ? 0 : lower_bound_off(w.pt_off, w.n_pt, t_lo);
  const int p_hi = (blockIdx.x == gridDim.x - 1) ? w.n_pt : lower_bound_off(w.pt_off, w.n_pt, t_hi);

  double cost = 0.0;

  int p = p_lo;
  while (p < p_hi) {
    // ---- chunk bounds: as many whole points as fit kTP points / kTO observations -------------
    if (tid < kTP) {
      int q = p + tid + 1;
      bool ok = (q <= p_hi) && (__ldg(w.pt_off + q) - __ldg(w.pt_off + p) <= kTO);
      unsigned m = __ballot_sync(kFull, ok);
      if (lane == 0) sm.off[kTP - 1 - wid] = __popc(m);   // scratch: counts of the two warps
    }
    __syncthreads();
    if (tid == 0) {
      int c0 = sm.off[kTP - 1], c1 = sm.off[kTP - 2];
      int cnt = (c0 == 32) ? 32 + c1 : c0;
      sm.npts = cnt;
      sm.o_lo = __ldg(w.pt_off + p);
      sm.cmin = 0x7fffffff;
      sm.cmax = -1;
      if (cnt == 0) ctl->solve_fail = 2;      // a single point with more than kTO observations
    }
    __syncthreads();
    int npts = sm.npts;
    const int o_lo = sm.o_lo;
    if (npts == 0) { p += 1; __syncthreads(); continue; }
    if (tid <= npts) sm.off[tid] = __ldg(w.pt_off + p + tid) - o_lo;
    __syncthreads();
    const int nobs_c = sm.off[npts];

    // ---- pass 1: per-observation V / g_l contributions, cost ---------------------------------
    for (int i = tid; i < nobs_c; i += kTT) {
      const int o = o_lo + i;
      const int cam = __ldg(w.obs_cam + o);
      const int pj = __ldg(w.obs_pt + o) - p;
      float2 uv = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
      const double* Xj = X + 3 * (size_t)(p + pj);
      ObsLin ol;
      linearize_obs<false>(w.K, pose + 12 * cam, pose + 12 * cam + 9, Xj[0], Xj[1], Xj[2],
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
      cost += huber_rho(ol.ru, delta) + huber_rho(ol.rv, delta);
      if (kWindowed) { atomicMin(&sm.cmin, cam); atomicMax(&sm.cmax, cam); }
    }
    __syncthreads();
    // camera window of this chunk: local index = cam - cmin, L cameras
    const int cmin = kWindowed ? sm.cmin : 0;
    const int L = kWindowed ? (sm.cmax - sm.cmin + 1) : ncam;
    if (kWindowed && L > kNCmax) {
      // cameras too far apart for the local window (e.g. a loop closure): scatter this chunk
      __syncthreads();
      double (*sW)[19] = reinterpret_cast<double (*)[19]>(sm.Y) + wid * 32;
      int* sCam = sm.slot + wid * 32;
      double dummy = 0.0;
      for (int pj = wid; pj < npts; pj += kTT / 32)
        scatter_point(w, p + pj, lane, lambda, delta, pose, X, sW, sCam, dummy);
      p += npts;
      __syncthreads();
      continue;
    }
    for (int i = tid; i < npts * L; i += kTT) {
      int pj = i / L, c = i - pj * L;
      sm.slot[pj * kNCmax + c] = -1;
    }
    // ---- per point: V, g_l -> C_j (C C^T = V^-1), z_j = C^T g_l -----------------------------
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
      // C = [ s0  m10 s1  m20 s2 ; 0  s1  m21 s2 ; 0 0 s2 ]
      const double c00 = s0, c01 = m10 * s1, c02 = m20 * s2, c11 = s1, c12 = m21 * s2, c22 = s2;
      P[0] = c00; P[1] = c01; P[2] = c02; P[3] = c11; P[4] = c12; P[5] = c22;
      // z = C^T g_l
      P[6] = c00 * gl[0];
      P[7] = c01 * gl[0] + c11 * gl[1];
      P[8] = c02 * gl[0] + c12 * gl[1] + c22 * gl[2];
    }
    __syncthreads();
    // ---- pass 2: Y = W C, sqrt-weighted pose rows, slot table --------------------------------
    for (int i = tid; i < nobs_c; i += kTT) {
      const int o = o_lo + i;
      const int cam = __ldg(w.obs_cam + o);
      short nx = -1;
      if (w.fixed[cam] == 0) {
        const int pj = sm.ptl[i];
        float2 uv = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
        const double* Xj = X + 3 * (size_t)(p + pj);
        ObsLin ol;
        linearize_obs<true>(w.K, pose + 12 * cam, pose + 12 * cam + 9, Xj[0], Xj[1], Xj[2],
                            (double)uv.x, (double)uv.y, delta, ol);
        const double* P = sm.Pt + pj * 9;
        const double c00 = P[0], c01 = P[1], c02 = P[2], c11 = P[3], c12 = P[4], c22 = P[5];
        double* y = sm.Y + i * kYS;
        double* jt = sm.J + i * kJS;
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
        nx = (short)atomicExch(&sm.slot[pj * kNCmax + (cam - cmin)], i);
      }
      sm.next[i] = nx;
    }
    __syncthreads();
    // ---- merge repeated (point, camera) observations: the head's Y becomes the block sum ------
    for (int i = tid; i < npts * L; i += kTT) {
      int pj = i / L, c = i - pj * L;
      int h = sm.slot[pj * kNCmax + c];
      if (h >= 0 && sm.next[h] >= 0) {
        double* yh = sm.Y + h * kYS;
        for (int s2 = sm.next[h]; s2 >= 0; s2 = sm.next[s2]) {
          const double* ys = sm.Y + s2 * kYS;
#pragma unroll
          for (int q = 0; q < 18; ++q) yh[q] += ys[q];
        }
      }
    }
    __syncthreads();

    // ---- phase B: register-accumulated blocks ------------------------------------------------
    // task list: L(L-1)/2 off-diagonal blocks of the window (cameras without a slot, e.g. fixed
    // ones, never contribute and are skipped at flush), then two tasks per camera for the diagonal
    const int n_off = L * (L - 1) / 2;
    const int n_task = n_off + 2 * L;
    for (int task = tid; task < (kDbg == 2 ? 0 : n_task); task += kTT) {
      if (task < n_off) {
        // block (a, b), a > b: task = a(a-1)/2 + b
        int a = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)task)) * 0.5f);
        while (a * (a - 1) / 2 > task) --a;
        while ((a + 1) * a / 2 <= task) ++a;
        const int b = task - a * (a - 1) / 2;
        double acc[36];
#pragma unroll
        for (int q = 0; q < 36; ++q) acc[q] = 0.0;
        bool any = false;
        for (int j = 0; j < npts; ++j) {
          const int sa = sm.slot[j * kNCmax + a], sb = sm.slot[j * kNCmax + b];
          if ((sa | sb) < 0) continue;
          any = true;
          const double* Ya = sm.Y + sa * kYS;
          const double* Yb = sm.Y + sb * kYS;
          double ya[18];
#pragma unroll
          for (int q = 0; q < 18; ++q) ya[q] = Ya[q];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double b0 = Yb[3 * c], b1 = Yb[3 * c + 1], b2 = Yb[3 * c + 2];
#pragma unroll
            for (int r = 0; r < 6; ++r)
              acc[6 * r + c] += ya[3 * r] * b0 + ya[3 * r + 1] * b1 + ya[3 * r + 2] * b2;
          }
        }
        if (any && (kDbg != 1 || acc[7] == 1.2345e300)) {
          double* Sd = s_block_ptr(w, cmin + a, cmin + b);
#pragma unroll
          for (int q = 0; q < 36; ++q) atomicAdd(Sd + q, -acc[q]);
        }
      } else {
        // diagonal work of camera a, split in two: part 0 = S rows 0..3 + diag(U); part 1 = S rows 4,5 + b
        const int t2 = task - n_off;
        const int a = t2 >> 1, part = t2 & 1;
        double acc[17];
#pragma unroll
        for (int q = 0; q < 17; ++q) acc[q] = 0.0;
        bool any = false;
        for (int j = 0; j < npts; ++j) {
          const int h = sm.slot[j * kNCmax + a];
          if (h < 0) continue;
          any = true;
          const double* y = sm.Y + h * kYS;
          const double* z = sm.Pt + j * 9 + 6;
          if (part == 0) {
            int e = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c <= r; ++c)
                acc[e++] -= y[3 * r] * y[3 * c] + y[3 * r + 1] * y[3 * c + 1] + y[3 * r + 2] * y[3 * c + 2];
          } else {
            int e = 0;
#pragma unroll
            for (int r = 4; r < 6; ++r)
#pragma unroll
              for (int c = 0; c <= r; ++c)
                acc[e++] -= y[3 * r] * y[3 * c] + y[3 * r + 1] * y[3 * c + 1] + y[3 * r + 2] * y[3 * c + 2];
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[11 + r] += y[3 * r] * z[0] + y[3 * r + 1] * z[1] + y[3 * r + 2] * z[2];
          }
          for (int s2 = h; s2 >= 0; s2 = sm.next[s2]) {
            const double* jt = sm.J + s2 * kJS;
            if (part == 0) {
              int e = 0;
#pragma unroll
              for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) acc[e++] += jt[r] * jt[c] + jt[6 + r] * jt[6 + c];
#pragma unroll
              for (int r = 0; r < 6; ++r) acc[10 + r] += jt[r] * jt[r] + jt[6 + r] * jt[6 + r];
            } else {
              int e = 0;
#pragma unroll
              for (int r = 4; r < 6; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) acc[e++] += jt[r] * jt[c] + jt[6 + r] * jt[6 + c];
#pragma unroll
              for (int r = 0; r < 6; ++r) acc[11 + r] -= jt[r] * jt[12] + jt[6 + r] * jt[13];
            }
          }
        }
        if (any && (kDbg != 1 || acc[7] == 1.2345e300)) {
          double* Sd = s_block_ptr(w, cmin + a, cmin + a);
          if (part == 0) {
            int e = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c <= r; ++c) atomicAdd(Sd + 6 * r + c, acc[e++]);
#pragma unroll
            for (int r = 0; r < 6; ++r) atomicAdd(w.diagU + 6 * (size_t)(cmin + a) + r, acc[10 + r]);
          } else {
            int e = 0;
#pragma unroll
            for (int r = 4; r < 6; ++r)
#pragma unroll
              for (int c = 0; c <= r; ++c) atomicAdd(Sd + 6 * r + c, acc[e++]);
#pragma unroll
            for (int r = 0; r < 6; ++r) atomicAdd(w.b + 6 * (size_t)(cmin + a) + r, acc[11 + r]);
          }
        }
      }
    }
    p += npts;
    __syncthreads();
  }

  cost = warp_sum(cost);
  if (lane == 0) sm.red[wid] = cost;
  __syncthreads();
  if (tid == 0) {
    double c = 0.0;
#pragma unroll
    for (int i = 0; i < kTT / 32; ++i) c += sm.red[i];
    if (c != 0.0) atomicAdd(w.cost_lin, c);
  }
}

int configure_tiled() {
  const int bytes = (int)sizeof(TiledSmem);
  cudaError_t e = cudaFuncSetAttribute(k_linearize_tiled<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linearize_tiled<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linearize_tiled<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_linearize_tiled<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

int launch_linearize_tiled(const Work& w, int sm_count, cudaStream_t s) {
  int blocks = sm_count;
  int by_obs = (w.n_obs + 127) / 128;            // do not spread tiny windows over idle CTAs
  if (blocks > by_obs) blocks = by_obs;
  if (blocks < 1) blocks = 1;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("BA_TILED_DBG"); dbg = e ? atoi(e) : 0; }
  if (w.n_cam > kNCmax) k_linearize_tiled<0, true><<<blocks, kTT, sizeof(TiledSmem), s>>>(w);
  else if (dbg == 1) k_linearize_tiled<1, false><<<blocks, kTT, sizeof(TiledSmem), s>>>(w);
  else if (dbg == 2) k_linearize_tiled<2, false><<<blocks, kTT, sizeof(TiledSmem), s>>>(w);
  else k_linearize_tiled<0, false><<<blocks, kTT, sizeof(TiledSmem), s>>>(w);
  return 1;
}

}  // namespace ba
