// Batched pose-only (motion-only) LM: one CTA per frame, the whole LM loop inside the kernel.
//
// Replaces the per-frame pose refinement the reference delegates to OpenCV
// (visual_slam/utils/motion_estimation.py:107-116, called from tracking.py:350-356) with the same
// objective as local BA (torch_optimizer.py:152-177) restricted to one free pose and fixed points.
// Per iteration: every thread linearises its matches and accumulates the 6x6 normal equations in
// float64 registers; a shuffle + shared-memory reduction gives H, g; one thread solves the damped
// 6x6 system; all threads evaluate the candidate cost; thread 0 accepts or rejects.
#include "ba_kernels.cuh"

namespace ba {

static constexpr int kPoseThreads = 256;
static constexpr int kPoseWarps = kPoseThreads / 32;

// in-place Cholesky solve of a 6x6 SPD system (lower triangle of H used); returns false if not SPD
__device__ bool chol6_solve(double H[36], const double g[6], double x[6]) {
  for (int k = 0; k < 6; ++k) {
    double akk = H[7 * k];
    if (!(akk > 0.0) || !isfinite(akk)) return false;
    double d = sqrt(akk), inv = 1.0 / d;
    H[7 * k] = d;
    for (int i = k + 1; i < 6; ++i) H[6 * i + k] *= inv;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) H[6 * i + j] -= H[6 * i + k] * H[6 * j + k];
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double s = g[i];
    for (int j = 0; j < i; ++j) s -= H[6 * i + j] * y[j];
    y[i] = s / H[7 * i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
    for (int j = i + 1; j < 6; ++j) s -= H[6 * j + i] * x[j];
    x[i] = s / H[7 * i];
  }
  bool ok = true;
  for (int i = 0; i < 6; ++i) ok = ok && isfinite(x[i]);
  return ok;
}

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* s_red /* [kPoseWarps][N] */,
                                          double* s_out /* [N] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // butterfly with the N sums side by side (independent shuffle/add chains per step)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double tmp[N];
#pragma unroll
    for (int i = 0; i < N; ++i) tmp[i] = __shfl_xor_sync(kFull, v[i], o);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += tmp[i];
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) s_red[wid * N + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double a = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < kPoseWarps; ++w2) a += s_red[w2 * N + threadIdx.x];
    s_out[threadIdx.x] = a;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPoseThreads)
k_pose_only(PoseWork w) {
  __shared__ double s_red[kPoseWarps * 28];
  __shared__ double s_sum[28];
  __shared__ double s_pose[2][12];
  __shared__ double s_cost, s_lambda;
  __shared__ int s_cur, s_fail;
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  const int m0 = w.frame_off[f], m1 = w.frame_off[f + 1];
  const double delta = w.huber_delta;

  if (tid < 12) {
    double v = tid < 9 ? w.R[9 * f + tid] : w.t[3 * f + tid - 9];
    s_pose[0][tid] = v;
    s_pose[1][tid] = v;
  }
  if (tid == 0) { s_cur = 0; s_lambda = w.lambda0; s_fail = 0; }
  __syncthreads();

  // initial cost
  {
    double c[1] = {0.0};
    const double* P = s_pose[0];
    for (int m = m0 + tid; m < m1; m += kPoseThreads) {
      float2 uv = __ldg(reinterpret_cast<const float2*>(w.uv) + m);
      double ru, rv;
      residual_only(w.K, P, P + 9, (double)__ldg(w.X + 3 * m), (double)__ldg(w.X + 3 * m + 1),
                    (double)__ldg(w.X + 3 * m + 2), (double)uv.x, (double)uv.y, ru, rv);
      c[0] += huber_rho(ru, delta) + huber_rho(rv, delta);
    }
    block_sum<1>(c, s_red, s_sum);
    if (tid == 0) {
      s_cost = s_sum[0];
      if (w.cost_hist) w.cost_hist[(size_t)f * (w.max_iter + 1)] = s_sum[0];
    }
    __syncthreads();
  }

  for (int it = 0; it < w.max_iter; ++it) {
    const int cur = s_cur;
    const double* P = s_pose[cur];
    // normal equations: 21 unique entries of H (lower, row-major) + 6 of g
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.0;
    for (int m = m0 + tid; m < m1; m += kPoseThreads) {
      float2 uv = __ldg(reinterpret_cast<const float2*>(w.uv) + m);
      ObsLin ol;
      linearize_obs<true, false>(w.K, P, P + 9, (double)__ldg(w.X + 3 * m),
                                 (double)__ldg(w.X + 3 * m + 1), (double)__ldg(w.X + 3 * m + 2),
                                 (double)uv.x, (double)uv.y, delta, ol);
      int e = 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double a = ol.wu * ol.pu[r], b = ol.wv * ol.pv[r];
#pragma unroll
        for (int c = 0; c <= r; ++c) acc[e++] += a * ol.pu[c] + b * ol.pv[c];
        acc[21 + r] += a * ol.ru + b * ol.rv;
      }
    }
    block_sum<27>(acc, s_red, s_sum);
    if (tid == 0) {
      double H[36], g[6], x[6];
      int e = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c <= r; ++c) { H[6 * r + c] = s_sum[e]; H[6 * c + r] = s_sum[e]; ++e; }
      for (int r = 0; r < 6; ++r) { H[7 * r] += damp(H[7 * r], s_lambda); g[r] = -s_sum[21 + r]; }
      bool ok = chol6_solve(H, g, x);
      s_fail = ok ? 0 : 1;
      double* Pn = s_pose[cur ^ 1];
      if (ok) {
        so3_exp_mul(x, P, Pn);
        for (int q = 0; q < 3; ++q) Pn[9 + q] = P[9 + q] + x[3 + q];
      } else {
        for (int q = 0; q < 12; ++q) Pn[q] = P[q];
      }
    }
    __syncthreads();
    // candidate cost
    double c[1] = {0.0};
    const double* Pn = s_pose[cur ^ 1];
    for (int m = m0 + tid; m < m1; m += kPoseThreads) {
      float2 uv = __ldg(reinterpret_cast<const float2*>(w.uv) + m);
      double ru, rv;
      residual_only(w.K, Pn, Pn + 9, (double)__ldg(w.X + 3 * m), (double)__ldg(w.X + 3 * m + 1),
                    (double)__ldg(w.X + 3 * m + 2), (double)uv.x, (double)uv.y, ru, rv);
      c[0] += huber_rho(ru, delta) + huber_rho(rv, delta);
    }
    block_sum<1>(c, s_red, s_sum);
    if (tid == 0) {
      double cand = s_sum[0];
      bool ok = (s_fail == 0) && isfinite(cand) && (cand < s_cost * (1.0 - w.accept_rtol));
      if (ok) {
        s_cur = cur ^ 1;
        s_cost = cand;
        s_lambda = fmax(s_lambda * w.lambda_down, w.lambda_min);
      } else {
        s_lambda = fmin(s_lambda * w.lambda_up, w.lambda_max);
      }
      if (w.cost_hist) w.cost_hist[(size_t)f * (w.max_iter + 1) + it + 1] = s_cost;
      if (w.accepted) w.accepted[(size_t)f * w.max_iter + it] = ok ? 1 : 0;
    }
    __syncthreads();
  }

  const double* P = s_pose[s_cur];
  if (tid < 9) w.R[9 * f + tid] = P[tid];
  else if (tid < 12) w.t[3 * f + tid - 9] = P[tid];
  if (w.flags) {
    for (int m = m0 + tid; m < m1; m += kPoseThreads) {
      float2 uv = __ldg(reinterpret_cast<const float2*>(w.uv) + m);
      double ru, rv;
      residual_only(w.K, P, P + 9, (double)__ldg(w.X + 3 * m), (double)__ldg(w.X + 3 * m + 1),
                    (double)__ldg(w.X + 3 * m + 2), (double)uv.x, (double)uv.y, ru, rv);
      w.flags[m] = (ru * ru + rv * rv) > w.chi2_thr ? 1 : 0;
    }
  }
}

int launch_pose_only(const PoseWork& w, cudaStream_t s) {
  if (w.n_frames <= 0) return 0;
  k_pose_only<<<w.n_frames, kPoseThreads, 0, s>>>(w);
  return 1;
}

}  // namespace ba
