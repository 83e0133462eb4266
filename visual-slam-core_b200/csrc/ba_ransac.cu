// Pose without a prior for the tracking call site: RANSAC over minimal three-point samples, one CTA per hypothesis.
//
// The reference estimates the current frame's pose with cv2.solvePnPRansac and NO initial guess
// (visual_slam/utils/motion_estimation.py:107-116, called from visual_slam/tracking.py:350-356); the batched pose-only
// LM (ba_pose.cu) needs a starting pose.  This kernel supplies it the way RANSAC does: each hypothesis takes three
// matches, solves the perspective-three-point problem (Grunert's quartic in v = s3 / s1, Haralick et al. 1994, eq. 6-9),
// turns every admissible root into a pose by aligning the three camera-frame points with their world points, and counts
// the matches of the WHOLE frame that reproject within the threshold -- the consensus OpenCV forms.  The best hypothesis
// is the LM's initial pose; the LM (Huber, all matches) then does what OpenCV's final solvePnP on the inliers does.
//
// Roots: the quartic is sampled on a logarithmic grid of depth ratios (2^-5 .. 2^5, 64 intervals, two per lane of warp 0)
// and every sign change is refined by bisection -- branch-free, no closed-form quartic, double roots (tangent cases) are
// missed, which costs a hypothesis, not correctness.
#include "ba_kernels.cuh"

namespace ba {

namespace {

constexpr int kRT = 128;                 // threads per hypothesis
constexpr int kOutD = 16;                // doubles per hypothesis: inliers | R (9) | t (3) | pad

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

struct P3P {
  double A4, A3, A2, A1, A0, q, ca, cb, cg, c2;
};
__device__ __forceinline__ double quartic(const P3P& p, double v) {
  return (((p.A4 * v + p.A3) * v + p.A2) * v + p.A1) * v + p.A0;
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ bool unit3(double a[3]) {
  const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  if (!(n > 1e-12)) return false;
  a[0] /= n; a[1] /= n; a[2] /= n;
  return true;
}

__global__ void __launch_bounds__(kRT)
k_p3p_ransac(RansacWork w) {
  __shared__ double s_root[8];
  __shared__ int s_nroot;
  __shared__ double s_P[3][3], s_f[3][3];
  __shared__ P3P s_p;
  __shared__ int s_ok;
  __shared__ int s_cnt[kRT / 32];
  __shared__ double s_best[kOutD];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int hyp = blockIdx.x;
  if (tid == 0) {
    s_nroot = 0;
    s_ok = 0;
    s_best[0] = -1.0;
    // three distinct matches
    int idx[3];
    unsigned st = hash32(w.seed * 0x9e3779b9u + (unsigned)hyp * 0x85ebca6bu + 1u);
    for (int k = 0; k < 3; ++k) {
      for (int tries = 0; tries < 16; ++tries) {
        st = hash32(st + 0x632be5abu);
        idx[k] = (int)(st % (unsigned)w.n_match);
        bool dup = false;
        for (int e = 0; e < k; ++e) dup |= idx[e] == idx[k];
        if (!dup) break;
      }
    }
    bool ok = idx[0] != idx[1] && idx[0] != idx[2] && idx[1] != idx[2];
    for (int k = 0; k < 3; ++k) {
      for (int d = 0; d < 3; ++d) s_P[k][d] = (double)w.X[3 * (size_t)idx[k] + d];
      double f[3] = {((double)w.uv[2 * (size_t)idx[k]] - w.K.cx) / w.K.fx, ((double)w.uv[2 * (size_t)idx[k] + 1] - w.K.cy) / w.K.fy, 1.0};
      ok = unit3(f) && ok;
      for (int d = 0; d < 3; ++d) s_f[k][d] = f[d];
    }
    auto dist2 = [&](int i, int j) {
      double s = 0.0;
      for (int d = 0; d < 3; ++d) { const double e = s_P[i][d] - s_P[j][d]; s += e * e; }
      return s;
    };
    const double a2 = dist2(1, 2), b2 = dist2(0, 2), c2 = dist2(0, 1);
    ok = ok && a2 > 1e-12 && b2 > 1e-12 && c2 > 1e-12;
    if (ok) {
      P3P p;
      auto dot = [&](int i, int j) { return s_f[i][0] * s_f[j][0] + s_f[i][1] * s_f[j][1] + s_f[i][2] * s_f[j][2]; };
      p.ca = dot(1, 2); p.cb = dot(0, 2); p.cg = dot(0, 1);
      p.q = (a2 - c2) / b2;
      const double p2 = (a2 + c2) / b2, ca2 = p.ca * p.ca, cg2 = p.cg * p.cg;
      p.A4 = (p.q - 1.0) * (p.q - 1.0) - 4.0 * c2 / b2 * ca2;
      p.A3 = 4.0 * (p.q * (1.0 - p.q) * p.cb - (1.0 - p2) * p.ca * p.cg + 2.0 * c2 / b2 * ca2 * p.cb);
      p.A2 = 2.0 * (p.q * p.q - 1.0 + 2.0 * p.q * p.q * p.cb * p.cb + 2.0 * (b2 - c2) / b2 * ca2 - 4.0 * p2 * p.ca * p.cb * p.cg +
                    2.0 * (b2 - a2) / b2 * cg2);
      p.A1 = 4.0 * (-p.q * (1.0 + p.q) * p.cb + 2.0 * a2 / b2 * cg2 * p.cb - (1.0 - p2) * p.ca * p.cg);
      p.A0 = (1.0 + p.q) * (1.0 + p.q) - 4.0 * a2 / b2 * cg2;
      p.c2 = c2;
      s_p = p;
      s_ok = 1;
    }
  }
  __syncthreads();
  if (s_ok && wid == 0) {
    // two grid intervals per lane; a sign change is bisected to float64 resolution
    const P3P p = s_p;
    for (int half = 0; half < 2; ++half) {
      const int k = 2 * lane + half;
      double lo = exp2(-5.0 + 10.0 * k / 64.0), hi = exp2(-5.0 + 10.0 * (k + 1) / 64.0);
      double flo = quartic(p, lo);
      const double fhi = quartic(p, hi);
      if ((flo <= 0.0) != (fhi <= 0.0)) {
        for (int it = 0; it < 60; ++it) {
          const double mid = 0.5 * (lo + hi), fm = quartic(p, mid);
          if ((fm <= 0.0) == (flo <= 0.0)) { lo = mid; flo = fm; } else { hi = mid; }
        }
        const int at = atomicAdd(&s_nroot, 1);
        if (at < 8) s_root[at] = 0.5 * (lo + hi);
      }
    }
  }
  __syncthreads();
  const int nroot = s_ok ? min(s_nroot, 8) : 0;
  for (int r = 0; r < nroot; ++r) {
    // pose of this root (every thread computes it: cheaper than a broadcast)
    const P3P p = s_p;
    const double v = s_root[r];
    const double den = 2.0 * (p.cg - v * p.ca);
    double R[9], t[3];
    bool ok = fabs(den) > 1e-12;
    double u = 0.0, s1 = 0.0;
    if (ok) {
      u = ((p.q - 1.0) * v * v - 2.0 * p.q * p.cb * v + 1.0 + p.q) / den;
      const double d1 = 1.0 + u * u - 2.0 * u * p.cg;
      ok = u > 0.0 && d1 > 1e-14;
      if (ok) s1 = sqrt(p.c2 / d1);
    }
    if (ok) {
      const double s[3] = {s1, u * s1, v * s1};
      double Q[3][3];
      for (int k = 0; k < 3; ++k)
        for (int d = 0; d < 3; ++d) Q[k][d] = s[k] * s_f[k][d];
      // orthonormal frames of the triangle in world and camera coordinates
      double ew1[3], ew2[3], ew3[3], ec1[3], ec2[3], ec3[3], tmp[3];
      for (int d = 0; d < 3; ++d) { ew1[d] = s_P[1][d] - s_P[0][d]; tmp[d] = s_P[2][d] - s_P[0][d]; }
      cross3(ew1, tmp, ew3);
      ok = unit3(ew1) && unit3(ew3);
      cross3(ew3, ew1, ew2);
      for (int d = 0; d < 3; ++d) { ec1[d] = Q[1][d] - Q[0][d]; tmp[d] = Q[2][d] - Q[0][d]; }
      cross3(ec1, tmp, ec3);
      ok = unit3(ec1) && unit3(ec3) && ok;
      cross3(ec3, ec1, ec2);
      if (ok) {
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) R[3 * i + j] = ec1[i] * ew1[j] + ec2[i] * ew2[j] + ec3[i] * ew3[j];
        for (int i = 0; i < 3; ++i)
          t[i] = Q[0][i] - (R[3 * i] * s_P[0][0] + R[3 * i + 1] * s_P[0][1] + R[3 * i + 2] * s_P[0][2]);
      }
    }
    // consensus over the whole frame (uniform branch: every thread computed the same `ok`)
    int cnt = 0;
    if (ok) {
      for (int m = tid; m < w.n_match; m += kRT) {
        const double X0 = (double)w.X[3 * (size_t)m], X1 = (double)w.X[3 * (size_t)m + 1], X2 = (double)w.X[3 * (size_t)m + 2];
        const double xc = R[0] * X0 + R[1] * X1 + R[2] * X2 + t[0];
        const double yc = R[3] * X0 + R[4] * X1 + R[5] * X2 + t[1];
        const double zc = R[6] * X0 + R[7] * X1 + R[8] * X2 + t[2];
        if (zc > 1e-9) {
          const double du = w.K.fx * xc / zc + w.K.cx - (double)w.uv[2 * (size_t)m];
          const double dv = w.K.fy * yc / zc + w.K.cy - (double)w.uv[2 * (size_t)m + 1];
          cnt += (du * du + dv * dv <= w.thr2) ? 1 : 0;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if (lane == 0) s_cnt[wid] = cnt;
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int i = 0; i < kRT / 32; ++i) total += s_cnt[i];
      if (ok && (double)total > s_best[0]) {
        s_best[0] = (double)total;
        for (int i = 0; i < 9; ++i) s_best[1 + i] = R[i];
        for (int i = 0; i < 3; ++i) s_best[10 + i] = t[i];
      }
    }
    __syncthreads();
  }
  if (tid < kOutD) w.out[(size_t)hyp * kOutD + tid] = (tid < 13) ? s_best[tid] : 0.0;
}

}  // namespace

int ransac_out_doubles() { return kOutD; }

int launch_pose_ransac(const RansacWork& w, cudaStream_t s) {
  k_p3p_ransac<<<w.n_hyp, kRT, 0, s>>>(w);
  return 1;
}

}  // namespace ba
