// Device-side building blocks shared by every kernel of the local-BA path.
//
// Math follows SURVEY.md Appendix B and the reference objective
// (visual_slam/optimization/torch_optimizer.py:152-177): Xc = R X + t, Z' = Xc_z + 1e-8,
// r = (fx Xc_x/Z' + cx - u, fy Xc_y/Z' + cy - v), per-component Huber.
// All arithmetic is float64 (B200 runs FP64 at half the FP32 rate; the path is bound by
// memory and latency, and bit-exact outlier flags need float64 residuals).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ba {

constexpr double kDepthEps = 1e-8;   // torch_optimizer.py:165
constexpr unsigned kFull = 0xffffffffu;

struct Intr { double fx, fy, cx, cy; };

// LM control block: lives in device memory so that the whole LM loop (linearise, solve,
// back-substitute, accept/reject) runs without a host round trip and can be graph-captured.
struct LmCtl {
  double lambda;
  double cost_cur;
  double cost_lin;      // accumulator: cost at the linearisation point
  double cost_cand;     // accumulator: cost at the candidate
  int cur;              // which of the two state buffers holds the current iterate
  int iter;
  int solve_fail;       // set by the solver for this iteration
  int n_accepted;
  unsigned ticket;      // last-block-done counter
  int pcg_iters_last;
  int pad0, pad1;
  // parameters
  double huber_delta, chi2_thr, lambda_up, lambda_down, lambda_min, lambda_max, accept_rtol;
  // history
  double cost_hist[65];
  double cand_hist[64];
  double lambda_hist[64];
  int solve_fail_hist[64];
  int pcg_iters_hist[64];
  unsigned char accepted[64];
  long long epoch_base;   // sharded runs with the peer-memory exchange: 64 * (solves so far), set by k_init
};

__device__ __forceinline__ double huber_rho(double r, double delta) {
  double a = fabs(r);
  return a < delta ? 0.5 * r * r : delta * (a - 0.5 * delta);
}
// Reciprocal without the division routine (25 instructions, a slow-path call and its convergence barriers per
// division; every observation costs three -- depth and two Huber weights -- in each of the four passes that linearise
// it): rcp.approx + two Newton steps, within an ulp of the quotient for the normal numbers met here.  The panel kernel's
// staging loop went from 0.951 to 0.923 ms with it; S, b, diag U of a linearisation agree with the oracle as before
// (3e-15 .. 1e-14 relative).
__device__ __forceinline__ double fast_rcp(double a) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}
__device__ __forceinline__ double huber_w(double r, double delta) {
  double a = fabs(r);
  return a < delta ? 1.0 : delta * fast_rcp(a);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// butterfly sums of a symmetric 3x3 (6) and a 3-vector side by side: nine independent
// shuffle/add chains per step instead of nine dependent five-step chains one after the other
__device__ __forceinline__ void warp_sum_6_3(double (&V)[6], double (&g)[3]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
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

// sin(x)/x with a Taylor branch below 1e-4 (same as oracle/lm_oracle.py:sinc)
__device__ __forceinline__ double sinc(double x) {
  double x2 = x * x;
  if (x < 1e-4) return 1.0 - x2 / 6.0 + x2 * x2 / 120.0;
  return sin(x) / x;
}

// R_out = exp([w]x) * R   (left increment, torch_optimizer.py:159)
__device__ __forceinline__ void so3_exp_mul(const double w[3], const double R[9], double Ro[9]) {
  double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double A = sinc(th);
  double h = sinc(0.5 * th);
  double B = 0.5 * h * h;
  double wx = w[0], wy = w[1], wz = w[2];
  double E[9];
  E[0] = 1.0 - B * (wy * wy + wz * wz);
  E[4] = 1.0 - B * (wx * wx + wz * wz);
  E[8] = 1.0 - B * (wx * wx + wy * wy);
  E[1] = B * wx * wy - A * wz;
  E[3] = B * wx * wy + A * wz;
  E[2] = B * wx * wz + A * wy;
  E[6] = B * wx * wz - A * wy;
  E[5] = B * wy * wz - A * wx;
  E[7] = B * wy * wz + A * wx;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      Ro[3 * i + j] = E[3 * i] * R[j] + E[3 * i + 1] * R[3 + j] + E[3 * i + 2] * R[6 + j];
}

// Inverse of a symmetric 3x3 given as a,b,c,d,e,f = (00,01,02,11,12,22) by explicit LDL^T.
// Same operation order as oracle/lm_oracle.py:sym3_inverse.  Non-positive pivot -> zero, false.
__device__ __forceinline__ bool sym3_inverse(const double V[6], double inv[6]) {
  double a = V[0], b = V[1], c = V[2], d = V[3], e = V[4], f = V[5];
  double i0 = fast_rcp(a);
  double l10 = b * i0, l20 = c * i0;
  double d1 = d - l10 * b;
  double e1 = e - l20 * b;
  double i1 = fast_rcp(d1);
  double l21 = e1 * i1;
  double d2 = f - l20 * c - l21 * e1;
  double i2 = fast_rcp(d2);
  double m10 = -l10, m20 = l10 * l21 - l20, m21 = -l21;
  inv[0] = i0 + m10 * m10 * i1 + m20 * m20 * i2;
  inv[1] = m10 * i1 + m20 * m21 * i2;
  inv[2] = m20 * i2;
  inv[3] = i1 + m21 * m21 * i2;
  inv[4] = m21 * i2;
  inv[5] = i2;
  bool ok = (a > 0.0) && (d1 > 0.0) && (d2 > 0.0);
#pragma unroll
  for (int k = 0; k < 6; ++k) ok = ok && isfinite(inv[k]);
  if (!ok) {
#pragma unroll
    for (int k = 0; k < 6; ++k) inv[k] = 0.0;
  }
  return ok;
}

__device__ __forceinline__ double damp(double d, double lambda) {
  return lambda * (d > 0.0 ? d : 1.0);
}

// One observation linearised: residual, IRLS weights, pose rows (dw,dt) and point rows.
struct ObsLin {
  double ru, rv, wu, wv;
  double pu[6], pv[6];   // d r_u / d(dw,dt), d r_v / d(dw,dt)
  double lu[3], lv[3];   // d r_u / dX,       d r_v / dX
};

__device__ __forceinline__ void residual_only(const Intr& K, const double* __restrict__ R,
                                              const double* __restrict__ t, double X0, double X1,
                                              double X2, double u, double v, double& ru, double& rv) {
  double Y0 = R[0] * X0 + R[1] * X1 + R[2] * X2;
  double Y1 = R[3] * X0 + R[4] * X1 + R[5] * X2;
  double Y2 = R[6] * X0 + R[7] * X1 + R[8] * X2;
  double xc = Y0 + t[0], yc = Y1 + t[1], zc = Y2 + t[2];
  double iz = fast_rcp(zc + kDepthEps);
  ru = K.fx * (xc * iz) + K.cx - u;
  rv = K.fy * (yc * iz) + K.cy - v;
}

template <bool kPose, bool kPoint = true>
__device__ __forceinline__ void linearize_obs(const Intr& K, const double* __restrict__ R,
                                              const double* __restrict__ t, double X0, double X1,
                                              double X2, double u, double v, double delta,
                                              ObsLin& o) {
  double Y0 = R[0] * X0 + R[1] * X1 + R[2] * X2;
  double Y1 = R[3] * X0 + R[4] * X1 + R[5] * X2;
  double Y2 = R[6] * X0 + R[7] * X1 + R[8] * X2;
  double xc = Y0 + t[0], yc = Y1 + t[1], zc = Y2 + t[2];
  double iz = fast_rcp(zc + kDepthEps);
  double xz = xc * iz, yz = yc * iz;
  o.ru = K.fx * xz + K.cx - u;
  o.rv = K.fy * yz + K.cy - v;
  o.wu = huber_w(o.ru, delta);
  o.wv = huber_w(o.rv, delta);
  double a = K.fx * iz, c = -K.fx * xz * iz;
  double b = K.fy * iz, d = -K.fy * yz * iz;
  if (kPose) {
    // J_pi * (-[Y]x), J_pi
    o.pu[0] = c * Y1;            o.pu[1] = a * Y2 - c * Y0;  o.pu[2] = -a * Y1;
    o.pu[3] = a;                 o.pu[4] = 0.0;              o.pu[5] = c;
    o.pv[0] = -b * Y2 + d * Y1;  o.pv[1] = -d * Y0;          o.pv[2] = b * Y0;
    o.pv[3] = 0.0;               o.pv[4] = b;                o.pv[5] = d;
  }
  if (kPoint) {
    // J_pi * R
    o.lu[0] = a * R[0] + c * R[6];  o.lu[1] = a * R[1] + c * R[7];  o.lu[2] = a * R[2] + c * R[8];
    o.lv[0] = b * R[3] + d * R[6];  o.lv[1] = b * R[4] + d * R[7];  o.lv[2] = b * R[5] + d * R[8];
  }
}

// accept / reject bookkeeping of one LM iteration (single thread, the tail of every iteration).  Everything it
// needs is read up front (lm_load, which a kernel can issue before it knows that it is the one to decide): the
// stores of lm_accept go to the same struct, so loads placed after them would each wait for their own L2 round
// trip.
struct LmIn {
  int iter, solve_fail, pcg_iters_last, n_accepted, cur;
  double cost_cur, lambda, accept_rtol, lambda_up, lambda_down, lambda_min, lambda_max;
};
__device__ __forceinline__ LmIn lm_load(const LmCtl* c) {
  LmIn in;
  in.iter = c->iter; in.solve_fail = c->solve_fail; in.pcg_iters_last = c->pcg_iters_last;
  in.n_accepted = c->n_accepted; in.cur = c->cur;
  in.cost_cur = c->cost_cur; in.lambda = c->lambda; in.accept_rtol = c->accept_rtol;
  in.lambda_up = c->lambda_up; in.lambda_down = c->lambda_down;
  in.lambda_min = c->lambda_min; in.lambda_max = c->lambda_max;
  return in;
}
__device__ __forceinline__ void lm_accept(LmCtl* c, const LmIn& in, double cand) {
  const int it = in.iter;
  const bool ok = (in.solve_fail == 0) && isfinite(cand) && (cand < in.cost_cur * (1.0 - in.accept_rtol));
  if (it < 64) {
    c->lambda_hist[it] = in.lambda;
    c->cand_hist[it] = cand;
    c->accepted[it] = ok ? 1 : 0;
    c->solve_fail_hist[it] = in.solve_fail;
    c->pcg_iters_hist[it] = in.pcg_iters_last;
  }
  if (ok) {
    c->cur = in.cur ^ 1;
    c->cost_cur = cand;
    c->n_accepted = in.n_accepted + 1;
    c->lambda = fmax(in.lambda * in.lambda_down, in.lambda_min);
  } else {
    c->lambda = fmin(in.lambda * in.lambda_up, in.lambda_max);
  }
  if (it + 1 < 65) c->cost_hist[it + 1] = ok ? cand : in.cost_cur;
  c->iter = it + 1;
  c->cost_cand = 0.0;
  c->cost_lin = 0.0;
  c->solve_fail = 0;
  c->pcg_iters_last = 0;
  c->ticket = 0u;
}

}  // namespace ba
