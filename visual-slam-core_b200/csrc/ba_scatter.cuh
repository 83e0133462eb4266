// Warp-per-point scatter routine (float64 RED atomics into the reduced camera matrix), shared by
// the scatter kernel and by the tiled kernel's fallback for chunks whose cameras do not fit its
// local window.  Works for any camera count and any number of observations per point.
#pragma once
#include "ba_kernels.cuh"

namespace ba {

struct ObsIn {
  int cam;
  double u, v;
};

__device__ __forceinline__ ObsIn load_obs(const Work& w, int o) {
  ObsIn r;
  r.cam = __ldg(w.obs_cam + o);
  float2 uv = __ldg(reinterpret_cast<const float2*>(w.obs_uv) + o);
  r.u = (double)uv.x;
  r.v = (double)uv.y;
  return r;
}

__device__ __forceinline__ void make_W(const ObsLin& o, double W[18]) {
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double a = o.wu * o.pu[r], b = o.wv * o.pv[r];
#pragma unroll
    for (int q = 0; q < 3; ++q) W[3 * r + q] = a * o.lu[q] + b * o.lv[q];
  }
}

// One point, one warp.  sW: 32 x 19 doubles and sCam: 32 ints of warp-private shared memory.
__device__ __forceinline__ void scatter_point(const Work& w, int j, int lane, double lambda, double delta,
                                              const double* __restrict__ pose, const double* __restrict__ X,
                                              double (*sW)[19], int* sCam, double& cost) {
  const int o0 = __ldg(w.pt_off + j);
  const int k = __ldg(w.pt_off + j + 1) - o0;
  if (k <= 0) return;
  const double X0 = X[3 * (size_t)j], X1 = X[3 * (size_t)j + 1], X2 = X[3 * (size_t)j + 2];

  // ---- pass 1: V_j, g_l,j, cost ---------------------------------------------------------
  double V[6] = {0, 0, 0, 0, 0, 0}, gl[3] = {0, 0, 0};
  for (int base = 0; base < k; base += 32) {
    int i = base + lane;
    if (i < k) {
      ObsIn in = load_obs(w, o0 + i);
      ObsLin ol;
      linearize_obs<false>(w.K, pose + 12 * (size_t)in.cam, pose + 12 * (size_t)in.cam + 9, X0, X1, X2,
                           in.u, in.v, delta, ol);
      V[0] += ol.wu * ol.lu[0] * ol.lu[0] + ol.wv * ol.lv[0] * ol.lv[0];
      V[1] += ol.wu * ol.lu[0] * ol.lu[1] + ol.wv * ol.lv[0] * ol.lv[1];
      V[2] += ol.wu * ol.lu[0] * ol.lu[2] + ol.wv * ol.lv[0] * ol.lv[2];
      V[3] += ol.wu * ol.lu[1] * ol.lu[1] + ol.wv * ol.lv[1] * ol.lv[1];
      V[4] += ol.wu * ol.lu[1] * ol.lu[2] + ol.wv * ol.lv[1] * ol.lv[2];
      V[5] += ol.wu * ol.lu[2] * ol.lu[2] + ol.wv * ol.lv[2] * ol.lv[2];
      double au = ol.wu * ol.ru, av = ol.wv * ol.rv;
#pragma unroll
      for (int q = 0; q < 3; ++q) gl[q] += au * ol.lu[q] + av * ol.lv[q];
      cost += huber_rho(ol.ru, delta) + huber_rho(ol.rv, delta);
    }
  }
  warp_sum_6_3(V, gl);
  V[0] += damp(V[0], lambda);
  V[3] += damp(V[3], lambda);
  V[5] += damp(V[5], lambda);
  double Vi[6];
  sym3_inverse(V, Vi);

  // ---- pass 2: pose blocks, Schur complement --------------------------------------------
  for (int baseA = 0; baseA < k; baseA += 32) {
    const int iA = baseA + lane;
    const bool validA = iA < k;
    int camA = -1;
    double WA[18], T[18];
    bool activeA = false;
    if (validA) {
      ObsIn in = load_obs(w, o0 + iA);
      camA = in.cam;
      activeA = (w.fixed[camA] == 0);
      ObsLin ol;
      linearize_obs<true>(w.K, pose + 12 * (size_t)camA, pose + 12 * (size_t)camA + 9, X0, X1, X2, in.u,
                          in.v, delta, ol);
      make_W(ol, WA);
      // T = W V^-1  (6x3)
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double w0 = WA[3 * r], w1 = WA[3 * r + 1], w2 = WA[3 * r + 2];
        T[3 * r + 0] = w0 * Vi[0] + w1 * Vi[1] + w2 * Vi[2];
        T[3 * r + 1] = w0 * Vi[1] + w1 * Vi[3] + w2 * Vi[4];
        T[3 * r + 2] = w0 * Vi[2] + w1 * Vi[4] + w2 * Vi[5];
      }
      if (activeA) {
        double* Sd = s_block_ptr(w, camA, camA);
        double au = ol.wu * ol.ru, av = ol.wv * ol.rv;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          double pur = ol.wu * ol.pu[r], pvr = ol.wv * ol.pv[r];
#pragma unroll
          for (int c = 0; c <= r; ++c) {
            double u_rc = pur * ol.pu[c] + pvr * ol.pv[c];
            double s_rc = T[3 * r] * WA[3 * c] + T[3 * r + 1] * WA[3 * c + 1] +
                          T[3 * r + 2] * WA[3 * c + 2];
            atomicAdd(Sd + 6 * r + c, u_rc - s_rc);
          }
          atomicAdd(w.diagU + 6 * (size_t)camA + r, pur * ol.pu[r] + pvr * ol.pv[r]);
          double br = -(au * ol.pu[r] + av * ol.pv[r]) + T[3 * r] * gl[0] +
                      T[3 * r + 1] * gl[1] + T[3 * r + 2] * gl[2];
          atomicAdd(w.b + 6 * (size_t)camA + r, br);
        }
      }
    }
    for (int baseB = 0; baseB < k; baseB += 32) {
      const int iB = baseB + lane;
      // stage W of chunk B
      if (baseB == baseA) {
        sCam[lane] = activeA ? camA : -1;
        if (activeA) {
#pragma unroll
          for (int q = 0; q < 18; ++q) sW[lane][q] = WA[q];
        }
      } else {
        int camB = -1;
        if (iB < k) {
          ObsIn in = load_obs(w, o0 + iB);
          if (w.fixed[in.cam] == 0) {
            camB = in.cam;
            ObsLin ol;
            linearize_obs<true>(w.K, pose + 12 * (size_t)camB, pose + 12 * (size_t)camB + 9, X0, X1, X2,
                                in.u, in.v, delta, ol);
            double WB[18];
            make_W(ol, WB);
#pragma unroll
            for (int q = 0; q < 18; ++q) sW[lane][q] = WB[q];
          }
        }
        sCam[lane] = camB;
      }
      __syncwarp();
      const int mB = min(32, k - baseB);
      if (activeA) {
        for (int m = 0; m < mB; ++m) {
          const int camB = sCam[m];
          if (camB < 0 || camB > camA || baseB + m == iA) continue;
          const double* Wb = sW[m];
          double* Sd = s_block_ptr(w, camA, camB);
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double b0 = Wb[3 * c], b1 = Wb[3 * c + 1], b2 = Wb[3 * c + 2];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              double s_rc = T[3 * r] * b0 + T[3 * r + 1] * b1 + T[3 * r + 2] * b2;
              atomicAdd(Sd + 6 * r + c, -s_rc);
            }
          }
        }
      }
      __syncwarp();
    }
  }
}

}  // namespace ba
