// Parameter-function PCA pre-transform ("parameterTrafoPCA"), fused in front of kernel (a).
//
// Reference: inside Emulator.predict, src/emulator.py:492-551 (and the identical copy in
// src/emulator_BAND.py:393-452): three groups of model parameters are replaced by the principal
// components of the 100-point curves they parametrise,
//     zeta/s(T)      columns [15,16,17,18]   T    in linspace(0, 0.5, 100)   src/emulator.py:100-106
//     eta/s(mu_B)    columns [12,13,14]      mu_B in linspace(0, 0.6, 100)   src/emulator.py:109-115
//     y_loss(y_init) columns [2,3,4]         y    in linspace(0, 6.2, 100)   src/emulator.py:118-124
// each curve going through StandardScaler.transform and PCA.transform; the new parameter vector is
// the untouched columns followed by the bulk, shear and y_loss components.  The reference does this
// with a Python triple loop over samples x 100 grid points.
//
// Scaler and PCA are folded on the host into one affine map per group,
//     PC_j = sum_t f_t * Wt[j][t] + b_j ,   Wt[j][t] = comp[j][t] / scale[t],
//     b_j  = - sum_t (mean[t] / scale[t] + pca_mean[t]) * comp[j][t],
// so a walker costs 300 curve evaluations and 300 * (k1 + k2 + k3) FMAs.  One warp per walker.
#pragma once
#include "common.cuh"

namespace gpbt {

constexpr int kPtMaxGroups = 3;
constexpr int kPtMaxArgs = 4;

struct ParamTrafoGroup {
  int kind;                  // 0 zeta/s(T), 1 eta/s(mu_B), 2 y_loss(y_init)
  int nargs;                 // parameters consumed (4, 3, 3)
  int idx[kPtMaxArgs];       // their columns in X
  int npts;                  // grid points (100)
  int ncomp;                 // principal components kept
  int out_off;               // first output column
  double g0, g1;             // grid = linspace(g0, g1, npts)
  const double* Wt;          // [ncomp][npts] device
  const double* b;           // [ncomp] device
};

struct ParamTrafoParams {
  const double* __restrict__ X;   // [N, p_in]
  double* __restrict__ theta;     // [N, p_out]
  const int* __restrict__ keep;   // [n_keep] columns of X copied through
  int64_t N;
  int p_in, p_out, n_keep, n_groups;
  ParamTrafoGroup grp[kPtMaxGroups];
};

// the three parametrisations, written as the reference evaluates them (strict / non-strict
// comparisons included)
__device__ __forceinline__ double curve_value(int kind, const double* a, double x) {
  if (kind == 0) {  // zeta_max, T_zeta0, sigma_plus, sigma_minus ; mu_B = 0
    const double T0 = a[1];
    const double sig = (x < T0) ? a[3] : a[2];
    const double d = x - T0;
    return a[0] * exp(-(d * d) / (2.0 * (sig * sig)));
  }
  if (kind == 1) {  // eta_0, eta_2, eta_4
    if (0.0 < x && x <= 0.2) return a[0] + (a[1] - a[0]) * (x / 0.2);
    if (0.2 < x && x < 0.4) return a[1] + (a[2] - a[1]) * ((x - 0.2) / 0.2);
    return a[2];
  }
  // yloss_2, yloss_4, yloss_6
  if (0.0 < x && x <= 2.0) return a[0] * (x / 2.0);
  if (2.0 < x && x < 4.0) return a[0] + (a[1] - a[0]) * ((x - 2.0) / 2.0);
  return a[1] + (a[2] - a[1]) * ((x - 4.0) / 2.0);
}

__global__ void __launch_bounds__(128) param_trafo_kernel(const ParamTrafoParams prm) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= prm.N) return;
  const double* x = prm.X + w * prm.p_in;
  double* th = prm.theta + w * prm.p_out;
  for (int c = lane; c < prm.n_keep; c += 32) th[c] = x[prm.keep[c]];
  for (int gi = 0; gi < prm.n_groups; gi++) {
    const ParamTrafoGroup& G = prm.grp[gi];
    double a[kPtMaxArgs];
#pragma unroll
    for (int i = 0; i < kPtMaxArgs; i++) a[i] = i < G.nargs ? x[G.idx[i]] : 0.0;
    // numpy.linspace: start + i * step, last point exactly the stop value
    const double step = (G.g1 - G.g0) / (double)(G.npts - 1);
    for (int j = 0; j < G.ncomp; j++) {
      double s = 0.0;
      for (int tt = lane; tt < G.npts; tt += 32) {
        const double gx = (tt == G.npts - 1) ? G.g1 : G.g0 + (double)tt * step;
        s = fma(curve_value(G.kind, a, gx), G.Wt[(size_t)j * G.npts + tt], s);
      }
      s = warp_sum(s);
      if (lane == 0) th[G.out_off + j] = s + G.b[j];
    }
  }
}

}  // namespace gpbt
