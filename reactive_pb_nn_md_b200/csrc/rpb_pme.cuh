// Warp-level PME building blocks shared by the principal-diabat kernels (kernels_pme.cu) and the
// per-diabat delta kernels (kernels_evb.cu).
#pragma once
#include "rpb_dev.cuh"

// B6 weights of one atom: lane l<18 holds w(dim=l/6, k=l%6); returns nearpt via np[]
__device__ __forceinline__ double spline_weight_lane(const Dev& d, const double u[3], int np[3], int lane) {
  np[0] = (int)floor(u[0]); np[1] = (int)floor(u[1]); np[2] = (int)floor(u[2]);
  double w = 0.0;
  if (lane < 18) {
    int dim = lane / 6, k = lane - 6 * dim;
    double arg = u[dim] - (double)(np[dim] - k);
    int idx = (int)ceil(arg / 6.0 * d.spline_grid);
    w = __ldg(&d.B6[idx - 1]);
  }
  return w;
}

// Spread one atom's charge over its 6^3 stencil (one full warp; fp64 RED atomics); sign=-1 subtracts.
// grid_Q / modify_Q_grid  pme.f90:218-259, 295-334: Q(n) +-= ((q*B6_1)*B6_2)*B6_3
__device__ __forceinline__ void spread_atom_warp(const Dev& d, double* Q, const double u[3], double q, double sign, int lane) {
  int np[3];
  double wt = spline_weight_lane(d, u, np, lane);
  int K = d.K;
#pragma unroll 1
  for (int it = 0; it < 7; it++) {
    int p = it * 32 + lane;
    int pp = p < 216 ? p : 215;
    int k1 = pp % 6, k2 = (pp / 6) % 6, k3 = pp / 36;
    double w1 = __shfl_sync(0xffffffffu, wt, k1);
    double w2 = __shfl_sync(0xffffffffu, wt, 6 + k2);
    double w3 = __shfl_sync(0xffffffffu, wt, 12 + k3);
    if (p < 216) {
      int n1 = np[0] - k1; if (n1 < 0) n1 += K;
      int n2 = np[1] - k2; if (n2 < 0) n2 += K;
      int n3 = np[2] - k3; if (n3 < 0) n3 += K;
      double val = q * w1 * w2 * w3;
      atomicAdd(&Q[(size_t)n1 + (size_t)K * n2 + (size_t)K * K * n3], sign * val);
    }
  }
}

// create_scaled_direct_coordinates general_routines.f90:497-524 for one atom (orthorhombic box)
__device__ __forceinline__ void scaled_coords(const Dev& d, const double x[3], double u[3]) {
  double K = (double)d.K;
#pragma unroll
  for (int l = 0; l < 3; l++) {
    double v = K * (d.kk[l] * x[l]);
    if (v < 0.0) v = v + K; else if (v >= K) v = v - K;
    if (fabs(fmod(v, 1.0)) < 1e-6) v = v + 1e-6;
    u[l] = v;
  }
}

// Force interpolation for one atom by one warp.  u: scaled coords, q: charge, theta: K^3 grid.
// Returns the Cartesian force in lane 0 (all lanes get the reduced value via xor-shuffle).
__device__ __forceinline__ void gather_atom_warp(const Dev& d, const double* __restrict__ theta, const double u[3], double q,
                                                 int lane, double F[3]) {
  int np[3] = {(int)floor(u[0]), (int)floor(u[1]), (int)floor(u[2])};
  double b6 = 0.0, dm = 0.0;
  if (lane < 18) {
    int dim = lane / 6, k = lane - 6 * dim;
    double arg1 = u[dim] - (double)(np[dim] - k);
    double arg2 = arg1 - 1.0;
    int g1n = (int)ceil(arg1 / 6.0 * d.spline_grid);
    b6 = __ldg(&d.B6[g1n - 1]);
    if (arg1 < 5.0) { int g = (int)ceil(arg1 / 5.0 * d.spline_grid); dm = __ldg(&d.B5[g - 1]); }
    if (0.0 < arg2) { int g = (int)ceil(arg2 / 5.0 * d.spline_grid); dm = dm - __ldg(&d.B5[g - 1]); }
  }
  int K = d.K;
  double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll 1
  for (int it = 0; it < 7; it++) {
    int p = it * 32 + lane;
    int pp = p < 216 ? p : 215;
    int k1 = pp % 6, k2 = (pp / 6) % 6, k3 = pp / 36;
    double w1 = __shfl_sync(0xffffffffu, b6, k1), w2 = __shfl_sync(0xffffffffu, b6, 6 + k2), w3 = __shfl_sync(0xffffffffu, b6, 12 + k3);
    double d1 = __shfl_sync(0xffffffffu, dm, k1), d2 = __shfl_sync(0xffffffffu, dm, 6 + k2), d3 = __shfl_sync(0xffffffffu, dm, 12 + k3);
    if (p < 216) {
      int n1 = np[0] - k1; if (n1 < 0) n1 += K;
      int n2 = np[1] - k2; if (n2 < 0) n2 += K;
      int n3 = np[2] - k3; if (n3 < 0) n3 += K;
      double th = __ldg(&theta[(size_t)n1 + (size_t)K * n2 + (size_t)K * K * n3]) * d.conv;
      f0 = fma(q * (d1 * w2 * w3), th, f0);
      f1 = fma(q * (d2 * w1 * w3), th, f1);
      f2 = fma(q * (d3 * w1 * w2), th, f2);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    f0 += __shfl_xor_sync(0xffffffffu, f0, o);
    f1 += __shfl_xor_sync(0xffffffffu, f1, o);
    f2 += __shfl_xor_sync(0xffffffffu, f2, o);
  }
  double Kd = (double)K;
  F[0] = -(Kd * d.kk[0]) * f0; F[1] = -(Kd * d.kk[1]) * f1; F[2] = -(Kd * d.kk[2]) * f2;
}

