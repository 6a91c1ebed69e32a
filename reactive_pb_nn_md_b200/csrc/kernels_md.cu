// Integrator, centre of mass / wrap, momentum removal, kinetic energy.  Replaces (reference file:line):
//   md_integrate_atomic (NVE)            src/md_integration.f90:469-532
//   subtract_center_of_mass_momentum     src/md_integration.f90:125-177
//   update_r_com / shift_molecules_into_box   src/general_routines.f90:420-440, 1145-1197
//   calculate_kinetic_energy             src/total_energy_forces.f90:106-121
// All O(N), HBM/latency bound.  (Neighbour list: kernels_nlist.cu.)
#include <algorithm>
#include "rpb_host.h"

#define TPB 256

__global__ void k_zero(double* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

// one thread per molecule: pos_com, then shift_molecules_into_box
__global__ void k_com_shift(Dev d, int do_shift) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= d.M) return;
  int f = d.mol_first[m], n = d.mol_natom[m];
  double c0 = 0, c1 = 0, c2 = 0, mt = 0;
  for (int a = 0; a < n; a++) {
    double4 p = d.xq[f + a];
    double ms = d.mass[f + a];
    c0 = c0 + p.x * ms; c1 = c1 + p.y * ms; c2 = c2 + p.z * ms;
    mt = mt + ms;
  }
  double rc[3] = {c0 / mt, c1 / mt, c2 / mt};
  if (do_shift) {
    double t[3];
    bool any = false;
    for (int k = 0; k < 3; k++) {
      double db = d.inv_box[k] * rc[k];
      double sh = 0.0;
      if (db < 0.0) sh = 1.0; else if (db > 1.0) sh = -1.0;
      t[k] = sh * d.box[k];
      any |= (sh != 0.0);
      rc[k] = rc[k] + t[k];
    }
    if (any)
      for (int a = 0; a < n; a++) {
        double4 p = d.xq[f + a];
        p.x = p.x + t[0]; p.y = p.y + t[1]; p.z = p.z + t[2];
        d.xq[f + a] = p;
      }
  }
  d.r_com[3 * m] = rc[0]; d.r_com[3 * m + 1] = rc[1]; d.r_com[3 * m + 2] = rc[2];
}

// First half kick  v += dt/2/m * F * conv ; x += v*dt  (md_integration.f90:478,484), centre of mass and shift into the box
// (k_com_shift with do_shift = 1) by ONE thread per molecule (its <= RPB_MA atoms are contiguous): one launch and one
// dependency edge at the head of every step instead of two.  The thread that consumed F_i also clears it (and thread 0
// the energy slots), so the force evaluation that follows needs no zeroing launches.
__global__ void k_integrate_first_mol(Dev d) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m == 0) { for (int k = 0; k < E_NSLOT; k++) d.en[k] = 0.0; }
  if (m >= d.M) return;
  const int f = d.mol_first[m], n = d.mol_natom[m];
  double c0 = 0, c1 = 0, c2 = 0, mt = 0;
  for (int a = 0; a < n; a++) {
    const int i = f + a;
    const double f0 = d.force[3 * i], f1 = d.force[3 * i + 1], f2 = d.force[3 * i + 2];
    d.force[3 * i] = 0.0; d.force[3 * i + 1] = 0.0; d.force[3 * i + 2] = 0.0;
    double4 p = d.xq[i];
    const double ms = d.mass[i];
    if (d.freeze[d.type[i]] != 1) {
      const double h = d.dt / 2.0 / ms;
      const double v0 = d.vel[3 * i] + h * f0 * d.conv_kin;
      const double v1 = d.vel[3 * i + 1] + h * f1 * d.conv_kin;
      const double v2 = d.vel[3 * i + 2] + h * f2 * d.conv_kin;
      d.vel[3 * i] = v0; d.vel[3 * i + 1] = v1; d.vel[3 * i + 2] = v2;
      p.x = p.x + v0 * d.dt; p.y = p.y + v1 * d.dt; p.z = p.z + v2 * d.dt;
      d.xq[i] = p;
    }
    c0 = c0 + p.x * ms; c1 = c1 + p.y * ms; c2 = c2 + p.z * ms;
    mt = mt + ms;
  }
  double rc[3] = {c0 / mt, c1 / mt, c2 / mt};
  double t[3];
  bool any = false;
  for (int k = 0; k < 3; k++) {
    const double db = d.inv_box[k] * rc[k];
    double sh = 0.0;
    if (db < 0.0) sh = 1.0; else if (db > 1.0) sh = -1.0;
    t[k] = sh * d.box[k];
    any |= (sh != 0.0);
    rc[k] = rc[k] + t[k];
  }
  if (any)
    for (int a = 0; a < n; a++) {
      double4 p = d.xq[f + a];
      p.x = p.x + t[0]; p.y = p.y + t[1]; p.z = p.z + t[2];
      d.xq[f + a] = p;
    }
  d.r_com[3 * m] = rc[0]; d.r_com[3 * m + 1] = rc[1]; d.r_com[3 * m + 2] = rc[2];
}

// second half kick + force sanity check + momentum partial sums   (md_integration.f90:507-529, 139-156)
__global__ void k_integrate_second(Dev d, double* psum /*[4]: px,py,pz,count*/) {
  __shared__ double sh[32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double p0 = 0, p1 = 0, p2 = 0, cnt = 0;
  if (i < d.N && d.freeze[d.type[i]] != 1) {
    double m = d.mass[i];
    double h = d.dt / 2.0 / m;
    double f0 = d.force[3 * i], f1 = d.force[3 * i + 1], f2 = d.force[3 * i + 2];
    double v0 = d.vel[3 * i] + h * f0 * d.conv_kin;
    double v1 = d.vel[3 * i + 1] + h * f1 * d.conv_kin;
    double v2 = d.vel[3 * i + 2] + h * f2 * d.conv_kin;
    d.vel[3 * i] = v0; d.vel[3 * i + 1] = v1; d.vel[3 * i + 2] = v2;
    if (!(fabs(f0) <= 10e4) || !(fabs(f1) <= 10e4) || !(fabs(f2) <= 10e4)) atomicMax(&d.err_flag[0], i + 1);
    p0 = m * v0; p1 = m * v1; p2 = m * v2; cnt = 1.0;
  }
  p0 = block_sum(p0, sh); p1 = block_sum(p1, sh); p2 = block_sum(p2, sh); cnt = block_sum(cnt, sh);
  // per-block partial sums, combined in a FIXED order by k_remove_com_momentum: the total momentum must come out bit-identical
  // on every rank of a state-sharded run (replicated integration), which a floating-point atomicAdd does not guarantee
  if (threadIdx.x == 0) { double* o = psum + 4 * blockIdx.x; o[0] = p0; o[1] = p1; o[2] = p2; o[3] = cnt; }
}

__global__ void k_remove_com_momentum(Dev d, const double* psum, int n_partial) {
  __shared__ double sh[32];
  __shared__ double tot[4];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < n_partial; b += blockDim.x)
    for (int k = 0; k < 4; k++) a[k] += psum[4 * b + k];
  for (int k = 0; k < 4; k++) { double v = block_sum(a[k], sh); if (threadIdx.x == 0) tot[k] = v; }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double ke = 0.0;
  if (i < d.N) {
    double v0 = d.vel[3 * i], v1 = d.vel[3 * i + 1], v2 = d.vel[3 * i + 2];
    const double m = d.mass[i];
    if (d.freeze[d.type[i]] != 1) {
      const double n = tot[3];
      v0 = v0 - (tot[0] / n) / m; v1 = v1 - (tot[1] / n) / m; v2 = v2 - (tot[2] / n) / m;
      d.vel[3 * i] = v0; d.vel[3 * i + 1] = v1; d.vel[3 * i + 2] = v2;
    }
    ke = 0.5 * m * (v0 * v0 + v1 * v1 + v2 * v2) / d.conv_kin;      // calculate_kinetic_energy total_energy_forces.f90:106-121
  }
  // the step's kinetic energy rides along (slot zeroed by k_integrate_first): rpb_get_energies after a step needs no kernel
  ke = block_sum(ke, sh);
  if (threadIdx.x == 0) atomicAdd(&d.en[E_KE], ke);
}

__global__ void k_kinetic(Dev d) {
  __shared__ double sh[32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0;
  if (i < d.N) {
    double v0 = d.vel[3 * i], v1 = d.vel[3 * i + 1], v2 = d.vel[3 * i + 2];
    e = 0.5 * d.mass[i] * (v0 * v0 + v1 * v1 + v2 * v2) / d.conv_kin;
  }
  e = block_sum(e, sh);
  if (threadIdx.x == 0) atomicAdd(&d.en[E_KE], e);
}


// ------------------------------------------------------------------------------------------------
static inline int nblk(int n, int t = TPB) { return (n + t - 1) / t; }

void launch_zero_forces(rpb_ctx* c) {
  k_zero<<<nblk(3 * c->d.N), TPB, 0, c->stream>>>(c->d.force, (size_t)3 * c->d.N);
  k_zero<<<1, 32, 0, c->stream>>>(c->d.en, E_NSLOT);
  c->n_launch += 2;
}

void launch_integrate_first(rpb_ctx* c) {
  ScopedTimer t(c, T_INTEGRATE);
  k_integrate_first_mol<<<nblk(c->d.M, 64), 64, 0, c->stream>>>(c->d);      // (small CTAs: a few thousand molecules spread over many SMs)
  c->n_launch += 1;
  c->forces_zeroed = true;
}

void launch_update_com_shift(rpb_ctx* c, bool shift) {
  k_com_shift<<<nblk(c->d.M), TPB, 0, c->stream>>>(c->d, shift ? 1 : 0);
  c->n_launch += 1;
}

void launch_integrate_second(rpb_ctx* c) {
  ScopedTimer t(c, T_INTEGRATE);
  const int nb = nblk(c->d.N);
  double* psum = c->d.maxd + 8 + 2 * ((c->d.N + 255) / 256 + 1);   // [nb][4] per-block momentum partials behind the displacement scratch
  k_integrate_second<<<nb, TPB, 0, c->stream>>>(c->d, psum);
  k_remove_com_momentum<<<nb, TPB, 0, c->stream>>>(c->d, psum, nb);
  c->n_launch += 2;
}

void launch_kinetic_energy(rpb_ctx* c) {
  k_zero<<<1, 32, 0, c->stream>>>(c->d.en + E_KE, 1);
  k_kinetic<<<nblk(c->d.N), TPB, 0, c->stream>>>(c->d);
  c->n_launch += 2;
}

// ------------------------------------------------------------------------------------------------
// fp64 FMA peak of this device (roofline denominator for the FP64-pipe-bound pair kernels): 8 independent
// DFMA chains per thread, 148*8 CTAs of 256 threads.
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, cc = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int measure_fp64_peak(rpb_ctx* c, double* tflops) {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
  double* buf = nullptr;
  if (cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) { c->err = "cudaMalloc failed"; return RPB_ERR_CUDA; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, c->stream);
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(buf, iters);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
  *tflops = best;
  return 0;
}
