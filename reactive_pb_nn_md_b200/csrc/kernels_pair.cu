// Real-space pair forces over the half Verlet list, and per-molecule bonded / intramolecular terms.
// Replaces:
//   pairwise_real_space_verlet            src/pair_int_real_space.f90:135-371
//   pairwise_real_space_{ewald,LJ,sapt}   src/pair_int_real_space.f90:621-759
//   intra_molecular_pairwise_energy_force src/pair_int_real_space.f90:386-588
//   intra_molecular_energy_force          src/intra_bonded_interactions.f90:17-552
//
// Pair kernel: one warp per i-atom over its row of the SYMMETRIC list (both directions of every listed pair, built
// next to the reference-ordered half list at every rebuild): each pair is evaluated from both of its atoms, F_i is
// reduced in registers + shuffles and stored once -- no atomics, so the forces are reproducible run to run -- and the
// energies are halved.  Per listed pair: minimum image with a reciprocal box (the shift can only differ from the
// reference's division for |dr| ~ L/2, far outside the cutoff) and the cutoff test.  Per in-cutoff pair: ONE rsqrt
// replaces the sqrt and the six divisions of pair_int_real_space.f90:621-645,698-759 (relative differences ~1e-16,
// the interpolated tables are continuous across bins), and the erfc / ewaldscale tables are read interleaved.
// Bound by the FP64 pipe -- see DESIGN.md.
#include "rpb_host.h"
#include "rpb_bonded.cuh"

#define PAIR_TPB 256

__global__ void __launch_bounds__(PAIR_TPB, 3) k_pair_verlet(Dev d) {
  extern __shared__ double sh_par[];           // [nT*nT][6] vdw parameters
  __shared__ int sh_vt[RPB_MAXT * RPB_MAXT];   // atype_vdw_type; 2 = SAPT row with all-zero coefficients (contributes exactly 0)
  __shared__ double sh_red[32];
  for (int k = threadIdx.x; k < d.nT * d.nT * 6; k += blockDim.x) sh_par[k] = d.vdw_param[k];
  for (int k = threadIdx.x; k < d.nT * d.nT; k += blockDim.x) {
    int vt = d.vdw_type[k];
    if (vt == 1) {
      const double* P = &d.vdw_param[6 * k];
      if (P[0] == 0.0 && P[2] == 0.0 && P[3] == 0.0 && P[4] == 0.0 && P[5] == 0.0) vt = 2;
    }
    sh_vt[k] = vt;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nwarp_total = (gridDim.x * blockDim.x) >> 5;
  const double ibx = d.inv_box[0], iby = d.inv_box[1], ibz = d.inv_box[2];
  const double bx = d.box[0], by = d.box[1], bz = d.box[2];
  double e_el = 0.0, e_vdw = 0.0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < d.N; i += nwarp_total) {
    const int vs = d.full_point[i], vf = d.full_point[i + 1];
    const double4 pi = d.xq[i];
    const int ti = d.type[i] * d.nT;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int v = vs + lane; v < vf; v += 32) {
      const int j = d.full_list[v];
      const double4 pj = d.xq[j];
      double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
      dx = fma(-bx, floor(fma(dx, ibx, 0.5)), dx);
      dy = fma(-by, floor(fma(dy, iby, 0.5)), dy);
      dz = fma(-bz, floor(fma(dz, ibz, 0.5)), dz);
      const double dr2 = fma(dz, dz, fma(dy, dy, dx * dx));
      if (dr2 < d.rc2) {
        const int pidx = ti + d.type[j];
        const int vt = sh_vt[pidx];
        const double inv_r = rsqrt(dr2);
        const double r = dr2 * inv_r, inv_r2 = inv_r * inv_r;
        // linear_interpolation_ewald_tables  pair_int_real_space.f90:740-759
        const double x1 = r * d.inv_erfc_dx;
        const double ci = ceil(x1);
        const int it = (int)ci;
        const double c2 = (x1 + 1.0) - ci, c1 = 1.0 - c2;
        const double2 t0 = __ldg(&d.es_t[it - 1]), t1 = __ldg(&d.es_t[it]);
        const double qr = (pi.w * pj.w) * inv_r;
        e_el = fma(qr, fma(c2, t1.x, c1 * t0.x), e_el);
        double fs = (qr * inv_r2) * fma(c2, t1.y, c1 * t0.y);
        if (vt == 0) {                       // pairwise_real_space_LJ :621-645
          const double c12 = sh_par[6 * pidx], c6 = sh_par[6 * pidx + 1];
          const double r6 = inv_r2 * inv_r2 * inv_r2, c12r6 = c12 * r6;
          e_vdw = fma(r6, c12r6 - c6, e_vdw);
          fs = fma(inv_r2 * r6, 12.0 * c12r6 - 6.0 * c6, fs);
        } else if (vt == 1) {                // pairwise_real_space_sapt :651-690 (generic path)
          double dr[3] = {dx, dy, dz}, ee, ev, f[3];
          pair_terms(d, dr, dr2, 0.0, 1, &sh_par[6 * pidx], false, ee, ev, f);
          e_vdw += ev;
          fx += f[0]; fy += f[1]; fz += f[2];
        }
        fx = fma(dx, fs, fx); fy = fma(dy, fs, fy); fz = fma(dz, fs, fz);
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) { d.force[3 * i] += fx; d.force[3 * i + 1] += fy; d.force[3 * i + 2] += fz; }   // this warp owns atom i
  }
  e_el = block_sum(e_el, sh_red);
  e_vdw = block_sum(e_vdw, sh_red);
  if (threadIdx.x == 0) { atomicAdd(&d.en[E_ELEC], 0.5 * e_el); atomicAdd(&d.en[E_VDW], 0.5 * e_vdw); }
}

// one thread per molecule: intramolecular non-bonded (exclusion correction, 1-4) + bonds/angles/dihedrals
__global__ void k_molecule_terms(Dev d) {
  __shared__ double sh_red[32];
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  MolEnergies E = {0, 0, 0, 0, 0};
  if (m < d.M) {
    int f0 = d.mol_first[m], n = d.mol_natom[m];
    const MolTypeDev& T = d.mt[d.mol_type[m]];
    double x[RPB_MA][3], q[RPB_MA], f[RPB_MA][3];
    int ty[RPB_MA];
    for (int a = 0; a < n; a++) {
      double4 p = d.xq[f0 + a];
      x[a][0] = p.x; x[a][1] = p.y; x[a][2] = p.z; q[a] = p.w;
      ty[a] = d.type[f0 + a];
      f[a][0] = f[a][1] = f[a][2] = 0.0;
    }
    molecule_terms(d, T, n, x, ty, q, f, E, true, true);
    for (int a = 0; a < n; a++)
      for (int k = 0; k < 3; k++) atomicAdd(&d.force[3 * (f0 + a) + k], f[a][k]);
  }
  double e;
  e = block_sum(E.e_elec, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_ELEC], e);
  e = block_sum(E.e_vdw, sh_red);  if (threadIdx.x == 0) atomicAdd(&d.en[E_VDW], e);
  e = block_sum(E.e_bond, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_BOND], e);
  e = block_sum(E.e_angle, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_ANGLE], e);
  e = block_sum(E.e_dih, sh_red);  if (threadIdx.x == 0) atomicAdd(&d.en[E_DIH], e);
}

void launch_pair_verlet(rpb_ctx* c) {
  ScopedTimer t(c, T_PAIR);
  int warps_per_block = PAIR_TPB / 32;
  int blocks = std::min((c->d.N + warps_per_block - 1) / warps_per_block, 148 * 8);
  size_t shmem = (size_t)c->d.nT * c->d.nT * 6 * sizeof(double);
  k_pair_verlet<<<blocks, PAIR_TPB, shmem, c->stream>>>(c->d);
  c->n_launch += 1;
}

void launch_molecule_terms(rpb_ctx* c) {
  ScopedTimer t(c, T_INTRA);
  k_molecule_terms<<<(c->d.M + 127) / 128, 128, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}
