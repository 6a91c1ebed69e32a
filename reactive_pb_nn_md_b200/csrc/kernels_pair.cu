// Real-space pair forces over the half Verlet list, and per-molecule bonded / intramolecular terms.
// Replaces:
//   pairwise_real_space_verlet            src/pair_int_real_space.f90:135-371
//   pairwise_real_space_{ewald,LJ,sapt}   src/pair_int_real_space.f90:621-759
//   intra_molecular_pairwise_energy_force src/pair_int_real_space.f90:386-588
//   intra_molecular_energy_force          src/intra_bonded_interactions.f90:17-552
//
// Pair kernel: one warp per i-atom, lanes stride over the CSR row (coalesced neighbour indices,
// 32-byte xq gathers that stay in L2), fp64 throughout; F_i is reduced in registers and shuffles,
// F_j goes out as fp64 RED atomics.  Bound by the FP64 pipe (div/sqrt sequences) -- see DESIGN.md.
#include "rpb_host.h"
#include "rpb_bonded.cuh"

#define PAIR_TPB 256

__global__ void __launch_bounds__(PAIR_TPB) k_pair_verlet(Dev d) {
  extern __shared__ double sh_par[];           // [nT*nT][6] vdw parameters
  __shared__ int sh_vt[RPB_MAXT * RPB_MAXT];
  __shared__ double sh_red[32];
  for (int k = threadIdx.x; k < d.nT * d.nT * 6; k += blockDim.x) sh_par[k] = d.vdw_param[k];
  for (int k = threadIdx.x; k < d.nT * d.nT; k += blockDim.x) sh_vt[k] = d.vdw_type[k];
  __syncthreads();
  int lane = threadIdx.x & 31;
  int nwarp_total = (gridDim.x * blockDim.x) >> 5;
  double e_el = 0.0, e_vdw = 0.0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < d.N; i += nwarp_total) {
    int vs = d.verlet_point[i] - 1, vf = d.verlet_point[i + 1] - 1;
    if (vf <= vs) continue;
    double4 pi = d.xq[i];
    int ti = d.type[i];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int v = vs + lane; v < vf; v += 32) {
      int j = d.neighbor_list[v] - 1;
      double4 pj = d.xq[j];
      double dr[3];
      dr[0] = min_image(pi.x - pj.x, d.box[0]);
      dr[1] = min_image(pi.y - pj.y, d.box[1]);
      dr[2] = min_image(pi.z - pj.z, d.box[2]);
      double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
      if (dr2 < d.rc2) {
        int pidx = ti * d.nT + d.type[j];
        double ee, ev, f[3];
        pair_terms(d, dr, dr2, pi.w * pj.w, sh_vt[pidx], &sh_par[6 * pidx], true, ee, ev, f);
        e_el += ee; e_vdw += ev;
        fx += f[0]; fy += f[1]; fz += f[2];
        atomicAdd(&d.force[3 * j], -f[0]);
        atomicAdd(&d.force[3 * j + 1], -f[1]);
        atomicAdd(&d.force[3 * j + 2], -f[2]);
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) {
      atomicAdd(&d.force[3 * i], fx);
      atomicAdd(&d.force[3 * i + 1], fy);
      atomicAdd(&d.force[3 * i + 2], fz);
    }
  }
  e_el = block_sum(e_el, sh_red);
  e_vdw = block_sum(e_vdw, sh_red);
  if (threadIdx.x == 0) { atomicAdd(&d.en[E_ELEC], e_el); atomicAdd(&d.en[E_VDW], e_vdw); }
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
