// MS-EVB Hamiltonian build, diagonalisation and Hellmann-Feynman mixing on the device.
// Replaces (reference file:line):
//   evb_conduct_proton_transfer_recursive / find_evb_reactive_neighbors   ms_evb.f90:498-607, 702-764
//   evb_change_data_structures_proton_transfer (+ reorder)               ms_evb.f90:843-1006   -> snapshots
//   ms_evb_diabat_force_energy (+ _update_real_space, _update_intra)     ms_evb.f90:1421-1954  -> item kernels
//   ms_evb_intermolecular_repulsion                                      ms_evb.f90:2259-2504
//   modify_Q_grid per diabat, calculate_reciprocal_space_pme,
//   update_reciprocal_space_force_dQ_dr                                  pme.f90:275-335, ms_evb.f90:1962-2248
//   evb_diabatic_coupling (+ geometric / function / electrostatics)      ms_evb.f90:1021-1403
//   diagonalize_evb_hamiltonian + jacobi                                 ms_evb.f90:242-351, general_routines.f90:2013-2088
//
// Structure (all diabats of this rank in flight; SURVEY 7 "two-pass" reciprocal formulation):
//   enumerate -> [host reads S + hop log: sizes the launches] -> snapshots -> item kernels (background + chain)
//   -> batched delta grids Q_s = Q_1 -/+ patches -> ONE batched D2Z/Z2D over the owned diabats with the k-space
//   energy fused into the CB pass -> per-diabat corrections for the few chain atoms -> coupling -> H elements.
//   After the (optional) all-reduce of H: warp-level cyclic Jacobi -> theta_mix = sum c_s^2 theta_s (streaming)
//   -> ONE gather for all atoms -> F = sum c_i c_j F_ij.
#include <algorithm>
#include <cstring>
#include "rpb_host.h"
#include "rpb_bonded.cuh"
#include "rpb_pme.cuh"

#define MAXS RPB_MAXS
#define MAXC RPB_MAXC
#define MA RPB_MA
#define CM RPB_CHAIN_MOLS
#define NLEV (RPB_MAXC + 1)

__host__ __device__ inline bool state_owned(int s, int rank, int world) {
  if (world <= 1) return true;
  if (s == 0) return rank == 0;
  return ((s - 1) % world) == rank;
}

// ================================================================================================
// K8: diabat enumeration, pre-order DFS exactly as evb_conduct_proton_transfer_recursive
// ================================================================================================
struct Frame { int mol, diabat, count, i_atom, n_nb, cursor; int nb[RPB_EVB_MAX_NEIGHBORS][2]; };

__global__ void k_evb_enumerate(Dev d, EvbDev e) {
  __shared__ Frame fr[MAXC + 2];
  __shared__ int depth, op, s_count, n_hit;
  __shared__ int hits[64][2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < MAXS * MAXC * 5; i++) e.proton_log[i] = -1;
    for (int i = 0; i < MAXS; i++) { e.parent[i] = -1; e.n_hops[i] = 0; }
    s_count = 1;
    depth = 0;
    fr[0].mol = *d.hydronium; fr[0].diabat = 0; fr[0].count = 0; fr[0].i_atom = -1; fr[0].n_nb = 0; fr[0].cursor = 0;
    op = 0;
  }
  __syncthreads();
  const int hyd = *d.hydronium;
  while (true) {
    if (tid == 0) {
      op = 0;
      while (true) {
        if (depth < 0) { op = 2; break; }
        Frame& f = fr[depth];
        if (f.count >= d.max_chain) { depth--; continue; }       // "if ( count < evb_max_chain )" :538
        if (f.cursor < f.n_nb) {
          int acc_mol = f.nb[f.cursor][0], acc_atom = f.nb[f.cursor][1];
          f.cursor++;
          if (s_count >= d.max_states) { atomicMax(&d.err_flag[2], 1); op = 2; break; }
          int da = s_count++;
          e.parent[da] = f.diabat;
          for (int h = 0; h < f.count; h++)
            for (int q = 0; q < 5; q++) e.proton_log[(da * MAXC + h) * 5 + q] = e.proton_log[(f.diabat * MAXC + h) * 5 + q];
          const MolTypeDev& T = d.mt[d.mol_type[f.mol]];
          int* L = &e.proton_log[(da * MAXC + f.count) * 5];
          L[0] = f.mol; L[1] = f.i_atom; L[2] = T.bonded_heavy[f.i_atom]; L[3] = acc_mol; L[4] = acc_atom;
          if (L[2] < 0) atomicMax(&d.err_flag[3], 1);
          e.n_hops[da] = f.count + 1;
          if (acc_mol != hyd) {                                     // flag_cycle :573,596
            Frame& g = fr[depth + 1];
            g.mol = acc_mol; g.diabat = da; g.count = f.count + 1; g.i_atom = -1; g.n_nb = 0; g.cursor = 0;
            depth++;
          }
          continue;
        }
        // next reactive proton of this donor (principal-topology molecule type) :542-546
        const MolTypeDev& T = d.mt[d.mol_type[f.mol]];
        int ia = f.i_atom + 1, n = d.mol_natom[f.mol];
        while (ia < n && T.reactive_proton[ia] != 1) ia++;
        if (ia >= n) { depth--; continue; }
        f.i_atom = ia; f.n_nb = 0; f.cursor = 0;
        n_hit = 0;
        op = 1;
        break;
      }
    }
    __syncthreads();
    if (op == 2) break;
    // find_evb_reactive_neighbors for (fr[depth].mol, fr[depth].i_atom)  :702-764
    {
      const Frame& f = fr[depth];
      int im = f.mol;
      double rci[3] = {d.r_com[3 * im], d.r_com[3 * im + 1], d.r_com[3 * im + 2]};
      double4 ph = d.xq[d.mol_first[im] + f.i_atom];
      double xh[3] = {ph.x, ph.y, ph.z};
      for (int jm = tid; jm < d.M; jm += blockDim.x) {
        if (jm == im) continue;
        double shift[3], dc[3];
        for (int k = 0; k < 3; k++) {
          double dr = d.r_com[3 * jm + k] - rci[k];
          shift[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
          dc[k] = d.r_com[3 * jm + k] - rci[k] - shift[k];
        }
        if (dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2] < d.cut_solv2) {
          const MolTypeDev& TJ = d.mt[d.mol_type[jm]];
          int fj = d.mol_first[jm], nj = d.mol_natom[jm];
          for (int ja = 0; ja < nj; ja++) {
            if (TJ.reactive_basic[ja] != 1) continue;
            double4 pj = d.xq[fj + ja];
            double r0 = pj.x - xh[0] - shift[0], r1 = pj.y - xh[1] - shift[1], r2 = pj.z - xh[2] - shift[2];
            if (r0 * r0 + r1 * r1 + r2 * r2 < d.cut_pair2) {
              int slot = atomicAdd(&n_hit, 1);
              if (slot < 64) { hits[slot][0] = jm; hits[slot][1] = ja; }
            }
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int n = min(n_hit, 64);
      for (int a = 1; a < n; a++) {   // ascending (molecule, atom) == the reference's loop order
        int m0 = hits[a][0], a0 = hits[a][1], b = a - 1;
        while (b >= 0 && (hits[b][0] > m0 || (hits[b][0] == m0 && hits[b][1] > a0))) { hits[b + 1][0] = hits[b][0]; hits[b + 1][1] = hits[b][1]; b--; }
        hits[b + 1][0] = m0; hits[b + 1][1] = a0;
      }
      Frame& f = fr[depth];
      f.n_nb = min(n, RPB_EVB_MAX_NEIGHBORS);   // evb_neighbor_list(evb_max_neighbors,2)
      for (int a = 0; a < f.n_nb; a++) { f.nb[a][0] = hits[a][0]; f.nb[a][1] = hits[a][1]; }
      f.cursor = 0;
    }
    __syncthreads();
  }
  if (tid == 0) *e.n_states = s_count;
}

// ================================================================================================
// snapshots: images of the chain molecules of diabat s at every topology level
// ================================================================================================
__device__ void load_principal_image(const Dev& d, int mol, MolImage& im) {
  im.mol = mol; im.n_atom = d.mol_natom[mol]; im.mtype = d.mol_type[mol];
  int f = d.mol_first[mol];
  for (int a = 0; a < im.n_atom; a++) {
    double4 p = d.xq[f + a];
    im.atom[a] = f + a; im.type[a] = d.type[f + a]; im.q[a] = p.w; im.mass[a] = d.mass[f + a];
    im.x[a][0] = p.x; im.x[a][1] = p.y; im.x[a][2] = p.z;
  }
  for (int k = 0; k < 3; k++) im.r_com[k] = d.r_com[3 * mol + k];
}

__device__ void image_pos_com(MolImage& im) {   // pos_com general_routines.f90:398-415
  double c0 = 0, c1 = 0, c2 = 0, mt = 0;
  for (int a = 0; a < im.n_atom; a++) {
    c0 = c0 + im.x[a][0] * im.mass[a]; c1 = c1 + im.x[a][1] * im.mass[a]; c2 = c2 + im.x[a][2] * im.mass[a];
    mt = mt + im.mass[a];
  }
  im.r_com[0] = c0 / mt; im.r_com[1] = c1 / mt; im.r_com[2] = c2 / mt;
}

__device__ void image_swap_atoms_rotate(MolImage& im, int i, int index) {  // move atom `index` to position i, shifting i..index-1 up
  int at = im.atom[index], ty = im.type[index];
  double q = im.q[index], ms = im.mass[index], x0 = im.x[index][0], x1 = im.x[index][1], x2 = im.x[index][2];
  for (int j = index - 1; j >= i; j--) {
    im.atom[j + 1] = im.atom[j]; im.type[j + 1] = im.type[j]; im.q[j + 1] = im.q[j]; im.mass[j + 1] = im.mass[j];
    im.x[j + 1][0] = im.x[j][0]; im.x[j + 1][1] = im.x[j][1]; im.x[j + 1][2] = im.x[j][2];
  }
  im.atom[i] = at; im.type[i] = ty; im.q[i] = q; im.mass[i] = ms; im.x[i][0] = x0; im.x[i][1] = x1; im.x[i][2] = x2;
}

// evb_change_data_structures_proton_transfer on images (ms_evb.f90:843-932)
__device__ void image_proton_transfer(const Dev& d, MolImage& D, MolImage& A, int i_atom_donor, int i_heavy_acceptor) {
  const EvbTables& E = *d.evb;
  // shift_array_data_donor_acceptor_transfer: the proton leaves the donor and is appended to the acceptor
  int last = A.n_atom;
  A.atom[last] = D.atom[i_atom_donor]; A.type[last] = D.type[i_atom_donor]; A.q[last] = D.q[i_atom_donor];
  A.mass[last] = D.mass[i_atom_donor];
  for (int k = 0; k < 3; k++) A.x[last][k] = D.x[i_atom_donor][k];
  for (int a = i_atom_donor; a < D.n_atom - 1; a++) {
    D.atom[a] = D.atom[a + 1]; D.type[a] = D.type[a + 1]; D.q[a] = D.q[a + 1]; D.mass[a] = D.mass[a + 1];
    for (int k = 0; k < 3; k++) D.x[a][k] = D.x[a + 1][k];
  }
  D.n_atom -= 1; A.n_atom += 1;
  // make_molecule_whole on the acceptor (general_routines.f90:1065-1086)
  for (int i = 1; i < A.n_atom; i++) {
    double sh[3];
    bool any = false;
    for (int k = 0; k < 3; k++) {
      double dr = A.x[i][k] - A.x[i - 1][k];
      sh[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
      any |= fabs(sh[k]) > 1e-6;
    }
    if (any) for (int k = 0; k < 3; k++) { double drij = A.x[i][k] - A.x[i - 1][k] - sh[k]; A.x[i][k] = A.x[i - 1][k] + drij; }
  }
  image_pos_com(D);
  image_pos_com(A);
  int acid_type = E.conj_pairs[A.mtype];
  A.type[last] = E.proton_index[acid_type];
  for (int a = 0; a < A.n_atom; a++) {
    int tn = (a != last) ? E.conj_atom[A.type[a]] : A.type[a];
    A.type[a] = tn; A.q[a] = E.atype_chg[tn];
  }
  A.type[i_heavy_acceptor] = E.heavy_acid_index[acid_type];
  for (int a = 0; a < D.n_atom; a++) { int tn = E.conj_atom[D.type[a]]; D.type[a] = tn; D.q[a] = E.atype_chg[tn]; }
  D.mtype = E.conj_pairs[D.mtype];
  A.mtype = acid_type;
  // reorder_molecule_data_structures (ms_evb.f90:941-1006)
  const MolTypeDev& T = d.mt[A.mtype];
  for (int i = 0; i < T.n_atom; i++) {
    if (T.atom_type[i] != A.type[i]) {
      int index = -1;
      for (int j = i + 1; j < A.n_atom; j++) if (T.atom_type[i] == A.type[j]) { index = j; break; }
      if (index < 0) { atomicMax(&d.err_flag[3], 2); return; }
      image_swap_atoms_rotate(A, i, index);
    }
  }
}

// only_state >= 0: build that diabat regardless of ownership (hop commit needs the new principal's images on every rank)
__global__ void k_evb_snapshots(Dev d, EvbDev e, int only_state) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int S = *e.n_states;
  if (only_state >= 0) { if (s != 0) return; s = only_state; }
  else if (s >= S || !(s == 0 || state_owned(s, d.rank, d.world))) return;
  Snapshot W;
  const int* L = &e.proton_log[s * MAXC * 5];
  int nh = e.n_hops[s];
  int m0 = *d.hydronium;
  W.n_mol = 1;
  load_principal_image(d, m0, W.m[0]);
  for (int h = 0; h < nh; h++) {
    int a = L[h * 5 + 3];
    bool found = false;
    for (int k = 0; k < W.n_mol; k++) found |= (W.m[k].mol == a);
    if (!found) { load_principal_image(d, a, W.m[W.n_mol]); W.n_mol++; }
  }
  for (int k = W.n_mol; k < CM; k++) { W.m[k].mol = -1; W.m[k].n_atom = 0; W.m[k].mtype = 0; }
  W.hydronium = 0;
  e.snap[s * NLEV + 0] = W;
  int cur = 0;  // slot of the current hydronium (= donor of the next hop)
  for (int h = 0; h < nh; h++) {
    int a = L[h * 5 + 3], as = 0;
    for (int k = 0; k < W.n_mol; k++) if (W.m[k].mol == a) as = k;
    image_proton_transfer(d, W.m[cur], W.m[as], L[h * 5 + 1], L[h * 5 + 4]);
    W.hydronium = as;
    cur = as;
    e.snap[s * NLEV + h + 1] = W;
  }
}

// ================================================================================================
// item kernels: real-space + EVB repulsion deltas of one (diabat, hop, topology)
// ================================================================================================
struct ItemShared {
  int n_chain; int chain_atoms[CM * MA];
  int nd, na, nh;                        // donor / acceptor / hydronium image sizes
  int d_atom[MA], a_atom[MA], h_atom[MA];
  int d_type[MA], a_type[MA], h_type[MA];
  double d_q[MA], a_q[MA];
  double d_x[MA][3], a_x[MA][3], h_x[MA][3];
  int h_heavy, h_type_H, h_type_heavy;
  int da_row[RPB_MAXT];                 // three-atom repulsion row per solvent atom type (-1 none)
  int pa_row[MA][RPB_MAXT];             // Born-Mayer row per (hydronium atom, solvent atom type)
};

__device__ void fill_item_shared(const Dev& d, const Snapshot& S, const EvbItem& it, ItemShared& sh) {
  // executed by thread 0
  sh.n_chain = 0;
  for (int k = 0; k < S.n_mol; k++)
    for (int a = 0; a < S.m[k].n_atom; a++) sh.chain_atoms[sh.n_chain++] = S.m[k].atom[a];
  sh.nd = sh.na = 0;
  if (it.donor_slot >= 0) {
    const MolImage& D = S.m[it.donor_slot];
    sh.nd = D.n_atom;
    for (int a = 0; a < D.n_atom; a++) { sh.d_atom[a] = D.atom[a]; sh.d_type[a] = D.type[a]; sh.d_q[a] = D.q[a]; for (int k = 0; k < 3; k++) sh.d_x[a][k] = D.x[a][k]; }
  }
  if (it.acceptor_slot >= 0) {
    const MolImage& A = S.m[it.acceptor_slot];
    sh.na = A.n_atom;
    for (int a = 0; a < A.n_atom; a++) { sh.a_atom[a] = A.atom[a]; sh.a_type[a] = A.type[a]; sh.a_q[a] = A.q[a]; for (int k = 0; k < 3; k++) sh.a_x[a][k] = A.x[a][k]; }
  }
  const MolImage& H = S.m[S.hydronium];
  const EvbTables& E = *d.evb;
  sh.nh = H.n_atom;
  for (int a = 0; a < H.n_atom; a++) { sh.h_atom[a] = H.atom[a]; sh.h_type[a] = H.type[a]; for (int k = 0; k < 3; k++) sh.h_x[a][k] = H.x[a][k]; }
  sh.h_heavy = d.mt[H.mtype].heavy_acid_atom;
  if (sh.h_heavy < 0) { atomicMax(&d.err_flag[3], 3); sh.h_heavy = 0; }
  sh.h_type_H = H.type[H.n_atom - 1];
  sh.h_type_heavy = H.type[sh.h_heavy];
  for (int t = 0; t < RPB_MAXT; t++) {
    int row = -1;
    for (int i = 0; i < RPB_MAXI; i++) {            // get_index_atom_set general_routines.f90:613-637
      if (E.da_int[i][0] < 0) break;
      if (E.da_int[i][0] == t && E.da_int[i][1] == sh.h_type_heavy && E.da_int[i][2] == sh.h_type_H) { row = i; break; }
    }
    sh.da_row[t] = row;
    for (int a = 0; a < H.n_atom; a++) {
      int r2 = -1;
      for (int i = 0; i < RPB_MAXI; i++) {
        if (E.pa_int[i][0] < 0) break;
        if (E.pa_int[i][0] == t && E.pa_int[i][1] == H.type[a]) { r2 = i; break; }
      }
      sh.pa_row[a][t] = r2;
    }
  }
}

__device__ __forceinline__ void repulsive_switch(double& sw, double& dsw, double r, double rs, double rc) {  // ms_evb.f90:2484-2504
  sw = 0.0; dsw = 0.0;
  if (r < rc) {
    if (r < rs) sw = 1.0;
    else {
      double c3 = (rc - rs) * (rc - rs) * (rc - rs);
      double term1 = (r - rs) * (r - rs) / c3;
      double term2 = 3.0 * rc - rs - 2.0 * r;
      sw = 1.0 - term1 * term2;
      dsw = -2.0 * (r - rs) * term2 / c3 + 2.0 * term1;
    }
  }
}

// EVB repulsion of the hydronium image with ONE solvent atom j (ms_evb.f90:2295-2478).
// fh[a][3] accumulates forces on hydronium atoms, fj on the solvent atom. Returns the energy.
__device__ inline double repulsion_with_atom(const Dev& d, const ItemShared& sh, const double xj[3], int tj, double (*fh)[3], double fj[3]) {
  const EvbTables& E = *d.evb;
  double en = 0.0;
  int row = sh.da_row[tj];
  if (row >= 0) {
    const double* P = E.da_par[row];
    double B = P[0], bl = P[1], d0 = P[2], blp = P[3], rs = P[4], rc = P[5];
    const double* xo = sh.h_x[sh.h_heavy];
    double shift[3], rO[3];
    for (int k = 0; k < 3; k++) {
      double dr = xj[k] - xo[k];
      shift[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
      rO[k] = -(xj[k] - xo[k] - shift[k]);
    }
    double r_OO = sqrt(rO[0] * rO[0] + rO[1] * rO[1] + rO[2] * rO[2]);
    if (r_OO < rc) {   // switch == dswitch == 0 beyond rc: every term below is exactly zero
      double sw, dsw;
      repulsive_switch(sw, dsw, r_OO, rs, rc);
      double fac_OO = B * exp(-bl * (r_OO - d0));
      double sum = 0.0;
      for (int a = 0; a < sh.nh; a++) {
        if (sh.h_type[a] != sh.h_type_H) continue;
        double q[3];
        for (int k = 0; k < 3; k++) {
          double rij = -(xj[k] - sh.h_x[a][k] - shift[k]);
          q[k] = (2.0 * xj[k] + rO[k]) / 2.0 - (xj[k] + rij);
        }
        double q2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
        double exp_q = exp(-blp * q2);
        sum = sum + exp_q;
        for (int k = 0; k < 3; k++) {
          fh[a][k] += sw * fac_OO * exp_q * -blp * 2.0 * q[k];
          double t = sw * fac_OO * exp_q * blp * q[k];
          fh[sh.h_heavy][k] += t;
          fj[k] += t;
        }
      }
      en += sw * fac_OO * sum;
      for (int k = 0; k < 3; k++) {
        double fij = rO[k] / r_OO * fac_OO * sum * (sw * bl - dsw);
        fh[sh.h_heavy][k] += fij;
        fj[k] -= fij;
      }
    }
  }
  for (int a = 0; a < sh.nh; a++) {
    int r2 = sh.pa_row[a][tj];
    if (r2 < 0) continue;
    const double* P = E.pa_par[r2];
    double C = P[0], cl = P[1], d0 = P[2], rs = P[3], rc = P[4];
    double rij[3];
    for (int k = 0; k < 3; k++) {
      double dr = xj[k] - sh.h_x[a][k];
      double shf = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
      rij[k] = -(xj[k] - sh.h_x[a][k] - shf);
    }
    double r = sqrt(rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2]);
    if (r < rc) {
      double sw, dsw;
      repulsive_switch(sw, dsw, r, rs, rc);
      double fac_OH = C * exp(-cl * (r - d0));
      en += sw * fac_OH;
      for (int k = 0; k < 3; k++) {
        double fij = rij[k] / r * fac_OH * (sw * cl - dsw);
        fh[a][k] += fij;
        fj[k] -= fij;
      }
    }
  }
  return en;
}

#define ITEM_TPB 256
// grid = (n_items, ceil(N/ITEM_TPB)): image atoms of one item against the background (non-chain) atoms
__global__ void __launch_bounds__(ITEM_TPB) k_evb_items_background(Dev d, EvbDev e, int n_items) {
  __shared__ ItemShared sh;
  __shared__ double red[32];
  __shared__ double facc[ITEM_TPB / 32][3 * MA][3];   // per-warp force accumulators: donor | acceptor | hydronium
  const int item_id = e.real_list[blockIdx.x];
  const EvbItem it = e.items[item_id];
  const Snapshot& S = e.snap[it.state * NLEV + it.level];
  if (threadIdx.x == 0) fill_item_shared(d, S, it, sh);
  for (int k = threadIdx.x; k < (ITEM_TPB / 32) * 3 * MA * 3; k += blockDim.x) (&facc[0][0][0])[k] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* outF = (it.state == 0) ? d.force : e.dF + (size_t)it.state * 3 * d.N;
  const double sign = it.sign;
  int j = blockIdx.y * blockDim.x + threadIdx.x;
  bool active = j < d.N;
  double xj[3] = {0, 0, 0}, qj = 0.0;
  int tj = 0;
  if (active) {
    for (int k = 0; k < sh.n_chain; k++) active &= (sh.chain_atoms[k] != j);
  }
  if (active) { double4 p = d.xq[j]; xj[0] = p.x; xj[1] = p.y; xj[2] = p.z; qj = p.w; tj = d.type[j]; }
  double en = 0.0, fj[3] = {0, 0, 0};
  // ---- real-space pairs (ms_evb.f90:1629-1841)
  for (int side = 0; side < 2; side++) {
    int n = side == 0 ? sh.nd : sh.na;
    for (int a = 0; a < n; a++) {
      const double* xi = side == 0 ? sh.d_x[a] : sh.a_x[a];
      double f[3] = {0, 0, 0};
      bool hit = false;
      if (active) {
        double dr[3] = {min_image(xi[0] - xj[0], d.box[0]), min_image(xi[1] - xj[1], d.box[1]), min_image(xi[2] - xj[2], d.box[2])};
        double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        if (dr2 < d.rc2) {
          int ti = side == 0 ? sh.d_type[a] : sh.a_type[a];
          double qi = side == 0 ? sh.d_q[a] : sh.a_q[a];
          int pidx = ti * d.nT + tj;
          double ee, ev;
          pair_terms(d, dr, dr2, qi * qj, d.vdw_type[pidx], &d.vdw_param[6 * pidx], true, ee, ev, f);
          en += ee + ev;
          fj[0] -= f[0]; fj[1] -= f[1]; fj[2] -= f[2];
          hit = true;
        }
      }
      if (__any_sync(0xffffffffu, hit)) {
        double s0 = warp_sum(f[0]), s1 = warp_sum(f[1]), s2 = warp_sum(f[2]);
        if (lane == 0) { double* t = facc[w][side * MA + a]; t[0] += s0; t[1] += s1; t[2] += s2; }
      }
    }
  }
  // ---- EVB repulsion of the hydronium image (ms_evb.f90:2259-2478)
  {
    double fh[MA][3];
    for (int a = 0; a < MA; a++) fh[a][0] = fh[a][1] = fh[a][2] = 0.0;
    bool hit = false;
    if (active && (sh.da_row[tj] >= 0 || true)) {
      double e0 = repulsion_with_atom(d, sh, xj, tj, fh, fj);
      hit = (e0 != 0.0);
      for (int a = 0; a < sh.nh && !hit; a++) hit |= (fh[a][0] != 0.0 || fh[a][1] != 0.0 || fh[a][2] != 0.0);
      en += e0;
    }
    if (__any_sync(0xffffffffu, hit)) {
      for (int a = 0; a < sh.nh; a++) {
        double s0 = warp_sum(fh[a][0]), s1 = warp_sum(fh[a][1]), s2 = warp_sum(fh[a][2]);
        if (lane == 0) { double* t = facc[w][2 * MA + a]; t[0] += s0; t[1] += s1; t[2] += s2; }
      }
    }
  }
  if (active && (fj[0] != 0.0 || fj[1] != 0.0 || fj[2] != 0.0)) {
    atomicAdd(&outF[3 * j], sign * fj[0]); atomicAdd(&outF[3 * j + 1], sign * fj[1]); atomicAdd(&outF[3 * j + 2], sign * fj[2]);
  }
  en = block_sum(en, red);
  if (threadIdx.x == 0 && en != 0.0) atomicAdd(&e.item_energy[item_id], en);
  __syncthreads();
  // fold the per-warp accumulators and push the image-atom forces out
  for (int k = threadIdx.x; k < 3 * MA * 3; k += blockDim.x) {
    int grp = k / (MA * 3), a = (k / 3) % MA, c = k % 3;
    int n = grp == 0 ? sh.nd : (grp == 1 ? sh.na : sh.nh);
    if (a >= n) continue;
    double s = 0.0;
    for (int ww = 0; ww < ITEM_TPB / 32; ww++) s += facc[ww][grp * MA + a][c];
    if (s != 0.0) {
      int atom = grp == 0 ? sh.d_atom[a] : (grp == 1 ? sh.a_atom[a] : sh.h_atom[a]);
      atomicAdd(&outF[3 * atom + c], sign * s);
    }
  }
}

// one thread per item: everything that involves only chain atoms (intramolecular terms of donor and acceptor,
// pairs among the chain molecules, repulsion with chain atoms, reference energy)
__global__ void k_evb_items_chain(Dev d, EvbDev e, int n_items) {
  int ir = blockIdx.x * blockDim.x + threadIdx.x;
  if (ir >= n_items) return;
  const int ii = e.real_list[ir];
  const EvbItem it = e.items[ii];
  const Snapshot& S = e.snap[it.state * NLEV + it.level];
  const EvbTables& E = *d.evb;
  double* outF = (it.state == 0) ? d.force : e.dF + (size_t)it.state * 3 * d.N;
  double en = 0.0;
  double fl[CM][MA][3];
  for (int k = 0; k < CM; k++) for (int a = 0; a < MA; a++) fl[k][a][0] = fl[k][a][1] = fl[k][a][2] = 0.0;
  const int ds = it.donor_slot, as = it.acceptor_slot;
  if (ds >= 0) {
    // reference energy of the acid of this topology (ms_evb.f90:1478,1520)
    en += (it.sign < 0) ? E.ref_energy[S.m[ds].mtype] : E.ref_energy[S.m[as].mtype];
    // intramolecular bonded + non-bonded of donor and acceptor (:1472, 1849-1855)
    for (int w = 0; w < 2; w++) {
      int sl = w == 0 ? ds : as;
      const MolImage& I = S.m[sl];
      MolEnergies ME;
      molecule_terms(d, d.mt[I.mtype], I.n_atom, I.x, I.type, I.q, fl[sl], ME, true, true);
      en += ME.e_bond + ME.e_angle + ME.e_dih + ME.e_elec + ME.e_vdw;
    }
    // pairs: donor atoms vs all other chain molecules; acceptor atoms vs chain molecules other than donor, acceptor
    for (int w = 0; w < 2; w++) {
      int sl = w == 0 ? ds : as;
      const MolImage& I = S.m[sl];
      for (int k = 0; k < S.n_mol; k++) {
        if (k == ds || (w == 1 && k == as)) continue;
        const MolImage& J = S.m[k];
        for (int a = 0; a < I.n_atom; a++)
          for (int b = 0; b < J.n_atom; b++) {
            double dr[3] = {min_image(I.x[a][0] - J.x[b][0], d.box[0]), min_image(I.x[a][1] - J.x[b][1], d.box[1]),
                            min_image(I.x[a][2] - J.x[b][2], d.box[2])};
            double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
            if (dr2 < d.rc2) {
              int pidx = I.type[a] * d.nT + J.type[b];
              double ee, ev, f[3];
              pair_terms(d, dr, dr2, I.q[a] * J.q[b], d.vdw_type[pidx], &d.vdw_param[6 * pidx], true, ee, ev, f);
              en += ee + ev;
              for (int c = 0; c < 3; c++) { fl[sl][a][c] += f[c]; fl[k][b][c] -= f[c]; }
            }
          }
      }
    }
  } else {
    en += E.ref_energy[S.m[S.hydronium].mtype];   // principal diabat: E_reference (ms_evb.f90:424)
  }
  // repulsion of the hydronium image with the other chain molecules' atoms
  {
    // build the small lookup locally (same code path as the background kernel, without shared memory)
    static_assert(sizeof(ItemShared) < 8192, "ItemShared too large for local use");
    ItemShared L;
    fill_item_shared(d, S, it, L);
    int hs = S.hydronium;
    for (int k = 0; k < S.n_mol; k++) {
      if (k == hs) continue;
      const MolImage& J = S.m[k];
      for (int b = 0; b < J.n_atom; b++) {
        double fj[3] = {0, 0, 0};
        en += repulsion_with_atom(d, L, J.x[b], J.type[b], fl[hs], fj);
        for (int c = 0; c < 3; c++) fl[k][b][c] += fj[c];
      }
    }
  }
  atomicAdd(&e.item_energy[ii], en);
  for (int k = 0; k < S.n_mol; k++)
    for (int a = 0; a < S.m[k].n_atom; a++)
      for (int c = 0; c < 3; c++)
        if (fl[k][a][c] != 0.0) atomicAdd(&outF[3 * S.m[k].atom[a] + c], it.sign * fl[k][a][c]);
}

// ================================================================================================
// K2: batched delta grids.  Q_slot = Q_principal for every owned diabat (pure streaming copy), then
// -/+ the B-spline patches of the donor / acceptor images (modify_Q_grid, |q| > 1e-6)
// ================================================================================================
__global__ void k_evb_broadcast_grid(const double* __restrict__ src, double* __restrict__ dst, size_t K3, int n_copies) {
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2;
  if (i >= K3) return;
  double2 v = *reinterpret_cast<const double2*>(src + i);
  for (int g = 0; g < n_copies; g++) *reinterpret_cast<double2*>(dst + (size_t)g * K3 + i) = v;
}

// grid = n_items * 2*MA warps; mode 0: spread into Q_slot ; mode 1: reciprocal force correction from theta_slot
__global__ void k_evb_item_pme(Dev d, EvbDev e, int n_items, const int* __restrict__ slot_of_state, int mode) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int ii = w / (2 * MA), r = w % (2 * MA);
  if (ii >= n_items) return;
  const EvbItem it = e.items[ii];
  if (it.donor_slot < 0) return;
  const Snapshot& S = e.snap[it.state * NLEV + it.level];
  const MolImage& I = S.m[r < MA ? it.donor_slot : it.acceptor_slot];
  int a = r % MA;
  if (a >= I.n_atom) return;
  double q = I.q[a];
  double u[3];
  scaled_coords(d, I.x[a], u);
  size_t K3 = (size_t)d.K * d.K * d.K;
  int slot = slot_of_state[it.state];
  if (mode == 0) {
    if (!(fabs(q) > 1e-6)) return;    // modify_Q_grid pme.f90:296
    spread_atom_warp(d, d.Q + K3 * slot, u, q, it.sign, lane);
  } else {
    double F[3];
    gather_atom_warp(d, d.theta + K3 * slot, u, q, lane, F);
    // slot of this atom in the diabat's chain-atom table (level-0 snapshot order)
    const Snapshot& S0 = e.snap[it.state * NLEV];
    int slot = -1, cnt = 0;
    for (int k = 0; k < S0.n_mol; k++) for (int b = 0; b < S0.m[k].n_atom; b++) { if (S0.m[k].atom[b] == I.atom[a]) slot = cnt; cnt++; }
    if (lane < 3 && slot >= 0) {
      double v = lane == 0 ? F[0] : (lane == 1 ? F[1] : F[2]);
      atomicAdd(&e.corr_f[((size_t)it.state * CM * MA + slot) * 3 + lane], it.sign * v);
      if (lane == 0) e.corr_atom[it.state * CM * MA + slot] = I.atom[a];
    }
  }
}

// ================================================================================================
// K11: off-diagonal coupling
// ================================================================================================
struct CouplingGeo {
  double A, Vconst, dA[3][3];
  double rz[3];
  int n_site; int site_atom[2 * MA]; double site_q[2 * MA]; double site_x[2 * MA][3];
  int atom_Od, atom_Oa, atom_H;
  int n_chain; int chain_atoms[CM * MA];
  int valid;
};

// evb_diabatic_coupling_function ms_evb.f90:1180-1266
__device__ void coupling_function(double& A, double& Vc, double dA[3][3], int ftype, const double* fp, const double q[3], const double rOO[3]) {
  double r = sqrt(rOO[0] * rOO[0] + rOO[1] * rOO[1] + rOO[2] * rOO[2]);
  double q2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  double qm = sqrt(q2);
  Vc = fp[0];
  if (ftype == 1) {
    double gamma = fp[1], P = fp[2], k = fp[3], D = fp[4], beta = fp[5], R0 = fp[6], Pp = fp[7], alpha = fp[8], rl0 = fp[9];
    double fac1 = exp(-gamma * q2);
    double g2 = exp(-k * ((r - D) * (r - D)));
    double fac2 = 1.0 + P * g2;
    double e3 = exp(-alpha * (r - rl0));
    double fac3 = 0.5 * (1.0 - tanh(beta * (r - R0))) + Pp * e3;
    double dfac1 = -gamma * 2.0 * qm * fac1;
    double dfac2 = P * -k * 2.0 * (r - D) * g2;
    double ch = cosh(beta * (r - R0));
    double dfac3 = -0.5 * beta / (ch * ch) - Pp * alpha * e3;
    A = fac1 * fac2 * fac3;
    for (int c = 0; c < 3; c++) {
      double tq = dfac1 * fac2 * fac3 * 0.5 * q[c] / qm;
      double t2 = fac1 * dfac2 * fac3 * rOO[c] / r, t3 = fac1 * fac2 * dfac3 * rOO[c] / r;
      dA[0][c] = tq + t2 + t3;
      dA[1][c] = tq - t2 - t3;
      dA[2][c] = dfac1 * fac2 * fac3 * -q[c] / qm;
    }
  } else {
    double gamma = fp[1], k = fp[2], D = fp[3];
    double fac1 = exp(-gamma * q2), fac2 = exp(-k * ((r - D) * (r - D)));
    double dfac1 = -gamma * 2.0 * qm * fac1, dfac2 = -k * 2.0 * (r - D) * fac2;
    A = fac1 * fac2;
    for (int c = 0; c < 3; c++) {
      double tq = dfac1 * fac2 * 0.5 * q[c] / qm, t2 = fac1 * dfac2 * rOO[c] / r;
      dA[0][c] = tq + t2; dA[1][c] = tq - t2; dA[2][c] = dfac1 * fac2 * -q[c] / qm;
    }
  }
}

// one thread per owned diabat s>=1: geometric factor, Zundel sites, and the Vex terms of the OTHER chain molecules
__global__ void k_evb_coupling_geo(Dev d, EvbDev e, CouplingGeo* geo) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int S = *e.n_states;
  if (s >= S) return;
  CouplingGeo& G = geo[s];
  G.valid = 0;
  if (s == 0 || !state_owned(s, d.rank, d.world)) return;
  const EvbTables& E = *d.evb;
  int nh = e.n_hops[s];
  const Snapshot& Sn = e.snap[s * NLEV + nh];
  const Snapshot& Sp = e.snap[s * NLEV + nh - 1];
  int as = Sn.hydronium, ds = Sp.hydronium;    // last acceptor / last donor
  const MolImage& D = Sn.m[ds];
  const MolImage& A = Sn.m[as];
  int iOd = d.mt[D.mtype].heavy_base_atom, iOa = d.mt[A.mtype].heavy_acid_atom, iH = A.n_atom - 1;
  if (iOd < 0 || iOa < 0) { atomicMax(&d.err_flag[3], 4); return; }
  // ---- geometric factor (ms_evb.f90:1117-1174)
  double rO1[3], rO2[3], rH[3], shift[3], rOO[3], q[3];
  for (int k = 0; k < 3; k++) {
    rO1[k] = D.x[iOd][k];
    double dr = A.x[iOa][k] - rO1[k];
    shift[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
    rO2[k] = rO1[k] + (A.x[iOa][k] - rO1[k] - shift[k]);
    rH[k] = rO1[k] + (A.x[iH][k] - rO1[k] - shift[k]);
    rOO[k] = rO1[k] - rO2[k];
    q[k] = (rO1[k] + rO2[k]) / 2.0 - rH[k];
  }
  int row = -1;
  for (int i = 0; i < RPB_MAXI; i++) {
    if (E.dc_int[i][0] < 0) break;
    if (E.dc_int[i][0] == D.type[iOd] && E.dc_int[i][1] == A.type[iOa] && E.dc_int[i][2] == A.type[iH]) { row = i; break; }
  }
  if (row < 0) { atomicMax(&d.err_flag[3], 5); return; }
  coupling_function(G.A, G.Vconst, G.dA, E.dc_type[row], E.dc_par[row], q, rOO);
  G.atom_Od = D.atom[iOd]; G.atom_Oa = A.atom[iOa]; G.atom_H = A.atom[iH];
  // ---- Zundel centre of mass and exchange-charge sites (ms_evb.f90:2946-2982, 1340-1392)
  double tmd = 0, tma = 0;
  for (int a = 0; a < D.n_atom; a++) tmd = tmd + D.mass[a];
  for (int a = 0; a < A.n_atom; a++) tma = tma + A.mass[a];
  double shifta[3];
  for (int k = 0; k < 3; k++) {
    double dr = A.r_com[k] - D.r_com[k];
    shifta[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
    double rca = D.r_com[k] + (A.r_com[k] - D.r_com[k] - shifta[k]);
    G.rz[k] = (tmd * D.r_com[k] + tma * rca) / (tmd + tma);
  }
  double qx = E.exch_proton[A.mtype][D.mtype];
  G.n_site = 0;
  for (int a = 0; a < D.n_atom; a++) {
    int n = G.n_site++;
    G.site_atom[n] = D.atom[a]; G.site_q[n] = E.exch_atomic[D.type[a]];
    for (int k = 0; k < 3; k++) { double dr = D.x[a][k] - G.rz[k] - 0.0; G.site_x[n][k] = G.rz[k] + dr; }
  }
  for (int a = 0; a < A.n_atom; a++) {
    int n = G.n_site++;
    G.site_atom[n] = A.atom[a]; G.site_q[n] = (a == A.n_atom - 1) ? qx : E.exch_atomic[A.type[a]];
    for (int k = 0; k < 3; k++) { double dr = A.x[a][k] - G.rz[k] - shifta[k]; G.site_x[n][k] = G.rz[k] + dr; }
  }
  G.n_chain = 0;
  for (int k = 0; k < Sn.n_mol; k++) for (int a = 0; a < Sn.m[k].n_atom; a++) G.chain_atoms[G.n_chain++] = Sn.m[k].atom[a];
  // ---- Vex with the other chain molecules (final-level charges, positions, centres of mass)
  double vex = 0.0;
  double* Fo = e.Foff + (size_t)s * 3 * d.N;
  for (int km = 0; km < Sn.n_mol; km++) {
    if (km == ds || km == as) continue;
    const MolImage& J = Sn.m[km];
    double sh[3];
    for (int k = 0; k < 3; k++) sh[k] = floor(d.inv_box[k] * (J.r_com[k] - G.rz[k]) + 0.5) * d.box[k];
    for (int b = 0; b < J.n_atom; b++)
      for (int n = 0; n < G.n_site; n++) {
        double r[3];
        for (int k = 0; k < 3; k++) r[k] = -(J.x[b][k] - G.site_x[n][k] - sh[k]);
        double rm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        double qq = G.site_q[n] * J.q[b];
        vex += qq / rm * d.conv;
        for (int k = 0; k < 3; k++) {
          double dV = -qq / (rm * rm * rm) * r[k] * d.conv;
          atomicAdd(&Fo[3 * G.site_atom[n] + k], -G.A * dV);
          atomicAdd(&Fo[3 * J.atom[b] + k], G.A * dV);
        }
      }
  }
  atomicAdd(&e.vex[s], vex);
  G.valid = 1;
}

// grid = (n_owned states list, ceil(N/256)): Vex between the Zundel sites and the background atoms
__global__ void __launch_bounds__(256) k_evb_coupling_vex(Dev d, EvbDev e, const CouplingGeo* geo, const int* state_list) {
  __shared__ double red[32];
  __shared__ double facc[8][2 * MA][3];
  int s = state_list[blockIdx.x];
  const CouplingGeo& G = geo[s];
  if (!G.valid) return;
  for (int k = threadIdx.x; k < 8 * 2 * MA * 3; k += blockDim.x) (&facc[0][0][0])[k] = 0.0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int j = blockIdx.y * blockDim.x + threadIdx.x;
  bool active = j < d.N;
  if (active) for (int k = 0; k < G.n_chain; k++) active &= (G.chain_atoms[k] != j);
  double vex = 0.0, fj[3] = {0, 0, 0};
  double xj[3] = {0, 0, 0}, qj = 0.0, sh[3] = {0, 0, 0};
  if (active) {
    double4 p = d.xq[j];
    xj[0] = p.x; xj[1] = p.y; xj[2] = p.z; qj = p.w;
    int jm = d.mol_of_atom[j];
    for (int k = 0; k < 3; k++) sh[k] = floor(d.inv_box[k] * (d.r_com[3 * jm + k] - G.rz[k]) + 0.5) * d.box[k];
  }
  for (int n = 0; n < G.n_site; n++) {
    double dV[3] = {0, 0, 0};
    if (active) {
      double r[3];
      for (int k = 0; k < 3; k++) r[k] = -(xj[k] - G.site_x[n][k] - sh[k]);
      double rm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      double qq = G.site_q[n] * qj;
      vex += qq / rm * d.conv;
      double g = -qq / (rm * rm * rm) * d.conv;
      for (int k = 0; k < 3; k++) { dV[k] = g * r[k]; fj[k] -= dV[k]; }
    }
    double s0 = warp_sum(dV[0]), s1 = warp_sum(dV[1]), s2 = warp_sum(dV[2]);
    if (lane == 0) { facc[w][n][0] += s0; facc[w][n][1] += s1; facc[w][n][2] += s2; }
  }
  double* Fo = e.Foff + (size_t)s * 3 * d.N;
  if (active) { Fo[3 * j] = -G.A * fj[0]; Fo[3 * j + 1] = -G.A * fj[1]; Fo[3 * j + 2] = -G.A * fj[2]; }  // only this thread writes atom j
  vex = block_sum(vex, red);
  if (threadIdx.x == 0) atomicAdd(&e.vex[s], vex);
  __syncthreads();
  for (int k = threadIdx.x; k < G.n_site * 3; k += blockDim.x) {
    int n = k / 3, c = k % 3;
    double t = 0.0;
    for (int ww = 0; ww < 8; ww++) t += facc[ww][n][c];
    atomicAdd(&Fo[3 * G.site_atom[n] + c], -G.A * t);
  }
}

// one thread per owned diabat: H_ss, H_parent,s and the geometric part of the coupling force
__global__ void k_evb_assemble(Dev d, EvbDev e, const CouplingGeo* geo, int n_items, const int* slot_of_state) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int S = *e.n_states;
  if (s >= MAXS) return;
  e.h_diag[s] = 0.0; e.h_diag[MAXS + s] = 0.0; e.h_diag[2 * MAXS + s] = 0.0;
  if (s >= S || !state_owned(s, d.rank, d.world)) return;
  if (s == 0) {
    // principal energy: calculate_total_force_energy + repulsion + reference (ms_evb.f90:411-436)
    double E_elec = d.en[E_ELEC] + d.en[E_RECIP] + d.ewald_self;
    double H11 = E_elec + d.en[E_VDW] + d.en[E_BOND] + d.en[E_ANGLE] + d.en[E_DIH];
    for (int ii = 0; ii < n_items; ii++) if (e.items[ii].state == 0) H11 = H11 + e.item_energy[ii];
    e.h_diag[0] = H11;
    return;
  }
  // energy delta of the last hop: acceptor-topology item minus donor-topology item (ms_evb.f90:1546)
  double dE = 0.0;
  for (int ii = 0; ii < n_items; ii++) {
    const EvbItem& it = e.items[ii];
    if (it.state == s && it.real && it.sign > 0) dE = e.item_energy[ii] - e.item_energy[ii - 1];
  }
  e.h_diag[s] = dE;
  e.h_diag[2 * MAXS + s] = e.e_recip[slot_of_state[s]] - e.e_recip[0];
  const CouplingGeo& G = geo[s];
  double pref = G.Vconst + e.vex[s];
  e.h_diag[MAXS + s] = pref * G.A;
  double* Fo = e.Foff + (size_t)s * 3 * d.N;
  for (int c = 0; c < 3; c++) {
    atomicAdd(&Fo[3 * G.atom_Od + c], -pref * G.dA[0][c]);
    atomicAdd(&Fo[3 * G.atom_Oa + c], -pref * G.dA[1][c]);
    atomicAdd(&Fo[3 * G.atom_H + c], -pref * G.dA[2][c]);
  }
}

// ================================================================================================
// K12: block-level Jacobi eigensolver for the (<= 80 x 80) EVB Hamiltonian, one CTA, matrix in shared memory.
// Same rotation formulas, thresholds and stopping logic as the reference's Numerical-Recipes routine
// (general_routines.f90:2035-2074), but the n/2 disjoint rotations of a round-robin round are applied
// concurrently (rows, then columns) instead of one (ip,iq) at a time; the ground-state eigenvector is unique up
// to sign, so only the rounding-level path differs.  Followed by the reference's selections (ms_evb.f90:279-328):
// ground state = first minimum eigenvalue, principal diabat = first maximum |c_i|.
// ================================================================================================
#define JAC_TPB 256
__global__ void __launch_bounds__(JAC_TPB) k_evb_jacobi(Dev d, EvbDev e, const double* coeff_override) {
  extern __shared__ double smem[];
  const int S = *e.n_states;
  const int n = S + (S & 1);          // padded to even for the round-robin pairing (dummy row/column stays zero)
  const int tid = threadIdx.x, nth = blockDim.x;
  double* a = smem;                   // [n*n] column-major, full symmetric
  double* v = a + n * n;              // [n*n]
  double* rc = v + n * n;             // [n/2] cos
  double* rs = rc + n / 2;            // [n/2] sin
  int* rp = (int*)(rs + n / 2);       // [n/2] p
  int* rq = rp + n / 2;               // [n/2] q   (q < 0: no rotation)
  __shared__ double red[32];
  __shared__ double s_sm;
  __shared__ int s_nrot;
  if (coeff_override) {
    for (int i = tid; i < S; i += nth) e.evec[i] = coeff_override[i];
  } else {
    for (int k = tid; k < n * n; k += nth) { a[k] = 0.0; v[k] = 0.0; }
    __syncthreads();
    for (int i = tid; i < n; i += nth) {
      v[i + n * i] = 1.0;
      if (i < S) {
        // H_ss = H_11 + sum of the hop deltas along the chain (root first) + (E_rec(s) - E_rec(1))   ms_evb.f90:1546, 2083
        int chain[MAXC + 1], nc = 0;
        for (int t = i; t > 0 && nc <= MAXC; t = e.parent[t]) chain[nc++] = t;
        double Hs = e.h_diag[0];
        for (int k = nc - 1; k >= 0; k--) Hs = Hs + e.h_diag[chain[k]];
        Hs = Hs + e.h_diag[2 * MAXS + i];
        a[i + n * i] = Hs;
        e.h_full[i] = Hs; e.h_full[MAXS + i] = (i > 0) ? e.h_diag[MAXS + i] : 0.0;
        if (i > 0) { int p = e.parent[i]; a[p + n * i] = e.h_diag[MAXS + i]; a[i + n * p] = e.h_diag[MAXS + i]; }
      }
    }
    __syncthreads();
    int status = 1;
    for (int it = 1; it <= 50; it++) {
      double sm = 0.0;
      for (int k = tid; k < n * n; k += nth) { int i = k % n, j = k / n; if (i < j) sm += fabs(a[k]); }
      sm = block_sum(sm, red);
      if (tid == 0) { s_sm = sm; s_nrot = 0; }
      __syncthreads();
      sm = s_sm;
      if (sm == 0.0) { status = 0; break; }
      const double tresh = (it < 4) ? 0.2 * sm / (double)(S * S) : 0.0;
      for (int r = 0; r < n - 1; r++) {
        // ---- phase A: rotation parameters of the n/2 disjoint pairs of this round
        if (tid < n / 2) {
          int t = tid, p, q;
          if (t == 0) { p = r; q = n - 1; }
          else { p = (r + t) % (n - 1); q = (r - t + (n - 1)) % (n - 1); }
          if (p > q) { int x = p; p = q; q = x; }
          double apq = a[p + n * q];
          int rot = 0;
          double cc = 1.0, sn = 0.0;
          if (apq != 0.0) {
            double app = a[p + n * p], aqq = a[q + n * q];
            double g = 100.0 * fabs(apq);
            if (it > 4 && (fabs(app) + g == fabs(app)) && (fabs(aqq) + g == fabs(aqq))) {
              a[p + n * q] = 0.0; a[q + n * p] = 0.0;
            } else if (fabs(apq) > tresh) {
              double h = aqq - app, tt;
              if (fabs(h) + g == fabs(h)) tt = apq / h;
              else {
                double theta = 0.5 * h / apq;
                tt = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                if (theta < 0.0) tt = -tt;
              }
              cc = 1.0 / sqrt(1 + tt * tt); sn = tt * cc;
              rot = 1;
            }
          }
          rc[t] = cc; rs[t] = sn; rp[t] = p; rq[t] = rot ? q : -1;
          if (rot) atomicAdd(&s_nrot, 1);
        }
        __syncthreads();
        // ---- phase B: rows p,q of A <- J^T A
        for (int w = tid; w < (n / 2) * n; w += nth) {
          int t = w / n, k = w - t * n, q = rq[t];
          if (q < 0) continue;
          int p = rp[t];
          double cc = rc[t], sn = rs[t];
          double x = a[p + n * k], y = a[q + n * k];
          a[p + n * k] = cc * x - sn * y;
          a[q + n * k] = sn * x + cc * y;
        }
        __syncthreads();
        // ---- phase C: columns p,q of A <- A J ; V <- V J
        for (int w = tid; w < (n / 2) * n; w += nth) {
          int t = w / n, k = w - t * n, q = rq[t];
          if (q < 0) continue;
          int p = rp[t];
          double cc = rc[t], sn = rs[t];
          double x = a[k + n * p], y = a[k + n * q];
          double xn = cc * x - sn * y, yn = sn * x + cc * y;
          if (k == p) yn = 0.0;       // a(p,q) = 0 exactly, as in the reference (:2062)
          if (k == q) xn = 0.0;
          a[k + n * p] = xn; a[k + n * q] = yn;
          double vx = v[k + n * p], vy = v[k + n * q];
          v[k + n * p] = cc * vx - sn * vy;
          v[k + n * q] = sn * vx + cc * vy;
        }
        __syncthreads();
      }
    }
    if (tid == 0) {
      int ground = 0;
      double e0 = a[0];
      for (int i = 1; i < S; i++) if (a[i + n * i] < e0) { e0 = a[i + n * i]; ground = i; }
      *e.e_ground = e0;
      int pd = 0;
      double coef = fabs(v[0 + n * ground]);
      for (int i = 0; i < S; i++) if (coef < fabs(v[i + n * ground])) { coef = fabs(v[i + n * ground]); pd = i; }
      int newh = *d.hydronium;
      for (int h = 0; h < d.max_chain; h++) {
        if (e.proton_log[(pd * MAXC + h) * 5] < 0) break;
        newh = e.proton_log[(pd * MAXC + h) * 5 + 3];
      }
      e.result[0] = pd; e.result[1] = newh; e.result[2] = status; e.result[3] = ground;
    }
    __syncthreads();
    int ground = e.result[3];
    for (int i = tid; i < S; i += nth) e.evec[i] = v[i + n * ground];
  }
  __syncthreads();
  // Hellmann-Feynman weights: c_s^2 (diagonal) and 2 c_parent c_s (coupling)   ms_evb.f90:298-303
  for (int i = tid; i < MAXS; i += nth) {
    double ci = i < S ? e.evec[i] : 0.0;
    e.coef2[i] = ci * ci;
    e.coef2[MAXS + i] = (i > 0 && i < S) ? 2.0 * e.evec[e.parent[i]] * ci : 0.0;
    e.coef2[2 * MAXS + i] = ci * ci;
  }
  __syncthreads();
  // weight of the last-hop force delta of diabat s = sum of c_t^2 over every diabat whose chain passes through s
  // (DFS pre-order => parent(s) < s, so one descending pass accumulates the subtrees)
  if (tid == 0) for (int i = S - 1; i >= 1; i--) e.coef2[2 * MAXS + e.parent[i]] += e.coef2[2 * MAXS + i];
}

// ================================================================================================
// K13: Hellmann-Feynman mixing
// ================================================================================================
// theta_mix = sum over owned grids of c_s^2 theta_slot   (streaming: reads n_slots*K^3, writes K^3)
__global__ void k_evb_theta_mix(Dev d, EvbDev e, const int* __restrict__ slot_state, int n_slots) {
  size_t K3 = (size_t)d.K * d.K * d.K;
  size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2;
  if (i >= K3) return;
  double ax = 0.0, ay = 0.0;
  for (int g = 0; g < n_slots; g++) {
    int s = slot_state[g];
    if (s < 0) continue;
    double w = e.coef2[s];
    double2 t = *reinterpret_cast<const double2*>(d.theta + (size_t)g * K3 + i);
    ax = fma(w, t.x, ax); ay = fma(w, t.y, ay);
  }
  *reinterpret_cast<double2*>(e.theta_mix + i) = make_double2(ax, ay);
}

// f_mix = [rank 0: principal force] + sum_s c_s^2 dF_s + 2 c_p c_s Foff_s   over owned diabats
__global__ void k_evb_mix_forces(Dev d, EvbDev e, const int* __restrict__ state_list, int n_list, int include_principal) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n3 = (size_t)3 * d.N;
  if (i >= n3) return;
  double f = include_principal ? e.dF[i] : 0.0;   // dF slot 0 holds the principal-diabat force (without F_rec)
  for (int k = 0; k < n_list; k++) {
    int s = state_list[k];
    f = fma(e.coef2[2 * MAXS + s], e.dF[(size_t)s * n3 + i], f);
    f = fma(e.coef2[MAXS + s], e.Foff[(size_t)s * n3 + i], f);
  }
  e.f_mix[i] = f;
}

// + sum_s c_s^2 * (reciprocal-space corrections of the chain atoms of diabat s)   ms_evb.f90:2103-2248
__global__ void k_evb_add_corr(Dev d, EvbDev e, const int* __restrict__ state_list, int n_list) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_list * CM * MA) return;
  int s = state_list[t / (CM * MA)], slot = t % (CM * MA);
  int atom = e.corr_atom[s * CM * MA + slot];
  if (atom < 0) return;
  double w = e.coef2[s];
  for (int c = 0; c < 3; c++) atomicAdd(&e.f_mix[3 * atom + c], w * e.corr_f[((size_t)s * CM * MA + slot) * 3 + c]);
}

__global__ void k_evb_gather_mix(Dev d, EvbDev e) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= d.N) return;
  double u[3] = {d.uscale[3 * w], d.uscale[3 * w + 1], d.uscale[3 * w + 2]};
  double F[3];
  gather_atom_warp(d, e.theta_mix, u, d.xq[w].w, lane, F);
  if (lane < 3) {
    double v = lane == 0 ? F[0] : (lane == 1 ? F[1] : F[2]);
    e.f_mix[3 * w + lane] += v;
  }
}

__global__ void k_copy(double* dst, const double* src, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// ================================================================================================
// K15: hop commit -- permute the per-atom arrays exactly as shift_array_data_donor_acceptor_transfer
// (ms_evb.f90:2677-2840) and retype / reorder the acceptor from the final-level snapshot.
// ================================================================================================
__global__ void k_evb_commit_permute(Dev d, const int* __restrict__ perm, double4* xq_new, double* vel_new, double* force_new,
                                     double* mass_new, int* type_new, int* moa_new, const int* __restrict__ moa_src) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.N) return;
  int o = perm[i];
  xq_new[i] = d.xq[o];
  for (int k = 0; k < 3; k++) { vel_new[3 * i + k] = d.vel[3 * o + k]; force_new[3 * i + k] = d.force[3 * o + k]; }
  mass_new[i] = d.mass[o];
  type_new[i] = d.type[o];
  moa_new[i] = moa_src[i];
}

// after the permutation: write snapshot data (positions made whole, charges, types, centres of mass) of the chain molecules
__global__ void k_evb_commit_patch(Dev d, EvbDev e, int state, int level, const int* __restrict__ new_first) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Snapshot& S = e.snap[state * NLEV + level];
  for (int k = 0; k < S.n_mol; k++) {
    const MolImage& I = S.m[k];
    int f = new_first[k];
    for (int a = 0; a < I.n_atom; a++) {
      d.xq[f + a] = make_double4(I.x[a][0], I.x[a][1], I.x[a][2], I.q[a]);
      d.type[f + a] = I.type[a];
    }
    for (int c = 0; c < 3; c++) d.r_com[3 * I.mol + c] = I.r_com[c];
    d.mol_first[I.mol] = f; d.mol_natom[I.mol] = I.n_atom; d.mol_type[I.mol] = I.mtype;
  }
  *d.hydronium = S.m[S.hydronium].mol;
}

// ================================================================================================
// host orchestration
// ================================================================================================
struct EvbScratch {   // device scratch owned by the context (allocated in evb_alloc)
  CouplingGeo* geo;
  int* slot_of_state;   // [MAXS]
  int* slot_state;      // [MAXS] inverse map (slot -> state), -1 unused
  int* state_list;      // [MAXS] owned diabats s>=1
  double* coeff_dev;    // [MAXS]
  int* perm; int* new_first;
  double4* xq2; double* vel2; double* force2; double* mass2; int* type2; int* moa2;
};
static std::map<rpb_ctx*, EvbScratch> g_scratch;

#define CKE(call)                                                                 \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
      return RPB_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

int evb_alloc(rpb_ctx* c) {
  if (c->e.n_states) return 0;
  EvbDev& e = c->e;
  const int N = c->d.N;
  const size_t K3 = (size_t)c->d.K * c->d.K * c->d.K;
  int rc;
#define AL(p, n) if ((rc = dev_alloc(c, &(p), (size_t)(n)))) return rc;
  AL(e.n_states, 1); AL(e.proton_log, MAXS * MAXC * 5); AL(e.parent, MAXS); AL(e.n_hops, MAXS);
  AL(e.snap, MAXS * NLEV); AL(e.items, RPB_MAX_ITEMS + 1); AL(e.n_items, 1); AL(e.item_energy, RPB_MAX_ITEMS + 1);
  AL(e.real_list, RPB_MAX_ITEMS + 1); AL(e.n_real, 1); AL(e.corr_f, (size_t)MAXS * CM * MA * 3); AL(e.corr_atom, MAXS * CM * MA); AL(e.h_full, 2 * MAXS);
  AL(e.dF, (size_t)MAXS * 3 * N); AL(e.Foff, (size_t)MAXS * 3 * N);
  AL(e.vex, MAXS); AL(e.e_recip, MAXS); AL(e.h_diag, 3 * MAXS); AL(e.f_mix, 3 * N); AL(e.evec, MAXS); AL(e.coef2, 3 * MAXS);
  AL(e.e_ground, 1); AL(e.result, 8); AL(e.theta_mix, K3);
  EvbScratch s;
  AL(s.geo, MAXS); AL(s.slot_of_state, MAXS); AL(s.slot_state, MAXS); AL(s.state_list, MAXS); AL(s.coeff_dev, MAXS);
  AL(s.perm, N); AL(s.new_first, CM);
  AL(s.xq2, N); AL(s.vel2, 3 * N); AL(s.force2, 3 * N); AL(s.mass2, N); AL(s.type2, N); AL(s.moa2, N);
#undef AL
  g_scratch[c] = s;
  CKE(cudaMallocHost(&c->eh.pinned, (16 + MAXS * (2 + MAXC * 5)) * sizeof(int) + 4 * MAXS * sizeof(double)));
  CKE(cudaMemset(e.n_states, 0, sizeof(int)));
  return 0;
}

static void host_items(rpb_ctx* c, std::vector<EvbItem>& items) {
  EvbHost& h = c->eh;
  items.clear();
  EvbItem p; p.state = 0; p.level = 0; p.donor_slot = -1; p.acceptor_slot = -1; p.sign = 1.0; p.real = 1; p.pad = 0;
  items.push_back(p);   // principal diabat: EVB repulsion + reference energy (ms_evb.f90:418-426)
  for (int s = 1; s < h.n_states; s++) {
    if (!state_owned(s, c->d.rank, c->d.world)) continue;
    int mols[CM], nm = 1;
    mols[0] = c->hydronium_mol;
    for (int k = 0; k < h.n_hops[s]; k++) {
      int a = h.proton_log[s][k][3];
      bool f = false;
      for (int q = 0; q < nm; q++) f |= (mols[q] == a);
      if (!f) mols[nm++] = a;
    }
    int cur = 0;
    for (int k = 0; k < h.n_hops[s]; k++) {
      int a = h.proton_log[s][k][3], as = 0;
      for (int q = 0; q < nm; q++) if (mols[q] == a) as = q;
      EvbItem it; it.state = s; it.donor_slot = cur; it.acceptor_slot = as; it.pad = 0;
      it.real = (k == h.n_hops[s] - 1) ? 1 : 0;   // earlier hops' real-space deltas are the ancestors' (same images, same background)
      it.level = k; it.sign = -1.0; items.push_back(it);
      it.level = k + 1; it.sign = 1.0; items.push_back(it);
      cur = as;
    }
  }
}

int evb_build(rpb_ctx* c) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = g_scratch[c];
  const int N = d.N;
  const size_t K3 = (size_t)d.K * d.K * d.K, n3 = (size_t)3 * N;
  int rc = calculate_total_force_energy(c, true);
  if (rc) return rc;
  {
    ScopedTimer t(c, T_EVB_ENUM);
    k_evb_enumerate<<<1, 256, 0, c->stream>>>(d, e);
    c->n_launch++;
    int* pin = h.pinned;
    CKE(cudaMemcpyAsync(pin, e.n_states, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaMemcpyAsync(pin + 16, e.n_hops, MAXS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaMemcpyAsync(pin + 16 + MAXS, e.parent, MAXS * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaMemcpyAsync(pin + 16 + 2 * MAXS, e.proton_log, MAXS * MAXC * 5 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaMemcpyAsync(c->h_flags, d.err_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CKE(cudaStreamSynchronize(c->stream));
    if (c->h_flags[1]) { c->err = "please increase size of verlet neighbor list"; return RPB_ERR_VERLET; }
    if (c->h_flags[2]) { c->err = "Found more diabat states than the current setting of evb_max_states"; return RPB_ERR_DIABATS; }
    if (c->h_flags[3]) { c->err = "error in subroutine find_bonded_atom_hydrogen"; return RPB_ERR_STATE; }
    h.n_states = pin[0];
    memcpy(h.n_hops, pin + 16, MAXS * sizeof(int));
    memcpy(h.parent, pin + 16 + MAXS, MAXS * sizeof(int));
    memcpy(h.proton_log, pin + 16 + 2 * MAXS, MAXS * MAXC * 5 * sizeof(int));
  }
  const int S = h.n_states;
  std::vector<EvbItem> items;
  host_items(c, items);
  h.n_items = (int)items.size();
  std::vector<int> real_list;
  for (int i = 0; i < h.n_items; i++) if (items[i].real) real_list.push_back(i);
  const int n_real = (int)real_list.size();
  CKE(cudaMemcpyAsync(e.real_list, real_list.data(), n_real * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  std::vector<int> slot_of_state(MAXS, 0), slot_state(MAXS, -1), state_list;
  slot_state[0] = 0;   // slot 0 = principal grid on every rank
  for (int s = 1; s < S; s++)
    if (state_owned(s, d.rank, d.world)) { state_list.push_back(s); slot_of_state[s] = (int)state_list.size(); slot_state[state_list.size()] = s; }
  const int n_own = (int)state_list.size();
  if (n_own + 1 > c->grid_capacity) { c->err = "grid capacity exceeded"; return RPB_ERR_DIABATS; }
  state_list.resize(MAXS, 0);
  CKE(cudaMemcpyAsync(e.items, items.data(), items.size() * sizeof(EvbItem), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(sc.slot_of_state, slot_of_state.data(), MAXS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(sc.slot_state, slot_state.data(), MAXS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(sc.state_list, state_list.data(), MAXS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  {
    CKE(cudaMemsetAsync(e.item_energy, 0, (RPB_MAX_ITEMS + 1) * sizeof(double), c->stream));
    CKE(cudaMemsetAsync(e.vex, 0, MAXS * sizeof(double), c->stream));
    CKE(cudaMemsetAsync(e.e_recip, 0, MAXS * sizeof(double), c->stream));
    CKE(cudaMemsetAsync(e.dF, 0, (size_t)S * n3 * sizeof(double), c->stream));
    CKE(cudaMemsetAsync(e.Foff, 0, (size_t)S * n3 * sizeof(double), c->stream));
    k_evb_snapshots<<<(S + 31) / 32, 32, 0, c->stream>>>(d, e, -1);
    CKE(cudaMemsetAsync(e.corr_f, 0, (size_t)S * CM * MA * 3 * sizeof(double), c->stream));
    CKE(cudaMemsetAsync(e.corr_atom, 0xff, (size_t)S * CM * MA * sizeof(int), c->stream));
    dim3 g(n_real, (N + ITEM_TPB - 1) / ITEM_TPB);
    { ScopedTimer t(c, T_EVB_ITEMS_BG); k_evb_items_background<<<g, ITEM_TPB, 0, c->stream>>>(d, e, n_real); }
    { ScopedTimer t(c, T_EVB_ITEMS_CHAIN); k_evb_items_chain<<<(n_real + 31) / 32, 32, 0, c->stream>>>(d, e, n_real); }
    c->n_launch += 3;
  }
  if (n_own > 0) {
    {
      { ScopedTimer t(c, T_EVB_BCAST); k_evb_broadcast_grid<<<(unsigned)((K3 / 2 + 255) / 256), 256, 0, c->stream>>>(d.Q, d.Q + K3, K3, n_own); }
      int warps = h.n_items * 2 * MA;
      ScopedTimer t(c, T_EVB_PATCH);
      k_evb_item_pme<<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(d, e, h.n_items, sc.slot_of_state, 0);
      c->n_launch += 2;
    }
    // e_recip[0] = principal E_rec (already in d.en[E_RECIP]); slots 1..n_own batched
    rc = launch_convolve(c, 1, n_own, e.e_recip, true);
    if (rc) return rc;
    {
      ScopedTimer t(c, T_EVB_CORR);
      int warps = h.n_items * 2 * MA;
      k_evb_item_pme<<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(d, e, h.n_items, sc.slot_of_state, 1);
      c->n_launch += 1;
    }
  }
  {
    ScopedTimer t(c, T_EVB_COUPLING);
    k_copy<<<1, 32, 0, c->stream>>>(e.e_recip, d.en + E_RECIP, 1);
    k_evb_coupling_geo<<<(S + 31) / 32, 32, 0, c->stream>>>(d, e, sc.geo);
    c->n_launch += 2;
    if (n_own > 0) {
      dim3 g(n_own, (N + 255) / 256);
      k_evb_coupling_vex<<<g, 256, 0, c->stream>>>(d, e, sc.geo, sc.state_list);
      c->n_launch += 1;
    }
    k_evb_assemble<<<(MAXS + 31) / 32, 32, 0, c->stream>>>(d, e, sc.geo, h.n_items, sc.slot_of_state);
    c->n_launch += 1;
  }
  // keep the principal-diabat force (incl. EVB repulsion, without reciprocal part) in dF slot 0: d.force is
  // overwritten with the adiabatic force at commit time
  k_copy<<<(unsigned)((n3 + 255) / 256), 256, 0, c->stream>>>(e.dF, d.force, n3);
  c->n_launch++;
  h.built = true;
  return 0;
}

int evb_mix(rpb_ctx* c, const double* coeff_override_host, double* force_out_host) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = g_scratch[c];
  if (!h.built) { c->err = "evb_mix before evb_build"; return RPB_ERR_STATE; }
  const int N = d.N, S = h.n_states;
  const size_t K3 = (size_t)d.K * d.K * d.K, n3 = (size_t)3 * N;
  int n_own = 0;
  for (int s = 1; s < S; s++) if (state_owned(s, d.rank, d.world)) n_own++;
  const double* coeff_dev = nullptr;
  if (coeff_override_host) {
    CKE(cudaMemcpyAsync(sc.coeff_dev, coeff_override_host, S * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    coeff_dev = sc.coeff_dev;
  }
  {
    ScopedTimer t(c, T_EVB_DIAG);
    const int np = S + (S & 1);
    size_t shmem = ((size_t)2 * np * np + 2 * np) * sizeof(double);
    if (shmem > 48 * 1024) cudaFuncSetAttribute(k_evb_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem);
    k_evb_jacobi<<<1, JAC_TPB, shmem, c->stream>>>(d, e, coeff_dev);
    c->n_launch++;
  }
  {
    int include_principal = (d.rank == 0) ? 1 : 0;
    // slot 0 (principal theta) only contributes on rank 0: mask it on the other ranks through slot_state
    if (!include_principal) { int m1 = -1; CKE(cudaMemcpyAsync(sc.slot_state, &m1, sizeof(int), cudaMemcpyHostToDevice, c->stream)); }
    { ScopedTimer t(c, T_EVB_THETAMIX); k_evb_theta_mix<<<(unsigned)((K3 / 2 + 255) / 256), 256, 0, c->stream>>>(d, e, sc.slot_state, n_own + 1); }
    { ScopedTimer t(c, T_EVB_MIXF); k_evb_mix_forces<<<(unsigned)((n3 + 255) / 256), 256, 0, c->stream>>>(d, e, sc.state_list, n_own, include_principal);
      if (n_own > 0) { k_evb_add_corr<<<(n_own * CM * MA + 127) / 128, 128, 0, c->stream>>>(d, e, sc.state_list, n_own); c->n_launch++; } }
    { ScopedTimer t(c, T_EVB_GATHERMIX); k_evb_gather_mix<<<(N * 32 + 255) / 256, 256, 0, c->stream>>>(d, e); }
    c->n_launch += 3;
  }
  if (coeff_override_host) {
    CKE(cudaStreamSynchronize(c->stream));
    CKE(cudaMemcpy(force_out_host, e.f_mix, n3 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
  }
  // read back what the host needs for the commit decision and the accessors
  double* pd = (double*)(h.pinned + 16 + MAXS * (2 + MAXC * 5));
  CKE(cudaMemcpyAsync(h.pinned + 1, e.result, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaMemcpyAsync(pd, e.e_ground, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaMemcpyAsync(pd + 1, e.evec, MAXS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaMemcpyAsync(pd + 1 + MAXS, e.h_full, 2 * MAXS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaMemcpyAsync(c->h_flags, d.err_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaMemcpyAsync(c->h_en, d.en, E_NSLOT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaStreamSynchronize(c->stream));
  if (c->h_flags[3]) { c->err = "couldn't find index in subroutine 'get_index_atom_set' (code " + std::to_string(c->h_flags[3]) + ")"; return RPB_ERR_STATE; }
  if (h.pinned[3]) { c->err = "too many iterations in jacobi"; return RPB_ERR_STATE; }
  h.principal_diabat = h.pinned[1]; h.new_hydronium = h.pinned[2];
  h.adiabatic_potential = pd[0];
  memcpy(h.evec, pd + 1, MAXS * sizeof(double));
  for (int i = 0; i < MAXS; i++) for (int j = 0; j < MAXS; j++) h.hamiltonian[i][j] = 0.0;
  for (int s = 0; s < S; s++) {
    h.hamiltonian[s][s] = pd[1 + MAXS + s];
    if (s > 0) h.hamiltonian[h.parent[s]][s] = pd[1 + 2 * MAXS + s];
  }
  return 0;
}

// shift_array_data_donor_acceptor_transfer on a permutation vector (ms_evb.f90:2677-2840)
static void host_shift(std::vector<int>& perm, std::vector<int>& first, std::vector<int>& natom, int m_from, int a_from, int m_to, int a_to) {
  int from_g = first[m_from] + a_from;
  int to_g = (m_from < m_to) ? first[m_to] + a_to - 1 : first[m_to] + a_to;
  int saved = perm[from_g];
  if (from_g < to_g) for (int i = from_g; i < to_g; i++) perm[i] = perm[i + 1];
  else for (int i = from_g; i > to_g; i--) perm[i] = perm[i - 1];
  perm[to_g] = saved;
  if (from_g < to_g) { for (int m = m_from + 1; m <= m_to; m++) first[m] -= 1; }
  else { for (int m = m_to + 1; m <= m_from; m++) first[m] += 1; }
  natom[m_to] += 1; natom[m_from] -= 1;
}

int evb_commit(rpb_ctx* c) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = g_scratch[c];
  const int N = d.N, M = d.M;
  k_copy<<<(3 * N + 255) / 256, 256, 0, c->stream>>>(d.force, e.f_mix, (size_t)3 * N);
  c->n_launch++;
  // energies as the reference leaves them: potential = adiabatic energy, components = principal diabat's
  {
    rpb_energies& en = c->last_en;
    const double* s = c->h_en;
    en.E_recip = s[E_RECIP];
    en.E_elec = s[E_ELEC] + s[E_RECIP] + c->cfg.ewald_self;
    en.E_vdw = s[E_VDW]; en.E_bond = s[E_BOND]; en.E_angle = s[E_ANGLE]; en.E_dihedral = s[E_DIH];
    en.potential_energy = h.adiabatic_potential;
  }
  if (h.new_hydronium == c->hydronium_mol) return 0;
  // ---- proton hop accepted: evb_change_diabat_data_structure_topology (ms_evb.f90:806-834)
  const int pdiab = h.principal_diabat;
  if (d.world > 1 && !state_owned(pdiab, d.rank, d.world)) {
    // every rank needs the final snapshot of the new principal diabat; non-owned diabats were not built in evb_build
    k_evb_snapshots<<<1, 32, 0, c->stream>>>(d, e, pdiab);
    c->n_launch++;
  }
  std::vector<int> perm(N), first = c->mol_first, natom = c->mol_natom;
  for (int i = 0; i < N; i++) perm[i] = i;
  int nh = h.n_hops[pdiab];
  // replay the hops on (first, n_atom) to obtain the atom permutation; the per-molecule reordering of the acceptor
  // (reorder_molecule_data_structures) is taken from the snapshot, whose atom[] lists give the final order
  int ima = c->hydronium_mol;
  std::vector<int> cur_first = first, cur_natom = natom;
  for (int k = 0; k < nh; k++) {
    int imd = ima;
    int i_atom_donor = h.proton_log[pdiab][k][1];
    ima = h.proton_log[pdiab][k][3];
    host_shift(perm, cur_first, cur_natom, imd, i_atom_donor, ima, cur_natom[ima]);
  }
  // chain molecule list (same construction as on the device)
  int mols[CM], nm = 1;
  mols[0] = c->hydronium_mol;
  for (int k = 0; k < nh; k++) {
    int a = h.proton_log[pdiab][k][3];
    bool f = false;
    for (int q = 0; q < nm; q++) f |= (mols[q] == a);
    if (!f) mols[nm++] = a;
  }
  // final atom order inside each chain molecule comes from the snapshot (download it: tiny)
  Snapshot snap;
  CKE(cudaMemcpyAsync(&snap, e.snap + pdiab * NLEV + nh, sizeof(Snapshot), cudaMemcpyDeviceToHost, c->stream));
  CKE(cudaStreamSynchronize(c->stream));
  int new_first[CM] = {0, 0, 0, 0};
  for (int k = 0; k < snap.n_mol; k++) {
    int m = snap.m[k].mol;
    new_first[k] = cur_first[m];
    for (int a = 0; a < snap.m[k].n_atom; a++) perm[cur_first[m] + a] = snap.m[k].atom[a];
  }
  std::vector<int> moa(N);
  for (int m = 0; m < M; m++) for (int a = 0; a < cur_natom[m]; a++) moa[cur_first[m] + a] = m;
  CKE(cudaMemcpyAsync(sc.perm, perm.data(), N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(sc.moa2, moa.data(), N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(sc.new_first, new_first, CM * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.mol_first, cur_first.data(), M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.mol_natom, cur_natom.data(), M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  k_evb_commit_permute<<<(N + 255) / 256, 256, 0, c->stream>>>(d, sc.perm, sc.xq2, sc.vel2, sc.force2, sc.mass2, sc.type2, d.mol_of_atom, sc.moa2);
  CKE(cudaMemcpyAsync(d.xq, sc.xq2, N * sizeof(double4), cudaMemcpyDeviceToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.vel, sc.vel2, 3 * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.force, sc.force2, 3 * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.mass, sc.mass2, N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CKE(cudaMemcpyAsync(d.type, sc.type2, N * sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  k_evb_commit_patch<<<1, 1, 0, c->stream>>>(d, e, pdiab, nh, sc.new_first);
  c->n_launch += 2;
  // host mirror
  c->mol_first = cur_first; c->mol_natom = cur_natom;
  for (int k = 0; k < snap.n_mol; k++) c->mol_type[snap.m[k].mol] = snap.m[k].mtype;
  c->hydronium_mol = h.new_hydronium;
  // construct_verlet_list + update_verlet_displacements(init)  (ms_evb.f90:223-225)
  launch_verlet_force_rebuild(c);
  CKE(cudaStreamSynchronize(c->stream));
  return 0;
}
