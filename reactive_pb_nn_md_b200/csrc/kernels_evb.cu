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
// Structure (all diabats of this rank in flight; SURVEY 7 "two-pass" reciprocal formulation).  The HOST is not part of a
// step: the enumeration kernel writes the diabat set AND the work lists every later kernel needs (EvbPlan: chain
// molecules, chain-molecule pairs, owned diabats, chain atoms) to device memory; every kernel is launched with a grid
// sized from a recent diabat count and loops over the device-side counts; the solver's hop decision is carried out by
// device kernels that exit at once when no hop was selected.  One step is therefore a fixed launch sequence that
// rpb_step replays as a CUDA graph.
//   enumerate (+ plan) -> snapshots -> candidate lists -> item kernels (real-space / repulsion / bonded deltas of every
//   diabat's last hop) | couplings (geometry, Vex) | reciprocal-space charge-delta algebra on the principal grid
//   -> tree solver (ground state, hop decision) -> F = sum c_i c_j F_ij | averaged grid -> ONE convolution -> ONE gather
//   -> hop commit (permutation + retyping + list rebuild) when the principal diabat changed.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "rpb_host.h"
#include "rpb_bonded.cuh"
#include "rpb_pme.cuh"
#include "rpb_commit.cuh"
#include "rpb_peer.cuh"

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
// K8: diabat enumeration.  The reference walks a pre-order DFS (evb_conduct_proton_transfer_recursive,
// ms_evb.f90:498-607) and runs find_evb_reactive_neighbors (:702-764) for every (donor, proton) it meets.  That
// search depends only on (molecule, proton) -- always principal-topology coordinates -- so it is memoised and done
// with the whole CTA, level by level (depth < max_chain):
//   1. compact the molecules whose centre of mass lies within max_chain * first-solvation cutoff of the hydronium
//      (nothing farther can ever be an acceptor);
//   2. per level:  a) (molecule of the level) x (compact molecule): the reference's centre-of-mass test -> short
//      "near" lists;  b) (molecule, reactive proton, near molecule): the proton-acceptor distance test on the basic
//      atoms -> hit lists;  c) hits sorted into the reference's (molecule, atom) loop order and truncated at
//      evb_max_neighbors;  d) unseen acceptors join the next level;
//   3. one thread replays the DFS over the memoised lists (shared memory only) and writes the hop logs.
// ================================================================================================
#define ENUM_TPB 512
#define ENUM_MAXMOL RPB_MAXS   // distinct molecules whose protons are searched (each is the acceptor of some diabat)
#define ENUM_MAXP 4            // reactive protons per molecule
#define ENUM_NEAR 40           // molecules inside the first-solvation cutoff of one molecule (~17 in water at 5 A)
#define ENUM_HITS 12           // acceptor atoms inside the reactive-pair distance of one proton (10 are kept)
#define ENUM_COMPACT 2048
#define ENUM_SMEM_BYTES (ENUM_COMPACT * (3 * sizeof(double) + 3 * sizeof(int)))

__global__ void __launch_bounds__(ENUM_TPB) k_evb_enumerate(Dev d, EvbDev e, int* __restrict__ cand_n) {
  extern __shared__ double compact_com[];        // dynamic: [ENUM_COMPACT][3] centres of mass of the compacted molecules, then their
  int* compact_first = reinterpret_cast<int*>(compact_com + 3 * ENUM_COMPACT);   // first atom, molecule type and atom count
  int* compact_type = compact_first + ENUM_COMPACT;
  int* compact_nat = compact_type + ENUM_COMPACT;
  __shared__ int compact[ENUM_COMPACT];
  __shared__ int vis_mol[ENUM_MAXMOL];
  __shared__ unsigned char prot_n[ENUM_MAXMOL], prot[ENUM_MAXMOL][ENUM_MAXP], heavy[ENUM_MAXMOL][ENUM_MAXP];
  __shared__ int near_n[ENUM_MAXMOL];
  __shared__ int near_mol[ENUM_MAXMOL][ENUM_NEAR];
  __shared__ int nb_n[ENUM_MAXMOL][ENUM_MAXP];
  __shared__ int nb[ENUM_MAXMOL][ENUM_MAXP][ENUM_HITS];          // acceptor molecule * 16 + acceptor atom
  __shared__ unsigned char nb_v[ENUM_MAXMOL][ENUM_MAXP][RPB_EVB_MAX_NEIGHBORS];
  __shared__ int s_ncomp, s_nvis, s_fail;
  __shared__ double vis_com[ENUM_MAXMOL][3];
  const int tid = threadIdx.x, nth = blockDim.x;
  const int hyd = *d.hydronium;
  EvbPlan& plan = *e.plan;
  __shared__ int s_S, s_ncmol, s_npair;
  __shared__ signed char mt_rp[RPB_MAXM][MA], mt_rb[RPB_MAXM][MA], mt_bh[RPB_MAXM][MA];   // reactive proton / basic atom flags, bonded heavy atom per molecule type
  for (int k = tid; k < d.nMT * MA; k += nth) {
    const MolTypeDev& T = d.mt[k / MA];
    mt_rp[k / MA][k % MA] = (signed char)T.reactive_proton[k % MA]; mt_rb[k / MA][k % MA] = (signed char)T.reactive_basic[k % MA];
    mt_bh[k / MA][k % MA] = (signed char)T.bonded_heavy[k % MA];
  }
  // the chain molecules of the previous step give their slots back
  for (int k = tid; k < plan.n_cmol; k += nth) e.mol_slot[plan.cmol[k]] = -1;
  for (int i = tid; i < MAXS * MAXC * 5; i += nth) e.proton_log[i] = -1;
  for (int i = tid; i < MAXS; i += nth) { e.parent[i] = -1; e.n_hops[i] = 0; }
  for (int i = tid; i < ENUM_MAXMOL * ENUM_MAXP; i += nth) (&nb_n[0][0])[i] = 0;
  for (int i = tid; i < ENUM_MAXMOL; i += nth) near_n[i] = 0;
  if (tid == 0) { s_ncomp = 0; s_nvis = 1; vis_mol[0] = hyd; s_fail = 0; }
  __syncthreads();
  // ---- 1. compact list (superset: min-image COM distance below max_chain * cutoff + 0.5 A)
  {
    double R = (double)d.max_chain * sqrt(d.cut_solv2) + 0.5;
    double R2 = R * R;
    double c0 = d.r_com[3 * hyd], c1 = d.r_com[3 * hyd + 1], c2 = d.r_com[3 * hyd + 2];
    for (int jm = tid; jm < d.M; jm += nth) {      // (a superset with a 0.5 A margin: the reciprocal-box minimum image is exact enough)
      double a0 = d.r_com[3 * jm] - c0, a1 = d.r_com[3 * jm + 1] - c1, a2 = d.r_com[3 * jm + 2] - c2;
      a0 -= d.box[0] * floor_fp64pipe(fma(a0, d.inv_box[0], 0.5));
      a1 -= d.box[1] * floor_fp64pipe(fma(a1, d.inv_box[1], 0.5));
      a2 -= d.box[2] * floor_fp64pipe(fma(a2, d.inv_box[2], 0.5));
      if (a0 * a0 + a1 * a1 + a2 * a2 < R2) {
        int slot = atomicAdd(&s_ncomp, 1);
        if (slot < ENUM_COMPACT) {
          compact[slot] = jm;
          compact_com[3 * slot] = d.r_com[3 * jm]; compact_com[3 * slot + 1] = d.r_com[3 * jm + 1]; compact_com[3 * slot + 2] = d.r_com[3 * jm + 2];
          compact_first[slot] = d.mol_first[jm]; compact_type[slot] = d.mol_type[jm]; compact_nat[slot] = d.mol_natom[jm];
        }
      }
    }
  }
  __syncthreads();
  if (s_ncomp > ENUM_COMPACT) {
    if (tid == 0) { atomicMax(&d.err_flag[3], 9); *e.n_states = 1; }
    return;
  }
  const int ncomp = s_ncomp;
  // ---- 2. memoised neighbour searches, level by level
  int lvl_begin = 0, lvl_end = 1;
  for (int L = 0; L < d.max_chain; L++) {
    const int nlvl = lvl_end - lvl_begin;
    // reactive protons of the molecules of this level (principal-topology molecule type, ms_evb.f90:542-546)
    for (int vv = lvl_begin + tid; vv < lvl_end; vv += nth) {
      int mol = vis_mol[vv];
      const int mty = d.mol_type[mol];
      int n = d.mol_natom[mol], np = 0;
      for (int ia = 0; ia < n; ia++)
        if (mt_rp[mty][ia] == 1) {
          if (np >= ENUM_MAXP) { s_fail = 2; break; }
          prot[vv][np] = (unsigned char)ia;
          int hv = mt_bh[mty][ia];
          heavy[vv][np] = (unsigned char)(hv < 0 ? 255 : hv);
          np++;
        }
      prot_n[vv] = (unsigned char)np;
    }
    // a) centre-of-mass test of find_evb_reactive_neighbors (:724-733), flattened over (molecule, compact entry)
    for (int vv = lvl_begin + tid; vv < lvl_end; vv += nth)
      for (int k = 0; k < 3; k++) vis_com[vv][k] = d.r_com[3 * vis_mol[vv] + k];
    __syncthreads();
    for (int idx = tid; idx < nlvl * ncomp; idx += nth) {
      const int vv = lvl_begin + idx / ncomp, cs = idx % ncomp, jm = compact[cs], im = vis_mol[vv];
      if (jm == im) continue;
      double dc2 = 0.0;
#pragma unroll
      for (int k = 0; k < 3; k++) {      // same operands, same order as before: the values now come from shared memory
        double dr = compact_com[3 * cs + k] - vis_com[vv][k];
        double shift = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
        double dc = compact_com[3 * cs + k] - vis_com[vv][k] - shift;
        dc2 = k == 0 ? dc * dc : dc2 + dc * dc;
      }
      if (dc2 < d.cut_solv2) {
        int slot = atomicAdd(&near_n[vv], 1);
        if (slot < ENUM_NEAR) near_mol[vv][slot] = cs; else s_fail = 3;      // compact slot of the near molecule
      }
    }
    __syncthreads();
    // b) proton -> basic-atom distance test (:734-754), flattened over (molecule, proton, near molecule)
    for (int idx = tid; idx < nlvl * ENUM_MAXP * ENUM_NEAR; idx += nth) {
      const int vv = lvl_begin + idx / (ENUM_MAXP * ENUM_NEAR), ip = (idx / ENUM_NEAR) % ENUM_MAXP, kn = idx % ENUM_NEAR;
      if (ip >= prot_n[vv] || kn >= min(near_n[vv], ENUM_NEAR)) continue;
      const int im = vis_mol[vv], cs = near_mol[vv][kn], jm = compact[cs];
      double4 ph = d.xq[d.mol_first[im] + prot[vv][ip]];
      double shift[3];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double dr = compact_com[3 * cs + k] - vis_com[vv][k];
        shift[k] = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
      }
      const int tyj = compact_type[cs];
      const int fj = compact_first[cs], nj = compact_nat[cs];
      for (int ja = 0; ja < nj; ja++) {
        if (mt_rb[tyj][ja] != 1) continue;
        double4 pj = d.xq[fj + ja];
        double r0 = pj.x - ph.x - shift[0], r1 = pj.y - ph.y - shift[1], r2 = pj.z - ph.z - shift[2];
        if (r0 * r0 + r1 * r1 + r2 * r2 < d.cut_pair2) {
          int slot = atomicAdd(&nb_n[vv][ip], 1);
          if (slot < ENUM_HITS) nb[vv][ip][slot] = jm * 16 + ja; else s_fail = 3;
        }
      }
    }
    __syncthreads();
    // c) ascending (molecule, atom) == the reference's loop order; keep evb_max_neighbors   (evb_neighbor_list(10,2))
    for (int idx = tid; idx < nlvl * ENUM_MAXP; idx += nth) {
      const int vv = lvl_begin + idx / ENUM_MAXP, ip = idx % ENUM_MAXP;
      int n = min(nb_n[vv][ip], ENUM_HITS);
      for (int a = 1; a < n; a++) {
        int key = nb[vv][ip][a], b2 = a - 1;
        while (b2 >= 0 && nb[vv][ip][b2] > key) { nb[vv][ip][b2 + 1] = nb[vv][ip][b2]; b2--; }
        nb[vv][ip][b2 + 1] = key;
      }
      nb_n[vv][ip] = min(n, RPB_EVB_MAX_NEIGHBORS);
    }
    __syncthreads();
    // d) next level: acceptors not seen before (their protons are searched only if a diabat can still hop from them).
    //    Every (molecule, proton, neighbour) entry looks its acceptor up in parallel; one thread then appends the unseen
    //    ones (few) in entry order.
    if (L + 1 < d.max_chain) {
      const int nv0 = s_nvis;
      for (int idx = tid; idx < nlvl * ENUM_MAXP * RPB_EVB_MAX_NEIGHBORS; idx += nth) {
        const int vv = lvl_begin + idx / (ENUM_MAXP * RPB_EVB_MAX_NEIGHBORS), ip = (idx / RPB_EVB_MAX_NEIGHBORS) % ENUM_MAXP, k = idx % RPB_EVB_MAX_NEIGHBORS;
        if (ip >= prot_n[vv] || k >= nb_n[vv][ip]) continue;
        const int acc = nb[vv][ip][k] >> 4;
        int at = 255;
        if (acc == hyd) at = 0;
        else for (int q = 0; q < nv0; q++) if (vis_mol[q] == acc) { at = q; break; }
        nb_v[vv][ip][k] = (unsigned char)at;
      }
      __syncthreads();
      if (tid == 0) {
        int nv = nv0;
        for (int vv = lvl_begin; vv < lvl_end; vv++)
          for (int ip = 0; ip < prot_n[vv]; ip++)
            for (int k = 0; k < nb_n[vv][ip]; k++) {
              if (nb_v[vv][ip][k] != 255) continue;
              const int acc = nb[vv][ip][k] >> 4;
              int at = -1;
              for (int q = nv0; q < nv; q++) if (vis_mol[q] == acc) { at = q; break; }
              if (at < 0) {
                if (nv >= ENUM_MAXMOL) { s_fail = 1; at = 0; }
                else { vis_mol[nv] = acc; at = nv++; }
              }
              nb_v[vv][ip][k] = (unsigned char)at;
            }
        s_nvis = nv;
      }
    }
    __syncthreads();
    lvl_begin = lvl_end; lvl_end = s_nvis;
  }
  // ---- 3. the reference's pre-order DFS numbering (new diabat id = ++counter, :557), level-parallel.  A diabat is a path
  //         root -> child j1 -> child j2 -> child j3 over the memoised lists: the children of a node whose acceptor is the
  //         visited molecule v are its (proton, neighbour) pairs in (ip, k) order; a node recurses unless its acceptor is
  //         the hydronium or the chain is full (flag_cycle :573,596 ; "if ( count < evb_max_chain )" :538).  With
  //         size(node) = 1 + sum of its children's sizes, id(first child) = id(node) + 1 and id(next sibling) = id + size.
  static_assert(MAXC == 3, "three levels of diabats are spelled out below");
  __shared__ int s_size1[ENUM_MAXP * RPB_EVB_MAX_NEIGHBORS], s_id1[ENUM_MAXP * RPB_EVB_MAX_NEIGHBORS + 1];
  auto n_children = [&](int v) { int t = 0; for (int ip = 0; ip < prot_n[v]; ip++) t += nb_n[v][ip]; return t; };
  auto decode = [&](int v, int j, int& ip, int& k) { ip = 0; while (j >= nb_n[v][ip]) { j -= nb_n[v][ip]; ip++; } k = j; };
  auto make_row = [&](int v, int ip, int k, int row[5]) {
    const int pk = nb[v][ip][k];
    row[0] = vis_mol[v]; row[1] = prot[v][ip]; row[2] = heavy[v][ip]; row[3] = pk >> 4; row[4] = pk & 15;
    if (row[2] == 255) { atomicMax(&d.err_flag[3], 1); row[2] = -1; }   // find_bonded_atom_hydrogen failed
  };
  auto write_state = [&](int id, int parent, int nh, const int (*rows)[5]) {
    if (id >= d.max_states || id >= MAXS) return;
    e.parent[id] = parent; e.n_hops[id] = nh;
    for (int h = 0; h < nh; h++) for (int q = 0; q < 5; q++) e.proton_log[(id * MAXC + h) * 5 + q] = rows[h][q];
  };
  const bool enum_ok = (s_fail == 0);
  if (tid == 0) {
    if (s_fail == 1) atomicMax(&d.err_flag[2], 1);              // more molecules than evb_max_states diabats
    else if (s_fail) atomicMax(&d.err_flag[3], 9);              // compiled enumeration limits exceeded
  }
  const int n1 = enum_ok ? n_children(0) : 0;
  for (int j1 = tid; j1 < n1; j1 += nth) {                       // subtree sizes of the root's children
    int ip1, k1;
    decode(0, j1, ip1, k1);
    const int acc1 = nb[0][ip1][k1] >> 4;
    int size = 1;
    if (acc1 != hyd && 1 < d.max_chain) {
      const int v1 = nb_v[0][ip1][k1], n2 = n_children(v1);
      for (int j2 = 0; j2 < n2; j2++) {
        int ip2, k2;
        decode(v1, j2, ip2, k2);
        const int acc2 = nb[v1][ip2][k2] >> 4;
        size += 1 + ((acc2 != hyd && 2 < d.max_chain) ? n_children(nb_v[v1][ip2][k2]) : 0);
      }
    }
    s_size1[j1] = size;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 1;
    for (int j1 = 0; j1 < n1; j1++) { s_id1[j1] = run; run += s_size1[j1]; }
    s_id1[n1] = run;
    if (run > d.max_states) { atomicMax(&d.err_flag[2], 1); run = d.max_states; }   // "Found more diabat states than ... evb_max_states"
    *e.n_states = run;
    s_S = run; s_ncmol = 1; s_npair = 0;
    e.mol_slot[hyd] = 0;
  }
  __syncthreads();
  constexpr int NCH = ENUM_MAXP * RPB_EVB_MAX_NEIGHBORS;        // children of one node at most
  for (int idx = tid; idx < n1 * (NCH + 1); idx += nth) {
    const int j1 = idx / (NCH + 1), j2 = idx % (NCH + 1) - 1;    // j2 == -1: the level-1 diabat itself
    int rows[MAXC][5];
    int ip1, k1;
    decode(0, j1, ip1, k1);
    make_row(0, ip1, k1, rows[0]);
    const int id1 = s_id1[j1];
    if (j2 < 0) { write_state(id1, 0, 1, rows); continue; }
    const int acc1 = rows[0][3];
    if (!(acc1 != hyd && 1 < d.max_chain)) continue;
    const int v1 = nb_v[0][ip1][k1];
    if (j2 >= n_children(v1)) continue;
    int id2 = id1 + 1;                                           // ids of the earlier siblings' subtrees
    for (int j = 0; j < j2; j++) {
      int ipj, kj;
      decode(v1, j, ipj, kj);
      const int accj = nb[v1][ipj][kj] >> 4;
      id2 += 1 + ((accj != hyd && 2 < d.max_chain) ? n_children(nb_v[v1][ipj][kj]) : 0);
    }
    int ip2, k2;
    decode(v1, j2, ip2, k2);
    make_row(v1, ip2, k2, rows[1]);
    write_state(id2, id1, 2, rows);
    const int acc2 = rows[1][3];
    if (!(acc2 != hyd && 2 < d.max_chain)) continue;
    const int v2 = nb_v[v1][ip2][k2], n3 = n_children(v2);
    for (int j3 = 0; j3 < n3; j3++) {
      int ip3, k3;
      decode(v2, j3, ip3, k3);
      make_row(v2, ip3, k3, rows[2]);
      write_state(id2 + 1 + j3, id2, 3, rows);
    }
  }
  // ---- 4. this step's work lists (what the host used to derive from the read-back hop logs).  The enumeration's own
  //         dynamic shared arrays are dead from here on: the pair-seen bit matrix, the ownership marks and the atom
  //         offsets alias the compact list.
  __syncthreads();
  const int S = s_S;
  unsigned int* seen = reinterpret_cast<unsigned int*>(compact_com);     // bit (i * RA_MOLS + j): pair already listed
  constexpr int SEEN_WORDS = (RPB_RA_MOLS * RPB_RA_MOLS + 31) / 32;
  int* own_mol = reinterpret_cast<int*>(seen + SEEN_WORDS);              // [RA_MOLS] chain molecule of a diabat this rank owns
  int* uniq_base = own_mol + RPB_RA_MOLS;                                // [RA_MOLS + 1] first slot of the molecule's atoms in uniq_atom
  for (int k = tid; k < SEEN_WORDS + 2 * RPB_RA_MOLS + 1; k += nth) seen[k] = 0u;
  for (int k = tid; k < RPB_CAND_SLOTS; k += nth) cand_n[k] = 0;
  // a) distinct chain molecules over all diabats: the first thread to claim a molecule gives it the next slot
  if (tid >= 1 && tid < S) {
    const int nh = e.n_hops[tid];
    for (int h = 0; h < nh; h++) {
      const int a = e.proton_log[(tid * MAXC + h) * 5 + 3];
      if (atomicCAS(&e.mol_slot[a], -1, -2) == -1) {
        const int slot = atomicAdd(&s_ncmol, 1);
        if (slot < RPB_RA_MOLS) { plan.cmol[slot] = a; e.mol_slot[a] = slot; }
        else { e.mol_slot[a] = 0; atomicMax(&d.err_flag[2], 1); }
      }
    }
  }
  if (tid == 0) plan.cmol[0] = hyd;
  __syncthreads();
  // b) ordered pairs of chain molecules that share a diabat; chain molecules of the diabats this rank owns
  if (tid == 0) own_mol[0] = 1;
  if (tid >= 1 && tid < S) {
    int idx[CM], ni = 1;
    idx[0] = 0;
    const int nh = e.n_hops[tid];
    for (int h = 0; h < nh; h++) {
      const int mi = e.mol_slot[e.proton_log[(tid * MAXC + h) * 5 + 3]];
      bool dup = false;
      for (int q = 0; q < ni; q++) dup |= (idx[q] == mi);
      if (!dup && ni < CM) idx[ni++] = mi;
    }
    const bool owned = state_owned(tid, d.rank, d.world);
    for (int a = 0; a < ni; a++) {
      if (owned) own_mol[idx[a]] = 1;
      for (int b = 0; b < ni; b++) {
        const int key = idx[a] * RPB_RA_MOLS + idx[b];
        const unsigned int bit = 1u << (key & 31);
        if (!(atomicOr(&seen[key >> 5], bit) & bit)) {
          const int p = atomicAdd(&s_npair, 1);
          if (p < RPB_RA_MAXPAIR) plan.molpair[p] = key; else atomicMax(&d.err_flag[2], 1);
        }
      }
    }
  }
  __syncthreads();
  // c) owned diabats (ascending) and the atoms of their chain molecules
  const int ncm = min(s_ncmol, RPB_RA_MOLS);
  if (tid == 0) {
    int n_own = 0;
    for (int s2 = 1; s2 < S; s2++) if (state_owned(s2, d.rank, d.world)) plan.state_list[n_own++] = s2;
    int nu = 0;
    for (int k = 0; k < ncm; k++) { uniq_base[k] = nu; if (own_mol[k]) nu += d.mol_natom[plan.cmol[k]]; }
    uniq_base[ncm] = nu;
    if (nu > RPB_CAND_SLOTS) { atomicMax(&d.err_flag[2], 1); nu = 0; }
    plan.n_own = n_own; plan.n_uniq = nu; plan.n_cmol = ncm; plan.n_pair = min(s_npair, RPB_RA_MAXPAIR); plan.hop = 0;
  }
  __syncthreads();
  if (uniq_base[ncm] <= RPB_CAND_SLOTS)
    for (int t = tid; t < ncm * MA; t += nth) {
      const int k = t / MA, a = t % MA;
      if (!own_mol[k]) continue;
      const int m = plan.cmol[k];
      if (a < d.mol_natom[m]) plan.uniq_atom[uniq_base[k] + a] = d.mol_first[m] + a;
    }
}

// ================================================================================================
// snapshots: images of the chain molecules of diabat s at every topology level
// ================================================================================================
__device__ void load_principal_image(const Dev& d, int mol, MolImage& im) {
  im.mol = mol; im.n_atom = d.mol_natom[mol]; im.mtype = d.mol_type[mol];
  int f = d.mol_first[mol];
  for (int a = 0; a < im.n_atom; a++) {
    double4 p = d.xq[f + a];
    im.atom[a] = f + a; im.ratom[a] = f + a; im.type[a] = d.type[f + a]; im.q[a] = p.w; im.mass[a] = d.mass[f + a];
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
  A.atom[last] = D.atom[i_atom_donor]; A.ratom[last] = D.ratom[i_atom_donor]; A.type[last] = D.type[i_atom_donor]; A.q[last] = D.q[i_atom_donor];
  A.mass[last] = D.mass[i_atom_donor];
  for (int k = 0; k < 3; k++) A.x[last][k] = D.x[i_atom_donor][k];
  for (int a = i_atom_donor; a < D.n_atom - 1; a++) {
    D.atom[a] = D.atom[a + 1]; D.ratom[a] = D.ratom[a + 1]; D.type[a] = D.type[a + 1]; D.q[a] = D.q[a + 1]; D.mass[a] = D.mass[a + 1];
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

// warp-cooperative copy of a struct whose size is a multiple of 8 bytes
template <typename T>
__device__ __forceinline__ void warp_copy_struct(T* dst, const T* src, int lane) {
  static_assert(sizeof(T) % 8 == 0, "struct size must be a multiple of 8");
  const double* s = reinterpret_cast<const double*>(src);
  double* t = reinterpret_cast<double*>(dst);
  for (int k = lane; k < (int)(sizeof(T) / 8); k += 32) t[k] = s[k];
}

// One warp per diabat, working copy in shared memory: the lanes load the principal images of the chain molecules
// and store every level's snapshot cooperatively; lane 0 replays the hops on the shared copy.
// only_state >= 0: build that diabat regardless of ownership (hop commit needs the new principal's images on every rank)
#define SNAP_WPB 4
__global__ void __launch_bounds__(32 * SNAP_WPB) k_evb_snapshots(Dev d, EvbDev e, int only_state, int all_states) {
  static_assert(CM * MA == 32, "one lane per (chain molecule, atom)");
  __shared__ Snapshot Wsh[SNAP_WPB];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int s = blockIdx.x * SNAP_WPB + w;
  const int S = *e.n_states;
  if (only_state >= 0) { if (s != 0) return; s = only_state; }
  else if (s >= S || !(s == 0 || all_states || state_owned(s, d.rank, d.world))) return;
  Snapshot& W = Wsh[w];
  const int* L = &e.proton_log[s * MAXC * 5];
  const int nh = e.n_hops[s];
  if (lane == 0) {
    int nm = 1;
    W.m[0].mol = *d.hydronium;
    for (int h = 0; h < nh; h++) {
      int a = L[h * 5 + 3];
      bool found = false;
      for (int k = 0; k < nm; k++) found |= (W.m[k].mol == a);
      if (!found) W.m[nm++].mol = a;
    }
    for (int k = nm; k < CM; k++) { W.m[k].mol = -1; W.m[k].n_atom = 0; W.m[k].mtype = 0; }
    W.n_mol = nm; W.hydronium = 0;
  }
  __syncwarp();
  {
    const int k = lane / MA, a = lane % MA;
    if (k < W.n_mol) {
      MolImage& im = W.m[k];
      const int mol = im.mol, f = d.mol_first[mol], n = d.mol_natom[mol];
      if (a == 0) { im.n_atom = n; im.mtype = d.mol_type[mol]; for (int c = 0; c < 3; c++) im.r_com[c] = d.r_com[3 * mol + c]; }
      if (a < n) {
        double4 p = d.xq[f + a];
        im.atom[a] = f + a; im.ratom[a] = f + a; im.type[a] = d.type[f + a]; im.q[a] = p.w; im.mass[a] = d.mass[f + a];
        im.x[a][0] = p.x; im.x[a][1] = p.y; im.x[a][2] = p.z;
      }
    }
  }
  __syncwarp();
  warp_copy_struct(&e.snap[s * NLEV + 0], &W, lane);
  int cur = 0;  // slot of the current hydronium (= donor of the next hop)
  for (int h = 0; h < nh; h++) {
    __syncwarp();
    int as = 0;
    const int a = L[h * 5 + 3];
    for (int k = 0; k < W.n_mol; k++) if (W.m[k].mol == a) as = k;
    if (lane == 0) {
      image_proton_transfer(d, W.m[cur], W.m[as], L[h * 5 + 1], L[h * 5 + 4]);
      W.hydronium = as;
    }
    cur = as;
    __syncwarp();
    warp_copy_struct(&e.snap[s * NLEV + h + 1], &W, lane);
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

// executed by the whole CTA (>= 96 threads; one thread per copied element, threads t < RPB_MAXT resolve the parameter
// rows); caller syncs
__device__ void fill_item_shared(const Dev& d, const Snapshot& S, const int donor_slot, const int acceptor_slot, ItemShared& sh, int tid) {
  const MolImage& H = S.m[S.hydronium];
  const EvbTables& E = *d.evb;
  if (tid < CM * MA) {                    // chain atoms, molecule by molecule
    const int k = tid / MA, a = tid % MA;
    if (k < S.n_mol && a < S.m[k].n_atom) {
      int off = 0;
      for (int k2 = 0; k2 < k; k2++) off += S.m[k2].n_atom;
      sh.chain_atoms[off + a] = S.m[k].atom[a];
    }
    if (tid == 0) {
      int n = 0;
      for (int k2 = 0; k2 < S.n_mol; k2++) n += S.m[k2].n_atom;
      sh.n_chain = n;
      sh.nd = donor_slot >= 0 ? S.m[donor_slot].n_atom : 0;
      sh.na = acceptor_slot >= 0 ? S.m[acceptor_slot].n_atom : 0;
      sh.nh = H.n_atom;
      sh.h_heavy = d.mt[H.mtype].heavy_acid_atom;
      if (sh.h_heavy < 0) { atomicMax(&d.err_flag[3], 3); sh.h_heavy = 0; }
      sh.h_type_H = H.type[H.n_atom - 1];
      sh.h_type_heavy = H.type[sh.h_heavy];
    }
  }
  if (tid >= 64 && tid < 64 + 3 * MA) {   // donor / acceptor / hydronium images
    const int grp = (tid - 64) / MA, a = (tid - 64) % MA;
    if (grp == 0 && donor_slot >= 0) {
      const MolImage& D = S.m[donor_slot];
      if (a < D.n_atom) { sh.d_atom[a] = D.atom[a]; sh.d_type[a] = D.type[a]; sh.d_q[a] = D.q[a]; for (int k = 0; k < 3; k++) sh.d_x[a][k] = D.x[a][k]; }
    } else if (grp == 1 && acceptor_slot >= 0) {
      const MolImage& A = S.m[acceptor_slot];
      if (a < A.n_atom) { sh.a_atom[a] = A.atom[a]; sh.a_type[a] = A.type[a]; sh.a_q[a] = A.q[a]; for (int k = 0; k < 3; k++) sh.a_x[a][k] = A.x[a][k]; }
    } else if (grp == 2) {
      if (a < H.n_atom) { sh.h_atom[a] = H.atom[a]; sh.h_type[a] = H.type[a]; for (int k = 0; k < 3; k++) sh.h_x[a][k] = H.x[a][k]; }
    }
  }
  if (tid >= 32 && tid < 32 + RPB_MAXT) {
    const int t = tid - 32;
    int hv = d.mt[H.mtype].heavy_acid_atom;
    if (hv < 0) hv = 0;
    const int tH = H.type[H.n_atom - 1], tO = H.type[hv];
    int row = -1;
    for (int i = 0; i < RPB_MAXI; i++) {            // get_index_atom_set general_routines.f90:613-637
      if (E.da_int[i][0] < 0) break;
      if (E.da_int[i][0] == t && E.da_int[i][1] == tO && E.da_int[i][2] == tH) { row = i; break; }
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
#ifndef ITEM_MINB
#define ITEM_MINB 3            // CTAs per SM the item kernel is compiled for (registers <= 85)
#endif
#define CAND_CAP 2048        // atoms inside the candidate radius of one chain atom (~420 at 10 A in water)
#define CAND_SLOTS RPB_CAND_SLOTS

// ---- candidate lists: for every distinct chain atom g (principal index; listed by the enumeration kernel, EvbPlan)
// the atoms j whose minimum-image distance from g is below the candidate radius (real-space cutoff + margin, or the
// EVB repulsion reach if that is larger), from the CURRENT positions -- so the set is exact, unlike a Verlet row
// between rebuilds (the reference scans all N atoms per image atom, ms_evb.f90:1629-1841).  The item kernel applies
// the reference's own cutoff test to the image position; this pass only removes the ~95 % of atoms that are far away,
// once per chain atom instead of once per (diabat, topology, image atom).
// grid = (ceil(N/256), a bound of the number of chain atoms; the slots beyond it are reached by the loop)
__global__ void __launch_bounds__(256) k_evb_candidates(Dev d, EvbDev e, double r2cand,
                                                        int* __restrict__ chain_slot, int* __restrict__ cand, int* __restrict__ cand_n) {
  const int lane = threadIdx.x & 31;
  const int n_uniq = e.plan->n_uniq;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  double4 pj = make_double4(0.0, 0.0, 0.0, 0.0);
  if (j < d.N) pj = d.xq[j];
  for (int slot = blockIdx.y; slot < n_uniq; slot += gridDim.y) {
    const int g = e.plan->uniq_atom[slot];
    if (blockIdx.x == 0 && threadIdx.x == 0) chain_slot[g] = slot;
    const double4 pg = d.xq[g];
    bool hit = false;
    if (j < d.N && j != g) {
      double dx = min_image(pg.x - pj.x, d.box[0]), dy = min_image(pg.y - pj.y, d.box[1]), dz = min_image(pg.z - pj.z, d.box[2]);
      hit = dx * dx + dy * dy + dz * dz < r2cand;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m) {
      int leader = __ffs(m) - 1, base = 0;
      if (lane == leader) base = atomicAdd(&cand_n[slot], __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (hit) {
        int p = base + __popc(m & ((1u << lane) - 1u));
        if (p < CAND_CAP) cand[(size_t)slot * CAND_CAP + p] = j;
      }
    }
  }
}

#define ITEM_SPLIT 7          // task CTAs per item (the candidate chunks of an item are dealt round-robin to 7 x 8 warps)
struct ItemBlock {
  ItemShared sh;
  int d_ci[MA], a_ci[MA], h_ci[MA];       // position of every image atom in sh.chain_atoms
  int task_first[2 * MA + 1];             // prefix sum of 32-candidate chunks per image atom
  double fl[CM * MA][3];                  // force accumulators of the chain atoms (shared-memory atomics)
  double red[32];
};

// ITEM_SPLIT CTAs per (diabat, last hop, topology) item; item ids as in rpb_evb.cuh (0: principal diabat, 2s-1 / 2s: donor
// / acceptor topology of the last hop of diabat s); the grid covers a bound of the item count, the loop the rest.
//   every warp of every CTA : (image atom, 32-candidate chunk) tasks -- real-space pairs of the donor/acceptor image atoms
//                             with the background atoms (ms_evb.f90:1629-1841); the chunks of the hydronium's heavy atom
//                             also carry the EVB repulsion (ms_evb.f90:2259-2478)
//   CTA 0 only, thread 0/32 : bonded + intramolecular terms of donor / acceptor   (ms_evb.f90:1472, 1849-1855)
//   CTA 0 only, threads 64+ : pairs among the chain molecules                      (ms_evb.f90:1629-1841 restricted to chain atoms)
//   CTA 0 only, warp 7      : repulsion of the hydronium image with chain atoms    (ms_evb.f90:2259-2478)
struct ItemDesc { int state, level, donor_slot, acceptor_slot; double sign; };
__device__ __forceinline__ bool item_describe(const Dev& d, const EvbDev& e, int item_id, int S, ItemDesc& it) {
  if (item_id == 0) {     // sharded runs: the principal diabat's repulsion / reference energy is counted on rank 0
    it.state = 0; it.level = 0; it.donor_slot = -1; it.acceptor_slot = -1; it.sign = 1.0;
    return d.rank == 0;
  }
  const int s = (item_id + 1) >> 1, side = (item_id + 1) & 1;     // 2s-1 -> side 0 (donor topology), 2s -> side 1
  if (s >= S || !state_owned(s, d.rank, d.world)) return false;
  const int nh = e.n_hops[s];
  it.state = s; it.level = nh - 1 + side; it.sign = side ? 1.0 : -1.0;
  it.donor_slot = e.snap[s * NLEV + nh - 1].hydronium;            // the donor of hop h is the hydronium of level h
  it.acceptor_slot = e.snap[s * NLEV + nh].hydronium;
  return true;
}

__global__ void __launch_bounds__(ITEM_TPB, ITEM_MINB) k_evb_items(Dev d, EvbDev e, const int* __restrict__ chain_slot,
                                                        const int* __restrict__ cand, const int* __restrict__ cand_n,
                                                        double rcand, double rep_reach) {
  __shared__ ItemBlock B;
  ItemShared& sh = B.sh;
  const int n_items = 2 * (*e.n_states) - 1;
  for (int item_id = blockIdx.x; item_id < n_items; item_id += gridDim.x) {
  ItemDesc it;
  if (!item_describe(d, e, item_id, *e.n_states, it)) continue;      // (uniform over the CTA)
  __syncthreads();                                                    // the previous item of this CTA is done with B
  const Snapshot& S = e.snap[it.state * NLEV + it.level];
  const EvbTables& E = *d.evb;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // the last slice of an item does the chain-internal terms, the others share the (image atom, candidate chunk) tasks
  const int part = blockIdx.y, nparts = gridDim.y - 1;
  const bool internal = part == nparts;
  fill_item_shared(d, S, it.donor_slot, it.acceptor_slot, sh, tid);
  for (int k = tid; k < CM * MA * 3; k += blockDim.x) (&B.fl[0][0])[k] = 0.0;
  __syncthreads();
  const int ds = it.donor_slot, as = it.acceptor_slot;
  const bool hyd_is_donor = (ds >= 0 && S.hydronium == ds);
  // image list: donor atoms, acceptor atoms (principal item: the hydronium atoms, repulsion only)
  const int n_img = (ds >= 0) ? sh.nd + sh.na : sh.nh;
  if (tid < 3 * MA) {   // position of every image atom in sh.chain_atoms
    int grp = tid / MA, a = tid % MA;
    int n = grp == 0 ? sh.nd : (grp == 1 ? sh.na : sh.nh);
    if (a < n) {
      int g = grp == 0 ? sh.d_atom[a] : (grp == 1 ? sh.a_atom[a] : sh.h_atom[a]), ci = 0;
      for (int k = 0; k < sh.n_chain; k++) if (sh.chain_atoms[k] == g) ci = k;
      (grp == 0 ? B.d_ci : (grp == 1 ? B.a_ci : B.h_ci))[a] = ci;
      if (grp == 2) {   // the heavy atom's candidate list must cover the reach of the repulsion terms of every hydronium atom
        const double* xo = sh.h_x[sh.h_heavy];
        double dx = sh.h_x[a][0] - xo[0], dy = sh.h_x[a][1] - xo[1], dz = sh.h_x[a][2] - xo[2];
        if (sqrt(dx * dx + dy * dy + dz * dz) + rep_reach > rcand) atomicMax(&d.err_flag[3], 8);
      }
    }
  }
  if (tid == 32) {
    int acc = 0;
    for (int ia = 0; ia < n_img; ia++) {
      B.task_first[ia] = acc;
      int g = (ds >= 0) ? (ia < sh.nd ? sh.d_atom[ia] : sh.a_atom[ia - sh.nd]) : sh.h_atom[ia];
      if (ds < 0 && ia != sh.h_heavy) continue;       // principal item: only the EVB repulsion, centred on the heavy atom
      int nc = cand_n[chain_slot[g]];
      if (nc > CAND_CAP) { atomicMax(&d.err_flag[3], 7); nc = CAND_CAP; }
      acc += (nc + 31) >> 5;
    }
    B.task_first[n_img] = acc;
  }
  __syncthreads();
  double* outF = (it.state == 0) ? d.force : e.dF + (size_t)it.state * 3 * d.N;
  const double sign = it.sign;
  double en = 0.0;

  // ---------------- image atoms x candidate atoms ----------------
  const int n_tasks = B.task_first[n_img];
  for (int t = internal ? n_tasks : part * (ITEM_TPB / 32) + w; t < n_tasks; t += nparts * (ITEM_TPB / 32)) {
    int ia = 0;
    while (t >= B.task_first[ia + 1]) ia++;
    int side, a;
    if (ds >= 0) { side = ia < sh.nd ? 0 : 1; a = side == 0 ? ia : ia - sh.nd; } else { side = 2; a = ia; }
    const double* xi = side == 0 ? sh.d_x[a] : (side == 1 ? sh.a_x[a] : sh.h_x[a]);
    const int g = side == 0 ? sh.d_atom[a] : (side == 1 ? sh.a_atom[a] : sh.h_atom[a]);
    const int ti = side == 0 ? sh.d_type[a] : (side == 1 ? sh.a_type[a] : sh.h_type[a]);
    const double qi = side == 0 ? sh.d_q[a] : (side == 1 ? sh.a_q[a] : 0.0);
    const int ci = side == 0 ? B.d_ci[a] : (side == 1 ? B.a_ci[a] : B.h_ci[a]);
    // is this image atom the heavy atom of the hydronium image?
    const bool in_hyd = (side == 2) || (side == 0 && hyd_is_donor) || (side == 1 && !hyd_is_donor);
    const bool heavy = in_hyd && (a == sh.h_heavy);
    const int slot = chain_slot[g];
    const int nc = min(cand_n[slot], CAND_CAP);
    const int c = (t - B.task_first[ia]) * 32 + lane;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    double fh[MA][3];
    if (heavy) for (int q = 0; q < MA; q++) fh[q][0] = fh[q][1] = fh[q][2] = 0.0;
    bool live = c < nc;
    int j = 0;
    if (live) {
      j = cand[(size_t)slot * CAND_CAP + c];
      for (int k = 0; k < sh.n_chain; k++) live &= (sh.chain_atoms[k] != j);
    }
    if (live) {
      double4 pj = d.xq[j];
      int tj = d.type[j];
      double xj[3] = {pj.x, pj.y, pj.z};
      double fj[3] = {0.0, 0.0, 0.0};
      if (side != 2) {
        double dr[3] = {min_image(xi[0] - xj[0], d.box[0]), min_image(xi[1] - xj[1], d.box[1]), min_image(xi[2] - xj[2], d.box[2])};
        double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        if (dr2 < d.rc2) {
          int pidx = ti * d.nT + tj;
          double ee, ev, f[3];
          pair_terms(d, dr, dr2, qi * pj.w, d.vdw_type[pidx], &d.vdw_param[6 * pidx], true, ee, ev, f);
          en += ee + ev;
          fx += f[0]; fy += f[1]; fz += f[2];
          fj[0] -= f[0]; fj[1] -= f[1]; fj[2] -= f[2];
        }
      }
      if (heavy) en += repulsion_with_atom(d, sh, xj, tj, fh, fj);
      if (fj[0] != 0.0 || fj[1] != 0.0 || fj[2] != 0.0) {
        atomicAdd(&outF[3 * j], sign * fj[0]); atomicAdd(&outF[3 * j + 1], sign * fj[1]); atomicAdd(&outF[3 * j + 2], sign * fj[2]);
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) { atomicAdd(&B.fl[ci][0], fx); atomicAdd(&B.fl[ci][1], fy); atomicAdd(&B.fl[ci][2], fz); }
    if (heavy) {
      for (int q = 0; q < sh.nh; q++) {
        double s0 = warp_sum(fh[q][0]), s1 = warp_sum(fh[q][1]), s2 = warp_sum(fh[q][2]);
        if (lane == 0) { int cq = B.h_ci[q]; atomicAdd(&B.fl[cq][0], s0); atomicAdd(&B.fl[cq][1], s1); atomicAdd(&B.fl[cq][2], s2); }
      }
    }
  }

  // ---------------- chain-internal terms (last CTA slice of the item) ----------------
  if (!internal) {
    // nothing
  } else if (ds >= 0) {
    if (tid == 0) en += (sign < 0) ? E.ref_energy[S.m[ds].mtype] : E.ref_energy[S.m[as].mtype];   // ms_evb.f90:1478,1520
    if (tid < 64) {
      // bonded + intramolecular terms of the donor (warp 0) and acceptor (warp 1) images: one term (bond, angle, dihedral,
      // atom pair) per lane
      int sl = w == 0 ? ds : as;
      const MolImage& I = S.m[sl];
      const MolTypeDev& T = d.mt[I.mtype];
      const int* cidx = w == 0 ? B.d_ci : B.a_ci;
      const int n_terms = molecule_term_count(T, I.n_atom, true, true);
      for (int term = lane; term < n_terms; term += 32) {
        double f[MA][3];
        for (int q = 0; q < MA; q++) f[q][0] = f[q][1] = f[q][2] = 0.0;
        MolEnergies ME = {0.0, 0.0, 0.0, 0.0, 0.0};
        molecule_term(d, T, I.n_atom, I.x, I.type, I.q, f, ME, true, true, term);
        en += ME.e_bond + ME.e_angle + ME.e_dih + ME.e_elec + ME.e_vdw;
        for (int q = 0; q < I.n_atom; q++) for (int c = 0; c < 3; c++) if (f[q][c] != 0.0) atomicAdd(&B.fl[cidx[q]][c], f[q][c]);
      }
    }
    if (tid >= 64) {
      // donor atoms vs every other chain molecule; acceptor atoms vs chain molecules other than donor and acceptor
      const int per = CM * MA * MA;
      for (int t = tid - 64; t < 2 * per; t += ITEM_TPB - 64) {
        int wv = t / per, r = t % per, k = r / (MA * MA), a = (r / MA) % MA, b2 = r % MA;
        int sl = wv == 0 ? ds : as;
        if (k >= S.n_mol || k == ds || (wv == 1 && k == as)) continue;
        const MolImage& I = S.m[sl];
        const MolImage& J = S.m[k];
        if (a >= I.n_atom || b2 >= J.n_atom) continue;
        double dr[3] = {min_image(I.x[a][0] - J.x[b2][0], d.box[0]), min_image(I.x[a][1] - J.x[b2][1], d.box[1]),
                        min_image(I.x[a][2] - J.x[b2][2], d.box[2])};
        double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        if (dr2 < d.rc2) {
          int pidx = I.type[a] * d.nT + J.type[b2];
          double ee, ev, f[3];
          pair_terms(d, dr, dr2, I.q[a] * J.q[b2], d.vdw_type[pidx], &d.vdw_param[6 * pidx], true, ee, ev, f);
          en += ee + ev;
          int c1 = (wv == 0 ? B.d_ci : B.a_ci)[a], c2 = 0;
          for (int q = 0; q < sh.n_chain; q++) if (sh.chain_atoms[q] == J.atom[b2]) c2 = q;
          for (int c = 0; c < 3; c++) { atomicAdd(&B.fl[c1][c], f[c]); atomicAdd(&B.fl[c2][c], -f[c]); }
        }
      }
    }
  } else if (tid == 0) {
    en += E.ref_energy[S.m[S.hydronium].mtype];   // principal diabat: E_reference (ms_evb.f90:424)
  }
  if (internal && w == 7) {
    // repulsion of the hydronium image with the atoms of the other chain molecules
    int hs = S.hydronium;
    for (int t = lane; t < S.n_mol * MA; t += 32) {
      int k = t / MA, b2 = t % MA;
      if (k == hs || b2 >= S.m[k].n_atom) continue;
      const MolImage& J = S.m[k];
      double fh[MA][3], fj[3] = {0.0, 0.0, 0.0};
      for (int q = 0; q < MA; q++) fh[q][0] = fh[q][1] = fh[q][2] = 0.0;
      en += repulsion_with_atom(d, sh, J.x[b2], J.type[b2], fh, fj);
      int c2 = 0;
      for (int q = 0; q < sh.n_chain; q++) if (sh.chain_atoms[q] == J.atom[b2]) c2 = q;
      for (int c = 0; c < 3; c++) if (fj[c] != 0.0) atomicAdd(&B.fl[c2][c], fj[c]);
      for (int q = 0; q < sh.nh; q++) for (int c = 0; c < 3; c++) if (fh[q][c] != 0.0) atomicAdd(&B.fl[B.h_ci[q]][c], fh[q][c]);
    }
  }
  en = block_sum(en, B.red);
  if (tid == 0 && en != 0.0) atomicAdd(&e.item_energy[item_id], en);
  __syncthreads();
  for (int t = tid; t < sh.n_chain * 3; t += blockDim.x) {
    int k = t / 3, c = t % 3;
    double v = B.fl[k][c];
    if (v != 0.0) atomicAdd(&outF[3 * sh.chain_atoms[k] + c], sign * v);
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

// one warp per owned diabat s>=1: geometric factor and Zundel sites (lane 0, on a shared-memory copy of the final-level
// snapshot), then the Vex terms of the OTHER chain molecules with (atom, site) pairs dealt to the lanes
#define GEO_WPB 4
__global__ void __launch_bounds__(32 * GEO_WPB) k_evb_coupling_geo(Dev d, EvbDev e, CouplingGeo* geo) {
  __shared__ Snapshot Ssh[GEO_WPB];
  __shared__ CouplingGeo Gsh[GEO_WPB];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = *e.n_states;
  for (int s = blockIdx.x * GEO_WPB + w; s < S; s += gridDim.x * GEO_WPB) {     // (a warp keeps its own shared slot: no CTA barrier inside)
  __syncwarp();
  if (s == 0 || !state_owned(s, d.rank, d.world)) { if (lane == 0) geo[s].valid = 0; continue; }
  const EvbTables& E = *d.evb;
  const int nh = e.n_hops[s];
  Snapshot& Sn = Ssh[w];
  CouplingGeo& G = Gsh[w];
  warp_copy_struct(&Sn, &e.snap[s * NLEV + nh], lane);
  const int ds = e.snap[s * NLEV + nh - 1].hydronium;    // last donor
  __syncwarp();
  const int as = Sn.hydronium;                            // last acceptor
  const MolImage& D = Sn.m[ds];
  const MolImage& A = Sn.m[as];
  if (lane == 0) {
    G.valid = 0;
    int iOd = d.mt[D.mtype].heavy_base_atom, iOa = d.mt[A.mtype].heavy_acid_atom, iH = A.n_atom - 1;
    if (iOd < 0 || iOa < 0) atomicMax(&d.err_flag[3], 4);
    else {
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
      if (row < 0) atomicMax(&d.err_flag[3], 5);
      else {
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
        int ns = 0;
        for (int a = 0; a < D.n_atom; a++) {
          int n = ns++;
          G.site_atom[n] = D.atom[a]; G.site_q[n] = E.exch_atomic[D.type[a]];
          for (int k = 0; k < 3; k++) { double dr = D.x[a][k] - G.rz[k] - 0.0; G.site_x[n][k] = G.rz[k] + dr; }
        }
        for (int a = 0; a < A.n_atom; a++) {
          int n = ns++;
          G.site_atom[n] = A.atom[a]; G.site_q[n] = (a == A.n_atom - 1) ? qx : E.exch_atomic[A.type[a]];
          for (int k = 0; k < 3; k++) { double dr = A.x[a][k] - G.rz[k] - shifta[k]; G.site_x[n][k] = G.rz[k] + dr; }
        }
        G.n_site = ns;
        int ncn = 0;
        for (int k = 0; k < Sn.n_mol; k++) for (int a = 0; a < Sn.m[k].n_atom; a++) G.chain_atoms[ncn++] = Sn.m[k].atom[a];
        G.n_chain = ncn;
        G.valid = 1;
      }
    }
  }
  __syncwarp();
  if (G.valid) {
    // ---- Vex with the other chain molecules (final-level charges, positions, centres of mass)
    double vex = 0.0;
    double* Fo = e.Foff + (size_t)s * 3 * d.N;
    const int nsite = G.n_site;
    for (int t = lane; t < CM * MA * nsite; t += 32) {
      const int km = t / (MA * nsite), b = (t / nsite) % MA, n = t % nsite;
      if (km >= Sn.n_mol || km == ds || km == as) continue;
      const MolImage& J = Sn.m[km];
      if (b >= J.n_atom) continue;
      double r[3];
      for (int k = 0; k < 3; k++) {
        double sh = floor(d.inv_box[k] * (J.r_com[k] - G.rz[k]) + 0.5) * d.box[k];
        r[k] = -(J.x[b][k] - G.site_x[n][k] - sh);
      }
      double rm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      double qq = G.site_q[n] * J.q[b];
      vex += qq / rm * d.conv;
      for (int k = 0; k < 3; k++) {
        double dV = -qq / (rm * rm * rm) * r[k] * d.conv;
        atomicAdd(&Fo[3 * G.site_atom[n] + k], -G.A * dV);
        atomicAdd(&Fo[3 * J.atom[b] + k], G.A * dV);
      }
    }
    vex = warp_sum(vex);
    if (lane == 0) atomicAdd(&e.vex[s], vex);
  }
  __syncwarp();
  warp_copy_struct(&geo[s], &G, lane);
  }
}

// grid = (bound of the owned diabats, ceil(N / (256 * VEX_APT))): Vex between the Zundel sites and the background atoms
// (evb_diabatic_coupling_electrostatics, ms_evb.f90:1324-1397: no cutoff, minimum image by molecule).  Every thread
// keeps VEX_APT atoms in registers, so that the per-site warp reductions of the site forces are paid once per
// VEX_APT * 32 atoms; q/r and q/r^3 come from ONE rsqrt instead of a sqrt and two divisions (relative differences ~1e-16).
#define VEX_APT 4
__global__ void __launch_bounds__(256) k_evb_coupling_vex(Dev d, EvbDev e, const CouplingGeo* geo) {
  __shared__ double red[32];
  __shared__ double facc[8][2 * MA][3];
  const int n_own = e.plan->n_own;
  for (int io = blockIdx.x; io < n_own; io += gridDim.x) {
  const int s = e.plan->state_list[io];
  const CouplingGeo& G = geo[s];
  if (!G.valid) continue;
  __syncthreads();
  for (int k = threadIdx.x; k < 8 * 2 * MA * 3; k += blockDim.x) (&facc[0][0][0])[k] = 0.0;
  __syncthreads();
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  bool active[VEX_APT];
  int jj[VEX_APT];
  double xj[VEX_APT][3], qj[VEX_APT], fj[VEX_APT][3];
#pragma unroll
  for (int q = 0; q < VEX_APT; q++) {
    const int j = (blockIdx.y * VEX_APT + q) * blockDim.x + threadIdx.x;
    jj[q] = j;
    active[q] = j < d.N;
    if (active[q]) for (int k = 0; k < G.n_chain; k++) active[q] &= (G.chain_atoms[k] != j);
    xj[q][0] = xj[q][1] = xj[q][2] = 0.0; qj[q] = 0.0; fj[q][0] = fj[q][1] = fj[q][2] = 0.0;
    if (active[q]) {
      const double4 p = d.xq[j];
      const int jm = d.mol_of_atom[j];
      // the molecule's minimum-image shift is folded into the stored position
      xj[q][0] = p.x - floor(d.inv_box[0] * (d.r_com[3 * jm] - G.rz[0]) + 0.5) * d.box[0];
      xj[q][1] = p.y - floor(d.inv_box[1] * (d.r_com[3 * jm + 1] - G.rz[1]) + 0.5) * d.box[1];
      xj[q][2] = p.z - floor(d.inv_box[2] * (d.r_com[3 * jm + 2] - G.rz[2]) + 0.5) * d.box[2];
      qj[q] = p.w;
    }
  }
  double vex = 0.0;
  for (int n = 0; n < G.n_site; n++) {
    const double sx = G.site_x[n][0], sy = G.site_x[n][1], sz = G.site_x[n][2], sq = G.site_q[n] * d.conv;
    double dV0 = 0.0, dV1 = 0.0, dV2 = 0.0;
#pragma unroll
    for (int q = 0; q < VEX_APT; q++) {
      if (active[q]) {
        const double r0 = sx - xj[q][0], r1 = sy - xj[q][1], r2 = sz - xj[q][2];
        const double inv = rsqrt(r0 * r0 + r1 * r1 + r2 * r2);
        const double qq = sq * qj[q];
        const double v = qq * inv;
        vex += v;
        const double g = -v * (inv * inv);
        const double t0 = g * r0, t1 = g * r1, t2 = g * r2;
        dV0 += t0; dV1 += t1; dV2 += t2;
        fj[q][0] -= t0; fj[q][1] -= t1; fj[q][2] -= t2;
      }
    }
    const double s0 = warp_sum(dV0), s1 = warp_sum(dV1), s2 = warp_sum(dV2);
    if (lane == 0) { facc[w][n][0] += s0; facc[w][n][1] += s1; facc[w][n][2] += s2; }
  }
  double* Fo = e.Foff + (size_t)s * 3 * d.N;
#pragma unroll
  for (int q = 0; q < VEX_APT; q++)   // only this thread writes atom j
    if (active[q]) { const int j = jj[q]; Fo[3 * j] = -G.A * fj[q][0]; Fo[3 * j + 1] = -G.A * fj[q][1]; Fo[3 * j + 2] = -G.A * fj[q][2]; }
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
}

// one thread per owned diabat: H_ss, H_parent,s and the geometric part of the coupling force
__device__ void assemble_state(const Dev& d, EvbDev& e, const CouplingGeo* geo, int s, bool defer_principal = false) {
  int S = *e.n_states;
  if (s >= MAXS) return;
  if (s == 0 && defer_principal) return;      // H_11 is assembled by k_evb_finalize_principal once the pair forces are done
  e.h_diag[s] = 0.0; e.h_diag[MAXS + s] = 0.0; e.h_diag[2 * MAXS + s] = 0.0;
  if (s == 0) {
    // principal energy: calculate_total_force_energy + repulsion + reference (ms_evb.f90:411-436).  Sharded runs: every
    // rank holds the pair energies of its slice of atoms; the bonded terms, E_rec, the Ewald self term and the EVB
    // repulsion / reference energy are counted on rank 0 only.  The energy slots travel behind the Hamiltonian elements.
    double* en_x = e.h_diag + 3 * MAXS;
    for (int k = 0; k < E_NSLOT; k++) en_x[k] = (d.rank == 0 || k == E_ELEC || k == E_VDW) ? d.en[k] : 0.0;
    double H11 = d.en[E_ELEC] + d.en[E_VDW];
    if (d.rank == 0) {
      double E_elec = d.en[E_ELEC] + d.en[E_RECIP] + d.ewald_self;
      H11 = E_elec + d.en[E_VDW] + d.en[E_BOND] + d.en[E_ANGLE] + d.en[E_DIH];
      H11 = H11 + e.item_energy[0];     // item 0 = the principal diabat's EVB repulsion + reference energy
    }
    e.h_diag[0] = H11;
    return;
  }
  if (s >= S || !state_owned(s, d.rank, d.world)) return;
  // energy delta of the last hop: acceptor-topology item minus donor-topology item (ms_evb.f90:1546)
  const int ii = 2 * s;             // acceptor-topology item of the last hop; the donor-topology item precedes it
  double dE = e.item_energy[ii] - e.item_energy[ii - 1];
  e.h_diag[s] = dE;
  e.h_diag[2 * MAXS + s] = e.rcp_dE[s];
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
__global__ void k_evb_assemble(Dev d, EvbDev e, const CouplingGeo* geo) {
  assemble_state(d, e, geo, blockIdx.x * blockDim.x + threadIdx.x);
}

// Hellmann-Feynman weights from the ground-state vector e.evec (whole CTA; caller has synchronised):
// c_s^2 (diagonal), 2 c_parent c_s (coupling) (ms_evb.f90:298-303), and the subtree sums used by the hop-tree
// de-duplication of the real-space deltas.
__device__ void hellmann_feynman_weights(const Dev& d, EvbDev& e, int S, int tid, int nth, bool with_status = true) {
  // the hop decision for the commit kernels, and how many diabats' accumulators the next step clears ahead of its enumeration
  if (tid == 0) { e.plan->hop = (e.result[1] != *d.hydronium) ? 1 : 0; e.plan->n_clear = min(MAXS, S + 8); }
  // error flags and energy slots ride along in the solver's read-back block
  if (with_status && tid < 4) e.status_copy[tid] = (double)d.err_flag[tid];
  if (with_status && tid < E_NSLOT) e.status_copy[4 + tid] = (d.world > 1) ? e.h_diag[3 * MAXS + tid] : d.en[tid];
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
// K12a: ground state of the EVB Hamiltonian from its TREE structure (default solver).
// A diabat couples only to its DFS parent (ms_evb.f90:675-681), so H is a symmetric tree matrix of depth
// <= evb_max_chain: diagonal H_ss, one off-diagonal beta_s = H_parent(s),s per diabat.  Eliminating the leaves first,
// the pivots of H - x are   d_s(x) = delta_s - x - sum_{children c} beta_c^2 / d_c(x)   (delta_s = H_ss - H_11, formed
// from the hop deltas directly, so the 1e5 kJ/mol common offset never enters the arithmetic), every level of the tree
// in parallel.  det(H - x) = prod d_s, and all d_s > 0 exactly when x is below the lowest eigenvalue (Sylvester), so:
//   * lowest eigenvalue: Laguerre's iteration on the characteristic polynomial through sum d'/d and its derivative
//     (d', d'' by the same recurrence), started left of the spectrum (Gershgorin, or the previous step's value):
//     monotone from the left, cubic, every pivot stays positive -- 4-6 evaluations; a step that rounding pushes onto
//     the root retreats geometrically;
//   * eigenvector: inverse iteration with the factorisation at the converged (valid) shift, 2-3 tree solves;
//   * energy: Rayleigh quotient of that vector.
// One evaluation costs ~(depth+1) dependent divisions, independent of S; the whole solve is ~10 us where the cyclic
// Jacobi sweeps (K12b, kept as the general-purpose cross-check, RPB_EVB_SOLVER=jacobi) need 100-300 us of barriers.
// Same selections afterwards as the reference (ms_evb.f90:279-328): principal diabat = first maximum |c_i|.
// ================================================================================================
#define TREE_TPB 96
static_assert(MAXS <= TREE_TPB, "one thread per diabat");
struct TreeShared {
  double delta[MAXS], beta[MAXS], b2[MAXS], dv[MAXS], d1[MAXS], d2[MAXS], invd[MAXS], y[MAXS], b[MAXS];
  int parent[MAXS], level[MAXS], child_start[MAXS + 1], child[MAXS];
  double red[2][4];
  int redi[4];
};

// sums two values over the CTA; every thread returns the same (deterministically ordered) totals
__device__ __forceinline__ void tree_sum2(TreeShared& T, double& a, double& b, int tid) {
  a = warp_sum(a); b = warp_sum(b);
  __syncthreads();
  if ((tid & 31) == 0) { T.red[0][tid >> 5] = a; T.red[1][tid >> 5] = b; }
  __syncthreads();
  a = T.red[0][0] + T.red[0][1] + T.red[0][2];
  b = T.red[1][0] + T.red[1][1] + T.red[1][2];
}

// pivots (and their first two derivatives) of H - x, leaves first; returns true when every pivot is positive
__device__ __forceinline__ bool tree_pivots(TreeShared& T, int S, int maxlev, int i, double x) {
  for (int L = maxlev; L >= 0; L--) {
    if (i < S && T.level[i] == L) {
      double s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int k = T.child_start[i]; k < T.child_start[i + 1]; k++) {
        const int c = T.child[k];
        const double r = T.invd[c], br = T.b2[c] * r, q = T.d1[c] * r;
        s0 += br; s1 += br * q; s2 += br * (T.d2[c] * r - 2.0 * q * q);
      }
      const double di = T.delta[i] - x - s0;
      T.dv[i] = di; T.d1[i] = -1.0 + s1; T.d2[i] = s2; T.invd[i] = 1.0 / di;
    }
    __syncthreads();
  }
  return __syncthreads_and(i >= S || T.dv[i] > 0.0) != 0;
}

// geo != nullptr (single rank): the Hamiltonian elements are assembled here first (one launch less on the critical path)
// defer_principal: the ground state only needs the diabats' energies RELATIVE to H_11, so the solver can run while the
// principal diabat's pair forces are still being computed; H_11 itself, the absolute energies and the status / energy
// slots of the read-back block are then filled in by k_evb_finalize_principal.
struct PeerArgsH { PeerArgs a; int on; };     // on = 0: no exchange in the solver (single rank, or done by k_peer_allreduce)

__global__ void __launch_bounds__(TREE_TPB) k_evb_tree_solver(Dev d, EvbDev e_in, const double* coeff_override, const CouplingGeo* geo,
                                                                int defer_principal, const PeerArgsH px) {
  __shared__ TreeShared T;
  EvbDev e = e_in;
  const int S = *e.n_states;
  const int tid = threadIdx.x, nth = blockDim.x, i = tid;
  if (geo) {
    assemble_state(d, e, geo, tid, defer_principal != 0);
    __syncthreads();     // h_diag is read back below by the same CTA
  }
  if (px.on) {
    // state-sharded step: this rank's partial Hamiltonian block (just assembled into its exchange arena) is summed with the
    // peers' over NVLink right here -- no assembly kernel, no all-reduce kernel, no launch gaps between them and the solver
    peer_allreduce_cta(px.a);
    e.h_diag = px.a.out;
  }
  const double h11 = defer_principal ? 0.0 : e.h_diag[0];
  if (coeff_override) {
    for (int k = tid; k < S; k += nth) e.evec[k] = coeff_override[k];
    __syncthreads();
    hellmann_feynman_weights(d, e, S, tid, nth);
    return;
  }
  if (i < S) {
    // H_ss = H_11 + sum of the hop deltas along the chain (root first) + (E_rec(s) - E_rec(1))   ms_evb.f90:1546, 2083
    int chain[MAXC + 1], nc = 0;
    for (int t = i; t > 0 && nc <= MAXC; t = e.parent[t]) chain[nc++] = t;
    double Hs = h11, dl = 0.0;
    for (int k = nc - 1; k >= 0; k--) { Hs = Hs + e.h_diag[chain[k]]; dl = dl + e.h_diag[chain[k]]; }
    Hs = Hs + e.h_diag[2 * MAXS + i]; dl = dl + e.h_diag[2 * MAXS + i];
    e.h_full[i] = Hs; e.h_full[MAXS + i] = (i > 0) ? e.h_diag[MAXS + i] : 0.0;
    T.delta[i] = (i > 0) ? dl : 0.0;
    T.beta[i] = (i > 0) ? e.h_diag[MAXS + i] : 0.0;
    T.b2[i] = T.beta[i] * T.beta[i];
    T.parent[i] = (i > 0) ? e.parent[i] : -1;
    T.level[i] = nc;
    T.y[i] = 1.0;
  }
  __syncthreads();
  // children of every diabat in ascending order (deterministic summation), levels, Gershgorin bound: one thread per diabat
  if (i < S) {
    int before = 0, nch = 0;
    const int p = T.parent[i];
    for (int k = 1; k < S; k++) { const int pk = T.parent[k]; before += (k < i && pk == p) ? 1 : 0; nch += (pk == i) ? 1 : 0; }
    T.child_start[i + 1] = nch;        // counts first; scanned below
    T.b[i] = (double)before;           // (scratch) rank among the siblings
  }
  if (tid == 0) T.child_start[0] = 0;
  __syncthreads();
  if (tid == 0) for (int k = 0; k < S; k++) T.child_start[k + 1] += T.child_start[k];
  __syncthreads();
  if (i >= 1 && i < S) T.child[T.child_start[T.parent[i]] + (int)T.b[i]] = i;
  __syncthreads();
  // deepest level, Gershgorin lower bound of the spectrum and the magnitude of the problem (block reductions in shared memory)
  {
    double g = 1e300, m = 0.0;
    int lv = 0;
    if (i < S) {
      g = T.delta[i] - fabs(T.beta[i]);
      for (int q = T.child_start[i]; q < T.child_start[i + 1]; q++) g -= fabs(T.beta[T.child[q]]);
      m = fabs(T.delta[i]); lv = T.level[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g = fmin(g, __shfl_xor_sync(0xffffffffu, g, o)); m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o)); lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, o));
    }
    if ((tid & 31) == 0) { T.red[0][tid >> 5] = g; T.red[1][tid >> 5] = m; T.redi[tid >> 5] = lv; }
    __syncthreads();
  }
  const double gers = fmin(T.red[0][0], fmin(T.red[0][1], T.red[0][2]));
  const double mag = fmax(T.red[1][0], fmax(T.red[1][1], T.red[1][2]));
  const int mlev = max(T.redi[0], max(T.redi[1], T.redi[2]));
  __syncthreads();
  double lo = gers - 1e-3 * (1.0 + fabs(gers)), hi = 1e300;
  const double tol = 1e-10 * fmax(fmax(mag, fabs(lo)), 1.0);   // the inverse iteration + Rayleigh quotient finish the job
  double x = lo;
  const double warm = e.tree_mu[0];
  bool warm_try = (e.tree_mu[1] == 1.0) && (warm - 2.0 > lo);
  if (warm_try) x = warm - 2.0;
  int n_eval = 0, retreat = 0, status = 1;
  if (S == 1) { status = 0; x = 0.0; }
  else {
    for (int it = 0; it < 120; it++) {
      const bool ok = tree_pivots(T, S, mlev, i, x);
      n_eval++;
      if (!ok) {
        if (warm_try && it == 0) { x = lo; warm_try = false; continue; }    // the previous value is no lower bound any more
        hi = fmin(hi, x);
        if (hi - lo <= tol) { x = lo; tree_pivots(T, S, mlev, i, x); n_eval++; status = 0; break; }
        x = fmax(hi - 0.5 * tol * pow(8.0, (double)retreat), 0.5 * (lo + hi));
        retreat++;
        continue;
      }
      lo = x;
      if (hi - lo <= tol) { status = 0; break; }
      double G = 0.0, S2 = 0.0;
      if (i < S) { const double q = T.d1[i] * T.invd[i]; G = q; S2 = T.d2[i] * T.invd[i] - q * q; }
      tree_sum2(T, G, S2, tid);
      const double n = (double)S;
      const double disc = fmax((n - 1.0) * (n * (-S2) - G * G), 0.0);
      const double a = n / (G - sqrt(disc));          // G < 0 left of the spectrum: a < 0, the step goes right
      double xn = x - a;
      if (xn - x <= tol) { status = 0; break; }
      if (!(xn < hi)) xn = 0.5 * (lo + hi);
      x = xn;
    }
  }
  // ---- eigenvector: inverse iteration with the factorisation at the (valid) shift x
  double mu = 0.0;
  if (S > 1) {
    for (int rep = 0; rep < 4; rep++) {
      if (i < S) T.b[i] = T.y[i];
      __syncthreads();
      for (int L = mlev; L >= 0; L--) {        // forward elimination, leaves first
        if (i < S && T.level[i] == L) {
          double acc = T.b[i];
          for (int k = T.child_start[i]; k < T.child_start[i + 1]; k++) { const int c = T.child[k]; acc -= T.beta[c] * (T.b[c] * T.invd[c]); }
          T.b[i] = acc;
        }
        __syncthreads();
      }
      for (int L = 0; L <= mlev; L++) {        // back substitution, root first
        if (i < S && T.level[i] == L) T.b[i] = (T.b[i] - (i > 0 ? T.beta[i] * T.b[T.parent[i]] : 0.0)) * T.invd[i];
        __syncthreads();
      }
      double nrm = (i < S) ? T.b[i] * T.b[i] : 0.0, zero = 0.0;
      tree_sum2(T, nrm, zero, tid);
      const double inv = rsqrt(nrm);
      double change = 0.0;
      if (i < S) { const double yn = T.b[i] * inv; change = fabs(fabs(yn) - fabs(T.y[i])); T.y[i] = yn; }
      n_eval++;
      // the shift sits ~1e-10 (relative) below the eigenvalue, so one solve contracts the error by ~1e-9: once a solve
      // changed the vector by less than 1e-8 the vector before it was already that close, and this one is converged
      if (__syncthreads_and(change < (rep >= 1 ? 1e-8 : 1e-15))) break;
    }
    double r1 = 0.0, r2 = 0.0;
    if (i < S) { r1 = T.delta[i] * T.y[i] * T.y[i]; if (i > 0) r2 = 2.0 * T.beta[i] * T.y[i] * T.y[T.parent[i]]; }
    tree_sum2(T, r1, r2, tid);
    mu = r1 + r2;
  }
  if (i < S) e.evec[i] = T.y[i];
  if (tid == 0) {
    *e.e_ground = h11 + mu;
    int pd = 0;
    double coef = fabs(T.y[0]);
    for (int k = 0; k < S; k++) if (coef < fabs(T.y[k])) { coef = fabs(T.y[k]); pd = k; }
    int newh = *d.hydronium;
    for (int h = 0; h < d.max_chain; h++) {
      if (e.proton_log[(pd * MAXC + h) * 5] < 0) break;
      newh = e.proton_log[(pd * MAXC + h) * 5 + 3];
    }
    e.result[0] = pd; e.result[1] = newh; e.result[2] = status; e.result[3] = 0; e.result[4] = n_eval;
    e.tree_mu[0] = mu; e.tree_mu[1] = (status == 0) ? 1.0 : 0.0;
  }
  __syncthreads();
  hellmann_feynman_weights(d, e, S, tid, nth, defer_principal == 0);
}

// H_11 = energy of the principal diabat (calculate_total_force_energy + EVB repulsion + reference energy, ms_evb.f90:411-436)
// once its pair forces and bonded terms are complete; absolute H_ss, the adiabatic energy and the status / energy slots of
// the read-back block.  One CTA of MAXS threads, behind k_evb_tree_solver(defer_principal = 1).
__global__ void k_evb_finalize_principal(Dev d, EvbDev e) {
  const int tid = threadIdx.x, S = *e.n_states;
  const double E_elec = d.en[E_ELEC] + d.en[E_RECIP] + d.ewald_self;
  double H11 = E_elec + d.en[E_VDW] + d.en[E_BOND] + d.en[E_ANGLE] + d.en[E_DIH];
  H11 = H11 + e.item_energy[0];
  if (tid < S) e.h_full[tid] = H11 + e.h_full[tid];
  if (tid == 0) { e.h_diag[0] = H11; *e.e_ground = H11 + *e.e_ground; }
  if (tid < 4) e.status_copy[tid] = (double)d.err_flag[tid];
  if (tid < E_NSLOT) e.status_copy[4 + tid] = d.en[tid];
}

// ================================================================================================
// K12b: block-level Jacobi eigensolver for the (<= 80 x 80) EVB Hamiltonian, one CTA, matrix in shared memory.
// Same rotation definition, thresholds and stopping logic as the reference's Numerical-Recipes routine
// (general_routines.f90:2035-2074), but the n/2 disjoint rotations of a round-robin round are applied
// concurrently.  A round is two phases:
//   A  one thread per pair (p,q): rotation parameters and the 2x2 diagonal block (a_pp -= t a_pq, a_qq += t a_pq,
//      a_pq = 0, :2056-2062).  The parameters use the hypot form of the same angle, t = sgn(z) b / (|z| + sqrt(z^2+b^2))
//      with z = (a_qq-a_pp)/2, b = a_pq, and (c, s) = (u, sgn(z) b) / sqrt(u^2+b^2), u = |z| + sqrt(z^2+b^2): two
//      dependent long-latency operations (rsqrt) instead of five (div, sqrt, div, sqrt, div) -- this serial chain, not
//      the arithmetic volume, is what a round costs;
//   B  one thread per off-diagonal 2x2 block above the diagonal, A_IJ <- R_I^T A_IJ R_J (mirrored into the lower
//      triangle), and one thread per 2x2 block of V <- V R_J.
// The ground-state eigenvector is unique up to sign, so only the rounding-level path differs from the serial sweep.
// Followed by the reference's selections (ms_evb.f90:279-328): ground state = first minimum eigenvalue,
// principal diabat = first maximum |c_i|.
// ================================================================================================
#define JAC_TPB 512
__device__ __forceinline__ void jac_pair(int t, int r, int n, int& p, int& q) {
  if (t == 0) { p = r; q = n - 1; }
  else {
    p = r + t; if (p >= n - 1) p -= n - 1;
    q = r - t; if (q < 0) q += n - 1;
  }
  if (p > q) { int x = p; p = q; q = x; }
}

__global__ void __launch_bounds__(JAC_TPB) k_evb_jacobi(Dev d, EvbDev e, const double* coeff_override) {
  extern __shared__ double smem[];
  const int S = *e.n_states;
  const int n = S + (S & 1);          // padded to even for the round-robin pairing (dummy row/column stays zero)
  const int nb = n / 2;
  const int tid = threadIdx.x, nth = blockDim.x;
  double* a = smem;                   // [n*n] column-major, full symmetric
  double* v = a + n * n;              // [n*n]
  double* tmp = v + n * n;            // [n*n] warm-start scratch
  double* rc = tmp + n * n;           // [nb] cos
  double* rs = rc + nb;               // [nb] sin
  int* rp = (int*)(rs + nb);          // [nb] p
  int* rq = rp + nb;                  // [nb] q
  __shared__ double red[32];
  __shared__ double s_sm;
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
    // ---- warm start: when the diabat set is the one of the previous call (same hop logs), H changed by one MD step
    // only, so it is first rotated into the previous eigenbasis, A = V_prev^T H V_prev (nearly diagonal), and the
    // sweeps start from V = V_prev.  Same solver, same stopping rule; about half the sweeps.
    {
      const int chain_len = e.jac_sig[1 + MAXS * MAXC * 5];   // consecutive warm starts: bounded, so that the rounding-level
      int same = (e.jac_sig[0] == S) && chain_len < 64;         // loss of orthogonality of V_prev cannot accumulate
      if (same) for (int k = tid; k < S * MAXC * 5; k += nth) same &= (e.jac_sig[1 + k] == e.proton_log[k]);
      same = __syncthreads_and(same);
      if (same) {
        for (int k = tid; k < n * n; k += nth) v[k] = e.jac_v[k];
        __syncthreads();
        for (int k = tid; k < n * n; k += nth) {       // tmp = H V
          const int i = k % n, j = k / n;
          double acc = 0.0;
          for (int m = 0; m < n; m++) acc = fma(a[i + n * m], v[m + n * j], acc);
          tmp[k] = acc;
        }
        __syncthreads();
        for (int k = tid; k < n * n; k += nth) {       // A = V^T tmp, upper triangle mirrored
          const int i = k % n, j = k / n;
          if (i > j) continue;
          double acc = 0.0;
          for (int m = 0; m < n; m++) acc = fma(v[m + n * i], tmp[m + n * j], acc);
          a[i + n * j] = acc; a[j + n * i] = acc;
        }
        __syncthreads();
      } else {
        for (int k = tid; k < S * MAXC * 5; k += nth) e.jac_sig[1 + k] = e.proton_log[k];
        if (tid == 0) e.jac_sig[0] = S;
      }
      __syncthreads();
      if (tid == 0) e.jac_sig[1 + MAXS * MAXC * 5] = same ? chain_len + 1 : 0;
    }
    const int n_upper = nb * (nb - 1) / 2;
    int status = 1, sweeps = 0;
    for (int it = 1; it <= 50; it++) {
      double sm = 0.0;
      for (int k = tid; k < n * n; k += nth) { int i = k % n, j = k / n; if (i < j) sm += fabs(a[k]); }
      sm = block_sum(sm, red);
      if (tid == 0) s_sm = sm;
      __syncthreads();
      sm = s_sm;
      if (sm == 0.0) { status = 0; break; }
      sweeps = it;
      const double tresh = (it < 4) ? 0.2 * sm / (double)(S * S) : 0.0;
      for (int r = 0; r < n - 1; r++) {
        // ---- phase A: parameters + diagonal block of every pair
        if (tid < nb) {
          int p, q;
          jac_pair(tid, r, n, p, q);
          double cc = 1.0, sn = 0.0;
          const double apq = a[p + n * q];
          if (apq != 0.0) {
            const double app = a[p + n * p], aqq = a[q + n * q];
            const double g = 100.0 * fabs(apq);
            // (the reference applies this flush only after four sweeps, :2043; it is a no-op at working precision
            //  whenever its condition holds, and applying it from the first sweep lets a warm-started solve stop early)
            if ((fabs(app) + g == fabs(app)) && (fabs(aqq) + g == fabs(aqq))) {
              a[p + n * q] = 0.0; a[q + n * p] = 0.0;                                  // :2044-2045
            } else if (fabs(apq) > tresh) {
              const double z = 0.5 * (aqq - app), az = fabs(z);
              const double h2 = az * az + apq * apq;
              const double u = az + h2 * rsqrt(h2);
              const double sb = z < 0.0 ? -apq : apq;
              const double tt = sb / u;
              const double w = rsqrt(u * u + apq * apq);
              cc = u * w; sn = sb * w;
              a[p + n * p] = app - tt * apq; a[q + n * q] = aqq + tt * apq;            // :2056-2062
              a[p + n * q] = 0.0; a[q + n * p] = 0.0;
            }
          }
          rc[tid] = cc; rs[tid] = sn; rp[tid] = p; rq[tid] = q;
        }
        __syncthreads();
        // ---- phase B: off-diagonal blocks of A (upper triangle of the block grid, mirrored) and V
        for (int b = tid; b < n_upper + nb * nb; b += nth) {
          if (b < n_upper) {
            // (I,J), I < J, from the linear index of the strict upper triangle
            int I = 0, rem = b;
            while (rem >= nb - 1 - I) { rem -= nb - 1 - I; I++; }
            const int J = I + 1 + rem;
            const double cI = rc[I], sI = rs[I], cJ = rc[J], sJ = rs[J];
            if (sI == 0.0 && sJ == 0.0) continue;
            const int pI = rp[I], qI = rq[I], pJ = rp[J], qJ = rq[J];
            const double x00 = a[pI + n * pJ], x10 = a[qI + n * pJ], x01 = a[pI + n * qJ], x11 = a[qI + n * qJ];
            const double y00 = cI * x00 - sI * x10, y10 = sI * x00 + cI * x10;   // rows p_I, q_I <- R_I^T A
            const double y01 = cI * x01 - sI * x11, y11 = sI * x01 + cI * x11;
            const double z00 = cJ * y00 - sJ * y01, z01 = sJ * y00 + cJ * y01;   // columns p_J, q_J <- A R_J
            const double z10 = cJ * y10 - sJ * y11, z11 = sJ * y10 + cJ * y11;
            a[pI + n * pJ] = z00; a[pJ + n * pI] = z00; a[pI + n * qJ] = z01; a[qJ + n * pI] = z01;
            a[qI + n * pJ] = z10; a[pJ + n * qI] = z10; a[qI + n * qJ] = z11; a[qJ + n * qI] = z11;
          } else {
            const int bb = b - n_upper, I = bb / nb, J = bb - I * nb;
            const double cJ = rc[J], sJ = rs[J];
            if (sJ == 0.0) continue;
            const int pJ = rp[J], qJ = rq[J];
#pragma unroll
            for (int k = 2 * I; k < 2 * I + 2; k++) {
              const double vx = v[k + n * pJ], vy = v[k + n * qJ];
              v[k + n * pJ] = cJ * vx - sJ * vy;
              v[k + n * qJ] = sJ * vx + cJ * vy;
            }
          }
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
      e.result[0] = pd; e.result[1] = newh; e.result[2] = status; e.result[3] = ground; e.result[4] = sweeps;
    }
    __syncthreads();
    int ground = e.result[3];
    for (int i = tid; i < S; i += nth) e.evec[i] = v[i + n * ground];
    if (status == 0) for (int k = tid; k < n * n; k += nth) e.jac_v[k] = v[k];
    else if (tid == 0) e.jac_sig[0] = -1;
  }
  __syncthreads();
  hellmann_feynman_weights(d, e, S, tid, nth);
}

// ================================================================================================
// K13: Hellmann-Feynman mixing
// ================================================================================================
// f_mix = [rank 0: principal force] + sum_s c_s^2 dF_s + 2 c_p c_s Foff_s   over owned diabats
// in_place (single rank, ground-state mix): the principal-diabat force is still in d.force; it is saved to dF slot 0
// (the per-state debug accessors need it later) and d.force receives the mixed force -- no copies before or after.
__global__ void k_evb_mix_forces(Dev d, EvbDev e, int include_principal, int in_place) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n3 = (size_t)3 * d.N;
  if (i >= n3) return;
  const int n_list = e.plan->n_own;
  const int* __restrict__ state_list = e.plan->state_list;
  double f = 0.0;
  if (in_place) { f = d.force[i]; e.dF[i] = f; }
  else if (include_principal) f = e.dF[i];        // dF slot 0 holds the principal-diabat force (without F_rec)
  for (int k = 0; k < n_list; k++) {
    int s = state_list[k];
    f = fma(e.coef2[2 * MAXS + s], e.dF[(size_t)s * n3 + i], f);
    f = fma(e.coef2[MAXS + s], e.Foff[(size_t)s * n3 + i], f);
  }
  (in_place == 1 ? d.force : e.f_mix)[i] = f;     // in_place == 2 (sharded): principal force read and saved as for 1, partial sum into the exchange arena
}


// ================================================================================================
// Reciprocal space of the diabats WITHOUT per-diabat grids ("delta algebra").
//
// The reference copies the principal Q grid per diabat, re-spreads the donor/acceptor molecules, and runs one full
// 3-D FFT convolution per diabat (ms_evb.f90:1445-1556, 1962-2248).  A diabat differs from the principal one only in
// the CHARGES of the atoms of its chain molecules -- the positions, hence the B-spline footprints w_a(x), are the
// principal ones.  With dq_a(s) = q_a(s) - q_a(1), g = IDFT(CB) the real-space Green function of the convolution and
//     P_a  = <w_a, theta_1>          G_a  = <grad w_a, theta_1>          (one 216-point gather per chain atom)
//     M_ab = <w_a, g * w_b>          N_ab = <grad w_a, g * w_b>          (pairs of chain atoms that share a diabat)
// linearity of the convolution gives, exactly:
//     E_rec(s) - E_rec(1) = sum_a dq_a P_a + 1/2 sum_ab dq_a dq_b M_ab                     -> H_ss  (ms_evb.f90:2083)
//     sum_s c_s^2 theta_s = theta_1 + g * sum_a D_a w_a,   D_a = sum_s c_s^2 dq_a(s)        -> ONE more convolution of
//         the grid carrying the Hellmann-Feynman averaged charges, gathered once for all atoms (ms_evb.f90:292-309),
//     chain atom a:  F_a += -K kk sum_{s contains a} c_s^2 dq_a(s) [ G_a + sum_b dq_b(s) N_ab ]   (ms_evb.f90:2171-2228)
// M_ab needs no grid at all: w is a product of 1-D splines, so <w_a, g * w_b> = sum over the 11^3 displacements t of
// g(n_a - n_b - t) cx(t_x) cy(t_y) cz(t_z) with the 1-D cross-correlations c(t) = sum_{k-k'=t} w_a[k] w_b[k'].
// Two convolutions per step instead of S+1, no per-diabat grid traffic; differences from the per-diabat transforms are
// rounding only (~1e-13 relative, checked in tests/ against the oracle's literal per-diabat grids, every diabat's own force).
// ================================================================================================
#define RA_MOLS RPB_RA_MOLS           // distinct chain molecules of a step
#define RA_SLOTS (RA_MOLS * MA)       // chain-atom slots: (chain molecule, atom offset inside the molecule)
#define RA_MAXPAIR RPB_RA_MAXPAIR     // ordered pairs of chain molecules that share a diabat
#define RA_ENT (CM * MA)              // chain atoms of one diabat

struct RecipDev {
  const EvbPlan* plan;     // cmol[n_cmol]: distinct chain molecules; molpair[n_pair]: i * RA_MOLS + j   (enumeration kernel)
  const int* mol_slot;     // [M] molecule -> index in plan->cmol
  double* P; double* G;    // [RA_SLOTS], [RA_SLOTS][3]   (conv included)
  double* Mx; double* Nx;  // [RA_SLOTS][RA_SLOTS], [3][RA_SLOTS][RA_SLOTS]
  double* D;               // [RA_SLOTS] Hellmann-Feynman averaged charge deltas
  int* st_n; int* st_slot; double* st_dq;   // per diabat: chain atoms with their charge deltas  [MAXS], [MAXS][RA_ENT]
  int* sl_n; int* sl_state; double* sl_dq;  // the same table by chain-atom slot: diabats that change its charge  [RA_SLOTS], [RA_SLOTS][MAXS]
  const double* gtab;      // K^3 Green function g = IDFT(CB)
};

// spline weights (lanes 0..17) and derivative factors of the atom with scaled coordinates u, as gather_atom_warp
__device__ __forceinline__ void spline_pair_lane(const Dev& d, const double u[3], int np[3], int lane, double& b6, double& dm) {
  np[0] = (int)floor(u[0]); np[1] = (int)floor(u[1]); np[2] = (int)floor(u[2]);
  b6 = 0.0; dm = 0.0;
  if (lane < 18) {
    int dim = lane / 6, k = lane - 6 * dim;
    double arg1 = u[dim] - (double)(np[dim] - k);
    double arg2 = arg1 - 1.0;
    int g1n = (int)ceil(arg1 / 6.0 * d.spline_grid);
    b6 = __ldg(&d.B6[g1n - 1]);
    if (arg1 < 5.0) { int g = (int)ceil(arg1 / 5.0 * d.spline_grid); dm = __ldg(&d.B5[g - 1]); }
    if (0.0 < arg2) { int g = (int)ceil(arg2 / 5.0 * d.spline_grid); dm = dm - __ldg(&d.B5[g - 1]); }
  }
}

// warp per chain-atom slot: P_a, G_a from theta_1; also publishes the molecule -> slot map and clears D
__global__ void k_evb_rcp_atoms(Dev d, RecipDev r) {
  const int lane = threadIdx.x & 31;
  const int n_mol = r.plan->n_cmol;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_mol * MA; w += (gridDim.x * blockDim.x) >> 5) {
  const int im = w / MA, a = w % MA, mol = r.plan->cmol[im];
  if (lane == 0) r.sl_n[w] = 0;
  if (a >= d.mol_natom[mol]) continue;
  const int atom = d.mol_first[mol] + a;
  const double u[3] = {d.uscale[3 * atom], d.uscale[3 * atom + 1], d.uscale[3 * atom + 2]};
  int np[3];
  double b6, dm;
  spline_pair_lane(d, u, np, lane, b6, dm);
  const int K = d.K;
  double p = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll 1
  for (int it = 0; it < 7; it++) {
    int pt = it * 32 + lane;
    int pp = pt < 216 ? pt : 215;
    int k1 = pp % 6, k2 = (pp / 6) % 6, k3 = pp / 36;
    double w1 = __shfl_sync(0xffffffffu, b6, k1), w2 = __shfl_sync(0xffffffffu, b6, 6 + k2), w3 = __shfl_sync(0xffffffffu, b6, 12 + k3);
    double d1 = __shfl_sync(0xffffffffu, dm, k1), d2 = __shfl_sync(0xffffffffu, dm, 6 + k2), d3 = __shfl_sync(0xffffffffu, dm, 12 + k3);
    if (pt < 216) {
      int n1 = np[0] - k1; if (n1 < 0) n1 += K;
      int n2 = np[1] - k2; if (n2 < 0) n2 += K;
      int n3 = np[2] - k3; if (n3 < 0) n3 += K;
      double th = __ldg(&d.theta[(size_t)n1 + (size_t)K * n2 + (size_t)K * K * n3]) * d.conv;
      p = fma(w1 * w2 * w3, th, p);
      g0 = fma(d1 * w2 * w3, th, g0); g1 = fma(d2 * w1 * w3, th, g1); g2 = fma(d3 * w1 * w2, th, g2);
    }
  }
  p = warp_sum(p); g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
  if (lane == 0) { r.P[w] = p; r.G[3 * w] = g0; r.G[3 * w + 1] = g1; r.G[3 * w + 2] = g2; }
  }
}

// warp per (ordered pair of chain molecules, atom a of the first, atom b of the second): M_ab, N_ab
__global__ void __launch_bounds__(128) k_evb_rcp_pairs(Dev d, RecipDev r) {
  __shared__ double sh_c[4][2][3][11];      // per warp: cross-correlations cw (0) / cd (1) per dimension, t = -5..5
  __shared__ double sh_w[4][3][18];         // per warp: w_a, dw_a, w_b
  __shared__ int sh_i[4][3][11];            // per warp: wrapped, stride-scaled grid offset of displacement t per dimension
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pair = r.plan->n_pair;
  for (int w = blockIdx.x * 4 + wib; w < n_pair * MA * MA; w += gridDim.x * 4) {     // (warp-private shared slots: no CTA barrier inside)
  __syncwarp();
  int sa = 0, sb = 0, dn[3] = {0, 0, 0};
  bool act = false;
  {
    const int ip = w / (MA * MA), a = (w / MA) % MA, b = w % MA;
    const int mi = r.plan->molpair[ip] / RA_MOLS, mj = r.plan->molpair[ip] % RA_MOLS;
    const int moli = r.plan->cmol[mi], molj = r.plan->cmol[mj];
    act = a < d.mol_natom[moli] && b < d.mol_natom[molj];
    if (act) {
      sa = mi * MA + a; sb = mj * MA + b;
      const int ia = d.mol_first[moli] + a, ib = d.mol_first[molj] + b;
      const double ua[3] = {d.uscale[3 * ia], d.uscale[3 * ia + 1], d.uscale[3 * ia + 2]};
      const double ub[3] = {d.uscale[3 * ib], d.uscale[3 * ib + 1], d.uscale[3 * ib + 2]};
      int na[3], nb[3];
      double wa, da, wb, db;   // (db unused: the gradient acts on the first atom)
      spline_pair_lane(d, ua, na, lane, wa, da);
      spline_pair_lane(d, ub, nb, lane, wb, db);
      for (int c = 0; c < 3; c++) dn[c] = na[c] - nb[c];
      if (lane < 18) { sh_w[wib][0][lane] = wa; sh_w[wib][1][lane] = da; sh_w[wib][2][lane] = wb; }
    }
  }
  __syncwarp();
  if (!act) continue;
  // c(t) = sum_{k - k' = t} f_a[k] w_b[k'],  lane -> (dimension, t): 33 values, two rounds
  for (int v = lane; v < 33; v += 32) {
    const int dim = v / 11, t = v % 11 - 5;
    double cw = 0.0, cd = 0.0;
    for (int k = 0; k < 6; k++) {
      const int kp = k - t;
      if (kp >= 0 && kp < 6) { cw = fma(sh_w[wib][0][6 * dim + k], sh_w[wib][2][6 * dim + kp], cw); cd = fma(sh_w[wib][1][6 * dim + k], sh_w[wib][2][6 * dim + kp], cd); }
    }
    sh_c[wib][0][dim][t + 5] = cw; sh_c[wib][1][dim][t + 5] = cd;
  }
  __syncwarp();
  const int K = d.K;
  // wrapped grid offsets of the 11 displacements per dimension (no integer division in the hot loop)
  for (int v = lane; v < 33; v += 32) {
    const int dim = v / 11, t = v % 11 - 5;
    int gi = (dn[dim] - t) % K; if (gi < 0) gi += K;
    sh_i[wib][dim][t + 5] = gi * (dim == 0 ? 1 : (dim == 1 ? K : K * K));
  }
  __syncwarp();
  double m = 0.0, n0 = 0.0, n1 = 0.0, n2 = 0.0;
  // 1331 displacements, 6 batches of 7 per lane: the 7 Green-function values of a batch are in flight together
  // (a plain loop exposes one L2 round trip per value)
#pragma unroll 1
  for (int base = 0; base < 1331; base += 32 * 7) {
    double gv[7];
#pragma unroll
    for (int q = 0; q < 7; q++) {
      const int pt = base + 32 * q + lane;
      gv[q] = 0.0;
      if (pt < 1331) { const int tx = pt % 11, ty = (pt / 11) % 11, tz = pt / 121; gv[q] = __ldg(&r.gtab[sh_i[wib][0][tx] + sh_i[wib][1][ty] + sh_i[wib][2][tz]]); }
    }
#pragma unroll
    for (int q = 0; q < 7; q++) {
      const int pt = base + 32 * q + lane;
      if (pt < 1331) {
        const int tx = pt % 11, ty = (pt / 11) % 11, tz = pt / 121;
        const double g = gv[q];
        const double cx = sh_c[wib][0][0][tx], cy = sh_c[wib][0][1][ty], cz = sh_c[wib][0][2][tz];
        const double ex = sh_c[wib][1][0][tx], ey = sh_c[wib][1][1][ty], ez = sh_c[wib][1][2][tz];
        const double gyz = g * cy * cz;
        m = fma(gyz, cx, m);
        n0 = fma(gyz, ex, n0);
        n1 = fma(g * cx * cz, ey, n1);
        n2 = fma(g * cx * cy, ez, n2);
      }
    }
  }
  m = warp_sum(m); n0 = warp_sum(n0); n1 = warp_sum(n1); n2 = warp_sum(n2);
  if (lane == 0) {
    const size_t o = (size_t)sa * RA_SLOTS + sb, plane = (size_t)RA_SLOTS * RA_SLOTS;
    r.Mx[o] = m * d.conv; r.Nx[o] = n0 * d.conv; r.Nx[plane + o] = n1 * d.conv; r.Nx[2 * plane + o] = n2 * d.conv;
  }
  }
}

// warp per diabat s >= 1: chain atoms of the FINAL topology with their charge deltas, and E_rec(s) - E_rec(1)
__global__ void k_evb_rcp_energy(Dev d, EvbDev e, RecipDev r) {
  __shared__ int sh_slot[4][RA_ENT];
  __shared__ double sh_dq[4][RA_ENT];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = *e.n_states;
  for (int s = blockIdx.x * 4 + wib; s < S; s += gridDim.x * 4) {
  __syncwarp();
  if (s == 0) { if (lane == 0) r.st_n[0] = 0; continue; }
  static_assert(RA_ENT == 32, "one lane per (chain molecule, atom) of the snapshot");
  const Snapshot& Sn = e.snap[s * NLEV + e.n_hops[s]];
  const int k = lane / MA, a = lane % MA;
  int slot = -1;
  double dq = 0.0;
  if (k < Sn.n_mol && a < Sn.m[k].n_atom) {
    const int atom = Sn.m[k].atom[a];
    const int mol = d.mol_of_atom[atom];
    slot = r.mol_slot[mol] * MA + (atom - d.mol_first[mol]);
    dq = Sn.m[k].q[a] - d.xq[atom].w;
  }
  // compact the entries with a charge delta
  const unsigned keep = __ballot_sync(0xffffffffu, slot >= 0 && dq != 0.0);
  const int n = __popc(keep), pos = __popc(keep & ((1u << lane) - 1u));
  if (slot >= 0 && dq != 0.0) {
    sh_slot[wib][pos] = slot; sh_dq[wib][pos] = dq; r.st_slot[s * RA_ENT + pos] = slot; r.st_dq[s * RA_ENT + pos] = dq;
    const int k2 = atomicAdd(&r.sl_n[slot], 1);          // < MAXS: one entry per diabat at most
    r.sl_state[slot * MAXS + k2] = s; r.sl_dq[slot * MAXS + k2] = dq;
  }
  if (lane == 0) r.st_n[s] = n;
  __syncwarp();
  double acc = 0.0;
  if (lane < n) acc = sh_dq[wib][lane] * r.P[sh_slot[wib][lane]];
  for (int q0 = 0; q0 < n * n; q0 += 128) {          // pairs (a, b) dealt to the lanes, four independent loads in flight
    double mv[4], ww[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int q = q0 + 32 * u + lane;
      mv[u] = 0.0; ww[u] = 0.0;
      if (q < n * n) { const int a2 = q / n, b2 = q - a2 * n; ww[u] = 0.5 * sh_dq[wib][a2] * sh_dq[wib][b2]; mv[u] = r.Mx[(size_t)sh_slot[wib][a2] * RA_SLOTS + sh_slot[wib][b2]]; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) acc = fma(ww[u], mv[u], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) e.rcp_dE[s] = acc;
  }
}

// after the solver: thread per (diabat, chain-atom entry): the chain atom's own reciprocal force term
__global__ void k_evb_rcp_mix(Dev d, EvbDev e, RecipDev r, double* __restrict__ out) {
  const int S = *e.n_states;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < S * RA_ENT; t += gridDim.x * blockDim.x) {
  const int s = t / RA_ENT, k = t % RA_ENT;
  if (s < 1 || k >= r.st_n[s]) continue;
  const double w = e.coef2[s];
  const int sa = r.st_slot[s * RA_ENT + k];
  const double dqa = r.st_dq[s * RA_ENT + k];
  const size_t plane = (size_t)RA_SLOTS * RA_SLOTS;
  double f0 = r.G[3 * sa], f1 = r.G[3 * sa + 1], f2 = r.G[3 * sa + 2];
  const int n = r.st_n[s];
  for (int j = 0; j < n; j++) {
    const size_t o = (size_t)sa * RA_SLOTS + r.st_slot[s * RA_ENT + j];
    const double dqb = r.st_dq[s * RA_ENT + j];
    f0 = fma(dqb, r.Nx[o], f0); f1 = fma(dqb, r.Nx[plane + o], f1); f2 = fma(dqb, r.Nx[2 * plane + o], f2);
  }
  const int mol = r.plan->cmol[sa / MA], atom = d.mol_first[mol] + sa % MA;
  const double Kd = (double)d.K, c = w * dqa;
  atomicAdd(&out[3 * atom], -(Kd * d.kk[0]) * (c * f0));
  atomicAdd(&out[3 * atom + 1], -(Kd * d.kk[1]) * (c * f1));
  atomicAdd(&out[3 * atom + 2], -(Kd * d.kk[2]) * (c * f2));
  }
}

// Q_mix = Q_1 + sum_a D_a w_a,  D_a = sum_s c_s^2 dq_a(s): the copy of Q_1 is made early (k_copy); one warp per chain-atom
// slot collects its averaged charge delta from the diabats' tables and spreads it
__global__ void k_evb_rcp_patch(Dev d, EvbDev e, RecipDev r, double* __restrict__ Qmix) {
  const int lane = threadIdx.x & 31;
  const int n_mol = (*e.n_states > 1) ? r.plan->n_cmol : 0;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_mol * MA; w += (gridDim.x * blockDim.x) >> 5) {
  const int mol = r.plan->cmol[w / MA], a = w % MA;
  if (a >= d.mol_natom[mol]) continue;
  const int n = r.sl_n[w];
  double D = 0.0;
  for (int k = lane; k < n; k += 32) D = fma(e.coef2[r.sl_state[w * MAXS + k]], r.sl_dq[w * MAXS + k], D);
  D = __shfl_sync(0xffffffffu, warp_sum(D), 0);   // warp_sum leaves the total in lane 0
  if (D == 0.0) continue;
  const int atom = d.mol_first[mol] + a;
  const double u[3] = {d.uscale[3 * atom], d.uscale[3 * atom + 1], d.uscale[3 * atom + 2]};
  spread_atom_warp(d, Qmix, u, D, 1.0, lane);
  }
}

// Reference quirk (ms_evb.f90:2523-2656): every force array of diabat s -- its diagonal force, its reciprocal-space
// force, its coupling force -- is mapped back to the principal atom order by undoing the proton TRANSFERS only; the
// re-ordering of a protonated acceptor to its molecule-type template (reorder_molecule_data_structures, :941-1006, e.g. a
// sulfonate protonated on an oxygen that is not its last) is not undone, so the diabat's force on such an atom g is
// credited to the atom t that sits at g's position in the principal order.  Everything here is computed on the true
// atoms; this kernel moves the mixed contribution X_s(g) = c_s^2 F_s(g) + 2 c_p c_s Foff_s(g) from g to t, one warp
// per diabat, one lane per (chain molecule, atom) of its final topology.  (Water / hydronium never re-order.)
__global__ void k_evb_reorder_quirk(Dev d, EvbDev e, RecipDev r, double* __restrict__ out, int recip_algebra) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = *e.n_states;
  for (int s = blockIdx.x * 4 + wib; s < S; s += gridDim.x * 4) {
  if (s < 1) continue;
  const Snapshot& Sn = e.snap[s * NLEV + e.n_hops[s]];
  const int k = lane / MA, a = lane % MA;
  if (k >= Sn.n_mol || a >= Sn.m[k].n_atom) continue;
  const int g = Sn.m[k].atom[a], t = Sn.m[k].ratom[a];
  if (g == t) continue;
  const size_t n3 = (size_t)3 * d.N;
  const double w = e.coef2[s], wc = e.coef2[MAXS + s];
  double X[3] = {0.0, 0.0, 0.0};
  // diagonal force of diabat s on g: principal-diabat force (this rank's share) + the hop deltas along the chain (owned ones)
  for (int c = 0; c < 3; c++) X[c] = w * e.dF[3 * g + c];
  for (int q = s; q > 0; q = e.parent[q])
    if (state_owned(q, d.rank, d.world))
      for (int c = 0; c < 3; c++) X[c] = fma(w, e.dF[(size_t)q * n3 + 3 * g + c], X[c]);
  if (state_owned(s, d.rank, d.world))
    for (int c = 0; c < 3; c++) X[c] = fma(wc, e.Foff[(size_t)s * n3 + 3 * g + c], X[c]);
  if (recip_algebra && d.rank == 0) {
    // reciprocal-space force of diabat s on g: -K kk q_g(s) [ G_g + sum_b dq_b(s) N_gb ]
    const int mol = d.mol_of_atom[g];
    const int sg = r.mol_slot[mol] * MA + (g - d.mol_first[mol]);
    const size_t plane = (size_t)RA_SLOTS * RA_SLOTS;
    double f0 = r.G[3 * sg], f1 = r.G[3 * sg + 1], f2 = r.G[3 * sg + 2];
    const int n = r.st_n[s];
    for (int j = 0; j < n; j++) {
      const size_t o = (size_t)sg * RA_SLOTS + r.st_slot[s * RA_ENT + j];
      const double dqb = r.st_dq[s * RA_ENT + j];
      f0 = fma(dqb, r.Nx[o], f0); f1 = fma(dqb, r.Nx[plane + o], f1); f2 = fma(dqb, r.Nx[2 * plane + o], f2);
    }
    const double Kd = (double)d.K, qs = w * Sn.m[k].q[a];
    X[0] += -(Kd * d.kk[0]) * (qs * f0); X[1] += -(Kd * d.kk[1]) * (qs * f1); X[2] += -(Kd * d.kk[2]) * (qs * f2);
  }
  for (int c = 0; c < 3; c++) { atomicAdd(&out[3 * g + c], -X[c]); atomicAdd(&out[3 * t + c], X[c]); }
  }
}


// gather of the mixed grid for the atoms [i0, i1)
__global__ void k_evb_gather_range(Dev d, const double* __restrict__ theta, double* __restrict__ out, int i0, int i1) {
  int w = i0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= i1) return;
  double u[3] = {d.uscale[3 * w], d.uscale[3 * w + 1], d.uscale[3 * w + 2]};
  double F[3];
  gather_atom_warp(d, theta, u, d.xq[w].w, lane, F);
  if (lane < 3) {
    double v = lane == 0 ? F[0] : (lane == 1 ? F[1] : F[2]);
    out[3 * w + lane] += v;
  }
}

__global__ void k_copy(double* dst, const double* src, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// zero every accumulator the build adds into.  mode 0 (ahead of the enumeration): item energies, Vex, and the per-diabat
// force deltas / coupling forces of the diabats [0, plan.n_clear) -- the previous step's count plus a margin;
// mode 1 (behind the enumeration): the diabats [plan.n_clear, S) of a step that gained more than the margin
__global__ void k_evb_clear(Dev d, EvbDev e, int mode) {
  const int s_begin = mode ? e.plan->n_clear : 0, s_end = mode ? *e.n_states : e.plan->n_clear;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (size_t k = tid; k < RPB_MAX_ITEMS + 1; k += nth) e.item_energy[k] = 0.0;
    for (size_t k = tid; k < MAXS; k += nth) e.vex[k] = 0.0;
  }
  if (s_end <= s_begin) return;
  const size_t n3 = (size_t)3 * d.N, nbig = (size_t)(s_end - s_begin) * n3;
  double* dF = e.dF + (size_t)s_begin * n3; double* Fo = e.Foff + (size_t)s_begin * n3;
  for (size_t k = tid; k < nbig; k += nth) { dF[k] = 0.0; Fo[k] = 0.0; }
}

// ================================================================================================
// host orchestration: ENQUEUE ONLY.  Nothing below waits for the device inside a step; evb_readback (end of rpb_step /
// rpb_force_energy) is the one synchronising read of the last step's results.
// ================================================================================================
#define ENUM_BLOCK_INTS (16 + MAXS * (2 + MAXC * 5))
// result[8 ints] | e_ground | evec[MAXS] | h_full[2 MAXS] | err_flag[4] as doubles | en[E_NSLOT]
#define SOLVER_BLOCK_DOUBLES (5 + 3 * MAXS + 4 + E_NSLOT)

struct EvbScratch {   // device scratch owned by the context (allocated in evb_alloc, freed with it)
  CouplingGeo* geo;
  double* coeff_dev;    // [MAXS]
  CommitInfo* commit;
  int* chain_slot; int* cand; int* cand_n;
  double4* xq2; double* vel2; double* force2; double* mass2; int* type2; int* moa2;
  RecipDev rd; double* gtab;
  int hop_count_seen = 0;
};

#define CKE(call)                                                                 \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
      return RPB_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

static inline EvbScratch& scratch(rpb_ctx* c) { return *static_cast<EvbScratch*>(c->evb_scratch); }

int evb_alloc(rpb_ctx* c) {
  if (c->e.n_states) return 0;
  EvbDev& e = c->e;
  const int N = c->d.N;
  const size_t K3 = (size_t)c->d.K * c->d.K * c->d.K;
  int rc;
#define AL(p, n) if ((rc = dev_alloc(c, &(p), (size_t)(n)))) return rc;
  {  // enumeration results: one contiguous block, one copy (layout == EvbHost::pinned)
    int* blk;
    AL(blk, ENUM_BLOCK_INTS);
    e.n_states = blk; e.n_hops = blk + 16; e.parent = blk + 16 + MAXS; e.proton_log = blk + 16 + 2 * MAXS;
  }
  AL(e.snap, MAXS * NLEV); AL(e.item_energy, RPB_MAX_ITEMS + 1); AL(e.plan, 1); AL(e.mol_slot, c->d.M);
  AL(e.dF, (size_t)MAXS * 3 * N); AL(e.Foff, (size_t)MAXS * 3 * N);
  AL(e.vex, MAXS); AL(e.e_recip, 4); AL(e.h_diag, 3 * MAXS + E_NSLOT); AL(e.f_mix, 3 * N); AL(e.coef2, 3 * MAXS);
  {  // solver results: one contiguous block, one copy (SOLVER_BLOCK_DOUBLES)
    double* blk;
    AL(blk, SOLVER_BLOCK_DOUBLES);
    e.result = (int*)blk; e.e_ground = blk + 4; e.evec = blk + 5; e.h_full = blk + 5 + MAXS; e.status_copy = blk + 5 + 3 * MAXS;
  }
  AL(e.jac_v, MAXS * MAXS); AL(e.jac_sig, 2 + MAXS * MAXC * 5); AL(e.tree_mu, 2);
  EvbScratch* sp = new EvbScratch();
  c->evb_scratch = sp;
  EvbScratch& s = *sp;
  AL(s.geo, MAXS); AL(s.coeff_dev, MAXS); AL(s.commit, 1);
  AL(s.chain_slot, N); AL(s.cand, (size_t)CAND_SLOTS * CAND_CAP); AL(s.cand_n, CAND_SLOTS + 2);
  AL(s.xq2, N); AL(s.vel2, 3 * N); AL(s.force2, 3 * N); AL(s.mass2, N); AL(s.type2, N); AL(s.moa2, N);
  {
    RecipDev& r = s.rd;
    r.plan = e.plan; r.mol_slot = e.mol_slot;
    AL(r.P, RA_SLOTS); AL(r.G, 3 * RA_SLOTS); AL(r.Mx, (size_t)RA_SLOTS * RA_SLOTS); AL(r.Nx, (size_t)3 * RA_SLOTS * RA_SLOTS);
    AL(r.D, RA_SLOTS); AL(r.st_n, MAXS); AL(r.st_slot, MAXS * RA_ENT); AL(r.st_dq, MAXS * RA_ENT);
    AL(r.sl_n, RA_SLOTS); AL(r.sl_state, RA_SLOTS * MAXS); AL(r.sl_dq, RA_SLOTS * MAXS);
    AL(s.gtab, K3); r.gtab = s.gtab;
    AL(e.rcp_dE, MAXS);
    CKE(cudaMemset(e.rcp_dE, 0, MAXS * sizeof(double)));
  }
#undef AL
  // cuFFT path only (grid sizes with factors other than 2 and 3): the two plans a step needs (no plan is built inside a step)
  const int own_fft = fft_conv_supported(c);
  if (own_fft < 0) return own_fft;
  if (!own_fft) { cufftHandle pf, pi; if ((rc = pme_get_plans(c, 1, &pf, &pi))) return rc; }
  {   // Green function of the reciprocal-space convolution, g = IDFT(CB): the convolution of a unit charge at the origin
    if (!c->have_tables) { c->err = "rpb_set_tables must precede rpb_set_evb"; return RPB_ERR_STATE; }
    const double one = 1.0;
    CKE(cudaMemsetAsync(c->d.Q + K3, 0, K3 * sizeof(double), c->stream));
    CKE(cudaMemcpyAsync(c->d.Q + K3, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_convolve(c, 1, 1, e.e_recip, true))) return rc;
    CKE(cudaMemcpyAsync(s.gtab, c->d.theta + K3, K3 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CKE(cudaStreamSynchronize(c->stream));
  }
  CKE(cudaMallocHost(&c->eh.pinned, ENUM_BLOCK_INTS * sizeof(int) + (SOLVER_BLOCK_DOUBLES + 8) * sizeof(double) + sizeof(CommitInfo)));
  CKE(cudaMemset(e.n_states, 0, ENUM_BLOCK_INTS * sizeof(int)));
  CKE(cudaMemset(e.jac_sig, 0xff, (2 + MAXS * MAXC * 5) * sizeof(int)));
  CKE(cudaMemset(e.tree_mu, 0, 2 * sizeof(double)));
  CKE(cudaMemset(e.mol_slot, 0xff, c->d.M * sizeof(int)));
  CKE(cudaMemset(s.commit, 0, sizeof(CommitInfo)));
  {
    EvbPlan p0;
    memset(&p0, 0, sizeof(p0));
    p0.n_clear = MAXS;                  // the first step clears every diabat's accumulators
    CKE(cudaMemcpy(e.plan, &p0, sizeof(p0), cudaMemcpyHostToDevice));
  }
  c->d.commit_hop = &e.plan->hop;
  CKE(cudaFuncSetAttribute(k_evb_enumerate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENUM_SMEM_BYTES));
  CKE(cudaFuncSetAttribute(k_evb_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(((size_t)3 * MAXS * MAXS + 2 * MAXS) * sizeof(double))));
  { const char* sv = getenv("RPB_EVB_SOLVER"); c->evb_solver = (sv && std::string(sv) == "jacobi") ? 1 : 0; }
  return 0;
}

void evb_free(rpb_ctx* c) {
  if (!c->evb_scratch) return;
  delete static_cast<EvbScratch*>(c->evb_scratch);
  c->evb_scratch = nullptr;
}

// grid bounds from a recent diabat count (performance only: every kernel loops over the device-side counts)
static inline int s_bound(rpb_ctx* c) { return c->evb_s_bound_fixed ? c->evb_s_bound_fixed : std::min(MAXS, std::max(c->eh.s_hint, 8) + 16); }

int evb_enumerate_async(rpb_ctx* c, int part) {
  Dev& d = c->d; EvbDev& e = c->e;
  EvbScratch& sc = scratch(c);
  if (part == 0) {   // the kernel alone: the caller queues the pair kernel on the main stream before the rest
    ScopedTimer t(c, T_EVB_ENUM);
    k_evb_enumerate<<<1, ENUM_TPB, ENUM_SMEM_BYTES, c->stream>>>(d, e, sc.cand_n);
    CKE(cudaEventRecord(c->ev_sync[20], c->stream));            // "plan ready, candidate counters zeroed"
    c->n_launch += 1;
    return 0;
  }
  // diabat images for however many diabats the enumeration found (grid sized for evb_max_states, surplus warps exit)
  { ScopedTimer t(c, T_EVB_SNAP); k_evb_snapshots<<<(MAXS + SNAP_WPB - 1) / SNAP_WPB, 32 * SNAP_WPB, 0, c->stream>>>(d, e, -1, 1); }   // every rank needs the charges of every diabat
  c->n_launch += 1;
  CKE(cudaEventRecord(c->ev_sync[18], c->stream));            // "images ready"
  CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[11], 0));     // the early clears (aux[1])
  k_evb_clear<<<32, 256, 0, c->stream>>>(d, e, 1);            // the diabats beyond the early clears' margin (normally none)
  CKE(cudaEventRecord(c->ev_sync[14], c->stream));            // "images ready, every accumulator cleared"
  c->n_launch += 1;
  {
    // geometry factors of the couplings need the images and the clears (they add the Vex terms of the other chain
    // molecules); on aux[3] so that aux[0] is free for the Vex kernel's wait
    StreamScope ss(c, c->aux[3]);
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[14], 0));
    {
      ScopedTimer t(c, T_EVB_COUPLING_GEO);
      const int sb = s_bound(c);
      k_evb_coupling_geo<<<(sb + GEO_WPB - 1) / GEO_WPB, 32 * GEO_WPB, 0, c->stream>>>(d, e, sc.geo);
    }
    CKE(cudaEventRecord(c->ev_sync[19], c->stream));          // "clears complete, coupling geometry ready"
    c->n_launch += 1;
  }
  return 0;
}

// Accumulators of the build for the diabats [0, previous S + margin): ahead of the enumeration, on a stream with slack
void evb_clear_early(rpb_ctx* c) {
  k_evb_clear<<<148 * 2, 256, 0, c->stream>>>(c->d, c->e, 0);
  c->n_launch += 1;
}

int evb_build(rpb_ctx* c) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = scratch(c);
  const int N = d.N;
  const size_t K3 = (size_t)d.K * d.K * d.K, n3 = (size_t)3 * N;
  int rc = calculate_total_force_energy(c, true);   // principal diabat (+ enumeration, images, coupling geometry)
  if (rc) return rc;
  {
    // the principal grid is the only one convolved before the solver -- queued right behind the spreading; its copy in
    // slot 1 later receives the Hellmann-Feynman averaged charge deltas (evb_mix)
    StreamScope ss(c, c->aux[1]);
    k_copy<<<(unsigned)((K3 + 255) / 256), 256, 0, c->stream>>>(d.Q + K3, d.Q, K3);
    c->n_launch += 1;
    rc = launch_convolve(c, 0, 1, e.e_recip, true);
    if (rc) return rc;
    k_copy<<<1, 32, 0, c->stream>>>(d.en + E_RECIP, e.e_recip, 1);   // E_rec of the principal diabat (pme.f90:127)
    c->n_launch += 1;
  }
  const int sb = s_bound(c);
  // Branches from here (joined before the Hamiltonian is assembled):
  //   aux[4] : candidate lists -> real-space / repulsion / bonded deltas of every (diabat, last hop, topology)
  //   aux[0] : [enumeration, images] -> Vex of the couplings
  //   aux[1] : [principal grid spread + convolution] -> reciprocal-space algebra (P_a, G_a; energies)
  //   aux[2] : [bonded terms of the principal diabat] -> pair matrix of the chain atoms
  {
    StreamScope ss(c, c->aux[4]);
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[20], 0));      // the candidate lists need the plan only: next to the images
    {
      ScopedTimer t(c, T_EVB_CAND);
      dim3 g((N + 255) / 256, std::min(CAND_SLOTS, 4 * sb + 8));
      k_evb_candidates<<<g, 256, 0, c->stream>>>(d, e, c->evb_rcand * c->evb_rcand, sc.chain_slot, sc.cand, sc.cand_n);
    }
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[14], 0));      // images, clears -- not the coupling geometry
    { ScopedTimer t(c, T_EVB_ITEMS_BG); k_evb_items<<<dim3(2 * sb - 1, ITEM_SPLIT + 1), ITEM_TPB, 0, c->stream>>>(d, e, sc.chain_slot, sc.cand, sc.cand_n, c->evb_rcand, c->evb_rep_reach); }
    c->n_launch += 2;
  }
  {
    {   // the pair matrix needs the plan and the scaled coordinates only, not theta_1
      StreamScope ss(c, c->aux[2]);
      CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[18], 0));    // enumeration (plan)
      CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[15], 0));    // scaled coordinates
      ScopedTimer t(c, T_EVB_CORR);
      k_evb_rcp_pairs<<<std::min(4 * sb * MA * MA / 4 + 1, 148 * 8), 128, 0, c->stream>>>(d, sc.rd);
      CKE(cudaEventRecord(c->ev_sync[16], c->stream));
    }
    StreamScope ss(c, c->aux[1]);
    ScopedTimer t(c, T_EVB_CORR);
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[18], 0));      // enumeration (plan), images
    k_evb_rcp_atoms<<<((sb + 1) * MA * 32 + 255) / 256, 256, 0, c->stream>>>(d, sc.rd);
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[16], 0));
    k_evb_rcp_energy<<<(sb + 3) / 4, 128, 0, c->stream>>>(d, e, sc.rd);
    c->n_launch += 3;
  }
  {
    StreamScope ss(c, c->aux[0]);
    CKE(cudaStreamWaitEvent(c->stream, c->ev_sync[19], 0));       // coupling geometry (aux[3])
    ScopedTimer t(c, T_EVB_COUPLING);
    const int own_bound = (sb + d.world - 1) / d.world + 1;
    dim3 g(own_bound, (N + 256 * VEX_APT - 1) / (256 * VEX_APT));
    k_evb_coupling_vex<<<g, 256, 0, c->stream>>>(d, e, sc.geo);
    c->n_launch += 1;
  }
  // Single rank, tree solver: the ground state needs only energies RELATIVE to H_11, so the solver (and the averaged-grid
  // convolution behind it) must not wait for the principal diabat's pair forces -- the longest kernel of the step on large
  // boxes.  The branches join on aux[0], where evb_mix runs the solver; the main stream (pair forces) is joined only by
  // k_evb_finalize_principal and the force mixing.
  c->evb_overlap_solver = (d.world == 1 && c->evb_solver == 0);
  if (c->evb_overlap_solver) {
    stream_depend(c, 7, c->aux[1], c->aux[0]);
    stream_depend(c, 13, c->aux[4], c->aux[0]);
  } else {
    stream_depend(c, 6, c->aux[0], c->main_stream);
    stream_depend(c, 7, c->aux[1], c->main_stream);
    stream_depend(c, 13, c->aux[4], c->main_stream);
  }
  if (d.rank == 0) stream_depend(c, 9, c->aux[2], c->main_stream);   // bonded terms of the principal diabat (+ pair matrix)
  else stream_depend(c, 9, c->aux[2], c->evb_overlap_solver ? c->aux[0] : c->main_stream);
  if ((d.world > 1 && !c->evb_h_exchange_in_solver) || c->evb_solver != 0) {
    ScopedTimer t(c, T_EVB_ASSEMBLE);
    k_evb_assemble<<<(MAXS + 31) / 32, 32, 0, c->stream>>>(d, e, sc.geo);
    c->n_launch += 1;
    c->evb_assemble_pending = false;
  } else c->evb_assemble_pending = true;   // single rank, tree solver: assembled in the solver's prologue
  // keep the principal-diabat force (incl. EVB repulsion, without reciprocal part) in dF slot 0: d.force is
  // overwritten with the adiabatic force
  // (saved to dF slot 0 by k_evb_mix_forces itself)
  h.built = true;
  return 0;
}

int evb_mix(rpb_ctx* c, const double* coeff_override_host, double* force_out_host) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = scratch(c);
  if (!h.built) { c->err = "evb_mix before evb_build"; return RPB_ERR_STATE; }
  const int N = d.N;
  const size_t K3 = (size_t)d.K * d.K * d.K, n3 = (size_t)3 * N;
  const double* coeff_dev = nullptr;
  if (coeff_override_host) {
    CKE(cudaMemcpyAsync(sc.coeff_dev, coeff_override_host, MAXS * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    coeff_dev = sc.coeff_dev;
  }
  const bool fuse = (c->evb_solver == 0 && d.world == 1 && !coeff_override_host && c->evb_assemble_pending);
  const bool fuse_peer = (c->evb_solver == 0 && d.world > 1 && c->evb_h_exchange_in_solver && !coeff_override_host && c->evb_assemble_pending);
  const bool overlap = fuse && c->evb_overlap_solver;
  PeerArgsH no_px;
  memset(&no_px, 0, sizeof(no_px));
  const int sb = s_bound(c);
  if (overlap) {
    {
      StreamScope ss(c, c->aux[0]);
      ScopedTimer t(c, T_EVB_DIAG);
      k_evb_tree_solver<<<1, TREE_TPB, 0, c->stream>>>(d, e, nullptr, sc.geo, 1, no_px);
      CKE(cudaEventRecord(c->ev_sync[6], c->stream));               // "ground state known"
    }
    // H_11, absolute energies and the status block: behind the solver and behind what the main stream holds so far (pair
    // forces; bonded terms were joined by evb_build) -- on aux[3], not in the chain of the force mixing
    CKE(cudaEventRecord(c->ev_sync[0], c->main_stream));
    CKE(cudaStreamWaitEvent(c->aux[3], c->ev_sync[6], 0));
    CKE(cudaStreamWaitEvent(c->aux[3], c->ev_sync[0], 0));
    k_evb_finalize_principal<<<1, MAXS, 0, c->aux[3]>>>(d, e);
    CKE(cudaStreamWaitEvent(c->main_stream, c->ev_sync[6], 0));     // main: [pair forces, bonded terms] + ground state -> mixing
    c->n_launch += 2;
    c->evb_assemble_pending = false;
  } else {
    ScopedTimer t(c, T_EVB_DIAG);
    if (c->evb_solver == 0 && fuse_peer) {
      // sharded: assembly of this rank's partial block into its arena + the exchange + the solver, one kernel
      PeerArgsH px;
      const EvbDev e_part = c->e;                  // (h_diag = this step's arena parity; peer_args_h redirects c->e to the sum)
      peer_args_h(c, &px.a);
      px.on = 1;
      k_evb_tree_solver<<<1, TREE_TPB, 0, c->stream>>>(d, e_part, nullptr, sc.geo, 0, px);
      c->evb_assemble_pending = false;
    } else if (c->evb_solver == 0) {
      k_evb_tree_solver<<<1, TREE_TPB, 0, c->stream>>>(d, e, coeff_dev, fuse ? sc.geo : nullptr, 0, no_px);
      c->evb_assemble_pending = false;
    } else {
      const size_t shmem = ((size_t)3 * MAXS * MAXS + 2 * MAXS) * sizeof(double);
      k_evb_jacobi<<<1, JAC_TPB, shmem, c->stream>>>(d, e, coeff_dev);
    }
    c->n_launch++;
    CKE(cudaEventRecord(c->ev_sync[6], c->stream));
  }
  {
    const int include_principal = 1;      // every rank holds its share of the principal-diabat force (pair forces are sharded by clusters)
    const int in_place = coeff_override_host ? 0 : (d.world == 1 ? 1 : 2);
    double* out = in_place == 1 ? d.force : e.f_mix;
    CKE(cudaStreamWaitEvent(c->aux[1], c->ev_sync[6], 0));   // the averaged grid needs the ground state, not the pair forces
    // aux[1]: averaged charge deltas -> patch the copy of the principal grid -> ONE convolution; main: force mixing and the
    // chain atoms' own reciprocal terms; then the mixed grid is gathered once (sharded runs: this rank's slice of atoms)
    {
      StreamScope ss(c, c->aux[1]);
      if (coeff_override_host) { k_copy<<<(unsigned)((K3 + 255) / 256), 256, 0, c->stream>>>(d.Q + K3, d.Q, K3); c->n_launch++; }   // debug re-mix: fresh copy
      {
        ScopedTimer t(c, T_EVB_PATCH);
        k_evb_rcp_patch<<<((sb + 1) * MA * 32 + 255) / 256, 256, 0, c->stream>>>(d, e, sc.rd, d.Q + K3);
        c->n_launch += 1;
      }
      int rc2 = launch_convolve(c, 1, 1, e.e_recip, true);
      if (rc2) return rc2;
    }
    {
      ScopedTimer t(c, T_EVB_MIXF);
      k_evb_mix_forces<<<(unsigned)((n3 + 255) / 256), 256, 0, c->stream>>>(d, e, include_principal, in_place);
      if (d.rank == 0) { k_evb_rcp_mix<<<(sb * RA_ENT + 127) / 128, 128, 0, c->stream>>>(d, e, sc.rd, out); c->n_launch++; }
      if (c->evb_any_multi_basic && c->evb_quirk_types_present) { k_evb_reorder_quirk<<<(sb + 3) / 4, 128, 0, c->stream>>>(d, e, sc.rd, out, 1); c->n_launch++; }
    }
    stream_depend(c, 1, c->aux[1], c->main_stream);
    const int i0 = (int)((long long)N * d.rank / d.world), i1 = (int)((long long)N * (d.rank + 1) / d.world);
    { ScopedTimer t(c, T_EVB_GATHERMIX); k_evb_gather_range<<<((i1 - i0) * 32 + 255) / 256, 256, 0, c->stream>>>(d, d.theta + K3, out, i0, i1); }
    c->n_launch += 2;
  }
  if (coeff_override_host) {
    CKE(cudaStreamSynchronize(c->stream));
    CKE(cudaMemcpy(force_out_host, e.f_mix, n3 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return 0;
}

// Hop commit (evb_change_diabat_data_structure_topology, ms_evb.f90:806-834, + the forced list rebuild of :223-225):
// device kernels that act only when the solver selected a new hydronium molecule.  Also queues the read-back of the
// step's results (enumeration block, solver block, commit counter) into pinned memory -- nobody waits for it here.
int evb_commit(rpb_ctx* c) {
  Dev& d = c->d; EvbDev& e = c->e; EvbHost& h = c->eh;
  EvbScratch& sc = scratch(c);
  const int N = d.N;
  if (d.world > 1 && !c->peer.f_reduced_in_place) { k_copy<<<(3 * N + 255) / 256, 256, 0, c->stream>>>(d.force, e.f_mix, (size_t)3 * N); c->n_launch++; }   // single rank: mixed in place; peer exchange: reduced into d.force
  c->peer.f_reduced_in_place = false;
  CommitArgs ca;
  ca.e = e; ca.ci = sc.commit; ca.xq2 = sc.xq2; ca.vel2 = sc.vel2; ca.force2 = sc.force2; ca.mass2 = sc.mass2; ca.type2 = sc.type2; ca.moa2 = sc.moa2;
  int rc = launch_commit_and_rebuild(c, &ca);   // permutation, retyping, construct_verlet_list + update_verlet_displacements(init): only after a hop
  if (rc) return rc;
  // read-back (aux[3], behind the main stream's commit kernels and its own finalize kernel)
  CKE(cudaEventRecord(c->ev_sync[17], c->stream));
  CKE(cudaStreamWaitEvent(c->aux[3], c->ev_sync[17], 0));
  int* pin = h.pinned;
  double* pd = (double*)(pin + ENUM_BLOCK_INTS + (ENUM_BLOCK_INTS & 1));
  CKE(cudaMemcpyAsync(pin, e.n_states, ENUM_BLOCK_INTS * sizeof(int), cudaMemcpyDeviceToHost, c->aux[3]));
  CKE(cudaMemcpyAsync(pd, e.result, SOLVER_BLOCK_DOUBLES * sizeof(double), cudaMemcpyDeviceToHost, c->aux[3]));
  CKE(cudaMemcpyAsync(pd + SOLVER_BLOCK_DOUBLES + 1, sc.commit, sizeof(CommitInfo), cudaMemcpyDeviceToHost, c->aux[3]));
  // the main stream joins these copies at the END of the step (evb_join_readback): the second half kick need not wait
  c->evb_join_pending = true;
  return 0;
}

void evb_join_readback(rpb_ctx* c) {
  if (!c->evb_join_pending) return;
  stream_depend(c, 2, c->aux[3], c->main_stream);
  c->evb_join_pending = false;
}

// The one synchronising read of a call: results of the LAST step (accessors), sticky error flags of all of them.
int evb_readback(rpb_ctx* c) {
  EvbHost& h = c->eh;
  EvbScratch& sc = scratch(c);
  CKE(cudaStreamSynchronize(c->main_stream));
  const int* pin = h.pinned;
  const double* pd = (const double*)(pin + ENUM_BLOCK_INTS + (ENUM_BLOCK_INTS & 1));
  const int* pres = (const int*)pd;
  h.n_states = pin[0];
  h.s_hint = std::max(1, h.n_states);
  memcpy(h.n_hops, pin + 16, MAXS * sizeof(int));
  memcpy(h.parent, pin + 16 + MAXS, MAXS * sizeof(int));
  memcpy(h.proton_log, pin + 16 + 2 * MAXS, MAXS * MAXC * 5 * sizeof(int));
  for (int k = 0; k < 4; k++) c->h_flags[k] = (int)pd[5 + 3 * MAXS + k];
  for (int k = 0; k < E_NSLOT; k++) c->h_en[k] = pd[5 + 3 * MAXS + 4 + k];
  if (c->h_flags[1] == 2) { c->err = "a molecule's atoms are farther apart than the box allows (r_cutoff + extent of three consecutive atoms >= L/2)"; return RPB_ERR_VERLET; }
  if (c->h_flags[1]) { c->err = "please increase size of verlet neighbor list"; return RPB_ERR_VERLET; }
  if (c->h_flags[2]) { c->err = "Found more diabat states than the current setting of evb_max_states"; return RPB_ERR_DIABATS; }
  if (c->h_flags[3] >= 30) { c->err = "peer-memory exchange: rank " + std::to_string(c->h_flags[3] - 30) + " did not arrive"; return RPB_ERR_CUDA; }
  if (c->h_flags[3] == 1) { c->err = "error in subroutine find_bonded_atom_hydrogen"; return RPB_ERR_STATE; }
  if (c->h_flags[3]) { c->err = "couldn't find index in subroutine 'get_index_atom_set' (code " + std::to_string(c->h_flags[3]) + ")"; return RPB_ERR_STATE; }
  if (pres[2]) { c->err = "too many iterations in jacobi"; return RPB_ERR_STATE; }
  static const bool dbg_jacobi = getenv("RPB_DEBUG_JACOBI") != nullptr;
  if (dbg_jacobi) fprintf(stderr, "[evb solver %s] S=%d %s=%d\n", c->evb_solver ? "jacobi" : "tree", h.n_states, c->evb_solver ? "sweeps" : "evaluations", pres[4]);
  const int S = h.n_states;
  h.principal_diabat = pres[0]; h.new_hydronium = pres[1];
  h.adiabatic_potential = pd[4];
  memcpy(h.evec, pd + 5, MAXS * sizeof(double));
  for (int i = 0; i < MAXS; i++) for (int j = 0; j < MAXS; j++) h.hamiltonian[i][j] = 0.0;
  for (int s = 0; s < S; s++) {
    h.hamiltonian[s][s] = pd[5 + MAXS + s];
    if (s > 0) h.hamiltonian[h.parent[s]][s] = pd[5 + 2 * MAXS + s];
  }
  // energies as the reference leaves them: potential = adiabatic energy, components = principal diabat's
  {
    rpb_energies& en = c->last_en;
    const double* s = c->h_en;
    en.E_recip = s[E_RECIP];
    en.E_elec = s[E_ELEC] + s[E_RECIP] + c->cfg.ewald_self;
    en.E_vdw = s[E_VDW]; en.E_bond = s[E_BOND]; en.E_angle = s[E_ANGLE]; en.E_dihedral = s[E_DIH];
    en.potential_energy = h.adiabatic_potential;
  }
  // committed hops permuted the per-atom / per-molecule tables: the host mirrors and the upload cache are stale
  CommitInfo ci;
  memcpy(&ci, pd + SOLVER_BLOCK_DOUBLES + 1, sizeof(ci));
  if (ci.hop_count != sc.hop_count_seen) { sc.hop_count_seen = ci.hop_count; c->state_cache_valid = false; c->mirror_stale = true; }
  return 0;
}
