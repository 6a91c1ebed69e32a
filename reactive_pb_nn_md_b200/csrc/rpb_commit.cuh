// K15: hop commit on the device -- when the solver selected a new hydronium molecule (plan.hop), permute the per-atom
// arrays exactly as shift_array_data_donor_acceptor_transfer (ms_evb.f90:2677-2840), hop by hop of the new principal
// diabat, and retype / re-order the chain molecules from its final-level snapshot (ms_evb.f90:806-1006).  The phases are
// device functions run by the cooperative commit-and-rebuild kernel (kernels_nlist.cu), which exits at once when no hop was
// selected: the fixed per-step launch list pays ONE idle launch for the whole commit + forced list rebuild.
#pragma once
#include "rpb_evb.cuh"

struct CommitInfo {
  int n_hops;
  int from_g[RPB_MAXC], to_g[RPB_MAXC];     // global atom index the proton leaves / arrives at, in the index space BEFORE that hop
  int m_from[RPB_MAXC], m_to[RPB_MAXC];
  int n_mol; int mol[RPB_CHAIN_MOLS]; int new_first[RPB_CHAIN_MOLS]; int n_atom[RPB_CHAIN_MOLS];   // chain molecules of the new principal diabat after all hops
  int hop_count;                    // committed hops since rpb_set_evb (the host notices permuted tables through it)
};

struct CommitArgs {                 // passed by value to the commit-and-rebuild kernel (all null without MS-EVB)
  EvbDev e;
  CommitInfo* ci;
  double4* xq2; double* vel2; double* force2; double* mass2; int* type2; int* moa2;
};

#ifdef __CUDACC__
#define NLEV_C (RPB_MAXC + 1)
// one thread: the hop parameters, replayed on the few molecules involved
__device__ inline void commit_prepare(const Dev& d, const EvbDev& e, CommitInfo* ci) {
  const int pdiab = e.result[0], nh = e.n_hops[pdiab];
  const int* L = &e.proton_log[pdiab * RPB_MAXC * 5];
  // first atom / atom count of molecule m after the hops applied so far
  auto first_now = [&](int m, int upto) {
    int f = d.mol_first[m];
    for (int h = 0; h < upto; h++) {
      if (ci->m_from[h] < ci->m_to[h]) { if (m > ci->m_from[h] && m <= ci->m_to[h]) f -= 1; }
      else { if (m > ci->m_to[h] && m <= ci->m_from[h]) f += 1; }
    }
    return f;
  };
  auto natom_now = [&](int m, int upto) {
    int n = d.mol_natom[m];
    for (int h = 0; h < upto; h++) { if (ci->m_from[h] == m) n -= 1; if (ci->m_to[h] == m) n += 1; }
    return n;
  };
  int ima = *d.hydronium;
  for (int k = 0; k < nh; k++) {
    const int imd = ima, a_from = L[k * 5 + 1];
    ima = L[k * 5 + 3];
    ci->m_from[k] = imd; ci->m_to[k] = ima;
    const int a_to = natom_now(ima, k);
    ci->from_g[k] = first_now(imd, k) + a_from;
    ci->to_g[k] = (imd < ima) ? first_now(ima, k) + a_to - 1 : first_now(ima, k) + a_to;
  }
  ci->n_hops = nh;
  const Snapshot& S = e.snap[pdiab * NLEV_C + nh];
  ci->n_mol = S.n_mol;
  for (int k = 0; k < S.n_mol; k++) { ci->mol[k] = S.m[k].mol; ci->new_first[k] = first_now(S.m[k].mol, nh); ci->n_atom[k] = S.m[k].n_atom; }
}

// index i < N: source of new position i (hops undone last to first; chain molecules in snapshot order); i < M: the
// molecule table and the atom -> molecule map after the hops
__device__ inline void commit_permute(const Dev& d, const CommitArgs& a, int i) {
  const EvbDev& e = a.e;
  const CommitInfo* ci = a.ci;
  const int nh = ci->n_hops;
  if (i < d.N) {
    int src = i;
    for (int h = nh - 1; h >= 0; h--) {
      const int fg = ci->from_g[h], tg = ci->to_g[h];
      if (src == tg) src = fg;
      else if (fg < tg) { if (src >= fg && src < tg) src += 1; }
      else { if (src > tg && src <= fg) src -= 1; }
    }
    const Snapshot& S = e.snap[e.result[0] * NLEV_C + nh];
    for (int k = 0; k < ci->n_mol; k++)         // re-ordered acceptor: the snapshot lists the principal index of every position
      if (i >= ci->new_first[k] && i < ci->new_first[k] + ci->n_atom[k]) src = S.m[k].atom[i - ci->new_first[k]];
    a.xq2[i] = d.xq[src];
    for (int k = 0; k < 3; k++) { a.vel2[3 * i + k] = d.vel[3 * src + k]; a.force2[3 * i + k] = d.force[3 * src + k]; }
    a.mass2[i] = d.mass[src];
    a.type2[i] = d.type[src];
  }
  if (i < d.M) {
    int f = d.mol_first[i], n = d.mol_natom[i];
    for (int h = 0; h < nh; h++) {
      if (ci->m_from[h] < ci->m_to[h]) { if (i > ci->m_from[h] && i <= ci->m_to[h]) f -= 1; }
      else { if (i > ci->m_to[h] && i <= ci->m_from[h]) f += 1; }
      if (ci->m_from[h] == i) n -= 1;
      if (ci->m_to[h] == i) n += 1;
    }
    d.mol_first[i] = f; d.mol_natom[i] = n;     // (only this index reads or writes entry i)
    for (int q = 0; q < n; q++) a.moa2[f + q] = i;
  }
}

// copy back, then the snapshot data (positions made whole, charges, types, centres of mass, molecule types) of the chain
// molecules and the new hydronium index
__device__ inline void commit_finish(const Dev& d, const CommitArgs& a, int i) {
  if (i >= d.N) return;
  const EvbDev& e = a.e;
  CommitInfo* ci = a.ci;
  const Snapshot& S = e.snap[e.result[0] * NLEV_C + ci->n_hops];
  double4 x = a.xq2[i];
  int ty = a.type2[i];
  for (int k = 0; k < ci->n_mol; k++)
    if (i >= ci->new_first[k] && i < ci->new_first[k] + ci->n_atom[k]) {
      const MolImage& I = S.m[k];
      const int q = i - ci->new_first[k];
      x = make_double4(I.x[q][0], I.x[q][1], I.x[q][2], I.q[q]);
      ty = I.type[q];
      if (q == 0) {
        for (int c = 0; c < 3; c++) d.r_com[3 * I.mol + c] = I.r_com[c];
        d.mol_type[I.mol] = I.mtype;
      }
    }
  d.xq[i] = x; d.type[i] = ty;
  for (int k = 0; k < 3; k++) { d.vel[3 * i + k] = a.vel2[3 * i + k]; d.force[3 * i + k] = a.force2[3 * i + k]; }
  d.mass[i] = a.mass2[i];
  d.mol_of_atom[i] = a.moa2[i];
  if (i == 0) { *d.hydronium = S.m[S.hydronium].mol; ci->hop_count += 1; }
}
#endif
