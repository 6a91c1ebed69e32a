// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_core.h).  PARITY UNPINNED.
// MS-EVB Hamiltonian build: literal restatement of src/ms_evb.f90:181-3123, including the
// per-diabat deep copies and the physical array shifts the reference performs for every proton
// hop (ms_evb.f90:770-798, 843-932, 2677-2840).  Deliberately NOT optimised: this is the
// anchor the patch-based CUDA path is compared against.
#include "oracle_md.h"
#include <algorithm>

namespace orc {

namespace {

struct Diabat {  // atom_data_diabat / molecule_data_diabat / system_data_diabat copies (ms_evb.f90:770-798)
  AtomData atoms;
  std::vector<Molecule> mol;
  SystemData sys;
  int hydronium;
};

void create_diabat(const Ctx& c, Diabat& D) {
  D.atoms = c.atoms; D.mol = c.mol; D.sys = c.sys; D.hydronium = c.hydronium_mol;
}

struct Arr { double* d; int* i; int w; };

// shift_array_data_donor_acceptor_transfer ms_evb.f90:2677-2840 (a_from/a_to 0-based within molecule)
void shift_transfer(std::vector<Molecule>& mol, int m_from, int a_from, int m_to, int a_to, std::vector<Arr>& arrays) {
  int from_g = mol[m_from].first + a_from;
  int to_g;
  if (m_from < m_to) to_g = mol[m_to].first + a_to - 1;  // atom_index(i_atom_transfer_to-1)
  else to_g = mol[m_to].first + a_to;                     // both branches of :2706-2711 evaluate the same
  for (auto& A : arrays) {
    double sd[3] = {0, 0, 0}; int si = 0;
    if (A.d) for (int k = 0; k < A.w; k++) sd[k] = A.d[A.w * from_g + k]; else si = A.i[from_g];
    if (from_g < to_g) {
      for (int i = from_g; i < to_g; i++) {
        if (A.d) for (int k = 0; k < A.w; k++) A.d[A.w * i + k] = A.d[A.w * (i + 1) + k]; else A.i[i] = A.i[i + 1];
      }
    } else {
      for (int i = from_g; i > to_g; i--) {
        if (A.d) for (int k = 0; k < A.w; k++) A.d[A.w * i + k] = A.d[A.w * (i - 1) + k]; else A.i[i] = A.i[i - 1];
      }
    }
    if (A.d) for (int k = 0; k < A.w; k++) A.d[A.w * to_g + k] = sd[k]; else A.i[to_g] = si;
  }
  if (from_g < to_g) { for (int im = m_from + 1; im <= m_to; im++) mol[im].first -= 1; }
  else { for (int im = m_to + 1; im <= m_from; im++) mol[im].first += 1; }
  mol[m_to].n_atom += 1;
  mol[m_from].n_atom -= 1;
}

// reorder_molecule_data_structures ms_evb.f90:941-1006
int reorder_molecule(const Ctx& c, Diabat& D, int i_mole) {
  Molecule& m = D.mol[i_mole];
  const MoleculeType& T = c.mt[m.type];
  int f = m.first;
  for (int i = 0; i < T.n_atom; i++) {
    if (T.atom_type[i] != D.atoms.type[f + i]) {
      int index = -1;
      for (int j = i + 1; j < m.n_atom; j++) if (T.atom_type[i] == D.atoms.type[f + j]) { index = j; break; }
      if (index < 0) return RPB_ERR_STATE;
      auto rot3 = [&](std::vector<double>& a) {
        double s[3] = {a[3 * (f + index)], a[3 * (f + index) + 1], a[3 * (f + index) + 2]};
        for (int j = index - 1; j >= i; j--) for (int k = 0; k < 3; k++) a[3 * (f + j + 1) + k] = a[3 * (f + j) + k];
        for (int k = 0; k < 3; k++) a[3 * (f + i) + k] = s[k];
      };
      auto rot1 = [&](std::vector<double>& a) {
        double s = a[f + index];
        for (int j = index - 1; j >= i; j--) a[f + j + 1] = a[f + j];
        a[f + i] = s;
      };
      rot3(D.atoms.xyz); rot3(D.atoms.vel); rot3(D.atoms.force); rot1(D.atoms.charge); rot1(D.atoms.mass);
      int st = D.atoms.type[f + index];
      for (int j = index - 1; j >= i; j--) D.atoms.type[f + j + 1] = D.atoms.type[f + j];
      D.atoms.type[f + i] = st;
    }
  }
  return 0;
}

// evb_change_data_structures_proton_transfer ms_evb.f90:843-932 (all indices 0-based)
int change_topology_proton_transfer(const Ctx& c, Diabat& D, int i_mole_donor, int i_atom_donor, int i_mole_acceptor,
                                    int i_atom_acceptor, int i_heavy_acceptor) {
  D.hydronium = i_mole_acceptor;
  std::vector<Arr> arrays = {{D.atoms.xyz.data(), nullptr, 3}, {D.atoms.vel.data(), nullptr, 3},
                             {D.atoms.force.data(), nullptr, 3}, {D.atoms.mass.data(), nullptr, 1},
                             {D.atoms.charge.data(), nullptr, 1}, {nullptr, D.atoms.type.data(), 1}};
  shift_transfer(D.mol, i_mole_donor, i_atom_donor, i_mole_acceptor, i_atom_acceptor, arrays);
  Molecule& md = D.mol[i_mole_donor];
  Molecule& ma = D.mol[i_mole_acceptor];
  make_molecule_whole(ma.n_atom, &D.atoms.xyz[3 * ma.first], D.sys);
  pos_com(md.r_com, &D.atoms.xyz[3 * md.first], &D.atoms.mass[md.first], md.n_atom);
  pos_com(ma.r_com, &D.atoms.xyz[3 * ma.first], &D.atoms.mass[ma.first], ma.n_atom);
  // change_proton_index_proton_transfer :2992-3004
  D.atoms.type[ma.first + i_atom_acceptor] = c.proton_index[c.conj_pairs[ma.type]];
  for (int a = 0; a < ma.n_atom; a++) {
    int t = D.atoms.type[ma.first + a];
    int tn = (a != i_atom_acceptor) ? c.conj_atom[t] : t;
    D.atoms.type[ma.first + a] = tn;
    D.atoms.charge[ma.first + a] = c.atype_chg[tn];
  }
  D.atoms.type[ma.first + i_heavy_acceptor] = c.heavy_acid_index[c.conj_pairs[ma.type]];
  for (int a = 0; a < md.n_atom; a++) {
    int tn = c.conj_atom[D.atoms.type[md.first + a]];
    D.atoms.type[md.first + a] = tn;
    D.atoms.charge[md.first + a] = c.atype_chg[tn];
  }
  md.type = c.conj_pairs[md.type];
  ma.type = c.conj_pairs[ma.type];
  return reorder_molecule(c, D, i_mole_acceptor);
}

// get_index_atom_set general_routines.f90:613-637 (tables hold 0-based types, -1 = empty row)
template <int W>
int get_index_atom_set(const int (*lookup)[W], const int* itype) {
  for (int i = 0; i < MAXI; i++) {
    if (lookup[i][0] < 0) break;
    bool ok = true;
    for (int j = 0; j < W; j++) if (lookup[i][j] != itype[j]) ok = false;
    if (ok) return i;
  }
  return -1;
}

// get_heavy_atom_transfer_acid / _base ms_evb.f90:2888-2938
int heavy_atom_acid(const Ctx& c, int type_acid) {
  int th = c.heavy_acid_index[type_acid];
  for (int a = 0; a < c.mt[type_acid].n_atom; a++) if (c.mt[type_acid].atom_type[a] == th) return a;
  return -1;
}
int heavy_atom_base(const Ctx& c, int type_base) { return heavy_atom_acid(c, c.conj_pairs[type_base]); }

// ms_evb_repulsive_switch ms_evb.f90:2484-2504
void repulsive_switch(double* sw, double* dsw, double r, double rs, double rc) {
  *sw = 0.0; *dsw = 0.0;
  if (r < rc) {
    if (r < rs) *sw = 1.0;
    else {
      double term1 = (r - rs) * (r - rs) / ((rc - rs) * (rc - rs) * (rc - rs));
      double term2 = 3.0 * rc - rs - 2.0 * r;
      *sw = 1.0 - term1 * term2;
      *dsw = -2.0 * (r - rs) * term2 / ((rc - rs) * (rc - rs) * (rc - rs)) + 2.0 * term1;
    }
  }
}

// ms_evb_intermolecular_repulsion ms_evb.f90:2259-2478 (three-atom term, then Born-Mayer)
int intermolecular_repulsion(const Ctx& c, double* force, double* E_rep, const Diabat& D) {
  *E_rep = 0.0;
  const int ih = D.hydronium;
  const Molecule& H = D.mol[ih];
  const double* xi = &D.atoms.xyz[3 * H.first];
  const int* ti = &D.atoms.type[H.first];
  double* fi = &force[3 * H.first];
  // ---- ms_evb_three_atom_repulsion :2295-2399
  int i_type_H = ti[H.n_atom - 1];
  int i_heavy = heavy_atom_acid(c, H.type);
  if (i_heavy < 0) return RPB_ERR_STATE;
  int i_type_heavy = ti[i_heavy];
  for (int jm = 0; jm < D.sys.n_mole; jm++) {
    if (jm == ih) continue;
    const Molecule& J = D.mol[jm];
    for (int ja = 0; ja < J.n_atom; ja++) {
      int jt = D.atoms.type[J.first + ja];
      int key[3] = {jt, i_type_heavy, i_type_H};
      int idx = get_index_atom_set<3>(c.da_int, key);
      if (idx < 0) continue;
      const double* P = c.da_par[idx];
      double B = P[0], bl = P[1], d0 = P[2], blp = P[3], rs = P[4], rc = P[5];
      const double* xj = &D.atoms.xyz[3 * (J.first + ja)];
      double* fj = &force[3 * (J.first + ja)];
      double shift[3], t[3], rij_O[3];
      pbc_shift(shift, &xi[3 * i_heavy], xj, D.sys);
      pbc_dr(t, &xi[3 * i_heavy], xj, shift);
      for (int k = 0; k < 3; k++) rij_O[k] = -t[k];
      double r_OO = std::sqrt(rij_O[0] * rij_O[0] + rij_O[1] * rij_O[1] + rij_O[2] * rij_O[2]);
      double sw, dsw;
      repulsive_switch(&sw, &dsw, r_OO, rs, rc);
      double fac_OO = B * std::exp(-bl * (r_OO - d0));
      double sum = 0.0;
      for (int ia = 0; ia < H.n_atom; ia++) {
        if (ti[ia] != i_type_H) continue;
        double rij[3];
        pbc_dr(t, &xi[3 * ia], xj, shift);
        for (int k = 0; k < 3; k++) rij[k] = -t[k];
        double q[3];
        for (int k = 0; k < 3; k++) q[k] = (2.0 * xj[k] + rij_O[k]) / 2.0 - (xj[k] + rij[k]);
        double q2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
        double exp_q = std::exp(-blp * q2);
        sum = sum + exp_q;
        for (int k = 0; k < 3; k++) {
          fi[3 * ia + k] = fi[3 * ia + k] + sw * fac_OO * exp_q * -blp * 2.0 * q[k];
          fi[3 * i_heavy + k] = fi[3 * i_heavy + k] + sw * fac_OO * exp_q * blp * q[k];
          fj[k] = fj[k] + sw * fac_OO * exp_q * blp * q[k];
        }
      }
      *E_rep = *E_rep + sw * fac_OO * sum;
      for (int k = 0; k < 3; k++) {
        double fij = rij_O[k] / r_OO * fac_OO * sum * (sw * bl - dsw);
        fi[3 * i_heavy + k] = fi[3 * i_heavy + k] + fij;
        fj[k] = fj[k] - fij;
      }
    }
  }
  // ---- ms_evb_born_mayer :2405-2478
  for (int ia = 0; ia < H.n_atom; ia++) {
    for (int jm = 0; jm < D.sys.n_mole; jm++) {
      if (jm == ih) continue;
      const Molecule& J = D.mol[jm];
      for (int ja = 0; ja < J.n_atom; ja++) {
        int key[2] = {D.atoms.type[J.first + ja], ti[ia]};
        int idx = get_index_atom_set<2>(c.pa_int, key);
        if (idx < 0) continue;
        const double* P = c.pa_par[idx];
        double C = P[0], cl = P[1], d0 = P[2], rs = P[3], rc = P[4];
        const double* xj = &D.atoms.xyz[3 * (J.first + ja)];
        double shift[3], t[3], rij[3];
        pbc_shift(shift, &xi[3 * ia], xj, D.sys);
        pbc_dr(t, &xi[3 * ia], xj, shift);
        for (int k = 0; k < 3; k++) rij[k] = -t[k];
        double r_ij = std::sqrt(rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2]);
        double fac_OH = C * std::exp(-cl * (r_ij - d0));
        double sw, dsw;
        repulsive_switch(&sw, &dsw, r_ij, rs, rc);
        *E_rep = *E_rep + sw * fac_OH;
        for (int k = 0; k < 3; k++) {
          double fij = rij[k] / r_ij * fac_OH * (sw * cl - dsw);
          fi[3 * ia + k] = fi[3 * ia + k] + fij;
          force[3 * (J.first + ja) + k] = force[3 * (J.first + ja) + k] - fij;
        }
      }
    }
  }
  return 0;
}

// ms_evb_diabat_force_energy_update_intra ms_evb.f90:1900-1954
int diabat_update_intra(const Ctx& c, double* force_atoms, double* E_intra, int imd, int ima, const Diabat& D) {
  *E_intra = 0.0;
  const Molecule& md = D.mol[imd];
  const Molecule& ma = D.mol[ima];
  double El;
  typedef int (*fn)(const Ctx&, double*, const double*, const int*, double*, int);
  fn fns[3] = {intra_molecular_bond_energy_force, intra_molecular_angle_energy_force, intra_molecular_dihedral_energy_force};
  for (int t = 0; t < 3; t++) {
    if (fns[t](c, &El, &D.atoms.xyz[3 * md.first], &D.atoms.type[md.first], &force_atoms[3 * md.first], md.type)) return RPB_ERR_ARG;
    *E_intra = *E_intra + El;
    if (fns[t](c, &El, &D.atoms.xyz[3 * ma.first], &D.atoms.type[ma.first], &force_atoms[3 * ma.first], ma.type)) return RPB_ERR_ARG;
    *E_intra = *E_intra + El;
  }
  return 0;
}

// one donor/acceptor atom against all atoms (ms_evb.f90:1629-1732 / 1738-1841)
void diabat_atom_vs_all(const Ctx& c, double* force_atoms, double* dE, int i_atom, const std::vector<int>& screen,
                        const Diabat& D, PairList& cut, PairList& lj, PairList& sapt) {
  const double rc2 = c.cfg.real_space_cutoff * c.cfg.real_space_cutoff;
  const int N = D.sys.total_atoms;
  cut.clear(); lj.clear(); sapt.clear();
  int ti = D.atoms.type[i_atom];
  double qi = D.atoms.charge[i_atom];
  for (int j = 0; j < N; j++) {
    if (screen[j] != 1) continue;
    double dr[3];
    for (int k = 0; k < 3; k++) {
      double d = D.atoms.xyz[3 * i_atom + k] - D.atoms.xyz[3 * j + k];
      dr[k] = d - D.sys.box[k] * std::floor(d / D.sys.box[k] + 0.5);
    }
    double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
    if (dr2 < rc2) {
      int tj = D.atoms.type[j];
      double par[6];
      for (int k = 0; k < 6; k++) par[k] = c.vdw_param[vdw_idx(ti, tj, k)];
      double qq = qi * D.atoms.charge[j];
      cut.push(j, dr, dr2, qq, par);
      int vt = c.vdw_type[ti + MAXT * tj];
      if (vt == 0) lj.push(j, dr, dr2, qq, par); else if (vt == 1) sapt.push(j, dr, dr2, qq, par);
    }
  }
  double El = cut.idx.empty() ? 0.0 : pairwise_real_space_ewald(c, cut);
  double Elj = lj.idx.empty() ? 0.0 : pairwise_real_space_LJ(lj);
  double Es = sapt.idx.empty() ? 0.0 : pairwise_real_space_sapt(c, sapt);
  *dE = *dE + El + Elj + Es;
  auto add = [&](const PairList& p) {
    for (size_t k = 0; k < p.idx.size(); k++) {
      int j = p.idx[k];
      for (int d = 0; d < 3; d++) {
        force_atoms[3 * i_atom + d] = force_atoms[3 * i_atom + d] + p.f[3 * k + d];
        force_atoms[3 * j + d] = force_atoms[3 * j + d] - p.f[3 * k + d];
      }
    }
  };
  add(lj); add(sapt); add(cut);
}

// ms_evb_diabat_force_energy_update_real_space ms_evb.f90:1566-1894
void diabat_update_real_space(const Ctx& c, double* force_atoms, double* dE, int imd, int ima, const Diabat& D) {
  const int N = D.sys.total_atoms;
  const Molecule& md = D.mol[imd];
  const Molecule& ma = D.mol[ima];
  *dE = 0.0;
  PairList cut, lj, sapt;
  std::vector<int> screen(N, 1);
  for (int a = 0; a < md.n_atom; a++) screen[md.first + a] = 0;
  for (int a = 0; a < md.n_atom; a++) diabat_atom_vs_all(c, force_atoms, dE, md.first + a, screen, D, cut, lj, sapt);
  for (int a = 0; a < ma.n_atom; a++) screen[ma.first + a] = 0;
  for (int a = 0; a < ma.n_atom; a++) diabat_atom_vs_all(c, force_atoms, dE, ma.first + a, screen, D, cut, lj, sapt);
  double Ee = 0, Ev = 0;
  intra_molecular_pairwise_energy_force(c, &force_atoms[3 * md.first], &Ee, &Ev, &D.atoms.xyz[3 * md.first],
                                        &D.atoms.charge[md.first], &D.atoms.type[md.first], md.type, md.n_atom);
  *dE = *dE + Ee + Ev;
  Ee = 0; Ev = 0;
  intra_molecular_pairwise_energy_force(c, &force_atoms[3 * ma.first], &Ee, &Ev, &D.atoms.xyz[3 * ma.first],
                                        &D.atoms.charge[ma.first], &D.atoms.type[ma.first], ma.type, ma.n_atom);
  *dE = *dE + Ee + Ev;
}

void modify_Q_molecule(const Ctx& c, double* Q, const Diabat& D, int im, const double kk[3][3], int op) {
  const Molecule& m = D.mol[im];
  std::vector<double> us(3 * m.n_atom);
  create_scaled_direct_coordinates(us.data(), &D.atoms.xyz[3 * m.first], m.n_atom, kk, c.cfg.pme_grid);
  spread_atoms(c, Q, &D.atoms.charge[m.first], us.data(), m.n_atom, op);
}

// ms_evb_diabat_force_energy ms_evb.f90:1421-1559
int diabat_force_energy(Ctx& c, Diabat& D, int i_diabat, int i_mole_principle) {
  const int N = D.sys.total_atoms;
  std::vector<double> Q_local = c.Q_grid;
  std::vector<double> dF(3 * N);
  double kk[3][3];
  int ima = i_mole_principle;
  for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
    if (c.proton_log[i_diabat][ih][0] < 0) break;
    int imd = ima;
    int i_atom_donor = c.proton_log[i_diabat][ih][1];
    ima = c.proton_log[i_diabat][ih][3];
    int i_heavy_acceptor = c.proton_log[i_diabat][ih][4];
    int i_atom_acceptor = D.mol[ima].n_atom;  // n_atom+1 in 1-based
    std::fill(dF.begin(), dF.end(), 0.0);
    double dE_d_intra, dE_d_real, E_d_rep, E_ref_d, dE_a_intra, dE_a_real, E_a_rep, E_ref_a;
    int rc = diabat_update_intra(c, dF.data(), &dE_d_intra, imd, ima, D); if (rc) return rc;
    diabat_update_real_space(c, dF.data(), &dE_d_real, imd, ima, D);
    rc = intermolecular_repulsion(c, dF.data(), &E_d_rep, D); if (rc) return rc;
    E_ref_d = c.ref_energy[D.mol[imd].type];
    reciprocal_lattice(kk, D.sys);
    modify_Q_molecule(c, Q_local.data(), D, imd, kk, -1);
    modify_Q_molecule(c, Q_local.data(), D, ima, kk, -1);
    for (int k = 0; k < 3 * N; k++) D.atoms.force[k] = D.atoms.force[k] - dF[k];
    std::fill(dF.begin(), dF.end(), 0.0);
    rc = change_topology_proton_transfer(c, D, imd, i_atom_donor, ima, i_atom_acceptor, i_heavy_acceptor); if (rc) return rc;
    rc = diabat_update_intra(c, dF.data(), &dE_a_intra, imd, ima, D); if (rc) return rc;
    diabat_update_real_space(c, dF.data(), &dE_a_real, imd, ima, D);
    rc = intermolecular_repulsion(c, dF.data(), &E_a_rep, D); if (rc) return rc;
    E_ref_a = c.ref_energy[D.mol[ima].type];
    modify_Q_molecule(c, Q_local.data(), D, imd, kk, +1);
    modify_Q_molecule(c, Q_local.data(), D, ima, kk, +1);
    for (int k = 0; k < 3 * N; k++) D.atoms.force[k] = D.atoms.force[k] + dF[k];
    D.sys.potential_energy = D.sys.potential_energy + E_ref_a + dE_a_intra + dE_a_real + E_a_rep - E_ref_d -
                             dE_d_intra - dE_d_real - E_d_rep;
  }
  c.Q_grid_diabats[i_diabat] = Q_local;
  return 0;
}

// map_diabat_force_to_principle_recursive ms_evb.f90:2608-2656
void map_force_to_principle(const Ctx& c, int diabat, int i_hop, int i_mole_principle, std::vector<Molecule>& mol_local,
                            double* force) {
  if (i_hop >= c.cfg.evb_max_chain) return;
  if (c.proton_log[diabat][i_hop][0] < 0) return;
  int imd = i_mole_principle;
  int i_atom_donor = c.proton_log[diabat][i_hop][1];
  int ima = c.proton_log[diabat][i_hop][3];
  map_force_to_principle(c, diabat, i_hop + 1, ima, mol_local, force);
  int i_atom_acceptor = mol_local[ima].n_atom - 1;
  std::vector<Arr> arrays = {{force, nullptr, 3}};
  shift_transfer(mol_local, ima, i_atom_acceptor, imd, i_atom_donor, arrays);
}

// evb_store_forces ms_evb.f90:2523-2590.  The reference hands out store slots from a shared counter inside
// an OMP CRITICAL section (order = thread arrival); slots here are fixed (0 principal, 2s-1 diagonal of diabat s,
// 2s its coupling to the parent) so the diabat loop can run in parallel without the critical section.
void store_forces(Ctx& c, int i_mole_principle, int d1, int d2, const Diabat& D, int slot) {
  c.evb_forces_lookup_index[d1][d2] = slot;
  c.evb_forces_store[slot] = D.atoms.force;
  int diabat = std::max(d1, d2);
  if (diabat > 0) {
    std::vector<Molecule> mol_local = D.mol;
    map_force_to_principle(c, diabat, 0, i_mole_principle, mol_local, c.evb_forces_store[slot].data());
  }
}

// zundel_r_com ms_evb.f90:2946-2982
void zundel_r_com(double out[3], double shiftd[3], double shifta[3], int imd, int ima, const Diabat& D) {
  const Molecule& md = D.mol[imd];
  const Molecule& ma = D.mol[ima];
  double tmd = 0, tma = 0;
  for (int a = 0; a < md.n_atom; a++) tmd = tmd + D.atoms.mass[md.first + a];
  for (int a = 0; a < ma.n_atom; a++) tma = tma + D.atoms.mass[ma.first + a];
  double shift[3], rda[3], r_com_a[3];
  pbc_shift(shift, md.r_com, ma.r_com, D.sys);
  pbc_dr(rda, md.r_com, ma.r_com, shift);
  for (int k = 0; k < 3; k++) r_com_a[k] = md.r_com[k] + rda[k];
  for (int k = 0; k < 3; k++) out[k] = (tmd * md.r_com[k] + tma * r_com_a[k]) / (tmd + tma);
  for (int k = 0; k < 3; k++) { shiftd[k] = 0.0; shifta[k] = shift[k]; }
}

// evb_diabatic_coupling_electrostatics ms_evb.f90:1276-1403
void coupling_electrostatics(const Ctx& c, double* Vex, double* dVex, int imd, int ima, const Diabat& D) {
  const double conv = c.cfg.conv_e2A_kJmol;
  *Vex = 0.0;
  std::fill(dVex, dVex + 3 * D.sys.total_atoms, 0.0);
  const Molecule& md = D.mol[imd];
  const Molecule& ma = D.mol[ima];
  double rz[3], shiftd[3], shifta[3];
  zundel_r_com(rz, shiftd, shifta, imd, ima, D);
  double q_exchange_transfer = c.exch_proton[ma.type][md.type];
  for (int jm = 0; jm < D.sys.n_mole; jm++) {
    if (jm == imd || jm == ima) continue;
    const Molecule& J = D.mol[jm];
    double shift[3];
    pbc_shift(shift, rz, J.r_com, D.sys);
    for (int ja = 0; ja < J.n_atom; ja++) {
      double q_j = D.atoms.charge[J.first + ja];
      const double* xj = &D.atoms.xyz[3 * (J.first + ja)];
      double* fj = &dVex[3 * (J.first + ja)];
      for (int side = 0; side < 2; side++) {
        const Molecule& S = side == 0 ? md : ma;
        const double* sh = side == 0 ? shiftd : shifta;
        for (int ia = 0; ia < S.n_atom; ia++) {
          double q_i;
          if (side == 1 && ia == S.n_atom - 1) q_i = q_exchange_transfer;
          else q_i = c.exch_atomic[D.atoms.type[S.first + ia]];
          double dr[3], xs[3], t[3], r_ij[3];
          pbc_dr(dr, rz, &D.atoms.xyz[3 * (S.first + ia)], sh);
          for (int k = 0; k < 3; k++) xs[k] = rz[k] + dr[k];
          pbc_dr(t, xs, xj, shift);
          for (int k = 0; k < 3; k++) r_ij[k] = -t[k];
          double r_mag = std::sqrt(r_ij[0] * r_ij[0] + r_ij[1] * r_ij[1] + r_ij[2] * r_ij[2]);
          *Vex = *Vex + q_i * q_j / r_mag * conv;
          double* fs = &dVex[3 * (S.first + ia)];
          for (int k = 0; k < 3; k++) {
            double dV = -q_i * q_j / (r_mag * r_mag * r_mag) * r_ij[k] * conv;
            fs[k] = fs[k] + dV;
            fj[k] = fj[k] - dV;
          }
        }
      }
    }
  }
}

// evb_diabatic_coupling_function ms_evb.f90:1180-1266
void coupling_function(double* A, double* Vconstij, double dA[3][3], int function_type, const double* fp, const double q[3],
                       const double r_OO[3]) {
  double r_OO_mag = std::sqrt(r_OO[0] * r_OO[0] + r_OO[1] * r_OO[1] + r_OO[2] * r_OO[2]);
  double q2_mag = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
  double q_mag = std::sqrt(q2_mag);
  if (function_type == 1) {
    *Vconstij = fp[0];
    double gamma = fp[1], P = fp[2], k = fp[3], Dd = fp[4], beta = fp[5], R0 = fp[6], Pp = fp[7], alpha = fp[8], rl0 = fp[9];
    double fac1 = std::exp(-gamma * q2_mag);
    double fac2 = 1.0 + P * std::exp(-k * ((r_OO_mag - Dd) * (r_OO_mag - Dd)));
    double fac3 = 0.5 * (1.0 - std::tanh(beta * (r_OO_mag - R0))) + Pp * std::exp(-alpha * (r_OO_mag - rl0));
    double dfac1 = -gamma * 2.0 * q_mag * std::exp(-gamma * q2_mag);
    double dfac2 = P * -k * 2.0 * (r_OO_mag - Dd) * std::exp(-k * ((r_OO_mag - Dd) * (r_OO_mag - Dd)));
    double ch = std::cosh(beta * (r_OO_mag - R0));
    double dfac3 = -0.5 * beta / (ch * ch) - Pp * alpha * std::exp(-alpha * (r_OO_mag - rl0));
    *A = fac1 * fac2 * fac3;
    for (int d = 0; d < 3; d++) {
      dA[0][d] = dfac1 * fac2 * fac3 * 0.5 * q[d] / q_mag;
      dA[0][d] = dA[0][d] + fac1 * dfac2 * fac3 * r_OO[d] / r_OO_mag;
      dA[0][d] = dA[0][d] + fac1 * fac2 * dfac3 * r_OO[d] / r_OO_mag;
      dA[1][d] = dfac1 * fac2 * fac3 * 0.5 * q[d] / q_mag;
      dA[1][d] = dA[1][d] + fac1 * dfac2 * fac3 * -r_OO[d] / r_OO_mag;
      dA[1][d] = dA[1][d] + fac1 * fac2 * dfac3 * -r_OO[d] / r_OO_mag;
      dA[2][d] = dfac1 * fac2 * fac3 * -q[d] / q_mag;
    }
  } else {
    *Vconstij = fp[0];
    double gamma = fp[1], k = fp[2], Dd = fp[3];
    double fac1 = std::exp(-gamma * q2_mag);
    double fac2 = std::exp(-k * ((r_OO_mag - Dd) * (r_OO_mag - Dd)));
    double dfac1 = -gamma * 2.0 * q_mag * std::exp(-gamma * q2_mag);
    double dfac2 = -k * 2.0 * (r_OO_mag - Dd) * std::exp(-k * ((r_OO_mag - Dd) * (r_OO_mag - Dd)));
    *A = fac1 * fac2;
    for (int d = 0; d < 3; d++) {
      dA[0][d] = dfac1 * fac2 * 0.5 * q[d] / q_mag;
      dA[0][d] = dA[0][d] + fac1 * dfac2 * r_OO[d] / r_OO_mag;
      dA[1][d] = dfac1 * fac2 * 0.5 * q[d] / q_mag;
      dA[1][d] = dA[1][d] + fac1 * dfac2 * -r_OO[d] / r_OO_mag;
      dA[2][d] = dfac1 * fac2 * -q[d] / q_mag;
    }
  }
}

// evb_diabatic_coupling ms_evb.f90:1021-1104 (+ _geometric :1117-1174)
int diabatic_coupling(Ctx& c, Diabat& D, int i_diabat, int i_mole_principle) {
  const int N = D.sys.total_atoms;
  D.sys.potential_energy = 0.0;
  std::fill(D.atoms.force.begin(), D.atoms.force.end(), 0.0);
  std::vector<double> dVex(3 * N);
  int ima = i_mole_principle, imd = -1;
  for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
    if (c.proton_log[i_diabat][ih][0] < 0) break;
    imd = ima;
    ima = c.proton_log[i_diabat][ih][3];
  }
  double Vex;
  coupling_electrostatics(c, &Vex, dVex.data(), imd, ima, D);
  const Molecule& md = D.mol[imd];
  const Molecule& ma = D.mol[ima];
  int i_atom_donor = heavy_atom_base(c, md.type);
  int i_atom_acceptor = heavy_atom_acid(c, ma.type);
  if (i_atom_donor < 0 || i_atom_acceptor < 0) return RPB_ERR_STATE;
  double r_O1[3], r_O2[3], r_H[3], shift[3], r_ij[3], r_OO[3], q[3];
  for (int k = 0; k < 3; k++) { r_O1[k] = D.atoms.xyz[3 * (md.first + i_atom_donor) + k]; r_O2[k] = D.atoms.xyz[3 * (ma.first + i_atom_acceptor) + k]; }
  pbc_shift(shift, r_O1, r_O2, D.sys);
  pbc_dr(r_ij, r_O1, r_O2, shift);
  for (int k = 0; k < 3; k++) r_O2[k] = r_O1[k] + r_ij[k];
  for (int k = 0; k < 3; k++) r_H[k] = D.atoms.xyz[3 * (ma.first + ma.n_atom - 1) + k];
  pbc_dr(r_ij, r_O1, r_H, shift);
  for (int k = 0; k < 3; k++) r_H[k] = r_O1[k] + r_ij[k];
  for (int k = 0; k < 3; k++) { r_OO[k] = r_O1[k] - r_O2[k]; q[k] = (r_O1[k] + r_O2[k]) / 2.0 - r_H[k]; }
  int key[3] = {D.atoms.type[md.first + i_atom_donor], D.atoms.type[ma.first + i_atom_acceptor],
                D.atoms.type[ma.first + ma.n_atom - 1]};
  int idx = get_index_atom_set<3>(c.dc_int, key);
  if (idx < 0) { c.err = "couldn't find index in subroutine 'get_index_atom_set'"; return RPB_ERR_STATE; }
  double A, Vconstij, dA[3][3];
  coupling_function(&A, &Vconstij, dA, c.dc_type[idx], c.dc_par[idx], q, r_OO);
  D.sys.potential_energy = (Vconstij + Vex) * A;
  for (int k = 0; k < 3; k++) D.atoms.force[3 * (md.first + i_atom_donor) + k] = -(Vconstij + Vex) * dA[0][k];
  for (int k = 0; k < 3; k++) D.atoms.force[3 * (ma.first + i_atom_acceptor) + k] = -(Vconstij + Vex) * dA[1][k];
  for (int k = 0; k < 3; k++) D.atoms.force[3 * (ma.first + ma.n_atom - 1) + k] = -(Vconstij + Vex) * dA[2][k];
  for (int k = 0; k < 3 * N; k++) D.atoms.force[k] = D.atoms.force[k] - dVex[k] * A;
  return 0;
}

// find_evb_reactive_neighbors ms_evb.f90:702-764
void find_reactive_neighbors(const Ctx& c, int i_mole, int i_atom, int nl[MAXN][2]) {
  for (int k = 0; k < MAXN; k++) nl[k][0] = nl[k][1] = -1;
  int index = 0;
  const Molecule& I = c.mol[i_mole];
  const double cut1 = c.cfg.evb_first_solvation_cutoff * c.cfg.evb_first_solvation_cutoff;
  const double cut2 = c.cfg.evb_reactive_pair_distance * c.cfg.evb_reactive_pair_distance;
  for (int jm = 0; jm < c.sys.n_mole; jm++) {
    if (jm == i_mole) continue;
    const Molecule& J = c.mol[jm];
    double shift[3], dr_com[3];
    pbc_shift(shift, I.r_com, J.r_com, c.sys);
    pbc_dr(dr_com, I.r_com, J.r_com, shift);
    if (dr_com[0] * dr_com[0] + dr_com[1] * dr_com[1] + dr_com[2] * dr_com[2] < cut1) {
      for (int ja = 0; ja < J.n_atom; ja++) {
        if (c.mt[J.type].reactive_basic[ja] == 1) {
          double rij[3];
          pbc_dr(rij, &c.atoms.xyz[3 * (I.first + i_atom)], &c.atoms.xyz[3 * (J.first + ja)], shift);
          if (rij[0] * rij[0] + rij[1] * rij[1] + rij[2] * rij[2] < cut2) {
            if (index < MAXN) { nl[index][0] = jm; nl[index][1] = ja; }  // reference has no bounds check (:749-751)
            index++;
          }
        }
      }
    }
  }
}

// find_bonded_atom_hydrogen general_routines.f90:575-602
int find_bonded_atom_hydrogen(const Ctx& c, int mtype, int h_atom) {
  int heavy = -1, count = 0;
  const MoleculeType& M = c.mt[mtype];
  for (size_t b = 0; b < M.bonds.size() / 2; b++) {
    if (M.bonds[2 * b] == h_atom) { heavy = M.bonds[2 * b + 1]; count++; }
    else if (M.bonds[2 * b + 1] == h_atom) { heavy = M.bonds[2 * b]; count++; }
  }
  return count == 1 ? heavy : -1;
}

// evb_conduct_proton_transfer_recursive ms_evb.f90:498-607
int conduct_proton_transfer_recursive(Ctx& c, int i_mole_donor, int diabat_index_donor) {
  int count = 0;
  for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) { if (c.proton_log[diabat_index_donor][ih][0] < 0) break; count++; }
  if (count >= c.cfg.evb_max_chain) return 0;
  int mtype = c.mol[i_mole_donor].type;
  for (int i_atom = 0; i_atom < c.mol[i_mole_donor].n_atom; i_atom++) {
    if (c.mt[mtype].reactive_proton[i_atom] != 1) continue;
    int nl[MAXN][2];
    find_reactive_neighbors(c, i_mole_donor, i_atom, nl);
    for (int im = 0; im < MAXN; im++) {
      if (nl[im][0] < 0) break;
      c.diabat_index++;
      if (c.diabat_index > c.cfg.evb_max_states) { c.err = "Found more diabat states than the current setting of evb_max_states"; return RPB_ERR_DIABATS; }
      int da = c.diabat_index - 1;  // 0-based id of new diabat
      c.coupling_matrix[da] = diabat_index_donor;
      int i_mole_acceptor = nl[im][0], i_atom_acceptor = nl[im][1];
      int flag_cycle = (c.hydronium_mol == i_mole_acceptor) ? 1 : -1;
      for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
        if (c.proton_log[diabat_index_donor][ih][0] < 0) break;
        for (int f = 0; f < 5; f++) c.proton_log[da][ih][f] = c.proton_log[diabat_index_donor][ih][f];
      }
      int j_atom = find_bonded_atom_hydrogen(c, mtype, i_atom);
      if (j_atom < 0) { c.err = "error in subroutine find_bonded_atom_hydrogen"; return RPB_ERR_STATE; }
      c.proton_log[da][count][0] = i_mole_donor;
      c.proton_log[da][count][1] = i_atom;
      c.proton_log[da][count][2] = j_atom;
      c.proton_log[da][count][3] = i_mole_acceptor;
      c.proton_log[da][count][4] = i_atom_acceptor;
      if (flag_cycle < 1) {
        int rc = conduct_proton_transfer_recursive(c, i_mole_acceptor, da);
        if (rc) return rc;
      }
    }
  }
  return 0;
}

// update_reciprocal_space_force_dQ_dr ms_evb.f90:2103-2248
int update_recip_force_dQ_dr(Ctx& c, double* pme_force, const double* theta, int i_diabat, int i_mole_principle,
                             const double kk[3][3]) {
  Diabat D;
  create_diabat(c, D);
  D.atoms.force.assign(pme_force, pme_force + 3 * c.sys.total_atoms);
  const int K = c.cfg.pme_grid;
  auto mol_force = [&](int im, double sign) {
    const Molecule& m = D.mol[im];
    std::vector<double> us(3 * m.n_atom);
    create_scaled_direct_coordinates(us.data(), &D.atoms.xyz[3 * m.first], m.n_atom, kk, K);
    for (int a = 0; a < m.n_atom; a++) {
      double f[3];
      derivative_grid_Q(c, f, theta, &D.atoms.charge[m.first], us.data(), a, kk, nullptr, nullptr);
      for (int d = 0; d < 3; d++) {
        if (sign < 0) D.atoms.force[3 * (m.first + a) + d] = D.atoms.force[3 * (m.first + a) + d] - f[d];
        else D.atoms.force[3 * (m.first + a) + d] = D.atoms.force[3 * (m.first + a) + d] + f[d];
      }
    }
  };
  int ima = i_mole_principle;
  for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
    if (c.proton_log[i_diabat][ih][0] < 0) break;
    int imd = ima;
    int i_atom_donor = c.proton_log[i_diabat][ih][1];
    ima = c.proton_log[i_diabat][ih][3];
    int i_heavy_acceptor = c.proton_log[i_diabat][ih][4];
    int i_atom_acceptor = D.mol[ima].n_atom;
    mol_force(imd, -1); mol_force(ima, -1);
    int rc = change_topology_proton_transfer(c, D, imd, i_atom_donor, ima, i_atom_acceptor, i_heavy_acceptor);
    if (rc) return rc;
    mol_force(imd, +1); mol_force(ima, +1);
  }
  std::copy(D.atoms.force.begin(), D.atoms.force.end(), pme_force);
  map_force_to_principle(c, i_diabat, 0, i_mole_principle, D.mol, pme_force);
  return 0;
}

bool owned(const Ctx& c, int s0) {  // s0: 0-based diabat id
  if (c.cfg.world_size <= 1) return true;
  if (s0 == 0) return c.cfg.rank == 0;
  return ((s0 - 1) % c.cfg.world_size) == c.cfg.rank;
}

// calculate_reciprocal_space_pme ms_evb.f90:1962-2095
int calculate_reciprocal_space_pme(Ctx& c, int i_mole_principle) {
  const int N = c.sys.total_atoms, K = c.cfg.pme_grid;
  const size_t K3 = (size_t)K * K * K;
  const int p3 = c.cfg.spline_order * c.cfg.spline_order * c.cfg.spline_order;
  const double conv = c.cfg.conv_e2A_kJmol;
  double kk[3][3];
  reciprocal_lattice(kk, c.sys);
  std::vector<int> todo;
  for (int s = 1; s < c.diabat_index; s++) if (owned(c, s)) todo.push_back(s);
  int nt = std::max(1, c.n_threads);
  int err = 0;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1)
  for (int it = 0; it < (int)todo.size(); it++) {
    int s = todo[it];
    std::vector<double> theta(K3), pf(3 * N, 0.0);
    double E_recip_local = pme_convolve(c, c.Q_grid_diabats[s].data(), theta.data());
    double dE_recip = E_recip_local - c.E_recip;
    for (int i = 0; i < N; i++)
      for (int j = 0; j < p3; j++) {
        const double* dq = &c.dQ_dr[(size_t)3 * p3 * i + 3 * j];
        const int* ix = &c.dQ_dr_index[(size_t)3 * p3 * i + 3 * j];
        double th = theta[(size_t)ix[0] + (size_t)K * ix[1] + (size_t)K * K * ix[2]];
        for (int d = 0; d < 3; d++) pf[3 * i + d] = pf[3 * i + d] + dq[d] * th * conv;
      }
    for (int i = 0; i < N; i++) {
      double t[3] = {0, 0, 0};
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) t[a] = t[a] - (double)K * kk[b][a] * pf[3 * i + b];
      for (int a = 0; a < 3; a++) pf[3 * i + a] = t[a];
    }
    for (int k = 0; k < 3 * N; k++) pf[k] = pf[k] - c.force_recip[k];
    int rc = update_recip_force_dQ_dr(c, pf.data(), theta.data(), s, i_mole_principle, kk);
    if (rc) err = rc;
    int index = c.evb_forces_lookup_index[s][s];
    c.evb_hamiltonian[s][s] = c.evb_hamiltonian[s][s] + dE_recip;
    std::vector<double>& st = c.evb_forces_store[index];
    for (int k = 0; k < 3 * N; k++) st[k] = st[k] + pf[k];
    c.theta_diabats[s] = theta;  // kept for rpb_get_pme
  }
  return err;
}

// jacobi general_routines.f90:2013-2088 (Numerical Recipes cyclic Jacobi)
int jacobi(std::vector<double>& a, std::vector<double>& d, std::vector<double>& v, int n) {
  auto A = [&](int i, int j) -> double& { return a[i + (size_t)n * j]; };
  auto V = [&](int i, int j) -> double& { return v[i + (size_t)n * j]; };
  std::vector<double> b(n), z(n, 0.0);
  v.assign((size_t)n * n, 0.0);
  d.assign(n, 0.0);
  for (int i = 0; i < n; i++) { V(i, i) = 1.0; b[i] = A(i, i); d[i] = b[i]; }
  for (int it = 1; it <= 50; it++) {
    double sm = 0.0;
    for (int j = 0; j < n; j++) for (int i = 0; i < j; i++) sm += std::fabs(A(i, j));
    if (sm == 0.0) return 0;
    double tresh = (it < 4) ? 0.2 * sm / (double)(n * n) : 0.0;
    for (int ip = 0; ip < n - 1; ip++) {
      for (int iq = ip + 1; iq < n; iq++) {
        double g = 100.0 * std::fabs(A(ip, iq));
        if (it > 4 && (std::fabs(d[ip]) + g == std::fabs(d[ip])) && (std::fabs(d[iq]) + g == std::fabs(d[iq]))) {
          A(ip, iq) = 0.0;
        } else if (std::fabs(A(ip, iq)) > tresh) {
          double h = d[iq] - d[ip], t;
          if (std::fabs(h) + g == std::fabs(h)) t = A(ip, iq) / h;
          else {
            double theta = 0.5 * h / A(ip, iq);
            t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
            if (theta < 0.0) t = -t;
          }
          double cc = 1.0 / std::sqrt(1 + t * t), s = t * cc, tau = s / (1.0 + cc);
          h = t * A(ip, iq);
          z[ip] = z[ip] - h; z[iq] = z[iq] + h; d[ip] = d[ip] - h; d[iq] = d[iq] + h;
          A(ip, iq) = 0.0;
          auto rot = [&](double& a1, double& a2) {
            double w = a1;
            a1 = a1 - s * (a2 + a1 * tau);
            a2 = a2 + s * (w - a2 * tau);
          };
          for (int k = 0; k < ip; k++) rot(A(k, ip), A(k, iq));
          for (int k = ip + 1; k < iq; k++) rot(A(ip, k), A(k, iq));
          for (int k = iq + 1; k < n; k++) rot(A(ip, k), A(iq, k));
          for (int k = 0; k < n; k++) rot(V(k, ip), V(k, iq));
        }
      }
    }
    for (int i = 0; i < n; i++) { b[i] = b[i] + z[i]; d[i] = b[i]; z[i] = 0.0; }
  }
  return RPB_ERR_STATE;  // 'too many iterations in jacobi'
}

}  // namespace

// construct_evb_hamiltonian ms_evb.f90:375-489 (restricted to the diabats this rank owns)
int evb_phase_build(Ctx& c) {
  const int N = c.sys.total_atoms;
  if (c.hydronium_mol < 0) { c.err = "need at least one hydronium molecule!"; return RPB_ERR_STATE; }
  const int i_mole_principle = c.hydronium_mol;
  for (int s = 0; s < MAXS; s++) {
    for (int h = 0; h < MAXC; h++) for (int f = 0; f < 5; f++) c.proton_log[s][h][f] = -1;
    for (int t = 0; t < MAXS; t++) c.evb_hamiltonian[s][t] = 0.0;
    c.coupling_matrix[s] = -1;
  }
  c.diabat_index = 1;
  int rc = calculate_total_force_energy(c, true);
  if (rc) return rc;
  Diabat P;
  create_diabat(c, P);
  double E_rep;
  rc = intermolecular_repulsion(c, c.atoms.force.data(), &E_rep, P);
  if (rc) return rc;
  P.atoms.force = c.atoms.force;
  c.sys.potential_energy = c.sys.potential_energy + E_rep;
  c.sys.potential_energy = c.sys.potential_energy + c.ref_energy[c.mol[i_mole_principle].type];
  c.evb_hamiltonian[0][0] = c.sys.potential_energy;
  rc = conduct_proton_transfer_recursive(c, i_mole_principle, 0);
  if (rc) return rc;
  for (int i = 0; i < MAXS; i++) for (int j = 0; j < MAXS; j++) c.evb_forces_lookup_index[i][j] = -1;
  c.evb_forces_store.assign(2 * c.diabat_index - 1, std::vector<double>());
  store_forces(c, i_mole_principle, 0, 0, P, 0);
  c.Q_grid_diabats.assign(c.diabat_index, std::vector<double>());
  c.theta_diabats.assign(c.diabat_index, std::vector<double>());
  c.Q_grid_diabats[0] = c.Q_grid;
  c.theta_diabats[0] = c.theta_conv_Q;
  // evb_hamiltonian_elements_donor_acceptor ms_evb.f90:623-692 (OMP DO over diabats 2..S)
  int nt = std::max(1, c.n_threads);
  int err = 0;
#pragma omp parallel for num_threads(nt) schedule(dynamic, 1)
  for (int s = 1; s < c.diabat_index; s++) {
    if (!owned(c, s)) continue;
    Diabat D;
    create_diabat(c, D);
    int r2 = diabat_force_energy(c, D, s, i_mole_principle);
    if (r2) { err = r2; continue; }
    store_forces(c, i_mole_principle, s, s, D, 2 * s - 1);
    c.evb_hamiltonian[s][s] = D.sys.potential_energy;
    r2 = diabatic_coupling(c, D, s, i_mole_principle);
    if (r2) { err = r2; continue; }
    int donor = c.coupling_matrix[s];
    store_forces(c, i_mole_principle, donor, s, D, 2 * s);
    c.evb_hamiltonian[donor][s] = D.sys.potential_energy;
  }
  if (err) return err;
  rc = calculate_reciprocal_space_pme(c, i_mole_principle);
  if (rc) return rc;
  // exchange buffer: [H_ss (80)] [H_parent(s),s (80)], zero where not owned
  c.xh.assign(2 * MAXS, 0.0);
  for (int s = 0; s < c.diabat_index; s++) {
    if (!owned(c, s)) continue;
    c.xh[s] = c.evb_hamiltonian[s][s];
    if (s > 0) c.xh[MAXS + s] = c.evb_hamiltonian[c.coupling_matrix[s]][s];
  }
  c.xf.assign(3 * N, 0.0);
  return 0;
}

// diagonalize_evb_hamiltonian ms_evb.f90:242-351.  coeff_override != nullptr: only re-mix the stored
// forces with that coefficient vector into force_out (rpb_debug_mix_forces).
int evb_phase_mix(Ctx& c, const double* coeff_override, double* force_out) {
  const int N = c.sys.total_atoms, S = c.diabat_index;
  std::vector<double> gs;
  if (!coeff_override) {
    for (int s = 0; s < S; s++) {
      c.evb_hamiltonian[s][s] = c.xh[s];
      if (s > 0) c.evb_hamiltonian[c.coupling_matrix[s]][s] = c.xh[MAXS + s];
    }
    std::vector<double> ham((size_t)S * S, 0.0), ev, evec;
    for (int i = 0; i < S; i++) for (int j = i; j < S; j++) { ham[i + (size_t)S * j] = c.evb_hamiltonian[i][j]; ham[j + (size_t)S * i] = ham[i + (size_t)S * j]; }
    int rc = jacobi(ham, ev, evec, S);
    if (rc) { c.err = "too many iterations in jacobi"; return rc; }
    int ground = 0;
    c.adiabatic_potential = ev[0];
    for (int s = 1; s < S; s++) if (ev[s] < c.adiabatic_potential) { c.adiabatic_potential = ev[s]; ground = s; }
    gs.resize(S);
    for (int s = 0; s < S; s++) gs[s] = evec[s + (size_t)S * ground];
    c.ground_state_eigenvector = gs;
    c.principle_diabat = 0;
    double coef = std::fabs(gs[0]);
    for (int s = 0; s < S; s++) if (coef < std::fabs(gs[s])) { coef = std::fabs(gs[s]); c.principle_diabat = s; }
    c.new_hydronium = c.hydronium_mol;
    for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
      if (c.proton_log[c.principle_diabat][ih][0] < 0) break;
      c.new_hydronium = c.proton_log[c.principle_diabat][ih][3];
    }
  } else {
    gs.assign(coeff_override, coeff_override + S);
  }
  std::vector<double> F(3 * N, 0.0);
  for (int i = 0; i < S; i++)
    for (int j = i; j < S; j++) {
      int index = c.evb_forces_lookup_index[i][j];
      if (!coeff_override && !owned(c, j)) continue;  // element (i,j) is mixed by the rank that built diabat j
      if (index >= 0) {
        double factor = (i != j) ? 2.0 * gs[i] * gs[j] : gs[i] * gs[j];
        const std::vector<double>& st = c.evb_forces_store[index];
        for (int k = 0; k < 3 * N; k++) F[k] = F[k] + factor * st[k];
      }
    }
  if (coeff_override) std::copy(F.begin(), F.end(), force_out);
  else c.xf = F;
  return 0;
}

// tail of ms_evb_calculate_total_force_energy ms_evb.f90:208-233
int evb_phase_commit(Ctx& c) {
  c.atoms.force = c.xf;
  c.sys.potential_energy = c.adiabatic_potential;
  if (c.new_hydronium != c.hydronium_mol) {
    // evb_change_diabat_data_structure_topology ms_evb.f90:806-834, on the real data structures
    Diabat D;
    create_diabat(c, D);
    int ima = c.hydronium_mol;
    for (int ih = 0; ih < c.cfg.evb_max_chain; ih++) {
      if (c.proton_log[c.principle_diabat][ih][0] < 0) break;
      int imd = ima;
      int i_atom_donor = c.proton_log[c.principle_diabat][ih][1];
      ima = c.proton_log[c.principle_diabat][ih][3];
      int i_heavy = c.proton_log[c.principle_diabat][ih][4];
      int rc = change_topology_proton_transfer(c, D, imd, i_atom_donor, ima, D.mol[ima].n_atom, i_heavy);
      if (rc) return rc;
    }
    c.atoms = D.atoms; c.mol = D.mol; c.hydronium_mol = D.hydronium;
    int rc = construct_verlet_list(c);
    if (rc) return rc;
    int junk;
    update_verlet_displacements(c, &junk, true);
  }
  return 0;
}

int ms_evb_calculate_total_force_energy(Ctx& c) {
  int rc = evb_phase_build(c); if (rc) return rc;
  rc = evb_phase_mix(c, nullptr, nullptr); if (rc) return rc;
  return evb_phase_commit(c);
}

}  // namespace orc
