// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_core.h).  PARITY UNPINNED.
// Non-reactive force path: restates src/total_energy_forces.f90, src/pair_int_real_space.f90,
// src/pme.f90, src/intra_bonded_interactions.f90:17-552, src/md_integration.f90:125-177,438-541
// and the hot-path utilities of src/general_routines.f90.
#include "oracle_md.h"
#include <algorithm>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------------------------
// PBC helpers (general_routines.f90:535-568), orthorhombic box (main_ms_evb.f90:62-68):
// dr_box = matmul(xyz_to_box_transform, dr) with a diagonal transform = inv_box(i)*dr(i).
// ---------------------------------------------------------------------------------------------
void pbc_shift(double out[3], const double ri[3], const double rj[3], const SystemData& s) {
  for (int i = 0; i < 3; i++) {
    double dr = rj[i] - ri[i];
    double dr_box = s.inv_box[i] * dr;
    double sh = std::floor(dr_box + 0.5);
    out[i] = sh * s.box[i];
  }
}
void pbc_dr(double out[3], const double ri[3], const double rj[3], const double shift[3]) {
  for (int i = 0; i < 3; i++) out[i] = rj[i] - ri[i] - shift[i];
}

// pos_com general_routines.f90:398-415
void pos_com(double out[3], const double* xyz, const double* mass, int n_atom) {
  double c[3] = {0, 0, 0}, m_tot = 0;
  for (int a = 0; a < n_atom; a++) {
    for (int k = 0; k < 3; k++) c[k] = c[k] + xyz[3 * a + k] * mass[a];
    m_tot = m_tot + mass[a];
  }
  for (int k = 0; k < 3; k++) out[k] = c[k] / m_tot;
}

// update_r_com general_routines.f90:420-440
void update_r_com(Ctx& c) {
  for (auto& m : c.mol) pos_com(m.r_com, &c.atoms.xyz[3 * m.first], &c.atoms.mass[m.first], m.n_atom);
}

// make_molecule_whole general_routines.f90:1065-1086
void make_molecule_whole(int n_atom, double* xyz, const SystemData& s) {
  const double small = 1e-6;
  for (int i = 1; i < n_atom; i++) {
    int j = i - 1;
    double shift[3];
    pbc_shift(shift, &xyz[3 * j], &xyz[3 * i], s);
    if (std::fabs(shift[0]) > small || std::fabs(shift[1]) > small || std::fabs(shift[2]) > small) {
      double drij[3];
      pbc_dr(drij, &xyz[3 * j], &xyz[3 * i], shift);
      for (int k = 0; k < 3; k++) xyz[3 * i + k] = xyz[3 * j + k] + drij[k];
    }
  }
}

// shift_molecules_into_box general_routines.f90:1145-1197
void shift_molecules_into_box(Ctx& c) {
  for (auto& m : c.mol) {
    double t[3];
    for (int i = 0; i < 3; i++) {
      double dr_box = c.sys.inv_box[i] * m.r_com[i];
      double sh = 0.0;
      if (dr_box < 0.0) sh = 1.0; else if (dr_box > 1.0) sh = -1.0;
      t[i] = sh * c.sys.box[i];
    }
    for (int i = 0; i < 3; i++) m.r_com[i] = m.r_com[i] + t[i];
    for (int a = 0; a < m.n_atom; a++)
      for (int i = 0; i < 3; i++) c.atoms.xyz[3 * (m.first + a) + i] = c.atoms.xyz[3 * (m.first + a) + i] + t[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Verlet list (general_routines.f90:1206-1595)
// ---------------------------------------------------------------------------------------------
int verlet_capacity(const Ctx& c) {  // allocate_verlet_list :1206-1247
  if (c.cfg.verlet_capacity > 0) return c.cfg.verlet_capacity;
  double volume = c.sys.box[0] * c.sys.box[0] * c.sys.box[0];  // pme.f90:646 (after periodic_box_change)
  double N = (double)c.sys.total_atoms;
  double rv = c.cfg.verlet_cutoff;
  long long size_verlet = (long long)std::floor(4.0 * c.cfg.pi * (rv * rv * rv) * (N * N) / 6.0 / volume);
  long long min_size = (long long)c.sys.total_atoms * 50;
  size_verlet = std::max(min_size, size_verlet);
  size_verlet = (long long)std::floor((double)size_verlet * c.cfg.safe_verlet);
  return (int)size_verlet;
}

static inline int wrap_cell(int ig, int n) {  // ig - floor(dble(ig-1)/n)*n, 1-based
  return ig - (int)std::floor((double)(ig - 1) / (double)n) * n;
}

int construct_verlet_list(Ctx& c) {  // construct_verlet_list_grid :1408-1595
  const int N = c.sys.total_atoms;
  const int nx = c.cfg.na_nslist, ny = c.cfg.nb_nslist, nz = c.cfg.nc_nslist;
  if (c.sys.box[0] < 2 * c.cfg.verlet_cutoff) { c.err = "box size less than twice verlet cutoff"; return RPB_ERR_ARG; }
  if (nx < 10 || ny < 10 || nz < 10 || nx > 99 || ny > 99 || nz > 99) {
    c.err = " na_nslist, nb_nslist, nc_nslist must be between 10 and 100 "; return RPB_ERR_ARG;
  }
  const double rv2 = c.cfg.verlet_cutoff * c.cfg.verlet_cutoff;
  const int cap = (int)c.neighbor_list.size();
  c.verlet_point.assign(N + 1, 0);
  std::vector<int> nslist(N, -1), index_molecule(N), cellx(N), celly(N), cellz(N);
  std::vector<int> head(nx * ny * nz, -1), endl(nx * ny * nz, -1);
  auto cid = [&](int ix, int iy, int iz) { return (ix - 1) + nx * ((iy - 1) + ny * (iz - 1)); };
  for (int im = 0; im < c.sys.n_mole; im++) {
    for (int a = 0; a < c.mol[im].n_atom; a++) {
      int ai = c.mol[im].first + a;
      index_molecule[ai] = im;
      double r0 = c.sys.inv_box[0] * c.atoms.xyz[3 * ai], r1 = c.sys.inv_box[1] * c.atoms.xyz[3 * ai + 1],
             r2 = c.sys.inv_box[2] * c.atoms.xyz[3 * ai + 2];
      int ix = (int)std::floor(r0 * nx) + 1, iy = (int)std::floor(r1 * ny) + 1, iz = (int)std::floor(r2 * nz) + 1;
      ix = wrap_cell(ix, nx); iy = wrap_cell(iy, ny); iz = wrap_cell(iz, nz);
      cellx[ai] = ix; celly[ai] = iy; cellz[ai] = iz;
      int ci = cid(ix, iy, iz);
      if (head[ci] < 0) { head[ci] = ai; endl[ci] = ai; }
      else { nslist[endl[ci]] = ai; endl[ci] = ai; }
    }
  }
  // rka = |row of xyz_to_box_transform| = inv_box for a diagonal transform (dsqrt of the square)
  double rka = std::sqrt(c.sys.inv_box[0] * c.sys.inv_box[0]), rkb = std::sqrt(c.sys.inv_box[1] * c.sys.inv_box[1]),
         rkc = std::sqrt(c.sys.inv_box[2] * c.sys.inv_box[2]);
  int dia = (int)std::floor(c.cfg.verlet_cutoff * rka * (double)nx) + 1;
  int dib = (int)std::floor(c.cfg.verlet_cutoff * rkb * (double)ny) + 1;
  int dic = (int)std::floor(c.cfg.verlet_cutoff * rkc * (double)nz) + 1;
  if (dia >= nx / 2 || dib >= ny / 2 || dic >= nz / 2) {  // :1605-1634
    c.err = "number of grid cells to search in each dimension must be less than half the total number of grid cells";
    return RPB_ERR_ARG;
  }
  int verlet_index = 1;
  for (int im = 0; im < c.sys.n_mole; im++) {
    for (int a = 0; a < c.mol[im].n_atom; a++) {
      int ai = c.mol[im].first + a;
      c.verlet_point[ai] = verlet_index;
      int ix = cellx[ai], iy = celly[ai], iz = cellz[ai];
      for (int ia = -dia; ia <= dia; ia++) {
        int g1 = wrap_cell(ix + ia, nx);
        for (int ib = -dib; ib <= dib; ib++) {
          int g2 = wrap_cell(iy + ib, ny);
          for (int ic = -dic; ic <= dic; ic++) {
            int g3 = wrap_cell(iz + ic, nz);
            int j = head[cid(g1, g2, g3)];
            while (j >= 0) {
              if (ai < j && index_molecule[j] != im) {
                double r[3];
                for (int k = 0; k < 3; k++) {
                  double rij = c.atoms.xyz[3 * ai + k] - c.atoms.xyz[3 * j + k];
                  double sh = c.sys.box[k] * std::floor(rij / c.sys.box[k] + 0.5);
                  r[k] = rij - sh;
                }
                double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                if (d2 < rv2) {
                  if (verlet_index > cap) { c.err = "please increase size of verlet neighbor list"; return RPB_ERR_VERLET; }
                  c.neighbor_list[verlet_index - 1] = j + 1;
                  verlet_index++;
                }
              }
              j = nslist[j];
            }
          }
        }
      }
    }
  }
  c.verlet_point[N] = verlet_index;
  return 0;
}

// update_verlet_displacements general_routines.f90:1259-1337
void update_verlet_displacements(Ctx& c, int* flag, bool initialize) {
  const int N = c.sys.total_atoms;
  if (initialize) {
    c.verlet_xyz_store = c.atoms.xyz;
    c.verlet_disp_store.assign(3 * N, 0.0);
    *flag = 0;
    return;
  }
  double max_d1 = 0, max_d2 = 0;
  for (int i = 0; i < N; i++) {
    double shift[3], dr[3];
    pbc_shift(shift, &c.verlet_xyz_store[3 * i], &c.atoms.xyz[3 * i], c.sys);
    pbc_dr(dr, &c.verlet_xyz_store[3 * i], &c.atoms.xyz[3 * i], shift);
    for (int k = 0; k < 3; k++) c.verlet_disp_store[3 * i + k] = c.verlet_disp_store[3 * i + k] + dr[k];
    const double* d = &c.verlet_disp_store[3 * i];
    double norm_dr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (norm_dr > max_d2) {
      max_d2 = norm_dr;
      if (max_d2 > max_d1) { double t = max_d1; max_d1 = max_d2; max_d2 = t; }
    }
  }
  c.verlet_xyz_store = c.atoms.xyz;
  double verlet_skin = c.cfg.verlet_thresh * (c.cfg.verlet_cutoff - c.cfg.real_space_cutoff);
  *flag = ((max_d1 + max_d2) > verlet_skin) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// pair kernels (pair_int_real_space.f90:621-816)
// ---------------------------------------------------------------------------------------------
void PairList::clear() { idx.clear(); dr.clear(); dr2.clear(); qq.clear(); par.clear(); f.clear(); }
void PairList::push(int j, const double d[3], double d2, double q, const double* p6) {
  idx.push_back(j); dr.insert(dr.end(), d, d + 3); dr2.push_back(d2); qq.push_back(q);
  if (p6) par.insert(par.end(), p6, p6 + 6); else par.insert(par.end(), 6, 0.0);
  f.insert(f.end(), 3, 0.0);
}

// linear_interpolation_ewald_tables :740-759
static inline void interp_tables(const Ctx& c, double r, double* erfc_v, double* scale_v) {
  double x1 = r / c.cfg.erfc_dx;
  int i_index = (int)std::ceil(x1);
  double coeff2 = (x1 + 1.0) - i_index;
  double coeff1 = 1.0 - coeff2;
  *erfc_v = coeff1 * c.erfc_t[i_index - 1] + coeff2 * c.erfc_t[i_index];
  *scale_v = coeff1 * c.scale_t[i_index - 1] + coeff2 * c.scale_t[i_index];
}

double pairwise_real_space_ewald(const Ctx& c, PairList& p) {  // :698-731
  double E = 0;
  size_t n = p.idx.size();
  std::vector<double> ev(n), sv(n), rm(n), rm3(n);
  for (size_t k = 0; k < n; k++) { rm[k] = std::sqrt(p.dr2[k]); rm3[k] = p.dr2[k] * rm[k]; }
  for (size_t k = 0; k < n; k++) interp_tables(c, rm[k], &ev[k], &sv[k]);
  for (size_t k = 0; k < n; k++) E += p.qq[k] / rm[k] * ev[k];
  for (size_t k = 0; k < n; k++)
    for (int d = 0; d < 3; d++) p.f[3 * k + d] = p.f[3 * k + d] + p.qq[k] / rm3[k] * p.dr[3 * k + d] * sv[k];
  return E;
}

double pairwise_real_space_LJ(PairList& p) {  // :621-645
  double E = 0;
  size_t n = p.idx.size();
  for (size_t k = 0; k < n; k++) {
    double dr6 = p.dr2[k] * p.dr2[k] * p.dr2[k], dr12 = dr6 * dr6;
    E += p.par[6 * k] / dr12 - p.par[6 * k + 1] / dr6;
  }
  for (size_t k = 0; k < n; k++) {
    double dr6 = p.dr2[k] * p.dr2[k] * p.dr2[k], dr12 = dr6 * dr6;
    double fac = 12.0 * p.par[6 * k] / dr12 - 6.0 * p.par[6 * k + 1] / dr6;
    for (int d = 0; d < 3; d++) p.f[3 * k + d] = p.f[3 * k + d] + p.dr[3 * k + d] / p.dr2[k] * fac;
  }
  return E;
}

double pairwise_real_space_sapt(const Ctx& c, PairList& p) {  // :651-690
  double E = 0;
  size_t n = p.idx.size();
  for (size_t k = 0; k < n; k++) {
    const double* L = &p.par[6 * k];
    double dr1 = std::sqrt(p.dr2[k]);
    double dr6 = p.dr2[k] * p.dr2[k] * p.dr2[k], dr8 = dr6 * p.dr2[k], dr10 = dr8 * p.dr2[k], dr12 = dr10 * p.dr2[k];
    int i_index = (int)std::ceil(L[1] * dr1 / c.cfg.tt_max * (double)c.cfg.tt_grid);
    const double* tt = &c.tt[4 * (i_index - 1)];
    double dtt[4];
    for (int q = 0; q < 4; q++) dtt[q] = L[1] * c.dtt[4 * (i_index - 1) + q];
    E += L[0] * std::exp(-1 * L[1] * dr1) - tt[0] * L[2] / dr6 - tt[1] * L[3] / dr8 - tt[2] * L[4] / dr10 -
         tt[3] * L[5] / dr12;
    double fac = dr1 * L[0] * L[1] * std::exp(-1 * L[1] * dr1) + dr1 * dtt[0] * L[2] / dr6 - tt[0] * 6.0 * L[2] / dr6 +
                 dr1 * dtt[1] * L[3] / dr8 - tt[1] * 8.0 * L[3] / dr8 + dr1 * dtt[2] * L[4] / dr10 -
                 tt[2] * 10.0 * L[4] / dr10 + dr1 * dtt[3] * L[5] / dr12 - tt[3] * 12.0 * L[5] / dr12;
    for (int d = 0; d < 3; d++) p.f[3 * k + d] = p.f[3 * k + d] + p.dr[3 * k + d] / p.dr2[k] * fac;
  }
  return E;
}

double intra_pme_exclusion(const Ctx& c, PairList& p) {  // :781-816
  const double small = 1e-8;
  const double alpha = c.cfg.alpha_sqrt, conv = c.cfg.conv_e2A_kJmol;
  const double erf_factor = 2.0 * alpha / c.cfg.pi_sqrt;
  double E = 0;
  for (size_t k = 0; k < p.idx.size(); k++) {
    double rm = std::sqrt(p.dr2[k]), rm3 = p.dr2[k] * rm;
    if (rm < small) {
      E = E - erf_factor * p.qq[k] * conv;
    } else {
      E = E + p.qq[k] * (std::erfc(rm * alpha) - 1.0) / rm * conv;
      double g = (std::erfc(rm * alpha) - 1.0) / rm3 + erf_factor * std::exp(-((rm * alpha) * (rm * alpha))) / p.dr2[k];
      for (int d = 0; d < 3; d++) p.f[3 * k + d] = p.f[3 * k + d] + p.qq[k] * p.dr[3 * k + d] * g * conv;
    }
  }
  return E;
}

static inline void add_pair_forces(double* force, int i, const PairList& p) {
  for (size_t k = 0; k < p.idx.size(); k++) {
    int j = p.idx[k];
    for (int d = 0; d < 3; d++) {
      force[3 * i + d] = force[3 * i + d] + p.f[3 * k + d];
      force[3 * j + d] = force[3 * j + d] - p.f[3 * k + d];
    }
  }
}

// pairwise_real_space_verlet pair_int_real_space.f90:135-371
void pairwise_real_space_verlet(Ctx& c) {
  const int N = c.sys.total_atoms;
  const double rc2 = c.cfg.real_space_cutoff * c.cfg.real_space_cutoff;
  int nt = std::max(1, c.n_threads);
  std::vector<std::vector<double>> temp_force(nt, std::vector<double>(3 * N, 0.0));
  double E_elec = 0, E_vdw = 0;
#pragma omp parallel num_threads(nt) reduction(+ : E_elec, E_vdw)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* local_force = temp_force[tid].data();
    PairList cut, lj, sapt;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < N; i++) {
      int vs = c.verlet_point[i], vf = c.verlet_point[i + 1] - 1;
      int n_neighbors = vf - vs + 1;
      if (n_neighbors <= 0) continue;
      cut.clear(); lj.clear(); sapt.clear();
      int ti = c.atoms.type[i];
      double qi = c.atoms.charge[i];
      for (int v = vs; v <= vf; v++) {
        int j = c.neighbor_list[v - 1] - 1;
        int tj = c.atoms.type[j];
        double dr[3];
        for (int k = 0; k < 3; k++) {
          double d = c.atoms.xyz[3 * i + k] - c.atoms.xyz[3 * j + k];
          dr[k] = d - c.sys.box[k] * std::floor(d / c.sys.box[k] + 0.5);
        }
        double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
        if (dr2 < rc2) {
          double qq = qi * c.atoms.charge[j];
          double par[6];
          for (int k = 0; k < 6; k++) par[k] = c.vdw_param[vdw_idx(ti, tj, k)];
          cut.push(j, dr, dr2, qq, par);
          int vt = c.vdw_type[ti + MAXT * tj];
          if (vt == 0) lj.push(j, dr, dr2, qq, par);
          else if (vt == 1) sapt.push(j, dr, dr2, qq, par);
        }
      }
      double El = cut.idx.empty() ? 0.0 : pairwise_real_space_ewald(c, cut);
      double Elj = lj.idx.empty() ? 0.0 : pairwise_real_space_LJ(lj);
      double Es = sapt.idx.empty() ? 0.0 : pairwise_real_space_sapt(c, sapt);
      E_elec = E_elec + El;
      E_vdw = E_vdw + Elj;
      E_vdw = E_vdw + Es;
      add_pair_forces(local_force, i, lj);
      add_pair_forces(local_force, i, sapt);
      add_pair_forces(local_force, i, cut);
    }
  }
  for (int t = 0; t < nt; t++)
    for (int k = 0; k < 3 * N; k++) c.atoms.force[k] = c.atoms.force[k] + temp_force[t][k];
  c.sys.E_elec = c.sys.E_elec + E_elec;
  c.sys.E_vdw = c.sys.E_vdw + E_vdw;
}

// intra_molecular_pairwise_energy_force pair_int_real_space.f90:386-588
// xyz/charge/type/force point at the molecule block (single_molecule_data pointers).
void intra_molecular_pairwise_energy_force(const Ctx& c, double* force_local, double* E_elec, double* E_vdw,
                                           const double* xyz, const double* charge, const int* type, int i_mole_type,
                                           int n_atom) {
  const double rc2 = c.cfg.real_space_cutoff * c.cfg.real_space_cutoff;
  const MoleculeType& M = c.mt[i_mole_type];
  PairList excl, nonex, nlj, nsapt, cut;
  for (int i = 0; i < n_atom; i++) {
    int ti = type[i];
    excl.clear(); nonex.clear(); nlj.clear(); nsapt.clear(); cut.clear();
    for (int j = i + 1; j < n_atom; j++) {
      int tj = type[j];
      double dr[3] = {xyz[3 * i] - xyz[3 * j], xyz[3 * i + 1] - xyz[3 * j + 1], xyz[3 * i + 2] - xyz[3 * j + 2]};
      double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
      double qq = charge[i] * charge[j];
      if (M.pair_excl[i][j] != 1) {
        nonex.push(j, dr, dr2, qq, nullptr);
        int vt = c.vdw_type[ti + MAXT * tj];
        double par[6];
        if (vt == 0) {
          for (int k = 0; k < 6; k++)
            par[k] = (M.pair_excl[i][j] == 2) ? c.vdw_param14[vdw_idx(ti, tj, k)] : c.vdw_param[vdw_idx(ti, tj, k)];
          nlj.push(j, dr, dr2, qq, par);
        } else if (vt == 1) {
          for (int k = 0; k < 6; k++) par[k] = c.vdw_param[vdw_idx(ti, tj, k)];
          nsapt.push(j, dr, dr2, qq, par);
        }
      } else {
        excl.push(j, dr, dr2, qq, nullptr);
      }
    }
    if (!excl.idx.empty()) { double El = intra_pme_exclusion(c, excl); *E_elec = *E_elec + El; }
    if (!nonex.idx.empty()) {
      for (size_t k = 0; k < nonex.idx.size(); k++)
        if (nonex.dr2[k] < rc2) cut.push(nonex.idx[k], &nonex.dr[3 * k], nonex.dr2[k], nonex.qq[k], nullptr);
      double El = pairwise_real_space_ewald(c, cut);
      double Elj = nlj.idx.empty() ? 0.0 : pairwise_real_space_LJ(nlj);
      double Es = nsapt.idx.empty() ? 0.0 : pairwise_real_space_sapt(c, nsapt);
      *E_elec = *E_elec + El;
      *E_vdw = *E_vdw + Elj;
      *E_vdw = *E_vdw + Es;
    }
    add_pair_forces(force_local, i, excl);
    add_pair_forces(force_local, i, cut);  // reference loops to n_nonexcluded (latent OOB if a pair is beyond r_c, :559-563)
    add_pair_forces(force_local, i, nlj);
    add_pair_forces(force_local, i, nsapt);
  }
}

// real_space_energy_force pair_int_real_space.f90:60-122
void real_space_energy_force(Ctx& c) {
  pairwise_real_space_verlet(c);
  for (auto& m : c.mol)
    if (m.n_atom > 1)
      intra_molecular_pairwise_energy_force(c, &c.atoms.force[3 * m.first], &c.sys.E_elec, &c.sys.E_vdw,
                                            &c.atoms.xyz[3 * m.first], &c.atoms.charge[m.first], &c.atoms.type[m.first],
                                            m.type, m.n_atom);
}

// ---------------------------------------------------------------------------------------------
// PME (pme.f90)
// ---------------------------------------------------------------------------------------------
// construct_reciprocal_lattice_vector general_routines.f90:473-490 with REAL*4 volume() :1936-1947
void reciprocal_lattice(double kk[3][3], const SystemData& s) {
  double a[3] = {s.box[0], 0, 0}, b[3] = {0, s.box[1], 0}, cc[3] = {0, 0, s.box[2]};
  double v = a[0] * (b[1] * cc[2] - b[2] * cc[1]) - a[1] * (b[0] * cc[2] - b[2] * cc[0]) + a[2] * (b[0] * cc[1] - b[1] * cc[0]);
  float v32 = (float)v;       // `real function volume`
  v32 = std::fabs(v32);
  double vol = (double)v32;
  auto cross = [](const double* x, const double* y, double* o) {
    o[0] = x[1] * y[2] - x[2] * y[1]; o[1] = -x[0] * y[2] + x[2] * y[0]; o[2] = x[0] * y[1] - x[1] * y[0];
  };
  double ka[3], kb[3], kc[3];
  cross(a, b, kc); cross(b, cc, ka); cross(cc, a, kb);
  for (int k = 0; k < 3; k++) { kk[0][k] = ka[k] / vol; kk[1][k] = kb[k] / vol; kk[2][k] = kc[k] / vol; }
}

// create_scaled_direct_coordinates general_routines.f90:497-524
void create_scaled_direct_coordinates(double* xyz_scale, const double* xyz, int n_atom, const double kk[3][3], int K) {
  const double small = 1e-6;
  for (int i = 0; i < n_atom; i++)
    for (int l = 0; l < 3; l++) {
      double dot = kk[l][0] * xyz[3 * i] + kk[l][1] * xyz[3 * i + 1] + kk[l][2] * xyz[3 * i + 2];
      double u = (double)K * dot;
      if (u < 0.0) u = u + (double)K; else if (u >= (double)K) u = u - (double)K;
      if (std::fabs(std::fmod(u, 1.0)) < small) u = u + small;
      xyz_scale[3 * i + l] = u;
    }
}

// grid_Q / modify_Q_grid pme.f90:184-335.  op: 0 = grid_Q (+, no charge test), +1/-1 = modify_Q_grid
void spread_atoms(const Ctx& c, double* Q, const double* chg, const double* u3, int n_atom, int op) {
  const int K = c.cfg.pme_grid, p = c.cfg.spline_order;
  const double sg = (double)c.cfg.spline_grid;
  for (int j = 0; j < n_atom; j++) {
    if (op != 0 && !(std::fabs(chg[j]) > 1e-6)) continue;
    const double* u = &u3[3 * j];
    int np[3] = {(int)std::floor(u[0]), (int)std::floor(u[1]), (int)std::floor(u[2])};
    for (int k3 = 0; k3 < p; k3++) {
      int n3 = np[2] - k3; double a3 = u[2] - (double)n3; if (n3 < 0) n3 += K;
      for (int k2 = 0; k2 < p; k2++) {
        int n2 = np[1] - k2; double a2 = u[1] - (double)n2; if (n2 < 0) n2 += K;
        for (int k1 = 0; k1 < p; k1++) {
          int n1 = np[0] - k1; double a1 = u[0] - (double)n1; if (n1 < 0) n1 += K;
          int s1 = (int)std::ceil(a1 / 6.0 * sg), s2 = (int)std::ceil(a2 / 6.0 * sg), s3 = (int)std::ceil(a3 / 6.0 * sg);
          double sum = chg[j] * c.B6[s1 - 1] * c.B6[s2 - 1] * c.B6[s3 - 1];
          size_t g = (size_t)n1 + (size_t)K * n2 + (size_t)K * K * n3;
          if (op < 0) Q[g] = Q[g] - sum; else Q[g] = Q[g] + sum;
        }
      }
    }
  }
}

// derivative_grid_Q pme.f90:346-498.  store != nullptr mimics the ms_evb dQ_dr storage (:470-479).
void derivative_grid_Q(const Ctx& c, double force[3], const double* FQ, const double* chg, const double* u3, int i_atom,
                       const double kk[3][3], double* dQ_dr_store, int* dQ_dr_index_store) {
  const int K = c.cfg.pme_grid, p = c.cfg.spline_order;
  const double sg = (double)c.cfg.spline_grid, conv = c.cfg.conv_e2A_kJmol;
  const double pm1 = (double)(p - 1), pp = (double)p;
  double f[3] = {0, 0, 0};
  double chg_i = chg[i_atom];
  const double* u = &u3[3 * i_atom];
  int np[3] = {(int)std::floor(u[0]), (int)std::floor(u[1]), (int)std::floor(u[2])};
  int count = 0;
  for (int k3 = 0; k3 < p; k3++) {
    int n3 = np[2] - k3; double a13 = u[2] - (double)n3, a23 = a13 - 1.0; if (n3 < 0) n3 += K;
    for (int k2 = 0; k2 < p; k2++) {
      int n2 = np[1] - k2; double a12 = u[1] - (double)n2, a22 = a12 - 1.0; if (n2 < 0) n2 += K;
      for (int k1 = 0; k1 < p; k1++) {
        int n1 = np[0] - k1; double a11 = u[0] - (double)n1, a21 = a11 - 1.0; if (n1 < 0) n1 += K;
        double arg1[3] = {a11, a12, a13}, arg2[3] = {a21, a22, a23};
        int g2[3], g1n[3], g1nmin[3];
        for (int d = 0; d < 3; d++) {
          g2[d] = (int)std::ceil(arg2[d] / pm1 * sg);
          g1n[d] = (int)std::ceil(arg1[d] / pp * sg);
          g1nmin[d] = (int)std::ceil(arg1[d] / pm1 * sg);
        }
        double fac[3] = {0, 0, 0};
        for (int d = 0; d < 3; d++) {
          int o1 = (d == 0) ? 1 : 0, o2 = (d == 2) ? 1 : 2;  // x:(2,3) y:(1,3) z:(1,2)
          if (arg1[d] < pm1) fac[d] = chg_i * (c.B5[g1nmin[d] - 1] * c.B6[g1n[o1] - 1] * c.B6[g1n[o2] - 1]);
          if (0.0 < arg2[d]) fac[d] = fac[d] + chg_i * (-c.B5[g2[d] - 1] * c.B6[g1n[o1] - 1] * c.B6[g1n[o2] - 1]);
        }
        double th = FQ[(size_t)n1 + (size_t)K * n2 + (size_t)K * K * n3];
        for (int d = 0; d < 3; d++) f[d] = f[d] + fac[d] * th * conv;
        if (dQ_dr_store) {
          for (int d = 0; d < 3; d++) dQ_dr_store[3 * count + d] = fac[d];
          dQ_dr_index_store[3 * count] = n1; dQ_dr_index_store[3 * count + 1] = n2; dQ_dr_index_store[3 * count + 2] = n3;
          count++;
        }
      }
    }
  }
  for (int i = 0; i < 3; i++) {
    double t = 0.0;
    for (int j = 0; j < 3; j++) t = t - (double)K * kk[j][i] * f[j];
    force[i] = t;
  }
}

// --- unnormalised complex 3-D DFT pair standing in for MKL DFTI (pme.f90:85,113; default scale 1) ---
namespace {
void fft_rec(const cplx* in, cplx* out, int n, int stride, const cplx* tw, int tws, int N, cplx* scratch) {
  if (n == 1) { out[0] = in[0]; return; }
  int p = n;
  for (int q = 2; q * q <= n; q++) if (n % q == 0) { p = q; break; }
  int m = n / p;
  for (int r = 0; r < p; r++) fft_rec(in + (size_t)r * stride, out + (size_t)r * m, m, stride * p, tw, tws * p, N, scratch);
  for (int k = 0; k < m; k++) {
    for (int q = 0; q < p; q++) {
      cplx acc = out[k];
      int kk2 = k + q * m;
      for (int r = 1; r < p; r++) acc += out[(size_t)r * m + k] * tw[((long long)r * kk2 * tws) % N];
      scratch[q] = acc;
    }
    for (int q = 0; q < p; q++) out[k + (size_t)q * m] = scratch[q];
  }
}
}  // namespace

void fft3d(cplx* a, int K, int sign) {
  std::vector<cplx> tw(K), line(K), res(K), scratch(K);
  for (int k = 0; k < K; k++) {
    long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)K;
    tw[k] = cplx((double)cosl(ang), (double)(sign * sinl(ang)));
  }
  size_t K2 = (size_t)K * K;
  for (int dim = 0; dim < 3; dim++) {
    size_t st = dim == 0 ? 1 : (dim == 1 ? (size_t)K : K2);
    for (int b = 0; b < K; b++)
      for (int cidx = 0; cidx < K; cidx++) {
        size_t base = dim == 0 ? (size_t)K * b + K2 * cidx : (dim == 1 ? (size_t)b + K2 * cidx : (size_t)b + (size_t)K * cidx);
        for (int i = 0; i < K; i++) line[i] = a[base + st * i];
        fft_rec(line.data(), res.data(), K, 1, tw.data(), 1, K, scratch.data());
        for (int i = 0; i < K; i++) a[base + st * i] = res[i];
      }
  }
}

// Q -> theta_conv_Q, returns E = .5*sum(Q*theta)*conv  (pme.f90:73-128 ; ms_evb.f90:2029-2048)
double pme_convolve(const Ctx& c, const double* Q, double* theta) {
  const int K = c.cfg.pme_grid;
  size_t K3 = (size_t)K * K * K;
  std::vector<cplx> q(K3);
  for (size_t i = 0; i < K3; i++) q[i] = cplx(Q[i], 0.0);
  fft3d(q.data(), K, -1);
  for (size_t i = 0; i < K3; i++) q[i] = q[i] * c.CB[i];
  fft3d(q.data(), K, +1);
  double s = 0.0;
  for (size_t i = 0; i < K3; i++) { theta[i] = q[i].real(); s += Q[i] * theta[i]; }
  return 0.5 * s * c.cfg.conv_e2A_kJmol;
}

// pme_reciprocal_space_energy_force pme.f90:28-179
void pme_reciprocal_space_energy_force(Ctx& c, bool store_dQ_dr) {
  const int N = c.sys.total_atoms, K = c.cfg.pme_grid;
  size_t K3 = (size_t)K * K * K;
  double kk[3][3];
  reciprocal_lattice(kk, c.sys);
  std::vector<double> xyz_scale(3 * N);
  create_scaled_direct_coordinates(xyz_scale.data(), c.atoms.xyz.data(), N, kk, K);
  c.Q_grid.assign(K3, 0.0);
  c.theta_conv_Q.assign(K3, 0.0);
  spread_atoms(c, c.Q_grid.data(), c.atoms.charge.data(), xyz_scale.data(), N, 0);
  double E = pme_convolve(c, c.Q_grid.data(), c.theta_conv_Q.data());
  c.sys.E_elec = c.sys.E_elec + E;
  c.E_recip = E;
  c.force_recip.assign(3 * N, 0.0);
  const int p3 = c.cfg.spline_order * c.cfg.spline_order * c.cfg.spline_order;
  if (store_dQ_dr) { c.dQ_dr.assign((size_t)3 * p3 * N, 0.0); c.dQ_dr_index.assign((size_t)3 * p3 * N, 0); }
  int nt = std::max(1, c.n_threads);
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int i = 0; i < N; i++) {
    double f[3];
    derivative_grid_Q(c, f, c.theta_conv_Q.data(), c.atoms.charge.data(), xyz_scale.data(), i, kk,
                      store_dQ_dr ? &c.dQ_dr[(size_t)3 * p3 * i] : nullptr,
                      store_dQ_dr ? &c.dQ_dr_index[(size_t)3 * p3 * i] : nullptr);
    for (int d = 0; d < 3; d++) c.force_recip[3 * i + d] = c.force_recip[3 * i + d] + f[d];
  }
  for (int k = 0; k < 3 * N; k++) c.atoms.force[k] = c.atoms.force[k] + c.force_recip[k];
}

// ---------------------------------------------------------------------------------------------
// bonded terms (intra_bonded_interactions.f90:84-552); type look-ups by atom types as in the source
// ---------------------------------------------------------------------------------------------
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

int pairwise_bond_energy_force(const Ctx& c, double* E, double f[3], int ti, int tj, const double r[3], double r_mag) {
  int bt = c.bond_type[ti + MAXT * tj];
  auto P = [&](int k) { return c.bond_param[ti + MAXT * tj + MAXT * MAXT * k]; };
  if (bt == 1) {
    double b0 = P(0), kb = P(1);
    *E = 0.5 * kb * ((r_mag - b0) * (r_mag - b0));
    for (int d = 0; d < 3; d++) f[d] = -kb * (r_mag - b0) * r[d] / r_mag;
  } else if (bt == 2) {
    double b0 = P(0), kb = P(1);
    double t = r_mag * r_mag - b0 * b0;
    *E = 0.25 * kb * (t * t);
    for (int d = 0; d < 3; d++) f[d] = -kb * t * r[d];
  } else if (bt == 3) {
    double D = P(0), beta = P(1), b0 = P(2);
    double e = std::exp(-beta * (r_mag - b0));
    *E = D * ((1.0 - e) * (1.0 - e));
    for (int d = 0; d < 3; d++) f[d] = -2.0 * D * beta * e * (1.0 - e) * r[d] / r_mag;
  } else {
    return RPB_ERR_ARG;
  }
  return 0;
}

int trimer_angle_energy_force(const Ctx& c, double* E, double f_ij[3], double f_kj[3], int ti, int tj, int tk,
                              const double r_ij[3], const double r_kj[3]) {
  const double small = 1e-4;
  double rij_mag = std::sqrt(dot3(r_ij, r_ij)), rkj_mag = std::sqrt(dot3(r_kj, r_kj));
  double cosine = dot3(r_ij, r_kj) / rij_mag / rkj_mag;
  size_t idx = ti + MAXT * tj + (size_t)MAXT * MAXT * tk;
  int at = c.angle_type[idx];
  double th0 = c.angle_param[idx], cth = c.angle_param[idx + (size_t)MAXT * MAXT * MAXT];
  double fac;
  if (at == 1) {
    double theta;
    if (cosine < -0.999999999) theta = c.cfg.pi; else if (cosine > 0.999999999) theta = 0.0; else theta = std::acos(cosine);
    *E = 0.5 * cth * ((theta - th0) * (theta - th0));
    if (std::fabs(theta - th0) < small) fac = 0.0; else fac = cth * (theta - th0) / std::sqrt(1.0 - cosine * cosine);
  } else if (at == 2) {
    double cosine0 = std::cos(th0);
    *E = 0.5 * cth * ((cosine - cosine0) * (cosine - cosine0));
    fac = -cth * (cosine - cosine0);
  } else {
    return RPB_ERR_ARG;
  }
  for (int d = 0; d < 3; d++) {
    f_ij[d] = fac * (r_kj[d] / rij_mag / rkj_mag - cosine * r_ij[d] / (rij_mag * rij_mag));
    f_kj[d] = fac * (r_ij[d] / rij_mag / rkj_mag - cosine * r_kj[d] / (rkj_mag * rkj_mag));
  }
  return 0;
}

int quartet_dihedral_energy_force(const Ctx& c, double* E, double f_ji[3], double f_kj[3], double f_lk[3], int ti,
                                  int tj, int tk, int tl, const double r_ji[3], const double r_kj[3], const double r_lk[3]) {
  const double small = 1e-4;
  double rji2 = dot3(r_ji, r_ji), rkj2 = dot3(r_kj, r_kj), rlk2 = dot3(r_lk, r_lk);
  double d_kj_ji = dot3(r_kj, r_ji), d_lk_kj = dot3(r_lk, r_kj), d_lk_ji = dot3(r_lk, r_ji);
  double a_dot_b = d_kj_ji * d_lk_kj - d_lk_ji * rkj2;
  double a_dot_a = rji2 * rkj2 - d_kj_ji * d_kj_ji;
  double b_dot_b = rlk2 * rkj2 - d_lk_kj * d_lk_kj;
  double dab_ji[3], dab_kj[3], dab_lk[3], daa_ji[3], daa_kj[3], daa_lk[3], dbb_ji[3], dbb_kj[3], dbb_lk[3];
  for (int d = 0; d < 3; d++) {
    dab_ji[d] = r_kj[d] * d_lk_kj - r_lk[d] * rkj2;
    dab_kj[d] = r_ji[d] * d_lk_kj + d_kj_ji * r_lk[d] - d_lk_ji * 2.0 * r_kj[d];
    dab_lk[d] = d_kj_ji * r_kj[d] - r_ji[d] * rkj2;
    daa_ji[d] = rkj2 * 2.0 * r_ji[d] - 2.0 * d_kj_ji * r_kj[d];
    daa_kj[d] = rji2 * 2.0 * r_kj[d] - 2.0 * d_kj_ji * r_ji[d];
    daa_lk[d] = 0.0;
    dbb_ji[d] = 0.0;
    dbb_kj[d] = rlk2 * 2.0 * r_kj[d] - 2.0 * d_lk_kj * r_lk[d];
    dbb_lk[d] = rkj2 * 2.0 * r_lk[d] - 2.0 * d_lk_kj * r_kj[d];
  }
  double sa = std::sqrt(a_dot_a), sb = std::sqrt(b_dot_b);
  double cosine = a_dot_b / sa / sb;
  double xi;
  if (cosine < -0.999999999) xi = c.cfg.pi; else if (cosine > 0.999999999) xi = 0.0; else xi = std::acos(cosine);
  size_t idx = ti + MAXT * tj + (size_t)MAXT * MAXT * tk + (size_t)MAXT * MAXT * MAXT * tl;
  const size_t T4 = (size_t)MAXT * MAXT * MAXT * MAXT;
  int dt = c.dihedral_type[idx];
  auto P = [&](int k) { return c.dihedral_param[idx + T4 * k]; };
  double fac = 0.0;
  int shift = 0;
  if (dt == 1) {
    double xi0 = P(0), kxi = P(1), n_mult = P(2);
    *E = kxi * (1.0 + std::cos(n_mult * xi - xi0));
    double cosine2 = cosine * cosine;
    if (std::fabs(cosine2 - 1.0) < small) {
      if (std::fabs(xi0) < small || std::fabs(xi0 - c.cfg.pi) < small) fac = 0.0; else return RPB_ERR_ARG;
    } else {
      fac = kxi * -std::sin(n_mult * xi - xi0) * n_mult / std::sqrt(1.0 - cosine2);
    }
  } else if (dt == 2) {
    if (xi > (c.cfg.pi / 2.0)) { xi = std::fabs(xi - c.cfg.pi); shift = 1; }
    double xi0 = P(0), kxi = P(1);
    *E = 0.5 * kxi * ((xi - xi0) * (xi - xi0));
    if (std::fabs(xi - xi0) < small) fac = 0.0; else fac = kxi * (xi - xi0) / std::sqrt(1.0 - cosine * cosine);
  } else if (dt == 3) {
    double c0 = P(0), c1 = P(1), c2 = P(2), c3 = P(3), c4 = P(4), c5 = P(5);
    double co2 = cosine * cosine, co3 = co2 * cosine, co4 = co3 * cosine, co5 = co4 * cosine;
    *E = c0 - c1 * cosine + c2 * co2 - c3 * co3 + c4 * co4 - c5 * co5;
    fac = c1 - 2.0 * c2 * cosine + 3.0 * c3 * co2 - 4.0 * c4 * co3 + 5.0 * c5 * co4;
  } else {
    *E = 0.0;
    for (int d = 0; d < 3; d++) f_ji[d] = f_kj[d] = f_lk[d] = 0.0;  // Select Case with no match: outputs undefined in source
    return 0;
  }
  double aa15 = std::pow(a_dot_a, 1.5), bb15 = std::pow(b_dot_b, 1.5);
  for (int d = 0; d < 3; d++) {
    f_ji[d] = fac * (dab_ji[d] / sa / sb - 0.5 * a_dot_b / aa15 / sb * daa_ji[d] - 0.5 * a_dot_b / sa / bb15 * dbb_ji[d]);
    f_kj[d] = fac * (dab_kj[d] / sa / sb - 0.5 * a_dot_b / aa15 / sb * daa_kj[d] - 0.5 * a_dot_b / sa / bb15 * dbb_kj[d]);
    f_lk[d] = fac * (dab_lk[d] / sa / sb - 0.5 * a_dot_b / aa15 / sb * daa_lk[d] - 0.5 * a_dot_b / sa / bb15 * dbb_lk[d]);
  }
  if (shift == 1) for (int d = 0; d < 3; d++) { f_ji[d] = -f_ji[d]; f_kj[d] = -f_kj[d]; f_lk[d] = -f_lk[d]; }
  return 0;
}

int intra_molecular_bond_energy_force(const Ctx& c, double* E_bond, const double* xyz, const int* type, double* force, int mtype) {
  *E_bond = 0.0;
  const MoleculeType& M = c.mt[mtype];
  for (size_t b = 0; b < M.bonds.size() / 2; b++) {
    int i = M.bonds[2 * b], j = M.bonds[2 * b + 1];
    double r[3] = {xyz[3 * i] - xyz[3 * j], xyz[3 * i + 1] - xyz[3 * j + 1], xyz[3 * i + 2] - xyz[3 * j + 2]};
    double r_mag = std::sqrt(dot3(r, r));
    double E, f[3];
    if (pairwise_bond_energy_force(c, &E, f, type[i], type[j], r, r_mag)) return RPB_ERR_ARG;
    *E_bond = *E_bond + E;
    for (int d = 0; d < 3; d++) { force[3 * i + d] = force[3 * i + d] + f[d]; force[3 * j + d] = force[3 * j + d] - f[d]; }
  }
  return 0;
}
int intra_molecular_angle_energy_force(const Ctx& c, double* E_angle, const double* xyz, const int* type, double* force, int mtype) {
  *E_angle = 0.0;
  const MoleculeType& M = c.mt[mtype];
  for (size_t a = 0; a < M.angles.size() / 3; a++) {
    int i = M.angles[3 * a], j = M.angles[3 * a + 1], k = M.angles[3 * a + 2];
    double r_ij[3], r_kj[3];
    for (int d = 0; d < 3; d++) { r_ij[d] = xyz[3 * i + d] - xyz[3 * j + d]; r_kj[d] = xyz[3 * k + d] - xyz[3 * j + d]; }
    double E, f_ij[3], f_kj[3];
    if (trimer_angle_energy_force(c, &E, f_ij, f_kj, type[i], type[j], type[k], r_ij, r_kj)) return RPB_ERR_ARG;
    *E_angle = *E_angle + E;
    for (int d = 0; d < 3; d++) {
      force[3 * i + d] = force[3 * i + d] + f_ij[d];
      force[3 * k + d] = force[3 * k + d] + f_kj[d];
      force[3 * j + d] = force[3 * j + d] - f_ij[d] - f_kj[d];
    }
  }
  return 0;
}
int intra_molecular_dihedral_energy_force(const Ctx& c, double* E_dih, const double* xyz, const int* type, double* force, int mtype) {
  *E_dih = 0.0;
  const MoleculeType& M = c.mt[mtype];
  for (size_t q = 0; q < M.dihedrals.size() / 4; q++) {
    int i = M.dihedrals[4 * q], j = M.dihedrals[4 * q + 1], k = M.dihedrals[4 * q + 2], l = M.dihedrals[4 * q + 3];
    double r_ji[3], r_kj[3], r_lk[3];
    for (int d = 0; d < 3; d++) {
      r_ji[d] = xyz[3 * j + d] - xyz[3 * i + d]; r_kj[d] = xyz[3 * k + d] - xyz[3 * j + d]; r_lk[d] = xyz[3 * l + d] - xyz[3 * k + d];
    }
    double E, f_ji[3], f_kj[3], f_lk[3];
    if (quartet_dihedral_energy_force(c, &E, f_ji, f_kj, f_lk, type[i], type[j], type[k], type[l], r_ji, r_kj, r_lk))
      return RPB_ERR_ARG;
    *E_dih = *E_dih + E;
    for (int d = 0; d < 3; d++) {
      force[3 * i + d] = force[3 * i + d] - f_ji[d];
      force[3 * j + d] = force[3 * j + d] + f_ji[d] - f_kj[d];
      force[3 * k + d] = force[3 * k + d] + f_kj[d] - f_lk[d];
      force[3 * l + d] = force[3 * l + d] + f_lk[d];
    }
  }
  return 0;
}

// intra_molecular_energy_force intra_bonded_interactions.f90:17-73
int intra_molecular_energy_force(Ctx& c) {
  double Eb = 0, Ea = 0, Ed = 0;
  for (auto& m : c.mol) {
    double e1, e2, e3;
    const double* x = &c.atoms.xyz[3 * m.first]; const int* t = &c.atoms.type[m.first]; double* f = &c.atoms.force[3 * m.first];
    if (intra_molecular_bond_energy_force(c, &e1, x, t, f, m.type)) { c.err = "bond type isn't implemented!"; return RPB_ERR_ARG; }
    if (intra_molecular_angle_energy_force(c, &e2, x, t, f, m.type)) { c.err = "requested angle type potential not implemented"; return RPB_ERR_ARG; }
    if (intra_molecular_dihedral_energy_force(c, &e3, x, t, f, m.type)) { c.err = "undefined dihedral force"; return RPB_ERR_ARG; }
    Eb = Eb + e1; Ea = Ea + e2; Ed = Ed + e3;
  }
  c.sys.E_bond = Eb; c.sys.E_angle = Ea; c.sys.E_dihedral = Ed;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// calculate_total_force_energy total_energy_forces.f90:19-99
// ---------------------------------------------------------------------------------------------
int calculate_total_force_energy(Ctx& c, bool ms_evb) {
  if (c.flag_verlet_list == 1) {
    int rc = construct_verlet_list(c);
    if (rc) return rc;
    int junk;
    update_verlet_displacements(c, &junk, true);
    c.flag_verlet_list = 0;
  } else {
    update_verlet_displacements(c, &c.flag_verlet_list, false);
  }
  std::fill(c.atoms.force.begin(), c.atoms.force.end(), 0.0);
  c.sys.potential_energy = 0; c.sys.E_elec = 0; c.sys.E_vdw = 0; c.sys.E_bond = 0; c.sys.E_angle = 0; c.sys.E_dihedral = 0;
  real_space_energy_force(c);
  pme_reciprocal_space_energy_force(c, ms_evb);
  c.sys.E_elec = c.sys.E_elec + c.cfg.ewald_self;
  int rc = intra_molecular_energy_force(c);
  if (rc) return rc;
  c.sys.potential_energy = c.sys.E_elec + c.sys.E_vdw + c.sys.E_bond + c.sys.E_angle + c.sys.E_dihedral;
  return 0;
}

// calculate_kinetic_energy total_energy_forces.f90:106-121
double calculate_kinetic_energy(const Ctx& c) {
  double KE = 0;
  for (int i = 0; i < c.sys.total_atoms; i++) {
    const double* v = &c.atoms.vel[3 * i];
    KE = KE + 0.5 * c.atoms.mass[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / c.cfg.conv_kJmol_ang2ps2gmol;
  }
  return KE;
}

// md_integrate_atomic, first half (md_integration.f90:469-491)
void md_step_begin(Ctx& c) {
  const double dt = c.cfg.delta_t, conv = c.cfg.conv_kJmol_ang2ps2gmol;
  for (int i = 0; i < c.sys.total_atoms; i++) {
    if (c.atype_freeze[c.atoms.type[i]] != 1) {
      for (int d = 0; d < 3; d++) {
        c.atoms.vel[3 * i + d] = c.atoms.vel[3 * i + d] + dt / 2.0 / c.atoms.mass[i] * c.atoms.force[3 * i + d] * conv;
        c.atoms.xyz[3 * i + d] = c.atoms.xyz[3 * i + d] + c.atoms.vel[3 * i + d] * dt;
      }
    }
  }
  update_r_com(c);
  shift_molecules_into_box(c);
}

// md_integrate_atomic, second half (md_integration.f90:507-532) + subtract_center_of_mass_momentum (:125-177)
int md_step_end(Ctx& c) {
  const double dt = c.cfg.delta_t, conv = c.cfg.conv_kJmol_ang2ps2gmol;
  for (int i = 0; i < c.sys.total_atoms; i++) {
    if (c.atype_freeze[c.atoms.type[i]] != 1) {
      for (int d = 0; d < 3; d++)
        c.atoms.vel[3 * i + d] = c.atoms.vel[3 * i + d] + dt / 2.0 / c.atoms.mass[i] * c.atoms.force[3 * i + d] * conv;
      const double* f = &c.atoms.force[3 * i];
      if (std::fabs(f[0]) > 10e4 || std::fabs(f[1]) > 10e4 || std::fabs(f[2]) > 10e4) {
        c.err = "force on atom " + std::to_string(i + 1) + " is too big";
        return RPB_ERR_FORCE;
      }
    }
  }
  int n_tot = 0;
  double rho[3] = {0, 0, 0};
  for (int i = 0; i < c.sys.total_atoms; i++)
    if (c.atype_freeze[c.atoms.type[i]] != 1) {
      n_tot++;
      for (int d = 0; d < 3; d++) rho[d] = rho[d] + c.atoms.mass[i] * c.atoms.vel[3 * i + d];
    }
  double ex[3] = {rho[0] / (double)n_tot, rho[1] / (double)n_tot, rho[2] / (double)n_tot};
  for (int i = 0; i < c.sys.total_atoms; i++)
    if (c.atype_freeze[c.atoms.type[i]] != 1)
      for (int d = 0; d < 3; d++) c.atoms.vel[3 * i + d] = c.atoms.vel[3 * i + d] - ex[d] / c.atoms.mass[i];
  return 0;
}

}  // namespace orc
