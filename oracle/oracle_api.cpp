// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_core.h).  PARITY UNPINNED.
// C-ABI (include/rpbmd.h) in front of the CPU restatement, so tests drive the oracle and the
// CUDA library through the same binding.
#include "oracle_md.h"
#include <algorithm>
#include <cstdio>

using namespace orc;
struct rpb_ctx : public orc::Ctx {};

extern "C" {

const char* rpb_backend(void) { return "oracle-cpu"; }
const char* rpb_last_error(const rpb_ctx* c) { return c ? c->err.c_str() : "null context"; }

int rpb_create(rpb_ctx** out, const rpb_config* cfg) {
  if (!out || !cfg) return RPB_ERR_ARG;
  rpb_ctx* c = new rpb_ctx();
  c->cfg = *cfg;
  *out = c;
  if (cfg->spline_order != 6) { c->err = "only spline_order=6 is self-consistent in the reference (pme.f90:247 divides by 6.D0)"; return RPB_ERR_UNSUPPORTED; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      if (i != j && std::fabs(cfg->box[i + 3 * j]) > 10e-6) { c->err = "code has been modified to assume orthorhombic box"; return RPB_ERR_UNSUPPORTED; }
  if (cfg->evb_max_chain > MAXC || cfg->evb_max_states > MAXS) { c->err = "evb limits exceed compiled maxima"; return RPB_ERR_ARG; }
  c->n_threads = std::max(1, cfg->n_threads);
  c->sys.n_mole = cfg->n_mole;
  c->sys.total_atoms = cfg->n_atoms;
  for (int i = 0; i < 3; i++) { c->sys.box[i] = cfg->box[i + 3 * i]; c->sys.inv_box[i] = 1.0 / c->sys.box[i]; }
  c->atoms.resize(cfg->n_atoms);
  c->mol.resize(cfg->n_mole);
  return 0;
}

void rpb_destroy(rpb_ctx* c) { delete c; }

int rpb_set_tables(rpb_ctx* c, const double* B6, const double* B5, const double* erfc_t, const double* scale_t,
                   const double* tt, const double* dtt, const double* CB) {
  const int K = c->cfg.pme_grid;
  c->B6.assign(B6, B6 + c->cfg.spline_grid);
  c->B5.assign(B5, B5 + c->cfg.spline_grid);
  c->erfc_t.assign(erfc_t, erfc_t + c->cfg.erfc_grid + 1);
  c->scale_t.assign(scale_t, scale_t + c->cfg.erfc_grid + 1);
  c->tt.assign(tt, tt + 4 * c->cfg.tt_grid);
  c->dtt.assign(dtt, dtt + 4 * c->cfg.tt_grid);
  c->CB.assign(CB, CB + (size_t)K * K * K);
  c->have_tables = true;
  return 0;
}

int rpb_set_forcefield(rpb_ctx* c, const double* vdw_parameter, const int* vdw_type, const double* vdw_parameter_14,
                       const double* atype_chg, const int* atype_freeze, const int* bond_type, const double* bond_parameter,
                       const int* angle_type, const double* angle_parameter, const int* dihedral_type,
                       const double* dihedral_parameter) {
  const size_t T2 = MAXT * MAXT, T3 = T2 * MAXT, T4 = T3 * MAXT;
  c->vdw_param.assign(vdw_parameter, vdw_parameter + T2 * 6);
  c->vdw_param14.assign(vdw_parameter_14, vdw_parameter_14 + T2 * 6);
  c->vdw_type.assign(vdw_type, vdw_type + T2);
  for (int i = 0; i < MAXT; i++) { c->atype_chg[i] = atype_chg[i]; c->atype_freeze[i] = atype_freeze[i]; }
  c->bond_type.assign(bond_type, bond_type + T2);
  c->bond_param.assign(bond_parameter, bond_parameter + T2 * 3);
  c->angle_type.assign(angle_type, angle_type + T3);
  c->angle_param.assign(angle_parameter, angle_parameter + T3 * 2);
  c->dihedral_type.assign(dihedral_type, dihedral_type + T4);
  c->dihedral_param.assign(dihedral_parameter, dihedral_parameter + T4 * 6);
  c->have_ff = true;
  return 0;
}

int rpb_set_molecule_types(rpb_ctx* c, const int* n_atom, const int* atom_type, const int* n_bond, const int* bonds,
                           const int* n_angle, const int* angles, const int* n_dihedral, const int* dihedrals,
                           const int* pair_exclusions, const int* evb_reactive_protons, const int* evb_reactive_basic_atoms) {
  int ob = 0, oa = 0, od = 0;
  for (int t = 0; t < c->cfg.n_mole_type; t++) {
    MoleculeType& M = c->mt[t];
    M.n_atom = n_atom[t];
    if (M.n_atom > MA) { c->err = "molecule type larger than RPB_MAX_MOLE_ATOMS"; return RPB_ERR_ARG; }
    for (int a = 0; a < MA; a++) {
      M.atom_type[a] = atom_type[t * MA + a] - 1;
      M.reactive_proton[a] = evb_reactive_protons ? evb_reactive_protons[t * MA + a] : 0;
      M.reactive_basic[a] = evb_reactive_basic_atoms ? evb_reactive_basic_atoms[t * MA + a] : 0;
      for (int b = 0; b < MA; b++) M.pair_excl[a][b] = pair_exclusions[t * MA * MA + a + MA * b];
    }
    M.bonds.clear(); M.angles.clear(); M.dihedrals.clear();
    for (int k = 0; k < 2 * n_bond[t]; k++) M.bonds.push_back(bonds[2 * ob + k] - 1);
    for (int k = 0; k < 3 * n_angle[t]; k++) M.angles.push_back(angles[3 * oa + k] - 1);
    for (int k = 0; k < 4 * n_dihedral[t]; k++) M.dihedrals.push_back(dihedrals[4 * od + k] - 1);
    ob += n_bond[t]; oa += n_angle[t]; od += n_dihedral[t];
  }
  c->have_mt = true;
  return 0;
}

int rpb_set_evb(rpb_ctx* c, const int* da_i, const double* da_p, const int* pa_i, const double* pa_p, const int* dc_i,
                const double* dc_p, const int* dc_t, const double* ex_a, const double* ex_p, const int* acid,
                const int* basic, const int* conj_pairs, const int* conj_atom, const double* ref_e, const int* proton_index,
                const int* heavy_acid_index) {
  for (int i = 0; i < MAXI; i++) {
    for (int j = 0; j < 3; j++) { c->da_int[i][j] = da_i[i + MAXI * j] - 1; c->dc_int[i][j] = dc_i[i + MAXI * j] - 1; }
    for (int j = 0; j < 2; j++) c->pa_int[i][j] = pa_i[i + MAXI * j] - 1;
    for (int j = 0; j < 6; j++) c->da_par[i][j] = da_p[i + MAXI * j];
    for (int j = 0; j < 5; j++) c->pa_par[i][j] = pa_p[i + MAXI * j];
    for (int j = 0; j < 10; j++) c->dc_par[i][j] = dc_p[i + MAXI * j];
    c->dc_type[i] = dc_t[i];
  }
  for (int i = 0; i < MAXT; i++) { c->exch_atomic[i] = ex_a[i]; c->conj_atom[i] = conj_atom[i] - 1; }
  for (int i = 0; i < MAXM; i++) {
    for (int j = 0; j < MAXM; j++) c->exch_proton[i][j] = ex_p[i + MAXM * j];
    c->acid_mol[i] = acid[i]; c->basic_mol[i] = basic[i]; c->conj_pairs[i] = conj_pairs[i] - 1;
    c->ref_energy[i] = ref_e[i]; c->proton_index[i] = proton_index[i] - 1; c->heavy_acid_index[i] = heavy_acid_index[i] - 1;
  }
  c->have_evb = true;
  return 0;
}

int rpb_upload_state(rpb_ctx* c, const double* xyz, const double* velocity, const double* mass, const double* charge,
                     const int* atom_type_index, const int* mol_first_atom, const int* mol_n_atom, const int* mol_type,
                     int hydronium_mol) {
  const int N = c->sys.total_atoms, M = c->sys.n_mole;
  c->atoms.xyz.assign(xyz, xyz + 3 * N);
  c->atoms.vel.assign(velocity, velocity + 3 * N);
  c->atoms.force.assign(3 * N, 0.0);
  c->atoms.mass.assign(mass, mass + N);
  c->atoms.charge.assign(charge, charge + N);
  for (int i = 0; i < N; i++) c->atoms.type[i] = atom_type_index[i] - 1;
  int expect = 0;
  for (int m = 0; m < M; m++) {
    c->mol[m].first = mol_first_atom[m] - 1; c->mol[m].n_atom = mol_n_atom[m]; c->mol[m].type = mol_type[m] - 1;
    if (c->mol[m].first != expect) { c->err = "molecules must be contiguous ascending atom ranges"; return RPB_ERR_ARG; }
    expect += mol_n_atom[m];
  }
  if (expect != N) { c->err = "molecule table does not cover all atoms"; return RPB_ERR_ARG; }
  c->hydronium_mol = hydronium_mol - 1;
  c->have_state = true;
  return 0;
}

int rpb_initialize(rpb_ctx* c) {
  if (!(c->have_tables && c->have_ff && c->have_mt && c->have_state)) { c->err = "tables/forcefield/molecule types/state must be set first"; return RPB_ERR_STATE; }
  update_r_com(*c);
  shift_molecules_into_box(*c);
  c->neighbor_list.assign(verlet_capacity(*c), 0);
  int rc = construct_verlet_list(*c);
  if (rc) return rc;
  int junk;
  update_verlet_displacements(*c, &junk, true);
  c->flag_verlet_list = 0;
  return 0;
}

int rpb_force_energy(rpb_ctx* c, int ms_evb) {
  if (ms_evb) {
    if (!c->have_evb) { c->err = "rpb_set_evb not called"; return RPB_ERR_STATE; }
    if (c->cfg.world_size > 1) { c->err = "world_size>1: use the phase calls"; return RPB_ERR_STATE; }
    return ms_evb_calculate_total_force_energy(*c);
  }
  return calculate_total_force_energy(*c, false);
}

int rpb_step_begin(rpb_ctx* c) { md_step_begin(*c); return 0; }
int rpb_step_end(rpb_ctx* c) { return md_step_end(*c); }
int rpb_evb_phase_build(rpb_ctx* c) { return evb_phase_build(*c); }
int rpb_evb_phase_mix(rpb_ctx* c) { return evb_phase_mix(*c, nullptr, nullptr); }
int rpb_evb_phase_commit(rpb_ctx* c) { return evb_phase_commit(*c); }
int rpb_evb_exchange_h(rpb_ctx* c, void** ptr, int* n) { *ptr = c->xh.data(); *n = (int)c->xh.size(); return 0; }
int rpb_evb_exchange_f(rpb_ctx* c, void** ptr, int* n) { *ptr = c->xf.data(); *n = (int)c->xf.size(); return 0; }

int rpb_step(rpb_ctx* c, int n_steps, int ms_evb);
int rpb_ensemble_step(rpb_ctx** replicas, int n_replicas, int n_steps, int ms_evb) {   // serial on the CPU
  if (!replicas || n_replicas < 1) return RPB_ERR_ARG;
  for (int r = 0; r < n_replicas; r++) { int rc = rpb_step(replicas[r], n_steps, ms_evb); if (rc) return rc; }
  return 0;
}

// peer-memory exchange is a device feature of the CUDA library
int rpb_peer_export(rpb_ctx* c, void*) { c->err = "peer-memory exchange: CUDA library only"; return RPB_ERR_UNSUPPORTED; }
int rpb_peer_import(rpb_ctx* c, const void*, int) { c->err = "peer-memory exchange: CUDA library only"; return RPB_ERR_UNSUPPORTED; }
int rpb_peer_attach_local(rpb_ctx**, int) { return RPB_ERR_UNSUPPORTED; }
int rpb_peer_enabled(rpb_ctx*) { return 0; }

int rpb_step(rpb_ctx* c, int n_steps, int ms_evb) {
  for (int s = 0; s < n_steps; s++) {
    md_step_begin(*c);
    int rc = rpb_force_energy(c, ms_evb);
    if (rc) return rc;
    rc = md_step_end(*c);
    if (rc) return rc;
  }
  return 0;
}

int rpb_get_energies(rpb_ctx* c, rpb_energies* e) {
  e->potential_energy = c->sys.potential_energy;
  e->kinetic_energy = calculate_kinetic_energy(*c);
  e->E_elec = c->sys.E_elec; e->E_vdw = c->sys.E_vdw; e->E_bond = c->sys.E_bond; e->E_angle = c->sys.E_angle;
  e->E_dihedral = c->sys.E_dihedral; e->E_recip = c->E_recip;
  return 0;
}

int rpb_download_state(rpb_ctx* c, double* xyz, double* velocity, double* force, double* mass, double* charge,
                       int* atom_type_index, int* mol_first_atom, int* mol_n_atom, int* mol_type, int* hydronium_mol) {
  const int N = c->sys.total_atoms, M = c->sys.n_mole;
  if (xyz) std::copy(c->atoms.xyz.begin(), c->atoms.xyz.end(), xyz);
  if (velocity) std::copy(c->atoms.vel.begin(), c->atoms.vel.end(), velocity);
  if (force) std::copy(c->atoms.force.begin(), c->atoms.force.end(), force);
  if (mass) std::copy(c->atoms.mass.begin(), c->atoms.mass.end(), mass);
  if (charge) std::copy(c->atoms.charge.begin(), c->atoms.charge.end(), charge);
  if (atom_type_index) for (int i = 0; i < N; i++) atom_type_index[i] = c->atoms.type[i] + 1;
  for (int m = 0; m < M; m++) {
    if (mol_first_atom) mol_first_atom[m] = c->mol[m].first + 1;
    if (mol_n_atom) mol_n_atom[m] = c->mol[m].n_atom;
    if (mol_type) mol_type[m] = c->mol[m].type + 1;
  }
  if (hydronium_mol) *hydronium_mol = c->hydronium_mol + 1;
  return 0;
}

int rpb_get_r_com(rpb_ctx* c, double* r_com) {
  for (int m = 0; m < c->sys.n_mole; m++) for (int k = 0; k < 3; k++) r_com[3 * m + k] = c->mol[m].r_com[k];
  return 0;
}

int rpb_get_neighbor_list(rpb_ctx* c, int* verlet_point, int* neighbor_list, int capacity, int* n_pairs, int* flag) {
  const int N = c->sys.total_atoms;
  int np = c->verlet_point[N] - 1;
  if (n_pairs) *n_pairs = np;
  if (flag) *flag = c->flag_verlet_list;
  if (verlet_point) std::copy(c->verlet_point.begin(), c->verlet_point.end(), verlet_point);
  if (neighbor_list) {
    if (capacity < np) { c->err = "neighbor_list buffer too small"; return RPB_ERR_ARG; }
    std::copy(c->neighbor_list.begin(), c->neighbor_list.begin() + np, neighbor_list);
  }
  return 0;
}

int rpb_debug_tile_pairs(rpb_ctx* c, int* pair_i, int* pair_j, long long capacity, long long* n_pairs, long long* n_tiles) {
  const int N = c->sys.total_atoms;
  const long long np = c->verlet_point[N] - 1;
  if (n_pairs) *n_pairs = np;
  if (n_tiles) *n_tiles = np;
  if (pair_i && pair_j) {
    if (capacity < np) { c->err = "pair buffer too small"; return RPB_ERR_ARG; }
    std::vector<std::pair<int, int>> pr;
    for (int i = 0; i < N; i++)
      for (int k = c->verlet_point[i] - 1; k < c->verlet_point[i + 1] - 1; k++) pr.emplace_back(i + 1, c->neighbor_list[k]);
    std::sort(pr.begin(), pr.end());
    for (size_t k = 0; k < pr.size(); k++) { pair_i[k] = pr[k].first; pair_j[k] = pr[k].second; }
  }
  return 0;
}

int rpb_get_pme(rpb_ctx* c, int state, double* Q_grid, double* theta, double* force_recip) {
  const std::vector<double>* Q = &c->Q_grid; const std::vector<double>* T = &c->theta_conv_Q;
  if (state > 1) {
    if (state > c->diabat_index || c->Q_grid_diabats[state - 1].empty()) { c->err = "diabat grid not available"; return RPB_ERR_ARG; }
    Q = &c->Q_grid_diabats[state - 1]; T = &c->theta_diabats[state - 1];
  }
  if (Q_grid) std::copy(Q->begin(), Q->end(), Q_grid);
  if (theta) std::copy(T->begin(), T->end(), theta);
  if (force_recip) std::copy(c->force_recip.begin(), c->force_recip.end(), force_recip);
  return 0;
}

int rpb_get_evb(rpb_ctx* c, int* n_states, double* hamiltonian, double* eigenvector, int* proton_log, int* coupling_matrix,
                int* principal_diabat, int* new_hydronium_mol, double* adiabatic_potential) {
  const int S = c->diabat_index;
  if (n_states) *n_states = S;
  if (hamiltonian) for (int i = 0; i < MAXS; i++) for (int j = 0; j < MAXS; j++) hamiltonian[i + MAXS * j] = c->evb_hamiltonian[i][j];
  if (eigenvector) for (int s = 0; s < S; s++) eigenvector[s] = c->ground_state_eigenvector.empty() ? 0.0 : c->ground_state_eigenvector[s];
  if (proton_log)
    for (int s = 0; s < MAXS; s++) for (int h = 0; h < MAXC; h++) for (int f = 0; f < 5; f++) {
      int v = c->proton_log[s][h][f];
      proton_log[s + MAXS * h + MAXS * MAXC * f] = v < 0 ? -1 : v + 1;
    }
  if (coupling_matrix) for (int s = 0; s < MAXS; s++) coupling_matrix[s] = c->coupling_matrix[s] < 0 ? -1 : c->coupling_matrix[s] + 1;
  if (principal_diabat) *principal_diabat = c->principle_diabat + 1;
  if (new_hydronium_mol) *new_hydronium_mol = c->new_hydronium + 1;
  if (adiabatic_potential) *adiabatic_potential = c->adiabatic_potential;
  return 0;
}

int rpb_debug_mix_forces(rpb_ctx* c, const double* coeff, double* force) { return evb_phase_mix(*c, coeff, force); }

int rpb_get_launch_counts(rpb_ctx*, long long* own, long long* fft) { if (own) *own = 0; if (fft) *fft = 0; return 0; }
int rpb_timers_enable(rpb_ctx*, int) { return 0; }
int rpb_timers_reset(rpb_ctx*) { return 0; }
int rpb_timer_count(void) { return 0; }
const char* rpb_timer_name(int) { return ""; }
int rpb_timers_get(rpb_ctx*, double*, long long*) { return 0; }
void* rpb_get_stream(rpb_ctx*) { return nullptr; }
int rpb_measure_fp64_peak(rpb_ctx*, double* t) { if (t) *t = 0.0; return 0; }

}  // extern "C"
