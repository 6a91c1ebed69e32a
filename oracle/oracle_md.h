// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_core.h).  PARITY UNPINNED.
#pragma once
#include "oracle_core.h"

namespace orc {

// pairwise_neighbor_data_type (pair_int_real_space.f90:16-26), one growable list
struct PairList {
  std::vector<int> idx;
  std::vector<double> dr, dr2, qq, par, f;
  void clear();
  void push(int j, const double d[3], double d2, double q, const double* p6);
};

void pbc_shift(double out[3], const double ri[3], const double rj[3], const SystemData& s);
void pbc_dr(double out[3], const double ri[3], const double rj[3], const double shift[3]);
void pos_com(double out[3], const double* xyz, const double* mass, int n_atom);
void update_r_com(Ctx& c);
void make_molecule_whole(int n_atom, double* xyz, const SystemData& s);
void shift_molecules_into_box(Ctx& c);
int verlet_capacity(const Ctx& c);
int construct_verlet_list(Ctx& c);
void update_verlet_displacements(Ctx& c, int* flag, bool initialize);

double pairwise_real_space_ewald(const Ctx& c, PairList& p);
double pairwise_real_space_LJ(PairList& p);
double pairwise_real_space_sapt(const Ctx& c, PairList& p);
double intra_pme_exclusion(const Ctx& c, PairList& p);
void intra_molecular_pairwise_energy_force(const Ctx& c, double* force_local, double* E_elec, double* E_vdw,
                                           const double* xyz, const double* charge, const int* type, int i_mole_type,
                                           int n_atom);
void real_space_energy_force(Ctx& c);

void reciprocal_lattice(double kk[3][3], const SystemData& s);
void create_scaled_direct_coordinates(double* xyz_scale, const double* xyz, int n_atom, const double kk[3][3], int K);
void spread_atoms(const Ctx& c, double* Q, const double* chg, const double* u3, int n_atom, int op);
void derivative_grid_Q(const Ctx& c, double force[3], const double* FQ, const double* chg, const double* u3, int i_atom,
                       const double kk[3][3], double* dQ_dr_store, int* dQ_dr_index_store);
void fft3d(cplx* a, int K, int sign);
double pme_convolve(const Ctx& c, const double* Q, double* theta);
void pme_reciprocal_space_energy_force(Ctx& c, bool store_dQ_dr);

int intra_molecular_bond_energy_force(const Ctx& c, double* E, const double* xyz, const int* type, double* force, int mtype);
int intra_molecular_angle_energy_force(const Ctx& c, double* E, const double* xyz, const int* type, double* force, int mtype);
int intra_molecular_dihedral_energy_force(const Ctx& c, double* E, const double* xyz, const int* type, double* force, int mtype);
int intra_molecular_energy_force(Ctx& c);

int calculate_total_force_energy(Ctx& c, bool ms_evb);
double calculate_kinetic_energy(const Ctx& c);
void md_step_begin(Ctx& c);
int md_step_end(Ctx& c);

// ms_evb.f90 (oracle_evb.cpp)
int evb_phase_build(Ctx& c);
int evb_phase_mix(Ctx& c, const double* coeff_override, double* force_out);
int evb_phase_commit(Ctx& c);
int ms_evb_calculate_total_force_energy(Ctx& c);

}  // namespace orc
