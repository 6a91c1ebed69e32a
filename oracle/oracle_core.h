// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path of
// jmcdaniel43/Reactive_PB_NN_MD, evaluated in source order, fp64, no FMA contraction
// (build with -ffp-contract=off), REAL*4 wherever the Fortran source is REAL*4.
// Nothing under oracle/ is linked into or called by the product library (librpbmd.so);
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.
//
// PARITY UNPINNED: the reference ships no golden vectors, no tests and cannot be compiled
// in this image (no Fortran compiler, no MKL).  This restatement is pinned
// only by self-validation (finite differences, Ewald invariances, Madelung constant,
// hand-worked enumeration cases) in tests/.
//
// Data model: plain structs mirroring the derived types of src/glob_v.f90:125-310.
#pragma once
#include <vector>
#include <string>
#include <complex>
#include <cmath>
#include <cstring>
#include "../include/rpbmd.h"

namespace orc {

typedef std::complex<double> cplx;
constexpr int MAXT = RPB_MAX_N_ATOM_TYPE;
constexpr int MAXM = RPB_MAX_N_MOLE_TYPE;
constexpr int MAXI = RPB_MAX_INTERACTION_TYPE;
constexpr int MAXS = RPB_EVB_MAX_STATES;
constexpr int MAXC = RPB_EVB_MAX_CHAIN;
constexpr int MAXN = RPB_EVB_MAX_NEIGHBORS;
constexpr int MA = RPB_MAX_MOLE_ATOMS;

// molecule_data_type (glob_v.f90:168-176).  atom_index is always the contiguous ascending
// range first..first+n_atom-1 (general_routines.f90:670-671, ms_evb.f90:2805-2836), so it is
// stored as (first, n_atom); `first` is 0-based.
struct Molecule {
  int first, n_atom, type;  // type 0-based
  double r_com[3];
};

// atom_data_type (glob_v.f90:157-165)
struct AtomData {
  std::vector<double> xyz, vel, force, mass, charge;
  std::vector<int> type;  // 0-based
  void resize(int n) {
    xyz.assign(3 * n, 0.0); vel.assign(3 * n, 0.0); force.assign(3 * n, 0.0);
    mass.assign(n, 0.0); charge.assign(n, 0.0); type.assign(n, 0);
  }
};

struct MoleculeType {
  int n_atom = 0;
  int atom_type[MA];
  std::vector<int> bonds, angles, dihedrals;  // 0-based atom indices, 2/3/4 per entry
  int pair_excl[MA][MA];
  int reactive_proton[MA], reactive_basic[MA];
};

struct SystemData {  // system_data_type (glob_v.f90:125-143)
  int n_mole, total_atoms;
  double box[3], inv_box[3];  // orthorhombic: box(i,i); xyz_to_box_transform(i,i)=1/box(i,i) (gaussj)
  double potential_energy, kinetic_energy, E_elec, E_vdw, E_bond, E_angle, E_dihedral;
};

struct Ctx {
  rpb_config cfg;
  std::string err;
  bool have_tables = false, have_ff = false, have_mt = false, have_evb = false, have_state = false;
  int n_threads = 1;

  SystemData sys;
  std::vector<Molecule> mol;
  AtomData atoms;
  MoleculeType mt[MAXM];
  int hydronium_mol = -1;  // 0-based; hydronium_molecule_index(1)

  // force field (Fortran layouts kept; index helpers below)
  std::vector<double> vdw_param, vdw_param14, bond_param, angle_param, dihedral_param;
  std::vector<int> vdw_type, bond_type, angle_type, dihedral_type;
  double atype_chg[MAXT];
  int atype_freeze[MAXT];

  // tables
  std::vector<double> B6, B5, erfc_t, scale_t, tt, dtt, CB;

  // verlet list (verlet_list_data_type glob_v.f90:236-251)
  std::vector<int> verlet_point, neighbor_list;  // 1-based values as in the reference
  std::vector<double> verlet_xyz_store, verlet_disp_store;
  int flag_verlet_list = 0;

  // PME (PME_data_type glob_v.f90:256-275)
  std::vector<double> Q_grid, theta_conv_Q, force_recip, dQ_dr;
  std::vector<int> dQ_dr_index;
  double E_recip = 0.0;

  // MS-EVB tables (glob_v.f90:77-120)
  int da_int[MAXI][3]; double da_par[MAXI][6];
  int pa_int[MAXI][2]; double pa_par[MAXI][5];
  int dc_int[MAXI][3]; double dc_par[MAXI][10]; int dc_type[MAXI];
  double exch_atomic[MAXT], exch_proton[MAXM][MAXM];
  int acid_mol[MAXM], basic_mol[MAXM], conj_pairs[MAXM], conj_atom[MAXT];
  double ref_energy[MAXM]; int proton_index[MAXM], heavy_acid_index[MAXM];

  // MS-EVB state (module variables ms_evb.f90:21-42)
  double evb_hamiltonian[MAXS][MAXS];
  int evb_forces_lookup_index[MAXS][MAXS];
  std::vector<std::vector<double>> evb_forces_store;
  int proton_log[MAXS][MAXC][5];  // 0-based contents, -1 = end
  int coupling_matrix[MAXS];
  std::vector<std::vector<double>> Q_grid_diabats, theta_diabats;
  int diabat_index = 0;  // number of diabats S
  int store_index = 0;
  std::vector<double> ground_state_eigenvector;
  double adiabatic_potential = 0.0;
  int principle_diabat = 0, new_hydronium = -1;  // 0-based
  // sharded exchange
  std::vector<double> xh, xf;
};

inline int vdw_idx(int ti, int tj, int k) { return ti + MAXT * tj + MAXT * MAXT * k; }

}  // namespace orc
