// MS-EVB working set on the device: diabat chains as PATCHES over the principal-diabat arrays.
//
// The reference deep-copies all atom/molecule data per diabat and physically shifts every per-atom
// array for each proton hop (ms_evb.f90:770-798, 2677-2840).  Here a diabat is described by
// "snapshots": for diabat s and topology level t (t = number of hops applied, 0..n_s) the images of
// the <= RPB_CHAIN_MOLS molecules of its hop chain -- their atoms (principal indices), current atom
// types, charges, (made-whole) positions, masses, centre of mass and molecule type.  Every other atom
// of the system is read from the principal arrays.  Forces always stay in principal atom order.
#pragma once
#include "rpb_dev.cuh"

struct MolImage {
  int mol;                    // principal molecule index (0-based)
  int n_atom;
  int mtype;                  // molecule type at this level
  int atom[RPB_MA];           // principal (global) atom index of each image atom, in diabat order
  int ratom[RPB_MA];          // principal index the REFERENCE attributes this position's force to: its back-mapping
                              // (map_diabat_force_to_principle_recursive, ms_evb.f90:2608-2656) undoes the proton transfers
                              // but not reorder_molecule_data_structures -- tracks the transfers, ignores the re-ordering
  int type[RPB_MA];
  double q[RPB_MA];
  double mass[RPB_MA];
  double x[RPB_MA][3];
  double r_com[3];
};

struct Snapshot {
  int n_mol;                  // distinct chain molecules of this diabat (all levels list the same set)
  int hydronium;              // slot of the hydronium molecule at this level
  MolImage m[RPB_CHAIN_MOLS];
};

// one "item" = (diabat, hop, side): side 0 = donor topology (level hop, sign -1), side 1 = acceptor
// topology (level hop+1, sign +1)       ms_evb.f90:1460-1552
struct EvbItem {
  int state;                  // 0-based diabat
  int level;                  // snapshot level
  int donor_slot, acceptor_slot;
  double sign;
  int real;                   // 1: this hop is the LAST hop of its diabat (or the principal item): real-space/chain terms are evaluated
  int pad;
};

#define RPB_MAX_ITEMS (2 * RPB_MAXS * RPB_MAXC)

struct EvbDev {
  int* n_states;              // device scalar S
  int* proton_log;            // [MAXS][MAXC][5] 0-based, -1 end
  int* parent;                // evb_diabat_coupling_matrix [MAXS]
  int* n_hops;                // [MAXS]
  Snapshot* snap;             // [MAXS][MAXC+1]
  EvbItem* items;             // [RPB_MAX_ITEMS]
  int* n_items;
  double* item_energy;        // [RPB_MAX_ITEMS] : E_ref + E_intra + E_real + E_rep of the item's topology
  int* real_list; int* n_real; // item indices with real==1
  double* dF;                 // [MAXS][3N]: slot 0 principal-diabat force; slot s: force delta of the LAST hop of diabat s
  double* corr_f; int* corr_atom; // [MAXS][CM*MA][3], [MAXS][CM*MA]: reciprocal-space corrections of the chain atoms of diabat s
  double* Foff;               // [MAXS][3N] off-diagonal coupling forces
  double* vex;                // [MAXS]
  double* e_recip;            // [MAXS] E_rec of each diabat grid
  double* rcp_dE;             // [MAXS] E_rec(s) - E_rec(1) from the delta algebra (kernels_evb.cu); unused on the per-diabat grid path
  double* h_diag;             // exchange buffer [3*MAXS]: (H_11, dE_s of the last hop) | H_parent(s),s | E_rec(s)-E_rec(1)
  double* h_full;             // [2*MAXS] assembled H_ss | H_parent(s),s (after the exchange)
  double* f_mix;              // exchange buffer [3N]
  double* evec;               // ground-state eigenvector [MAXS]
  double* coef2;              // [3*MAXS] c_s^2 | 2 c_parent c_s | sum of c_t^2 over the DFS subtree of s
  double* e_ground;           // adiabatic potential
  double* status_copy;        // [4 + E_NSLOT] error flags and energy slots, copied by the solver into its read-back block
  int* result;                // [0] principal diabat (0-based) [1] new hydronium molecule (0-based) [2] jacobi status
  double* coupling_geo;       // [MAXS][16] A, Vconst, dA[3][3], atoms...
  double* theta_mix;          // K^3
  double* tree_mu;            // [0] lowest eigenvalue (relative to H_11) of the previous tree solve, [1] 1.0 if valid
  double* jac_v; int* jac_sig; // eigenvector matrix and diabat-set signature (S, hop logs) of the previous Jacobi solve
};

struct EvbHost {
  int* pinned = nullptr;      // read-back area: [0]=S, [1]=n_items, [2..] n_hops, proton_log, parent, result
  int n_states = 0, n_items = 0;
  int n_states_prev = 0;      // S of the previous build: bounds what the early clear kernel covered
  int n_hops[RPB_MAXS];
  int proton_log[RPB_MAXS][RPB_MAXC][5];
  int parent[RPB_MAXS];
  int principal_diabat = 0, new_hydronium = -1;
  double adiabatic_potential = 0.0;
  double hamiltonian[RPB_MAXS][RPB_MAXS];
  double evec[RPB_MAXS];
  bool built = false;
  bool assemble_pending = false;   // evb_build left the Hamiltonian assembly to the solver kernel
  bool overlap_solver = false;     // the branches of evb_build were joined on aux[0] (not on the main stream): evb_mix runs the solver there
};
