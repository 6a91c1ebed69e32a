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

// Work items of the real-space / repulsion / bonded deltas (ms_evb.f90:1460-1552): item 0 is the principal diabat (EVB
// repulsion + reference energy, ms_evb.f90:418-426); items 2s-1 and 2s are the LAST hop of diabat s >= 1 evaluated in its
// donor topology (snapshot level n_hops-1, sign -1) and in its acceptor topology (level n_hops, sign +1).  Earlier hops of
// a chain are the ancestors' last hops (same images, same background: hop-tree de-duplication).
#define RPB_MAX_ITEMS (2 * RPB_MAXS)

// Per-step work lists, written on the DEVICE by the enumeration kernel (the host never sees the diabat set inside a step)
#define RPB_RA_MOLS (RPB_MAXS + 1)            // distinct chain molecules of a step
#define RPB_RA_MAXPAIR (8 * RPB_MAXS + 8)     // ordered pairs of chain molecules that share a diabat
#define RPB_CAND_SLOTS 1024                   // distinct chain atoms over all diabats of a step
struct EvbPlan {
  int n_cmol;                         // distinct chain molecules over ALL diabats; cmol[0] = hydronium
  int n_pair;                         // ordered pairs cmol index i * RPB_RA_MOLS + j sharing a diabat (incl. i == j)
  int n_own;                          // diabats s >= 1 owned by this rank, ascending, in state_list
  int n_uniq;                         // atoms of the chain molecules of the owned diabats (+ hydronium), in uniq_atom
  int n_clear;                        // accumulators of diabats [0, n_clear) were cleared ahead of this step's enumeration
  int hop;                            // 1: the solver selected a new hydronium molecule (commit kernels act)
  int pad[2];
  int cmol[RPB_RA_MOLS];
  int molpair[RPB_RA_MAXPAIR];
  int state_list[RPB_MAXS];
  int uniq_atom[RPB_CAND_SLOTS];
};

struct EvbDev {
  int* n_states;              // device scalar S
  int* proton_log;            // [MAXS][MAXC][5] 0-based, -1 end
  int* parent;                // evb_diabat_coupling_matrix [MAXS]
  int* n_hops;                // [MAXS]
  Snapshot* snap;             // [MAXS][MAXC+1]
  EvbPlan* plan;              // this step's work lists
  int* mol_slot;              // [M] molecule -> index in plan->cmol (-1: not a chain molecule)
  double* item_energy;        // [RPB_MAX_ITEMS] : E_ref + E_intra + E_real + E_rep of the item's topology
  double* dF;                 // [MAXS][3N]: slot 0 principal-diabat force; slot s: force delta of the LAST hop of diabat s
  double* Foff;               // [MAXS][3N] off-diagonal coupling forces
  double* vex;                // [MAXS]
  double* e_recip;            // [4] E_rec of the grids convolved this step (slot 0: principal diabat, slot 1: Hellmann-Feynman averaged grid)
  double* rcp_dE;             // [MAXS] E_rec(s) - E_rec(1) from the delta algebra (kernels_evb.cu)
  double* h_diag;             // exchange buffer [3*MAXS]: (H_11, dE_s of the last hop) | H_parent(s),s | E_rec(s)-E_rec(1)
  double* h_full;             // [2*MAXS] assembled H_ss | H_parent(s),s (after the exchange)
  double* f_mix;              // exchange buffer [3N]
  double* evec;               // ground-state eigenvector [MAXS]
  double* coef2;              // [3*MAXS] c_s^2 | 2 c_parent c_s | sum of c_t^2 over the DFS subtree of s
  double* e_ground;           // adiabatic potential
  double* status_copy;        // [4 + E_NSLOT] error flags and energy slots, copied by the solver into its read-back block
  int* result;                // [0] principal diabat (0-based) [1] new hydronium molecule (0-based) [2] jacobi status
  double* coupling_geo;       // [MAXS][16] A, Vconst, dA[3][3], atoms...
  double* tree_mu;            // [0] lowest eigenvalue (relative to H_11) of the previous tree solve, [1] 1.0 if valid
  double* jac_v; int* jac_sig; // eigenvector matrix and diabat-set signature (S, hop logs) of the previous Jacobi solve
};

struct EvbHost {   // what the host keeps of the LAST evaluated step (read back once per rpb_step / rpb_force_energy call)
  int* pinned = nullptr;      // read-back area: enumeration block, then the solver block
  int n_states = 0;
  int s_hint = 16;            // grid sizing only: a recent number of diabats (kernels loop over the device-side counts)
  int n_hops[RPB_MAXS];
  int proton_log[RPB_MAXS][RPB_MAXC][5];
  int parent[RPB_MAXS];
  int principal_diabat = 0, new_hydronium = -1;
  double adiabatic_potential = 0.0;
  double hamiltonian[RPB_MAXS][RPB_MAXS];
  double evec[RPB_MAXS];
  bool built = false;
};
