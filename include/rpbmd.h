/*
 * rpbmd.h -- C-ABI of the B200-native MS-EVB force path.
 *
 * This is the drop-in boundary for the per-timestep force path of
 * jmcdaniel43/Reactive_PB_NN_MD.  The reference has no FFI layer of its own; the
 * seam is the pair of Fortran module procedures
 *     calculate_total_force_energy          (src/total_energy_forces.f90:19-99)
 *     ms_evb_calculate_total_force_energy   (src/ms_evb.f90:181-235)
 * called from md_integrate_atomic (src/md_integration.f90:495-500) and
 * initialize_energy_force (src/initialize_routines.f90:268-273).  The entry points
 * below are what an ISO_C_BINDING shim in those two routines binds.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are caller-owned HOST memory in the
 *     reference's own Fortran (column-major) layout and 1-based index VALUES;
 *   - every call returns 0 on success, a negative rpb_status on failure;
 *     rpb_last_error() gives the message the Fortran shim prints before `stop`
 *     (the reference's only error mechanism, e.g. src/md_integration.f90:523-526);
 *   - one handle == one CUDA device + one stream set; calls on one handle are not
 *     thread-safe (the reference calls the path from its single main thread);
 *   - two libraries export this identical ABI: librpbmd.so (CUDA, the product) and
 *     oracle/librpbmd_oracle.so (CPU restatement, test infrastructure only).
 *
 * Fixed leading dimensions follow src/glob_v.f90:34,56-72:
 *   MAX_N_ATOM_TYPE=25, MAX_N_MOLE_TYPE=10, max_interaction_type=15,
 *   evb_max_states=80, evb_max_chain=3, evb_max_neighbors=10.
 */
#ifndef RPBMD_H
#define RPBMD_H

#ifdef __cplusplus
extern "C" {
#endif

#define RPB_MAX_N_ATOM_TYPE 25      /* glob_v.f90:34  */
#define RPB_MAX_N_MOLE_TYPE 10      /* glob_v.f90:34  */
#define RPB_MAX_INTERACTION_TYPE 15 /* glob_v.f90:72  */
#define RPB_EVB_MAX_STATES 80       /* glob_v.f90:60  */
#define RPB_EVB_MAX_CHAIN 3         /* glob_v.f90:65  */
#define RPB_EVB_MAX_NEIGHBORS 10    /* glob_v.f90:56  */
#define RPB_MAX_MOLE_ATOMS 8        /* largest molecule type + 1 accepted proton */

typedef enum {
  RPB_OK = 0,
  RPB_ERR_ARG = -1,          /* bad argument / inconsistent sizes                      */
  RPB_ERR_CUDA = -2,         /* CUDA / cuFFT runtime failure                            */
  RPB_ERR_STATE = -3,        /* call order violated (tables / force field not set ...)  */
  RPB_ERR_FORCE = -4,        /* |F_i| > 1e5: "force on atom ... is too big" md_integration.f90:523 */
  RPB_ERR_VERLET = -5,       /* "please increase size of verlet neighbor list" general_routines.f90:1562 */
  RPB_ERR_DIABATS = -6,      /* more than evb_max_states diabats  ms_evb.f90:3107-3121  */
  RPB_ERR_UNSUPPORTED = -7   /* e.g. spline_order != 6, non-orthorhombic box */
} rpb_status;

typedef struct rpb_ctx rpb_ctx;

/* Scalars of system_data_type / PME_data_type / verlet_list_data_type / integrator_data_type
 * (src/glob_v.f90:125-275) and the constants of initialize_constants (glob_v.f90:379-398). */
typedef struct {
  int n_atoms;             /* system_data%total_atoms */
  int n_mole;              /* system_data%n_mole */
  int n_atom_type;         /* n_atom_type */
  int n_mole_type;         /* n_molecule_type */
  int pme_grid;            /* PME_data%pme_grid  (K) */
  int spline_order;        /* PME_data%spline_order; only 6 is self-consistent in the reference (pme.f90:247) */
  int spline_grid;         /* 100000, glob_v.f90:397 */
  int erfc_grid;           /* 100000, glob_v.f90:398 */
  int tt_grid;             /* Tang_Toennies_grid=1000, glob_v.f90:348 */
  int na_nslist, nb_nslist, nc_nslist; /* verlet cell grid, glob_v.f90:245-247 */
  int verlet_capacity;     /* size(neighbor_list); 0 = reference formula general_routines.f90:1231-1239 */
  int device;              /* CUDA device ordinal (ignored by the oracle) */
  int rank, world_size;    /* diabatic-state sharding; 0,1 = single GPU */
  int n_threads;           /* oracle only: OpenMP threads, mirrors n_threads glob_v.f90:365 */
  int evb_max_chain;       /* glob_v.f90:65 (<= RPB_EVB_MAX_CHAIN) */
  int evb_max_states;      /* glob_v.f90:60 (<= RPB_EVB_MAX_STATES) */
  int reserved_i[4];
  double box[9];           /* system_data%box(3,3), column-major; must be orthorhombic (main_ms_evb.f90:62) */
  double alpha_sqrt;       /* PME_data%alpha_sqrt */
  double real_space_cutoff;
  double verlet_cutoff;
  double delta_t;          /* integrator_data%delta_t (ps) */
  double erfc_dx;          /* real_space_cutoff/erfc_grid, initialize_routines.f90:234 */
  double tt_max;           /* Tang_Toennies_max=50 */
  double pi, pi_sqrt;      /* 3.141592654d0, 1.772453851d0 (truncated, glob_v.f90:386-387) */
  double conv_e2A_kJmol;   /* dble(1389.35465_4), glob_v.f90:389 */
  double conv_kJmol_ang2ps2gmol; /* 100d0 */
  double safe_verlet;      /* dble(1.2_4) */
  double verlet_thresh;    /* 1.2d0 */
  double evb_first_solvation_cutoff; /* 5d0, glob_v.f90:54 */
  double evb_reactive_pair_distance; /* 2.5d0, glob_v.f90:55 */
  double ewald_self;       /* PME_data%Ewald_self, frozen at init (initialize_routines.f90:193) */
  double reserved_d[4];
} rpb_config;

/* Energies of system_data_type (glob_v.f90:136-142). */
typedef struct {
  double potential_energy, kinetic_energy;
  double E_elec, E_vdw, E_bond, E_angle, E_dihedral;
  double E_recip;          /* PME_data%E_recip (pme.f90:134) */
} rpb_energies;

const char* rpb_last_error(const rpb_ctx*);
const char* rpb_backend(void);   /* "cuda-sm100a" or "oracle-cpu" */

int  rpb_create(rpb_ctx** out, const rpb_config* cfg);
void rpb_destroy(rpb_ctx*);

/* Look-up tables built by initialize_energy_force (initialize_routines.f90:212-264)
 * and CB_array (pme.f90:537-573).  tt/dtt are Tang_Toennies_table(4,grid) column-major. */
int rpb_set_tables(rpb_ctx*, const double* B6_spline, const double* B5_spline,
                   const double* erfc_table, const double* ewaldscale_table,
                   const double* tt_table, const double* dtt_table, const double* CB);

/* Global atom-type force-field arrays, Fortran shapes of glob_v.f90:322-337:
 *   vdw_parameter(25,25,6) vdw_type(25,25) vdw_parameter_14(25,25,6) atype_chg(25) atype_freeze(25)
 *   bond_type(25,25) bond_parameter(25,25,3) angle_type(25,25,25) angle_parameter(25,25,25,2)
 *   dihedral_type(25,25,25,25) dihedral_parameter(25,25,25,25,6). */
int rpb_set_forcefield(rpb_ctx*, const double* vdw_parameter, const int* vdw_type,
                       const double* vdw_parameter_14, const double* atype_chg, const int* atype_freeze,
                       const int* bond_type, const double* bond_parameter,
                       const int* angle_type, const double* angle_parameter,
                       const int* dihedral_type, const double* dihedral_parameter);

/* molecule_type_data (glob_v.f90:299-317), flattened.  For molecule type t (0-based here,
 * 1-based in mol_type values): n_atom[t]; atom types at atom_type[t*RPB_MAX_MOLE_ATOMS + a];
 * bond/angle/dihedral lists are concatenated over types (counts in n_bond/n_angle/n_dihedral),
 * entries are 1-based atom indices within the molecule; pair_exclusions(a,b) at
 * [t*MA*MA + a + MA*b] with MA=RPB_MAX_MOLE_ATOMS (0 normal, 1 excluded, 2 special 1-4);
 * evb_reactive_protons / evb_reactive_basic_atoms at [t*MA + a] (ms_evb.f90:3225-3297). */
int rpb_set_molecule_types(rpb_ctx*, const int* n_atom, const int* atom_type,
                           const int* n_bond, const int* bonds,
                           const int* n_angle, const int* angles,
                           const int* n_dihedral, const int* dihedrals,
                           const int* pair_exclusions,
                           const int* evb_reactive_protons, const int* evb_reactive_basic_atoms);

/* MS-EVB parameter tables, Fortran shapes of glob_v.f90:77-120. */
int rpb_set_evb(rpb_ctx*,
                const int* donor_acceptor_interaction /*(15,3)*/, const double* donor_acceptor_parameters /*(15,6)*/,
                const int* proton_acceptor_interaction /*(15,2)*/, const double* proton_acceptor_parameters /*(15,5)*/,
                const int* diabat_coupling_interaction /*(15,3)*/, const double* diabat_coupling_parameters /*(15,10)*/,
                const int* diabat_coupling_type /*(15)*/,
                const double* exchange_charge_atomic /*(25)*/, const double* exchange_charge_proton /*(10,10)*/,
                const int* acid_molecule /*(10)*/, const int* basic_molecule /*(10)*/,
                const int* conjugate_pairs /*(10)*/, const int* conjugate_atom_index /*(25)*/,
                const double* reference_energy /*(10)*/, const int* proton_index /*(10)*/,
                const int* heavy_acid_index /*(10)*/);

/* atom_data / molecule_data (glob_v.f90:157-176): xyz,velocity (3,N); mass,charge,atom_type_index (N);
 * molecule i covers atoms mol_first_atom[i] .. +mol_n_atom[i]-1 (1-based, contiguous:
 * general_routines.f90:670-671).  hydronium_mol = hydronium_molecule_index(1) or 0. */
int rpb_upload_state(rpb_ctx*, const double* xyz, const double* velocity, const double* mass,
                     const double* charge, const int* atom_type_index,
                     const int* mol_first_atom, const int* mol_n_atom, const int* mol_type,
                     int hydronium_mol);

/* Tail of initialize_simulation (initialize_routines.f90:121-134): update_r_com,
 * shift_molecules_into_box, construct_verlet_list, update_verlet_displacements(init). */
int rpb_initialize(rpb_ctx*);

/* == calculate_total_force_energy (ms_evb=0) or ms_evb_calculate_total_force_energy (ms_evb=1). */
int rpb_force_energy(rpb_ctx*, int ms_evb);

/* == md_integrate_atomic, NVE branch (md_integration.f90:438-541), n_steps times, on device. */
int rpb_step(rpb_ctx*, int n_steps, int ms_evb);

/* Independent replicas on one device ("replicas only", BASELINE config 5): rpb_step on every context, one host thread
 * per context so that the replicas' launches overlap on the device.  Returns the first non-zero status. */
int rpb_ensemble_step(rpb_ctx** replicas, int n_replicas, int n_steps, int ms_evb);

/* Sharded variant of the MS-EVB force call (world_size>1): the caller runs the two
 * collectives between the phases:
 *   rpb_evb_phase_build   principal diabat, enumeration, owned states' matrix elements
 *   [all-reduce(sum) of the buffer returned by rpb_evb_exchange_h]
 *   rpb_evb_phase_mix     Jacobi, owned states' Hellmann-Feynman partial forces
 *   [all-reduce(sum) of the buffer returned by rpb_evb_exchange_f]
 *   rpb_evb_phase_commit  hop commit (+ Verlet rebuild) as ms_evb.f90:218-227
 * rpb_step_begin/rpb_step_end are the two halves of md_integrate_atomic around the force call. */
int rpb_step_begin(rpb_ctx*);
int rpb_step_end(rpb_ctx*);
int rpb_evb_phase_build(rpb_ctx*);
int rpb_evb_phase_mix(rpb_ctx*);
int rpb_evb_phase_commit(rpb_ctx*);
/* exchange buffers: device pointers for the CUDA library, host pointers for the oracle. */
int rpb_evb_exchange_h(rpb_ctx*, void** ptr, int* n_doubles);
int rpb_evb_exchange_f(rpb_ctx*, void** ptr, int* n_doubles);

/* Peer-memory exchange (CUDA library; NVLink / NVSwitch): with it rpb_step / rpb_force_energy run the whole sharded
 * MS-EVB step inside the library -- the two all-reduces become one kernel each that pulls the peers' partials over
 * peer memory and adds them in rank order (bit-identical on every rank), no host round trip, no library collective.
 *   one process per GPU:  rpb_peer_export -> exchange the RPB_PEER_HANDLE_BYTES-byte handles of all ranks (rank order)
 *                         by any means (MPI_Allgather, torch.distributed.all_gather) -> rpb_peer_import
 *   one process, several contexts:  rpb_peer_attach_local(contexts in rank order)
 * The oracle returns RPB_ERR_UNSUPPORTED. */
#define RPB_PEER_HANDLE_BYTES 64
#define RPB_MAX_RANKS 16
int rpb_peer_export(rpb_ctx*, void* handle_out /*RPB_PEER_HANDLE_BYTES*/);
int rpb_peer_import(rpb_ctx*, const void* handles /*world_size x RPB_PEER_HANDLE_BYTES*/, int world_size);
int rpb_peer_attach_local(rpb_ctx** ranks, int world_size);
int rpb_peer_enabled(rpb_ctx*);

/* ---- results ---- */
int rpb_get_energies(rpb_ctx*, rpb_energies* out);   /* KE from calculate_kinetic_energy total_energy_forces.f90:106 */
int rpb_download_state(rpb_ctx*, double* xyz, double* velocity, double* force,
                       double* mass, double* charge, int* atom_type_index,
                       int* mol_first_atom, int* mol_n_atom, int* mol_type, int* hydronium_mol);
int rpb_get_r_com(rpb_ctx*, double* r_com /*(3,M)*/);
/* verlet_point(N+1), neighbor_list(n_pairs): 1-based values exactly as general_routines.f90:1517,1568. */
int rpb_get_neighbor_list(rpb_ctx*, int* verlet_point, int* neighbor_list, int capacity, int* n_pairs,
                          int* flag_verlet_list);
/* The pair list the force kernel itself consumes, expanded into (i, j) pairs, i < j, 1-based, sorted by (i, j): as a SET
 * it must equal the half list above (the CUDA library keeps cluster-pair tiles with per-atom-pair masks and generates
 * the reference-ordered list only for the accessor; the oracle returns its own list re-sorted).  n_tiles: list words. */
int rpb_debug_tile_pairs(rpb_ctx*, int* pair_i, int* pair_j, long long capacity, long long* n_pairs, long long* n_tiles);
/* PME parity accessors: Q_grid, theta_conv_Q (K,K,K) of diabat `state` (1 = principal), force_recip(3,N). */
int rpb_get_pme(rpb_ctx*, int state, double* Q_grid, double* theta_conv_Q, double* force_recip);
/* evb_hamiltonian(80,80) upper triangle, ground-state eigenvector c(S), evb_diabat_proton_log(80,3,5),
 * evb_diabat_coupling_matrix(80), S=diabat_index, principal diabat and new hydronium molecule. */
int rpb_get_evb(rpb_ctx*, int* n_states, double* hamiltonian, double* eigenvector, int* proton_log,
                int* coupling_matrix, int* principal_diabat, int* new_hydronium_mol,
                double* adiabatic_potential);
/* Hellmann-Feynman mix with a caller-supplied coefficient vector c(S) instead of the ground state
 * (c = e_s gives the stored diagonal force of diabat s, evb_forces_store(:,:,idx(s,s))). */
int rpb_debug_mix_forces(rpb_ctx*, const double* c, double* force /*(3,N)*/);

/* ---- measurement hooks (bench.py) ---- */
/* number of kernels of this library launched since rpb_create (cuFFT launches counted separately) */
int rpb_get_launch_counts(rpb_ctx*, long long* own_kernels, long long* cufft_execs);
/* CUDA-event timers accumulated per phase since the last reset; names via rpb_timer_name. */
int rpb_timers_enable(rpb_ctx*, int on);
int rpb_timers_reset(rpb_ctx*);
int rpb_timer_count(void);
const char* rpb_timer_name(int i);
int rpb_timers_get(rpb_ctx*, double* ms /*rpb_timer_count()*/, long long* calls);
/* stream the library launches on (cudaStream_t as void*), for external event timing */
void* rpb_get_stream(rpb_ctx*);
/* measured fp64 FMA peak of the device (TFLOP/s): denominator for the FP64-pipe-bound pair kernels */
int rpb_measure_fp64_peak(rpb_ctx*, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* RPBMD_H */
