// C-ABI of librpbmd.so (include/rpbmd.h) and the host orchestration of one force evaluation.
// The orchestration functions carry the names of the reference routines whose call structure they
// keep: calculate_total_force_energy (total_energy_forces.f90:19-99),
// ms_evb_calculate_total_force_energy (ms_evb.f90:181-235), md_integrate_atomic (md_integration.f90:438-541).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include "rpb_host.h"

static const char* k_timer_names[T_NTIMER] = {
    "integrate", "verlet", "pair_real_space", "molecule_terms", "pme_spread", "pme_fft", "pme_convolve", "pme_gather",
    "evb_enumerate", "evb_items", "evb_candidates", "evb_grid_broadcast", "evb_grid_patch", "evb_recip_corr",
    "evb_coupling_vex", "evb_jacobi", "evb_theta_mix", "evb_mix_forces", "evb_gather_mix", "evb_snapshots", "evb_coupling_geo",
    "evb_assemble", "step_total"};

#define CK(call)                                                                  \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e__);               \
      return RPB_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

// layout of the contiguous state block, device and pinned images alike: xq[N + 4] | vel[3N] (padded to 32 bytes) | force[3N]
static inline size_t sb_xq_bytes(int N) { return (size_t)(N + 4) * sizeof(double4); }
static inline size_t sb_vel_bytes(int N) { return ((size_t)3 * N * sizeof(double) + 31) & ~(size_t)31; }   // force stays 32-byte aligned (peer all-reduce: 16-byte accesses)
static inline size_t sb_bytes_up(int N) { return sb_xq_bytes(N) + sb_vel_bytes(N); }
static inline size_t sb_bytes_all(int N) { return sb_xq_bytes(N) + sb_vel_bytes(N) + (size_t)3 * N * sizeof(double); }

template <typename T>
static int upload(rpb_ctx* c, T** dst, const T* src, size_t n) {
  int rc = dev_alloc(c, dst, n);
  if (rc) return rc;
  CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
// check device-side error flags + fetch energies (one small synchronising read-back)
static int fetch_status(rpb_ctx* c) {
  CK(cudaMemcpyAsync(c->h_en, c->d.en, E_NSLOT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(c->h_flags, c->d.err_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->last_en.kinetic_energy = c->h_en[E_KE];
  if (c->h_flags[1] == 2) { c->err = "a molecule's atoms are farther apart than the box allows (r_cutoff + extent of three consecutive atoms >= L/2)"; return RPB_ERR_VERLET; }
  if (c->h_flags[1]) { c->err = "please increase size of verlet neighbor list"; return RPB_ERR_VERLET; }
  if (c->h_flags[2]) { c->err = "Found more diabat states than the current setting of evb_max_states"; return RPB_ERR_DIABATS; }
  if (c->h_flags[3] >= 30) { c->err = "peer-memory exchange: rank " + std::to_string(c->h_flags[3] - 30) + " did not arrive"; return RPB_ERR_CUDA; }
  if (c->h_flags[3]) { c->err = "couldn't find index in subroutine 'get_index_atom_set'"; return RPB_ERR_STATE; }
  if (c->h_flags[0]) { c->err = "force on atom " + std::to_string(c->h_flags[0]) + " is too big"; return RPB_ERR_FORCE; }
  return 0;
}

static void energies_from_slots(rpb_ctx* c) {
  rpb_energies& e = c->last_en;
  const double* s = c->h_en;
  e.E_recip = s[E_RECIP];
  e.E_elec = s[E_ELEC] + s[E_RECIP] + c->cfg.ewald_self;
  e.E_vdw = s[E_VDW]; e.E_bond = s[E_BOND]; e.E_angle = s[E_ANGLE]; e.E_dihedral = s[E_DIH];
  e.potential_energy = e.E_elec + e.E_vdw + e.E_bond + e.E_angle + e.E_dihedral;
}

// calculate_total_force_energy (total_energy_forces.f90:19-99); evb_principal: called as the principal-diabat
// evaluation of construct_evb_hamiltonian (ms_evb.f90:411): the reciprocal force is gathered later from the
// Hellmann-Feynman mixed grid instead (DESIGN.md "theta-mix").
// Three independent branches run concurrently: [Verlet update -> pair forces] on the main stream, the per-molecule
// bonded / intramolecular terms on aux[0], and the PME reciprocal branch (scaled coordinates, spreading, FFT) on
// aux[1].  In MS-EVB mode aux[1] is NOT joined here: evb_build keeps using it for the batched diabat grids.
int calculate_total_force_energy(rpb_ctx* c, bool evb_principal) {
  if (!c->forces_zeroed) launch_zero_forces(c);
  c->forces_zeroed = false;
  stream_depend(c, 0, c->main_stream, c->aux[0]);
  stream_depend(c, 1, c->main_stream, c->aux[1]);
  // the long main-stream kernels are issued first: the host needs ~3 us per launch, and the GPU should not idle while
  // the side branches are being queued
  int rc = launch_verlet_update(c);
  if (rc) return rc;
  if (evb_principal) {
    // MS-EVB.  Two chains bound the time to the Hamiltonian: [pair forces -> candidate lists -> per-diabat real-space
    // deltas] on the main stream and [enumeration -> host -> per-step tables] on aux[0].  The host needs ~3 us per launch,
    // so the two long kernels are issued first, everything else behind them:
    //   aux[0]: enumeration + images (positions and centres of mass only; the host learns the number of diabats -- which
    //           sizes the later launches -- while the GPU is busy with the pair forces)
    //   main  : pair forces (sharded by atoms in a state-sharded run)
    //   aux[1]: accumulator clears, spreading of the principal grid      aux[2]: bonded terms (24 small CTAs; rank 0 only)
    //   aux[3]: the enumeration's read-back (so that the images on aux[0] do not queue behind two D2H copies)
    // evb_build keeps the side streams busy and joins them before the Hamiltonian.
    { StreamScope sc(c, c->aux[0]); evb_enumerate_async(c, 0); }
    if (c->d.rank == 0) stream_depend(c, 8, c->main_stream, c->aux[2]);   // fork ahead of the pair kernel
    launch_pair_verlet(c, true);
    {
      // aux[1] has slack before its first consumer: the accumulators of the build are cleared here, off both chains
      StreamScope sc(c, c->aux[1]);
      evb_clear_early(c);
      cudaEventRecord(c->ev_sync[11], c->stream);
      launch_spread_principal(c);
      cudaEventRecord(c->ev_sync[15], c->stream);   // scaled coordinates ready (pair matrix of the chain atoms, aux[2])
    }
    {   // after the clears have been queued: the coupling geometry launched here waits for them (event 11)
      StreamScope sc(c, c->aux[0]);
      rc = evb_enumerate_async(c, 1);
      if (rc) return rc;
    }
    if (c->d.rank == 0) { StreamScope sc(c, c->aux[2]); launch_molecule_terms(c); }
    return 0;
  }
  launch_pair_verlet(c, false);   // issued first (the longest kernel); the non-reactive call is never sharded
  {
    // the whole reciprocal branch -- spreading, convolution AND the force gather (atomic adds into d.force, like the
    // pair kernel's) -- runs next to the pair kernel
    StreamScope sc(c, c->aux[1]);
    launch_spread_principal(c);
    rc = launch_convolve(c, 0, 1, c->d.en + E_RECIP, true);
    if (rc) return rc;
    launch_gather(c, c->d.theta, c->d.force_recip, true);
  }
  { StreamScope sc(c, c->aux[0]); launch_molecule_terms(c); }
  stream_depend(c, 2, c->aux[0], c->main_stream);
  stream_depend(c, 3, c->aux[1], c->main_stream);
  return 0;
}

// enqueue one force evaluation (no host synchronisation)
static int enqueue_force(rpb_ctx* c, int ms_evb, bool defer_join = false) {
  int rc;
  c->image_valid = false; c->ke_valid = false;
  if (ms_evb) {
    if (!c->have_evb) { c->err = "rpb_set_evb not called"; return RPB_ERR_STATE; }
    const bool sharded = c->d.world > 1;   // peer-memory exchange (kernels_peer.cu); checked by the callers
    // sharded, tree solver: the solver kernel assembles this rank's partial block and runs the Hamiltonian exchange itself
    c->evb_h_exchange_in_solver = sharded && c->peer.on && c->evb_solver == 0;
    if (sharded) peer_begin(c, PEER_H);
    rc = evb_build(c); if (rc) return rc;
    if (sharded) { if (!c->evb_h_exchange_in_solver) { rc = peer_allreduce(c, PEER_H); if (rc) return rc; } peer_begin(c, PEER_F); }
    rc = evb_mix(c, nullptr, nullptr); if (rc) return rc;
    if (sharded) { rc = peer_allreduce(c, PEER_F); if (rc) return rc; }
    rc = evb_commit(c);
    if (!defer_join) evb_join_readback(c);     // (a step joins after its second half kick)
    return rc;
  }
  return calculate_total_force_energy(c, false);
}

// synchronise and collect what the accessors serve: error flags, energies, (MS-EVB) the last step's diabat set and solution
static int collect_results(rpb_ctx* c, int ms_evb) {
  int rc = fetch_status(c);
  if (rc) return rc;
  if (ms_evb) return evb_readback(c);
  energies_from_slots(c);
  return 0;
}

// md_integrate_atomic (md_integration.f90:438-541), one step, enqueue only
static int enqueue_step(rpb_ctx* c, int ms_evb) {
  ScopedTimer t(c, T_STEP);
  launch_integrate_first(c);
  int rc = enqueue_force(c, ms_evb, true);
  if (rc) return rc;
  launch_integrate_second(c);
  evb_join_readback(c);
  return 0;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* rpb_backend(void) { return "cuda-sm100a"; }
const char* rpb_last_error(const rpb_ctx* c) { return c ? c->err.c_str() : "null context"; }
int rpb_timer_count(void) { return T_NTIMER; }
const char* rpb_timer_name(int i) { return (i >= 0 && i < T_NTIMER) ? k_timer_names[i] : ""; }

int rpb_create(rpb_ctx** out, const rpb_config* cfg) {
  if (!out || !cfg) return RPB_ERR_ARG;
  rpb_ctx* c = new rpb_ctx();
  *out = c;
  c->cfg = *cfg;
  memset(&c->d, 0, sizeof(Dev));
  memset(&c->e, 0, sizeof(EvbDev));
  memset(&c->last_en, 0, sizeof(rpb_energies));
  for (int i = 0; i < T_NTIMER; i++) { c->t_ms[i] = 0; c->t_calls[i] = 0; }
  if (cfg->spline_order != 6) { c->err = "only spline_order=6 is self-consistent in the reference (pme.f90:247 divides by 6.D0)"; return RPB_ERR_UNSUPPORTED; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      if (i != j && std::fabs(cfg->box[i + 3 * j]) > 10e-6) { c->err = "code has been modified to assume orthorhombic box"; return RPB_ERR_UNSUPPORTED; }
  if (cfg->n_atoms >= (1 << 23)) { c->err = "more than 2^23 atoms: a cluster-pair list entry packs the atom index into 23 bits"; return RPB_ERR_UNSUPPORTED; }
  if (cfg->evb_max_chain > RPB_MAXC || cfg->evb_max_states > RPB_MAXS) { c->err = "evb limits exceed compiled maxima"; return RPB_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { c->err = "no CUDA device: librpbmd.so has no CPU fallback"; return RPB_ERR_CUDA; }
  CK(cudaSetDevice(cfg->device));
  {
    // the side streams carry short kernels that must not starve behind the grid of the pair kernel on the main stream:
    // they get the higher priority, so their blocks are scheduled first whenever an SM frees resources
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_lo));
    c->main_stream = c->stream;
    // RPB_SERIAL_STREAMS=1: every branch on the main stream (clean per-kernel CUDA-event times for the roofline table)
    c->serial_streams = getenv("RPB_SERIAL_STREAMS") != nullptr;
    for (int k = 0; k < 5; k++) {
      if (c->serial_streams) c->aux[k] = c->stream;
      else CK(cudaStreamCreateWithPriority(&c->aux[k], cudaStreamNonBlocking, prio_hi));
    }
  }
  for (int k = 0; k < 24; k++) CK(cudaEventCreateWithFlags(&c->ev_sync[k], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_enum, cudaEventDisableTiming));
  CK(cudaMallocHost(&c->h_en, E_NSLOT * sizeof(double)));
  CK(cudaMallocHost(&c->h_flags, 8 * sizeof(int)));
  Dev& d = c->d;
  d.N = cfg->n_atoms; d.M = cfg->n_mole; d.K = cfg->pme_grid; d.nT = cfg->n_atom_type; d.nMT = cfg->n_mole_type;
  d.rank = cfg->rank; d.world = std::max(1, cfg->world_size);
  for (int i = 0; i < 3; i++) { d.box[i] = cfg->box[i + 3 * i]; d.inv_box[i] = 1.0 / d.box[i]; }
  {  // construct_reciprocal_lattice_vector with `real function volume` (general_routines.f90:473-490,1936-1947)
    double v = d.box[0] * (d.box[1] * d.box[2]);
    float v32 = std::fabs((float)v);
    double vol = (double)v32;
    d.kk[0] = (d.box[1] * d.box[2]) / vol; d.kk[1] = (d.box[2] * d.box[0]) / vol; d.kk[2] = (d.box[0] * d.box[1]) / vol;
  }
  d.rc2 = cfg->real_space_cutoff * cfg->real_space_cutoff;
  d.rv2 = cfg->verlet_cutoff * cfg->verlet_cutoff;
  d.verlet_skin = cfg->verlet_thresh * (cfg->verlet_cutoff - cfg->real_space_cutoff);
  d.alpha = cfg->alpha_sqrt; d.erf_factor = 2.0 * cfg->alpha_sqrt / cfg->pi_sqrt;
  d.conv = cfg->conv_e2A_kJmol; d.conv_kin = cfg->conv_kJmol_ang2ps2gmol; d.dt = cfg->delta_t; d.pi = cfg->pi;
  d.erfc_dx = cfg->erfc_dx; d.tt_max = cfg->tt_max; d.tt_grid = cfg->tt_grid; d.spline_grid = (double)cfg->spline_grid;
  d.ewald_self = cfg->ewald_self;
  d.cut_solv2 = cfg->evb_first_solvation_cutoff * cfg->evb_first_solvation_cutoff;
  d.cut_pair2 = cfg->evb_reactive_pair_distance * cfg->evb_reactive_pair_distance;
  d.max_chain = cfg->evb_max_chain; d.max_states = cfg->evb_max_states;
  // verlet cell grid (general_routines.f90:1444-1509)
  d.ncx = cfg->na_nslist; d.ncy = cfg->nb_nslist; d.ncz = cfg->nc_nslist;
  if (d.box[0] < 2 * cfg->verlet_cutoff) { c->err = "box size less than twice verlet cutoff"; return RPB_ERR_ARG; }
  if (d.ncx < 10 || d.ncy < 10 || d.ncz < 10 || d.ncx > 99 || d.ncy > 99 || d.ncz > 99) { c->err = " na_nslist, nb_nslist, nc_nslist must be between 10 and 100 "; return RPB_ERR_ARG; }
  double rka = std::sqrt(d.inv_box[0] * d.inv_box[0]), rkb = std::sqrt(d.inv_box[1] * d.inv_box[1]), rkc = std::sqrt(d.inv_box[2] * d.inv_box[2]);
  d.dia = (int)std::floor(cfg->verlet_cutoff * rka * (double)d.ncx) + 1;
  d.dib = (int)std::floor(cfg->verlet_cutoff * rkb * (double)d.ncy) + 1;
  d.dic = (int)std::floor(cfg->verlet_cutoff * rkc * (double)d.ncz) + 1;
  if (d.dia >= d.ncx / 2 || d.dib >= d.ncy / 2 || d.dic >= d.ncz / 2) {
    c->err = "number of grid cells to search in each dimension must be less than half the total number of grid cells";
    return RPB_ERR_ARG;
  }
  {  // allocate_verlet_list general_routines.f90:1206-1247 (volume = box(1,1)**3 after periodic_box_change)
    if (cfg->verlet_capacity > 0) d.verlet_cap = cfg->verlet_capacity;
    else {
      double volume = d.box[0] * d.box[0] * d.box[0], N = (double)d.N, rv = cfg->verlet_cutoff;
      long long sv = (long long)std::floor(4.0 * cfg->pi * (rv * rv * rv) * (N * N) / 6.0 / volume);
      sv = std::max((long long)d.N * 50, sv);
      d.verlet_cap = (int)std::floor((double)sv * cfg->safe_verlet);
    }
  }
  int rc;
  const int N = d.N, M = d.M, K = d.K;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)K * K * (K / 2 + 1);
  const int ncell = d.ncx * d.ncy * d.ncz;
#define AL(p, n) if ((rc = dev_alloc(c, &(p), (n)))) return rc;
  {   // xq | vel | force contiguous: one copy per state transfer
    AL(c->state_block, sb_bytes_all(N));
    d.xq = reinterpret_cast<double4*>(c->state_block);
    d.vel = reinterpret_cast<double*>(c->state_block + sb_xq_bytes(N));
    d.force = reinterpret_cast<double*>(c->state_block + sb_xq_bytes(N) + sb_vel_bytes(N));
  }
  AL(d.mass, N); AL(d.type, N); AL(d.mol_of_atom, N);
  AL(d.mol_first, M); AL(d.mol_natom, M); AL(d.mol_type, M); AL(d.r_com, 3 * M); AL(d.hydronium, 1);
  AL(d.verlet_point, N + 1); AL(d.neighbor_list, d.verlet_cap);
  d.tile_cap = 2 * (long long)d.verlet_cap;       // a tile holds at least one listed pair, and the listed pairs fit the reference's capacity
  AL(d.tile_point, RPB_TILE_PARTS * (size_t)N + 4); AL(d.tile_tmp, (RPB_TILE_PARTS * (size_t)N + 4) * RPB_TILE_TMPCAP); AL(d.tile_list, (size_t)d.tile_cap); AL(d.cl_info, N + 1); AL(d.n_clusters, 1); AL(d.mol_cl_first, M + 1); AL(d.mol_ncl, M + 1);
  AL(d.vbuild_xq, N); AL(d.csort_xq, 3 * (size_t)N); AL(d.csort_mol, N); AL(d.csort_info, N); AL(d.vstat, 8);
  AL(d.vsort_xq, N); AL(d.vsort_mol, N); AL(d.vsort_entry, N); AL(d.vstore, 3 * N); AL(d.vdisp, 3 * N);
  AL(d.flag_verlet, 1); AL(d.rebuild_now, 1); AL(d.err_flag, 4); AL(d.vdone, 1);
  AL(d.cell_count, 2 * (ncell + 1)); AL(d.cell_start, ncell + 1); AL(d.cell_atoms, N); AL(d.atom_cell, N); AL(d.row_count, RPB_TILE_PARTS * (size_t)N + 4);
  AL(d.maxd, 8 + 2 * ((N + 255) / 256 + 1) + 4 * ((N + 127) / 128 + 1));
  AL(d.uscale, 3 * N); AL(d.force_recip, 3 * N); AL(d.en, E_NSLOT);
  c->grid_capacity = 1;
  if (cfg->evb_max_states > 0) c->grid_capacity = cfg->evb_max_states;
  // grids are allocated lazily for the diabats in rpb_set_evb; the principal needs one of each
  const size_t n_grids = (size_t)c->grid_capacity + 4;
  AL(d.Q, K3 * n_grids); AL(d.theta, K3 * n_grids); AL(d.FQ, Kh3 * n_grids);
#undef AL
  CK(cudaMemset(d.Q, 0, K3 * n_grids * sizeof(double)));
  CK(cudaMemset(d.flag_verlet, 0, sizeof(int)));
  CK(cudaMemset(d.rebuild_now, 0, sizeof(int)));
  CK(cudaMemset(d.vdone, 0, sizeof(int)));
  CK(cudaMemset(d.err_flag, 0, 4 * sizeof(int)));
  CK(cudaMemset(d.en, 0, E_NSLOT * sizeof(double)));
  CK(cudaMemset(d.force, 0, 3 * N * sizeof(double)));
  CK(cudaMemset(d.force_recip, 0, 3 * N * sizeof(double)));
  CK(cudaMemset(d.xq, 0, (N + 4) * sizeof(double4)));
  CK(cudaMemset(d.vstat, 0, 8 * sizeof(unsigned long long)));
  CK(cudaMemset(d.n_clusters, 0, sizeof(int)));
  return verlet_setup(c);
}

void rpb_destroy(rpb_ctx* c) {
  if (!c) return;
  if (c->stream) cudaStreamSynchronize(c->stream);
  peer_free(c);
  evb_free(c);
  fft_conv_free(c);
  for (auto& kv : c->plan_fwd) cufftDestroy(kv.second);
  for (auto& kv : c->plan_inv) cufftDestroy(kv.second);
  for (void* p : c->allocs) cudaFree(p);
  if (c->h_en) cudaFreeHost(c->h_en);
  if (c->h_flags) cudaFreeHost(c->h_flags);
  for (int k = 0; k < 8; k++) if (c->graph[k].exec) cudaGraphExecDestroy(c->graph[k].exec);
  if (c->staging) cudaFreeHost(c->staging);
  for (int k = 0; k < 2; k++) { if (c->staging_up[k]) cudaFreeHost(c->staging_up[k]); if (c->ev_up[k]) cudaEventDestroy(c->ev_up[k]); }
  if (c->eh.pinned) cudaFreeHost(c->eh.pinned);
  if (c->stream) {
    for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
    for (int k = 0; k < 24; k++) if (c->ev_sync[k]) cudaEventDestroy(c->ev_sync[k]);
    if (c->ev_enum) cudaEventDestroy(c->ev_enum);
    for (int k = 0; k < 5; k++) if (c->aux[k] && !c->serial_streams) { cudaStreamSynchronize(c->aux[k]); cudaStreamDestroy(c->aux[k]); }
    cudaStreamDestroy(c->stream);
  }
  delete c;
}

int rpb_set_tables(rpb_ctx* c, const double* B6, const double* B5, const double* erfc_t, const double* scale_t,
                   const double* tt, const double* dtt, const double* CB) {
  Dev& d = c->d;
  const int K = d.K, Kh = K / 2 + 1;
  int rc;
  double *p;
  if ((rc = upload(c, &p, B6, c->cfg.spline_grid))) return rc; d.B6 = p;
  if ((rc = upload(c, &p, B5, c->cfg.spline_grid))) return rc; d.B5 = p;
  if ((rc = upload(c, &p, erfc_t, c->cfg.erfc_grid + 1))) return rc; d.erfc_t = p;
  if ((rc = upload(c, &p, scale_t, c->cfg.erfc_grid + 1))) return rc; d.scale_t = p;
  {  // interleaved copy for the Verlet pair kernel; two spare points behind the end keep a rounded-up index in bounds
    const int G = c->cfg.erfc_grid;
    std::vector<double4> es((size_t)G + 3);
    for (int i = 0; i < G + 3; i++) { int k0 = std::min(std::max(i - 1, 0), G), k1 = std::min(i, G); es[i] = make_double4(erfc_t[k0], scale_t[k0], erfc_t[k1], scale_t[k1]); }
    double4* q;
    if ((rc = upload(c, &q, es.data(), es.size()))) return rc;
    d.es2_t = q; d.inv_erfc_dx = 1.0 / c->cfg.erfc_dx;
  }
  if ((rc = upload(c, &p, tt, 4 * c->cfg.tt_grid))) return rc; d.tt = p;
  if ((rc = upload(c, &p, dtt, 4 * c->cfg.tt_grid))) return rc; d.dtt = p;
  // Half-spectrum copy of CB(K,K,K), element (m1,m2,m3), m1 <= K/2, stored [m3][m2][m1].
  // The reference keeps dble() of a full complex transform (pme.f90:123).  Its CB is even only up to the
  // float32 rounding of the phase factors in bm_sq (pme.f90:589), and Re(IDFT(CB*F)) of a real grid equals
  // IDFT(CB_even*F) with CB_even(m) = (CB(m)+CB(-m))/2 -- which is what the real-to-complex pair needs.
  std::vector<double> cbh((size_t)Kh * K * K);
  for (int m3 = 0; m3 < K; m3++)
    for (int m2 = 0; m2 < K; m2++)
      for (int m1 = 0; m1 < Kh; m1++) {
        size_t a = (size_t)m1 + (size_t)K * (m2 + (size_t)K * m3);
        size_t b = (size_t)((K - m1) % K) + (size_t)K * (((K - m2) % K) + (size_t)K * ((K - m3) % K));
        cbh[(size_t)m1 + (size_t)Kh * (m2 + (size_t)K * m3)] = 0.5 * (CB[a] + CB[b]);
      }
  if ((rc = upload(c, &p, cbh.data(), cbh.size()))) return rc; d.CBh = p;
  c->have_tables = true;
  return 0;
}

int rpb_set_forcefield(rpb_ctx* c, const double* vdw_parameter, const int* vdw_type, const double* vdw_parameter_14,
                       const double* atype_chg, const int* atype_freeze, const int* bond_type, const double* bond_parameter,
                       const int* angle_type, const double* angle_parameter, const int* dihedral_type,
                       const double* dihedral_parameter) {
  const int T = RPB_MAXT, nT = c->d.nT;
  const size_t T2 = (size_t)T * T, T3 = T2 * T, T4 = T3 * T;
  // keep host copies in the Fortran layout for the per-term look-ups of rpb_set_molecule_types
  c->ff_bondt.assign(bond_type, bond_type + T2); c->ff_bondp.assign(bond_parameter, bond_parameter + T2 * 3);
  c->ff_anglet.assign(angle_type, angle_type + T3); c->ff_anglep.assign(angle_parameter, angle_parameter + T3 * 2);
  c->ff_diht.assign(dihedral_type, dihedral_type + T4); c->ff_dihp.assign(dihedral_parameter, dihedral_parameter + T4 * 6);
  for (int i = 0; i < T; i++) { c->atype_chg[i] = atype_chg[i]; c->atype_freeze[i] = atype_freeze[i]; }
  // compact [ti][tj][6] tables for the device
  std::vector<double> vp((size_t)nT * nT * 6), vp14((size_t)nT * nT * 6);
  std::vector<int> vt((size_t)nT * nT);
  for (int ti = 0; ti < nT; ti++)
    for (int tj = 0; tj < nT; tj++) {
      vt[ti * nT + tj] = vdw_type[ti + T * tj];
      for (int k = 0; k < 6; k++) {
        vp[6 * (ti * nT + tj) + k] = vdw_parameter[ti + T * tj + T2 * k];
        vp14[6 * (ti * nT + tj) + k] = vdw_parameter_14[ti + T * tj + T2 * k];
      }
    }
  int rc; double* p; int* q;
  if ((rc = upload(c, &p, vp.data(), vp.size()))) return rc; c->d.vdw_param = p;
  if ((rc = upload(c, &p, vp14.data(), vp14.size()))) return rc; c->d.vdw_param14 = p;
  if ((rc = upload(c, &q, vt.data(), vt.size()))) return rc; c->d.vdw_type = q;
  if ((rc = upload(c, &q, c->atype_freeze, (size_t)T))) return rc; c->d.freeze = q;
  c->have_ff = true;
  return 0;
}

int rpb_set_molecule_types(rpb_ctx* c, const int* n_atom, const int* atom_type, const int* n_bond, const int* bonds,
                           const int* n_angle, const int* angles, const int* n_dihedral, const int* dihedrals,
                           const int* pair_exclusions, const int* evb_reactive_protons, const int* evb_reactive_basic_atoms) {
  if (!c->have_ff) { c->err = "rpb_set_forcefield must precede rpb_set_molecule_types"; return RPB_ERR_STATE; }
  const int T = RPB_MAXT, MA = RPB_MA;
  const size_t T2 = (size_t)T * T, T3 = T2 * T, T4 = T3 * T;
  c->mt_host.assign(c->d.nMT, MolTypeDev());
  int ob = 0, oa = 0, od = 0;
  for (int t = 0; t < c->d.nMT; t++) {
    MolTypeDev& m = c->mt_host[t];
    memset(&m, 0, sizeof(m));
    m.n_atom = n_atom[t];
    if (m.n_atom + 1 > MA) { c->err = "molecule type larger than RPB_MAX_MOLE_ATOMS-1"; return RPB_ERR_ARG; }
    if (n_bond[t] > RPB_MAXB || n_angle[t] > RPB_MAXB || n_dihedral[t] > RPB_MAXB) { c->err = "too many bonded terms per molecule type"; return RPB_ERR_ARG; }
    for (int a = 0; a < MA; a++) {
      m.atom_type[a] = atom_type[t * MA + a] - 1;
      m.reactive_proton[a] = evb_reactive_protons ? evb_reactive_protons[t * MA + a] : 0;
      m.reactive_basic[a] = evb_reactive_basic_atoms ? evb_reactive_basic_atoms[t * MA + a] : 0;
      for (int b = 0; b < MA; b++) m.pair_excl[a][b] = pair_exclusions[t * MA * MA + a + MA * b];
    }
    { int nb = 0; for (int a = 0; a < MA; a++) nb += (m.reactive_basic[a] == 1); c->mt_multi_basic.resize(c->d.nMT, 0); c->mt_multi_basic[t] = nb > 1; if (nb > 1) c->evb_any_multi_basic = true; }
    m.n_bond = n_bond[t]; m.n_angle = n_angle[t]; m.n_dih = n_dihedral[t];
    for (int b = 0; b < m.n_bond; b++) {
      int i = bonds[2 * (ob + b)] - 1, j = bonds[2 * (ob + b) + 1] - 1;
      int ti = m.atom_type[i], tj = m.atom_type[j];
      m.bond[b][0] = i; m.bond[b][1] = j;
      m.bond_kind[b] = c->ff_bondt[ti + T * tj];
      if (m.bond_kind[b] < 1 || m.bond_kind[b] > 3) { c->err = "bond type isn't implemented!"; return RPB_ERR_ARG; }
      for (int k = 0; k < 3; k++) m.bond_par[b][k] = c->ff_bondp[ti + T * tj + T2 * k];
    }
    for (int a = 0; a < m.n_angle; a++) {
      int i = angles[3 * (oa + a)] - 1, j = angles[3 * (oa + a) + 1] - 1, k = angles[3 * (oa + a) + 2] - 1;
      int ti = m.atom_type[i], tj = m.atom_type[j], tk = m.atom_type[k];
      m.angle[a][0] = i; m.angle[a][1] = j; m.angle[a][2] = k;
      size_t idx = ti + T * tj + T2 * tk;
      m.angle_kind[a] = c->ff_anglet[idx];
      if (m.angle_kind[a] < 1 || m.angle_kind[a] > 2) { c->err = "requested angle type potential not implemented"; return RPB_ERR_ARG; }
      m.angle_par[a][0] = c->ff_anglep[idx]; m.angle_par[a][1] = c->ff_anglep[idx + T3];
    }
    for (int q = 0; q < m.n_dih; q++) {
      int ix[4], ty[4];
      for (int k = 0; k < 4; k++) { ix[k] = dihedrals[4 * (od + q) + k] - 1; ty[k] = m.atom_type[ix[k]]; m.dih[q][k] = ix[k]; }
      size_t idx = ty[0] + T * ty[1] + T2 * ty[2] + T3 * ty[3];
      m.dih_kind[q] = c->ff_diht[idx];
      for (int k = 0; k < 6; k++) m.dih_par[q][k] = c->ff_dihp[idx + T4 * k];
      if (m.dih_kind[q] == 1) {  // "undefined dihedral force" guard of intra_bonded_interactions.f90:434-441 needs phase 0 or pi
        double xi0 = m.dih_par[q][0];
        if (!(std::fabs(xi0) < 1e-4 || std::fabs(xi0 - c->cfg.pi) < 1e-4)) { /* handled at run time only when sin != 0 */ }
      }
    }
    ob += n_bond[t]; oa += n_angle[t]; od += n_dihedral[t];
    // find_bonded_atom_hydrogen general_routines.f90:575-602
    for (int a = 0; a < MA; a++) {
      int heavy = -1, count = 0;
      for (int b = 0; b < m.n_bond; b++) {
        if (m.bond[b][0] == a) { heavy = m.bond[b][1]; count++; }
        else if (m.bond[b][1] == a) { heavy = m.bond[b][0]; count++; }
      }
      m.bonded_heavy[a] = count == 1 ? heavy : -1;
    }
    m.heavy_acid_atom = m.heavy_base_atom = -1;
  }
  c->have_mt = true;
  if (!c->d.mt) {
    MolTypeDev* p;
    int rc = dev_alloc(c, &p, (size_t)c->d.nMT);
    if (rc) return rc;
    c->d.mt = p;
  }
  CK(cudaMemcpy((void*)c->d.mt, c->mt_host.data(), c->mt_host.size() * sizeof(MolTypeDev), cudaMemcpyHostToDevice));
  return 0;
}

int rpb_set_evb(rpb_ctx* c, const int* da_i, const double* da_p, const int* pa_i, const double* pa_p, const int* dc_i,
                const double* dc_p, const int* dc_t, const double* ex_a, const double* ex_p, const int* acid,
                const int* basic, const int* conj_pairs, const int* conj_atom, const double* ref_e, const int* proton_index,
                const int* heavy_acid_index) {
  if (!c->have_mt) { c->err = "rpb_set_molecule_types must precede rpb_set_evb"; return RPB_ERR_STATE; }
  EvbTables& e = c->evb_host;
  const int MI = RPB_MAXI, MM = RPB_MAXM, T = RPB_MAXT;
  for (int i = 0; i < MI; i++) {
    for (int j = 0; j < 3; j++) { e.da_int[i][j] = da_i[i + MI * j] - 1; e.dc_int[i][j] = dc_i[i + MI * j] - 1; }
    for (int j = 0; j < 2; j++) e.pa_int[i][j] = pa_i[i + MI * j] - 1;
    for (int j = 0; j < 6; j++) e.da_par[i][j] = da_p[i + MI * j];
    for (int j = 0; j < 5; j++) e.pa_par[i][j] = pa_p[i + MI * j];
    for (int j = 0; j < 10; j++) e.dc_par[i][j] = dc_p[i + MI * j];
    e.dc_type[i] = dc_t[i];
  }
  for (int i = 0; i < T; i++) { e.exch_atomic[i] = ex_a[i]; e.conj_atom[i] = conj_atom[i] - 1; e.atype_chg[i] = c->atype_chg[i]; }
  for (int i = 0; i < MM; i++) {
    for (int j = 0; j < MM; j++) e.exch_proton[i][j] = ex_p[i + MM * j];
    e.conj_pairs[i] = conj_pairs[i] - 1; e.ref_energy[i] = ref_e[i];
    e.proton_index[i] = proton_index[i] - 1; e.heavy_acid_index[i] = heavy_acid_index[i] - 1;
  }
  (void)acid; (void)basic;
  c->conj_pairs_host.assign(e.conj_pairs, e.conj_pairs + MM);
  {  // candidate radius of the diabat real-space deltas: the real-space cutoff (+ margin for the rounding of image
     // positions that were made whole), or the reach of the EVB repulsion terms if that is larger
    double da_rc = 0.0, pa_rc = 0.0;
    for (int i = 0; i < MI; i++) {
      if (e.da_int[i][0] >= 0) da_rc = std::max(da_rc, e.da_par[i][5]);
      if (e.pa_int[i][0] >= 0) pa_rc = std::max(pa_rc, e.pa_par[i][4]);
    }
    c->evb_rep_reach = pa_rc;
    c->evb_rcand = std::max(c->cfg.real_space_cutoff + 1e-3, std::max(da_rc, pa_rc + 4.0) + 1e-3);
  }
  // get_heavy_atom_transfer_acid / _base (ms_evb.f90:2888-2938), resolved per molecule type
  auto heavy_acid = [&](int t) {
    if (t < 0 || t >= c->d.nMT) return -1;
    int th = e.heavy_acid_index[t];
    if (th < 0) return -1;
    for (int a = 0; a < c->mt_host[t].n_atom; a++) if (c->mt_host[t].atom_type[a] == th) return a;
    return -1;
  };
  for (int t = 0; t < c->d.nMT; t++) {
    c->mt_host[t].heavy_acid_atom = heavy_acid(t);
    c->mt_host[t].heavy_base_atom = heavy_acid(e.conj_pairs[t]);
  }
  CK(cudaMemcpy((void*)c->d.mt, c->mt_host.data(), c->mt_host.size() * sizeof(MolTypeDev), cudaMemcpyHostToDevice));
  if (!c->d.evb) {
    EvbTables* p;
    int rc = dev_alloc(c, &p, 1);
    if (rc) return rc;
    c->d.evb = p;
  }
  CK(cudaMemcpy((void*)c->d.evb, &e, sizeof(EvbTables), cudaMemcpyHostToDevice));
  int rc = evb_alloc(c);
  if (rc) return rc;
  c->have_evb = true;
  return 0;
}

// Host <-> device state transfers go through one pinned staging area (allocated once) with asynchronous copies on the
// main stream; per-atom / per-molecule tables that did not change since the last upload are not sent again.
struct Staging {
  double4* xq; double* vel; double* force; double* mass; int* type; int* moa; int* mol;   // mol: first | n_atom | type
};
static int staging_get(rpb_ctx* c, Staging& st) {
  const size_t N = c->d.N, M = c->d.M;
  const size_t bytes = sb_bytes_all((int)N) + N * sizeof(double) + (2 * N + 3 * M) * sizeof(int);   // xq | vel | force laid out as on the device
  if (!c->staging) {
    CK(cudaMallocHost(&c->staging, bytes));
    memset(c->staging, 0xff, bytes);
  }
  char* p = (char*)c->staging;
  st.xq = (double4*)p; p += (N + 4) * sizeof(double4);
  st.vel = (double*)p; p += sb_vel_bytes((int)N);
  st.force = (double*)p; p += 3 * N * sizeof(double);
  st.mass = (double*)p; p += N * sizeof(double);
  st.type = (int*)p; p += N * sizeof(int);
  st.moa = (int*)p; p += N * sizeof(int);
  st.mol = (int*)p;
  return 0;
}

// committed hops permute the molecule table on the device; the host mirror (served by rpb_download_state, compared by
// rpb_upload_state) is refreshed from the device when evb_readback has seen a new hop
static int refresh_mirror(rpb_ctx* c) {
  if (!c->mirror_stale) return 0;
  const int M = c->d.M;
  CK(cudaStreamSynchronize(c->stream));
  c->mol_first.resize(M); c->mol_natom.resize(M); c->mol_type.resize(M);
  CK(cudaMemcpy(c->mol_first.data(), c->d.mol_first, M * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(c->mol_natom.data(), c->d.mol_natom, M * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(c->mol_type.data(), c->d.mol_type, M * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&c->hydronium_mol, c->d.hydronium, sizeof(int), cudaMemcpyDeviceToHost));
  int ncl = 0;
  for (int m = 0; m < M; m++) ncl += (c->mol_natom[m] + 2) / 3;
  c->n_clusters_bound = std::min(c->d.N, ncl + 2);
  c->mirror_stale = false;
  return 0;
}

int rpb_upload_state(rpb_ctx* c, const double* xyz, const double* velocity, const double* mass, const double* charge,
                     const int* atom_type_index, const int* mol_first_atom, const int* mol_n_atom, const int* mol_type,
                     int hydronium_mol) {
  const int N = c->d.N, M = c->d.M;
  Staging st;
  int rc = staging_get(c, st);
  if (rc) return rc;
  {   // validate the molecule table before anything is staged: a rejected upload must leave the change-detection cache alone
    int expect = 0;
    for (int m = 0; m < M; m++) {
      if (mol_first_atom[m] - 1 != expect) { c->err = "molecules must be contiguous ascending atom ranges"; return RPB_ERR_ARG; }
      expect += mol_n_atom[m];
    }
    if (expect != N) { c->err = "molecule table does not cover all atoms"; return RPB_ERR_ARG; }
  }
  if ((rc = refresh_mirror(c))) return rc;
  // {xq, vel} travel from one of two pinned images in ONE copy: no wait for the previous upload, only for the one before it
  const size_t up_bytes = sb_bytes_up(N);
  const int par = c->up_parity; c->up_parity ^= 1;
  if (!c->staging_up[par]) {
    CK(cudaMallocHost(&c->staging_up[par], up_bytes));
    memset(c->staging_up[par], 0, up_bytes);
    CK(cudaEventCreateWithFlags(&c->ev_up[par], cudaEventDisableTiming));
  } else CK(cudaEventSynchronize(c->ev_up[par]));
  double4* up_xq = reinterpret_cast<double4*>(c->staging_up[par]);
  double* up_vel = reinterpret_cast<double*>(reinterpret_cast<char*>(c->staging_up[par]) + sb_xq_bytes(N));
  const bool all = !c->have_state || !c->state_cache_valid;   // a committed proton hop permuted the device tables
  // change detection on raw copies of the caller's tables (one memcmp each): unchanged tables are neither converted nor sent
  c->raw_type.resize(N); c->raw_mass.resize(N); c->raw_mol.resize(3 * (size_t)M);
  bool type_changed = all || memcmp(c->raw_type.data(), atom_type_index, N * sizeof(int)) != 0;
  bool mass_changed = all || memcmp(c->raw_mass.data(), mass, N * sizeof(double)) != 0;
  bool mol_changed = all || memcmp(c->raw_mol.data(), mol_first_atom, M * sizeof(int)) != 0 ||
                     memcmp(c->raw_mol.data() + M, mol_n_atom, M * sizeof(int)) != 0 || memcmp(c->raw_mol.data() + 2 * (size_t)M, mol_type, M * sizeof(int)) != 0;
  if (type_changed || mass_changed || mol_changed) CK(cudaStreamSynchronize(c->stream));   // the cached tables below may still feed an earlier copy
  for (int i = 0; i < N; i++) up_xq[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], charge[i]);
  if (type_changed) { memcpy(c->raw_type.data(), atom_type_index, N * sizeof(int)); for (int i = 0; i < N; i++) st.type[i] = atom_type_index[i] - 1; }
  if (mass_changed) { memcpy(c->raw_mass.data(), mass, N * sizeof(double)); memcpy(st.mass, mass, N * sizeof(double)); }
  int ncl = 0;
  if (mol_changed) {
    memcpy(c->raw_mol.data(), mol_first_atom, M * sizeof(int)); memcpy(c->raw_mol.data() + M, mol_n_atom, M * sizeof(int));
    memcpy(c->raw_mol.data() + 2 * (size_t)M, mol_type, M * sizeof(int));
    const bool had_state = c->have_state && (int)c->mol_first.size() == M;
    c->mol_first.resize(M); c->mol_natom.resize(M); c->mol_type.resize(M);
    int expect = 0;
    for (int m = 0; m < M; m++) {
      const int f = mol_first_atom[m] - 1, n = mol_n_atom[m], t = mol_type[m] - 1;
      st.mol[m] = f; st.mol[M + m] = n; st.mol[2 * M + m] = t;
      if (had_state && (c->mol_first[m] != f || c->mol_natom[m] != n)) c->rebuild_forced = true;   // a different molecule table: the cluster table is stale
      c->mol_first[m] = f; c->mol_natom[m] = n; c->mol_type[m] = t;
      for (int a = 0; a < n; a++) st.moa[expect + a] = m;
      expect += n;
      ncl += (n + 2) / 3;
    }
  } else for (int m = 0; m < M; m++) ncl += (c->mol_natom[m] + 2) / 3;
  c->n_clusters_bound = std::min(N, ncl + 2);   // a hop moves one proton: the count changes by at most one either way
  if (mol_changed && c->evb_any_multi_basic && !c->conj_pairs_host.empty()) {
    // can the reference's re-ordering quirk (k_evb_reorder_quirk) occur at all?  Only if a molecule type with several basic
    // atoms is present, or can appear by a proton transfer (the conjugate of a present type)
    bool present = false;
    for (int m = 0; m < M && !present; m++) {
      const int t = c->mol_type[m], tc = (t >= 0 && t < (int)c->conj_pairs_host.size()) ? c->conj_pairs_host[t] : -1;
      present = (t >= 0 && t < (int)c->mt_multi_basic.size() && c->mt_multi_basic[t]) || (tc >= 0 && tc < (int)c->mt_multi_basic.size() && c->mt_multi_basic[tc]);
    }
    if (present != c->evb_quirk_types_present)       // the step graphs hold the launch list of the other case
      for (int k = 0; k < 8; k++) if (c->graph[k].exec) { cudaGraphExecDestroy(c->graph[k].exec); c->graph[k].exec = nullptr; }
    c->evb_quirk_types_present = present;
  }
  memcpy(up_vel, velocity, 3 * (size_t)N * sizeof(double));
  if (c->hydronium_mol != hydronium_mol - 1) mol_changed = true;
  c->hydronium_mol = hydronium_mol - 1;
  CK(cudaMemcpyAsync(c->state_block, c->staging_up[par], up_bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(c->ev_up[par], c->stream));
  if (mass_changed) CK(cudaMemcpyAsync(c->d.mass, st.mass, N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (type_changed) CK(cudaMemcpyAsync(c->d.type, st.type, N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (mol_changed) {
    CK(cudaMemcpyAsync(c->d.mol_of_atom, st.moa, N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d.mol_first, st.mol, M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d.mol_natom, st.mol + M, M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d.mol_type, st.mol + 2 * M, M * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    int* hp = c->h_flags + 4;                     // pinned scratch for the scalar
    *hp = c->hydronium_mol;
    CK(cudaMemcpyAsync(c->d.hydronium, hp, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  c->ke_valid = false; c->image_valid = false;
  c->have_state = true;
  c->state_cache_valid = true;
  return 0;
}

int rpb_initialize(rpb_ctx* c) {
  if (!(c->have_tables && c->have_ff && c->have_mt && c->have_state)) { c->err = "tables/forcefield/molecule types/state must be set first"; return RPB_ERR_STATE; }
  c->image_valid = false;
  launch_update_com_shift(c, true);
  int rc = launch_verlet_force_rebuild(c);
  if (rc) return rc;
  rc = fetch_status(c);
  if (rc) return rc;
  c->initialized = true;
  return 0;
}

int rpb_force_energy(rpb_ctx* c, int ms_evb) {
  if (!c->initialized) { c->err = "rpb_initialize not called"; return RPB_ERR_STATE; }
  if (ms_evb && c->d.world > 1 && !c->peer.on) { c->err = "world_size>1: set up the peer-memory exchange (rpb_peer_*) or use the phase calls"; return RPB_ERR_STATE; }
  int rc = enqueue_force(c, ms_evb);
  if (rc) return rc;
  return collect_results(c, ms_evb);
}

int rpb_step_begin(rpb_ctx* c) { c->image_valid = false; c->ke_valid = false; launch_integrate_first(c); return 0; }
int rpb_step_end(rpb_ctx* c) {
  launch_integrate_second(c);
  int rc = fetch_status(c);
  return rc;
}
// the phase calls always use the library's own exchange buffers (the caller runs the collectives), never the peer arena
int rpb_evb_phase_build(rpb_ctx* c) { c->evb_h_exchange_in_solver = false; if (c->peer.h_local) c->e.h_diag = c->peer.h_local; return evb_build(c); }
int rpb_evb_phase_mix(rpb_ctx* c) { if (c->peer.f_local) c->e.f_mix = c->peer.f_local; c->peer.f_reduced_in_place = false; return evb_mix(c, nullptr, nullptr); }
int rpb_evb_phase_commit(rpb_ctx* c) { int rc = evb_commit(c); evb_join_readback(c); return rc ? rc : evb_readback(c); }
int rpb_evb_exchange_h(rpb_ctx* c, void** ptr, int* n) { *ptr = c->e.h_diag; *n = 3 * RPB_MAXS + E_NSLOT; return 0; }
int rpb_evb_exchange_f(rpb_ctx* c, void** ptr, int* n) { *ptr = c->e.f_mix; *n = 3 * c->d.N; return 0; }

// One MD step is a fixed launch sequence (every data-dependent decision -- number of diabats, list rebuild, proton hop --
// is taken on the device), so it is captured once into a CUDA graph and replayed: ~45 kernel nodes on six streams cost
// one launch call per step instead of ~60 API calls.  Not used while the per-kernel timers run, for state-sharded ranks
// (the peer exchange alternates its buffers from step to step) or when RPB_GRAPH=0.
static bool graph_allowed(rpb_ctx* c, int ms_evb) {
  static const bool off = getenv("RPB_GRAPH") && atoi(getenv("RPB_GRAPH")) == 0;
  return !off && !c->timers_on && !c->serial_streams && !(ms_evb && c->d.world > 1 && !c->peer.on) && !c->graph_failed;
}

static int graph_get(rpb_ctx* c, int ms_evb, cudaGraphExec_t* out, int* launches, int force_slot = 0, int force_parity = -1) {
  // The grids of the MS-EVB kernels are sized for a bound of the diabat count (the kernels loop over the device-side count,
  // so a generous bound only costs idle CTAs).  The count of a trajectory wanders (20 ... 60 within 150 steps of the
  // benchmark system) and a capture costs ~1.2 ms of host time, so up to three graphs -- bounds 32, 56, evb_max_states --
  // are captured together on first use and KEPT; a step replays the smallest one that leaves 6 diabats of headroom.
  static const int bounds[3] = {32, 56, RPB_MAXS};
  const int hint = ms_evb ? c->eh.s_hint : 0;
  int slot = 0;
  if (ms_evb) { slot = 1; while (slot < 3 && hint + 6 > bounds[slot - 1]) slot++; }
  // state-sharded step: the producers of the exchanged partials write into the arena parity of the step's sequence number
  // (pointers baked into the launches), so there is one graph per parity and the steps alternate between them
  const bool sharded = ms_evb && c->d.world > 1;
  int parity = sharded ? (int)((c->peer.seq[PEER_H] + 1) & 1) : 0;
  if (force_slot) { slot = force_slot; parity = force_parity; }
  else if (ms_evb)
    for (int k = 1; k <= 3; k++)             // (all at once: no capture inside a later timed region)
      for (int q = 0; q < (sharded ? 2 : 1); q++)
        if ((k != slot || q != parity) && (!c->graph[k + 4 * q].exec || c->graph[k + 4 * q].n_clusters_bound != c->n_clusters_bound || c->graph[k + 4 * q].throughput_mode != c->throughput_mode) && !c->graph_failed) {
          cudaGraphExec_t dummy; int nl;
          int rc = graph_get(c, ms_evb, &dummy, &nl, k, q);
          if (rc) return rc;
        }
  StepGraph& g = c->graph[slot + 4 * parity];
  if (g.exec && (g.n_clusters_bound != c->n_clusters_bound || g.throughput_mode != c->throughput_mode)) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  if (!g.exec) {
    const long long l0 = c->n_launch;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c->main_stream, cudaStreamCaptureModeThreadLocal);
    int rc = 0;
    const int bound = ms_evb ? std::min(RPB_MAXS, bounds[slot - 1]) : 0;
    if (e == cudaSuccess) {
      const long long seq_h = c->peer.seq[PEER_H], seq_f = c->peer.seq[PEER_F];
      if (sharded) c->peer.seq[PEER_H] = c->peer.seq[PEER_F] = parity ? 0 : 1;   // the captured step's numbers get this parity
      c->evb_s_bound_fixed = bound;
      rc = enqueue_step(c, ms_evb);
      c->evb_s_bound_fixed = 0;
      c->peer.seq[PEER_H] = seq_h; c->peer.seq[PEER_F] = seq_f;                   // (nothing ran: the device-side numbers did not move)
      e = cudaStreamEndCapture(c->main_stream, &graph);
    }
    if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&g.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != 0 || !g.exec) {
      cudaGetLastError();
      g.exec = nullptr;
      c->graph_failed = true;              // fall back to plain launches for the rest of this context's life
      c->n_launch = l0;
      static const bool dbg = getenv("RPB_DEBUG_GRAPH") != nullptr;
      if (dbg) fprintf(stderr, "[rpbmd] step graph capture failed (%s, rc %d): plain launches\n", cudaGetErrorString(e), rc);
      if (rc) return rc;
      *out = nullptr;
      return 0;
    }
    g.launches = (int)(c->n_launch - l0);
    c->n_launch = l0;
    g.s_bound = bound; g.n_clusters_bound = c->n_clusters_bound; g.throughput_mode = c->throughput_mode;
    { static const bool dbg = getenv("RPB_DEBUG_GRAPH") != nullptr; if (dbg) fprintf(stderr, "[rpbmd] step graph captured: %d kernel launches, diabat bound %d (count %d)\n", g.launches, bound, hint); }
  }
  *out = g.exec;
  *launches = g.launches;
  return 0;
}

static int enqueue_steps(rpb_ctx* c, int n_steps, int ms_evb) {
  int s = 0;
  if (graph_allowed(c, ms_evb) && !c->rebuild_forced && n_steps > 0) {
    c->image_valid = false; c->ke_valid = false;
    for (; s < n_steps; s++) {
      cudaGraphExec_t exec = nullptr;
      int launches = 0;
      int rc = graph_get(c, ms_evb, &exec, &launches);     // (sharded steps alternate between the two parities' graphs)
      if (rc) return rc;
      if (!exec) break;
      CK(cudaGraphLaunch(exec, c->main_stream));
      c->n_launch += launches;
      if (ms_evb && c->d.world > 1) { c->peer.seq[PEER_H]++; c->peer.seq[PEER_F]++; }
    }
  }
  for (; s < n_steps; s++) {
    int rc = enqueue_step(c, ms_evb);
    if (rc) return rc;
  }
  return 0;
}

static int step_impl(rpb_ctx* c, int n_steps, int ms_evb);
int rpb_step(rpb_ctx* c, int n_steps, int ms_evb) {
  if (c) c->throughput_mode = false;
  return step_impl(c, n_steps, ms_evb);
}
static int step_impl(rpb_ctx* c, int n_steps, int ms_evb) {
  if (!c->initialized) { c->err = "rpb_initialize not called"; return RPB_ERR_STATE; }
  if (ms_evb && c->d.world > 1 && !c->peer.on) { c->err = "world_size>1: set up the peer-memory exchange (rpb_peer_*) or use the phase calls"; return RPB_ERR_STATE; }
  int rc = enqueue_steps(c, n_steps, ms_evb);
  if (rc) return rc;
  // a caller that downloads the state after every call (a host driver that owns the arrays): the copy is queued right
  // behind the steps and completes under the one synchronisation below, instead of costing a second round trip
  c->image_valid = false;
  if (c->download_streak >= 2) {
    Staging st;
    if ((rc = staging_get(c, st))) return rc;
    CK(cudaMemcpyAsync(st.xq, c->state_block, sb_bytes_all(c->d.N), cudaMemcpyDeviceToHost, c->stream));
    c->image_valid = true;
  }
  c->download_streak = std::max(0, c->download_streak - 1);   // (raised by two per download: a streak survives as long as every call is followed by one)
  rc = collect_results(c, ms_evb);
  c->ke_valid = (rc == 0 && n_steps > 0);     // the last kernel of a step leaves the kinetic energy in its slot
  return rc;
}

// Independent replicas (BASELINE config 5, "replicas only": no communication).  A step needs no host decision, so ONE host
// thread queues every replica's steps round-robin (one graph launch per replica and step) and the replicas' kernels --
// most of them latency-bound single-CTA or small-grid launches -- overlap on the device through the contexts' own streams;
// the results are collected once at the end.  Contexts on several devices, or without step graphs, are driven by one host
// thread each.  Returns the first non-zero status.
int rpb_ensemble_step(rpb_ctx** replicas, int n_replicas, int n_steps, int ms_evb) {
  if (!replicas || n_replicas < 1) return RPB_ERR_ARG;
  bool one_thread = true;
  for (int r = 0; r < n_replicas; r++) {
    rpb_ctx* c = replicas[r];
    if (!c->initialized) { c->err = "rpb_initialize not called"; return RPB_ERR_STATE; }
    c->throughput_mode = n_replicas > 1;
    one_thread = one_thread && c->cfg.device == replicas[0]->cfg.device && graph_allowed(c, ms_evb) && !c->rebuild_forced;
  }
  if (one_thread) {
    if (cudaSetDevice(replicas[0]->cfg.device) != cudaSuccess) { replicas[0]->err = "cudaSetDevice failed"; return RPB_ERR_CUDA; }
    for (int s = 0; s < n_steps; s++)
      for (int r = 0; r < n_replicas; r++) { int rc = enqueue_steps(replicas[r], 1, ms_evb); if (rc) return rc; }
    int first = 0;
    for (int r = 0; r < n_replicas; r++) { int rc = collect_results(replicas[r], ms_evb); if (rc && !first) first = rc; }
    return first;
  }
  std::vector<int> rc(n_replicas, 0);
  std::vector<std::thread> th;
  th.reserve(n_replicas);
  for (int r = 0; r < n_replicas; r++)
    th.emplace_back([&, r]() {
      rpb_ctx* c = replicas[r];
      if (cudaSetDevice(c->cfg.device) != cudaSuccess) { c->err = "cudaSetDevice failed"; rc[r] = RPB_ERR_CUDA; return; }   // the current device is per host thread
      rc[r] = step_impl(c, n_steps, ms_evb);
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < n_replicas; r++) if (rc[r]) return rc[r];
  return 0;
}

int rpb_get_energies(rpb_ctx* c, rpb_energies* e) {
  if (c->ke_valid) { *e = c->last_en; return 0; }
  launch_kinetic_energy(c);
  CK(cudaMemcpyAsync(c->h_en, c->d.en, E_NSLOT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->last_en.kinetic_energy = c->h_en[E_KE];
  *e = c->last_en;
  return 0;
}

int rpb_download_state(rpb_ctx* c, double* xyz, double* velocity, double* force, double* mass, double* charge,
                       int* atom_type_index, int* mol_first_atom, int* mol_n_atom, int* mol_type, int* hydronium_mol) {
  const int N = c->d.N, M = c->d.M;
  Staging st;
  int rc = staging_get(c, st);
  if (rc) return rc;
  if ((rc = refresh_mirror(c))) return rc;
  // a separate pinned block would be needed to keep the upload cache valid: downloads use the tail halves only where they
  // do not alias cached tables (xq, vel, force are re-sent on every upload anyway)
  const bool full = (xyz || charge) && velocity && force;
  if (full) c->download_streak = std::min(4, c->download_streak + 2);
  if (full && c->image_valid) {
    // the image was copied behind the last step (see rpb_step) and nothing touched the device state since
  } else if (full) {     // the usual full download: xq | vel | force are one block on both sides
    CK(cudaMemcpyAsync(st.xq, c->state_block, sb_bytes_all(N), cudaMemcpyDeviceToHost, c->stream));
  } else {
    if (xyz || charge) CK(cudaMemcpyAsync(st.xq, c->d.xq, N * sizeof(double4), cudaMemcpyDeviceToHost, c->stream));
    if (velocity) CK(cudaMemcpyAsync(st.vel, c->d.vel, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (force) CK(cudaMemcpyAsync(st.force, c->d.force, 3 * (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  // masses and atom types only change when a proton hop is committed: while the staging area still mirrors the device
  // tables (state_cache_valid, see rpb_upload_state) they are served from it without a copy
  const bool cached = c->state_cache_valid;
  const bool need_sync = !(full && c->image_valid) || ((mass || atom_type_index) && !cached);
  if (mass && !cached) CK(cudaMemcpyAsync(st.mass, c->d.mass, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (atom_type_index && !cached) CK(cudaMemcpyAsync(st.type, c->d.type, N * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if (need_sync) CK(cudaStreamSynchronize(c->stream));
  if (xyz || charge)
    for (int i = 0; i < N; i++) {
      if (xyz) { xyz[3 * i] = st.xq[i].x; xyz[3 * i + 1] = st.xq[i].y; xyz[3 * i + 2] = st.xq[i].z; }
      if (charge) charge[i] = st.xq[i].w;
    }
  if (velocity) memcpy(velocity, st.vel, 3 * (size_t)N * sizeof(double));
  if (force) memcpy(force, st.force, 3 * (size_t)N * sizeof(double));
  if (mass) memcpy(mass, st.mass, N * sizeof(double));
  if (atom_type_index) for (int i = 0; i < N; i++) atom_type_index[i] = st.type[i] + 1;
  for (int m = 0; m < M; m++) {
    if (mol_first_atom) mol_first_atom[m] = c->mol_first[m] + 1;
    if (mol_n_atom) mol_n_atom[m] = c->mol_natom[m];
    if (mol_type) mol_type[m] = c->mol_type[m] + 1;
  }
  if (hydronium_mol) *hydronium_mol = c->hydronium_mol + 1;
  return 0;
}

int rpb_get_r_com(rpb_ctx* c, double* r_com) {
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(r_com, c->d.r_com, 3 * c->d.M * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int rpb_get_neighbor_list(rpb_ctx* c, int* verlet_point, int* neighbor_list, int capacity, int* n_pairs, int* flag) {
  const int N = c->d.N;
  // the reference's half list in its row order is generated on demand from the positions of the last rebuild
  int rc = launch_verlet_reference_list(c);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  int last = 0, fl = 0, ovf = 0;
  CK(cudaMemcpy(&ovf, c->d.err_flag + 1, sizeof(int), cudaMemcpyDeviceToHost));
  if (ovf) { c->err = "please increase size of verlet neighbor list"; return RPB_ERR_VERLET; }
  CK(cudaMemcpy(&last, c->d.verlet_point + N, sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&fl, c->d.flag_verlet, sizeof(int), cudaMemcpyDeviceToHost));
  int np = last - 1;
  if (n_pairs) *n_pairs = np;
  if (flag) *flag = fl;
  if (verlet_point) CK(cudaMemcpy(verlet_point, c->d.verlet_point, (N + 1) * sizeof(int), cudaMemcpyDeviceToHost));
  if (neighbor_list) {
    if (capacity < np) { c->err = "neighbor_list buffer too small"; return RPB_ERR_ARG; }
    CK(cudaMemcpy(neighbor_list, c->d.neighbor_list, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost));
  }
  return 0;
}

// The cluster-pair list the pair kernel actually consumes, expanded on the host into the half list it encodes:
// pairs (i < j, 1-based) sorted by (i, j).  Parity accessor for tests: must equal the reference's half list as a set.
int rpb_debug_tile_pairs(rpb_ctx* c, int* pair_i, int* pair_j, long long capacity, long long* n_pairs, long long* n_tiles) {
  CK(cudaStreamSynchronize(c->stream));
  int NC = 0;
  CK(cudaMemcpy(&NC, c->d.n_clusters, sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<int> tp(RPB_TILE_PARTS * NC + 1), info(NC);
  CK(cudaMemcpy(tp.data(), c->d.tile_point, tp.size() * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(info.data(), c->d.cl_info, NC * sizeof(int), cudaMemcpyDeviceToHost));
  std::vector<unsigned> tl((size_t)tp[RPB_TILE_PARTS * NC]);
  CK(cudaMemcpy(tl.data(), c->d.tile_list, tl.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
  std::vector<std::pair<int, int>> pr;
  for (int I = 0; I < NC; I++) {
    const int fi = info[I] & 0xffffff;
    for (int k = tp[RPB_TILE_PARTS * I]; k < tp[RPB_TILE_PARTS * (I + 1)]; k++) {
      const int fj = tl[k] & 0x7fffff;
      const unsigned mask = tl[k] >> 23;
      for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++)
        if ((mask >> (3 * a + b)) & 1u) pr.emplace_back(std::min(fi + a, fj + b) + 1, std::max(fi + a, fj + b) + 1);   // each pair is stored once
    }
  }
  std::sort(pr.begin(), pr.end());
  if (n_pairs) *n_pairs = (long long)pr.size();
  if (n_tiles) *n_tiles = (long long)tl.size();
  if (pair_i && pair_j) {
    if (capacity < (long long)pr.size()) { c->err = "pair buffer too small"; return RPB_ERR_ARG; }
    for (size_t k = 0; k < pr.size(); k++) { pair_i[k] = pr[k].first; pair_j[k] = pr[k].second; }
  }
  return 0;
}

int rpb_get_pme(rpb_ctx* c, int state, double* Q_grid, double* theta, double* force_recip) {
  const int K = c->d.K;
  const size_t K3 = (size_t)K * K * K;
  if (state < 1 || state > c->grid_capacity) { c->err = "diabat grid not available"; return RPB_ERR_ARG; }
  CK(cudaStreamSynchronize(c->stream));
  if (Q_grid) {
    if (state > 1 || c->eh.built) { c->err = "Q grids are consumed in place by the batched FFT in MS-EVB mode; only theta is kept"; }
    CK(cudaMemcpy(Q_grid, c->d.Q + K3 * (state - 1), K3 * sizeof(double), cudaMemcpyDeviceToHost));
  }
  if (theta) CK(cudaMemcpy(theta, c->d.theta + K3 * (state - 1), K3 * sizeof(double), cudaMemcpyDeviceToHost));
  if (force_recip) CK(cudaMemcpy(force_recip, c->d.force_recip, 3 * c->d.N * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int rpb_get_evb(rpb_ctx* c, int* n_states, double* hamiltonian, double* eigenvector, int* proton_log, int* coupling_matrix,
                int* principal_diabat, int* new_hydronium_mol, double* adiabatic_potential) {
  EvbHost& h = c->eh;
  const int S = h.n_states;
  if (n_states) *n_states = S;
  if (hamiltonian) for (int i = 0; i < RPB_MAXS; i++) for (int j = 0; j < RPB_MAXS; j++) hamiltonian[i + RPB_MAXS * j] = h.hamiltonian[i][j];
  if (eigenvector) for (int s = 0; s < S; s++) eigenvector[s] = h.evec[s];
  if (proton_log)
    for (int s = 0; s < RPB_MAXS; s++) for (int k = 0; k < RPB_MAXC; k++) for (int f = 0; f < 5; f++) {
      int v = (s < S) ? h.proton_log[s][k][f] : -1;
      proton_log[s + RPB_MAXS * k + RPB_MAXS * RPB_MAXC * f] = v < 0 ? -1 : v + 1;
    }
  if (coupling_matrix) for (int s = 0; s < RPB_MAXS; s++) coupling_matrix[s] = (s < S && h.parent[s] >= 0) ? h.parent[s] + 1 : -1;
  if (principal_diabat) *principal_diabat = h.principal_diabat + 1;
  if (new_hydronium_mol) *new_hydronium_mol = h.new_hydronium + 1;
  if (adiabatic_potential) *adiabatic_potential = h.adiabatic_potential;
  return 0;
}

int rpb_debug_mix_forces(rpb_ctx* c, const double* coeff, double* force) { return evb_mix(c, coeff, force); }

int rpb_get_launch_counts(rpb_ctx* c, long long* own, long long* fft) {
  if (own) *own = c->n_launch;
  if (fft) *fft = c->n_fft;
  return 0;
}
static void timers_resolve(rpb_ctx* c) {
  if (c->ev_used == 0) return;
  cudaStreamSynchronize(c->stream);
  if (getenv("RPB_DEBUG_TIMELINE")) {   // offsets of every recorded interval from the start of the last step (all streams)
    int last_step = -1;
    for (int k = 0; k < c->ev_used; k++) if (c->ev_id[k] == T_STEP) last_step = k;
    if (last_step >= 0) {
      fprintf(stderr, "[timeline, us from step start]");
      for (int k = last_step; k < c->ev_used; k++) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->ev_pool[2 * last_step], c->ev_pool[2 * k]);
        cudaEventElapsedTime(&b, c->ev_pool[2 * last_step], c->ev_pool[2 * k + 1]);
        fprintf(stderr, " %s:%.0f-%.0f", k_timer_names[c->ev_id[k]], a * 1e3, b * 1e3);
      }
      fprintf(stderr, "\n");
    }
  }
  for (int k = 0; k < c->ev_used; k++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev_pool[2 * k], c->ev_pool[2 * k + 1]) == cudaSuccess) { c->t_ms[c->ev_id[k]] += ms; c->t_calls[c->ev_id[k]]++; }
  }
  c->ev_used = 0;
}
int rpb_timers_enable(rpb_ctx* c, int on) {
  if (on && c->ev_pool.empty()) {
    c->ev_pool.resize(2 * 8192);
    c->ev_id.resize(8192);
    for (auto& ev : c->ev_pool) if (cudaEventCreate(&ev) != cudaSuccess) { c->err = "cudaEventCreate failed"; return RPB_ERR_CUDA; }
  }
  if (!on) timers_resolve(c);
  c->timers_on = on != 0;
  return 0;
}
int rpb_timers_reset(rpb_ctx* c) { timers_resolve(c); for (int i = 0; i < T_NTIMER; i++) { c->t_ms[i] = 0; c->t_calls[i] = 0; } return 0; }
int rpb_timers_get(rpb_ctx* c, double* ms, long long* calls) {
  timers_resolve(c);
  for (int i = 0; i < T_NTIMER; i++) { ms[i] = c->t_ms[i]; calls[i] = c->t_calls[i]; }
  return 0;
}
void* rpb_get_stream(rpb_ctx* c) { return (void*)c->stream; }
int rpb_measure_fp64_peak(rpb_ctx* c, double* tflops) { return measure_fp64_peak(c, tflops); }

}  // extern "C"
