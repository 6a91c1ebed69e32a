// Host-side context of librpbmd.so and the launcher prototypes of every kernel file.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "rpb_dev.cuh"
#include "rpb_evb.cuh"

enum {
  T_INTEGRATE = 0, T_VERLET, T_PAIR, T_INTRA, T_SPREAD, T_FFT, T_CONV, T_GATHER, T_EVB_ENUM, T_EVB_ITEMS_BG,
  T_EVB_CAND, T_EVB_BCAST, T_EVB_PATCH, T_EVB_CORR, T_EVB_COUPLING, T_EVB_DIAG, T_EVB_THETAMIX, T_EVB_MIXF,
  T_EVB_GATHERMIX, T_EVB_SNAP, T_EVB_COUPLING_GEO, T_EVB_ASSEMBLE, T_STEP, T_NTIMER
};

// Peer-memory exchange of the state-sharded step (kernels_peer.cu)
enum { PEER_H = 0, PEER_F = 1 };
struct PeerExchange {
  bool on = false;
  int world = 1;
  double* arena = nullptr;                 // this rank's arena (flags + 2 parities of each partial)
  double* peer[RPB_MAX_RANKS] = {};        // every rank's arena as mapped here (peer[rank] == arena)
  bool opened[RPB_MAX_RANKS] = {};         // mapped with cudaIpcOpenMemHandle (closed in peer_free)
  double* h_total = nullptr;               // all-reduced Hamiltonian block
  double* h_local = nullptr; double* f_local = nullptr;   // the library's own (non-arena) exchange buffers
  long long seq[2] = {0, 0};               // collectives issued per kind (host mirror of seq_dev: selects the arena parity)
  unsigned long long* seq_dev = nullptr;   // device-side sequence numbers, ticked by a kernel ahead of each collective
  bool f_reduced_in_place = false;         // the force all-reduce of this step wrote d.force directly (evb_commit skips its copy)
  size_t off_flags = 0, off[2] = {0, 0}, arena_doubles = 0;
  int n[2] = {0, 0}, n_act[2] = {0, 0};    // stride / length in doubles of one partial
};

// one MD step captured as a CUDA graph (rpb_api.cu)
struct StepGraph { cudaGraphExec_t exec = nullptr; int launches = 0; int s_bound = 0; int n_clusters_bound = 0; bool throughput_mode = false; };

struct rpb_ctx {
  rpb_config cfg;
  std::string err;
  bool forces_zeroed = false;   // k_integrate_first cleared d.force / d.en: the next force evaluation skips its zeroing launches
  bool have_tables = false, have_ff = false, have_mt = false, have_evb = false, have_state = false, initialized = false;
  cudaStream_t stream = nullptr;        // stream the launchers use (normally the main stream; see StreamScope)
  cudaStream_t main_stream = nullptr;   // the library's main stream: host synchronisation and timing happen here
  cudaStream_t aux[5] = {};   // side streams for the independent branches of a force evaluation; MS-EVB: [2] bonded terms of the principal diabat, [3] read-backs, [4] per-diabat real-space deltas
  cudaEvent_t ev_sync[24] = {};   // fork / join points
  cudaEvent_t ev_enum = nullptr;        // enumeration results have reached pinned host memory
  Dev d;                       // device pointer table (host copy, passed by value to kernels)
  std::vector<void*> allocs;   // everything cudaMalloc'ed (freed in rpb_destroy)
  // host copies of small tables
  std::vector<MolTypeDev> mt_host;
  EvbTables evb_host;
  std::vector<double> ff_vdw, ff_vdw14, ff_bondp, ff_anglep, ff_dihp;
  std::vector<int> ff_vdwt, ff_bondt, ff_anglet, ff_diht;
  double atype_chg[RPB_MAXT]; int atype_freeze[RPB_MAXT];
  // host mirror of the molecule table (kept in sync on hop commit)
  std::vector<int> mol_first, mol_natom, mol_type;
  int hydronium_mol = -1;
  bool rebuild_forced = false;  // rpb_upload_state brought a different molecule table: the next force evaluation rebuilds the list
  int n_sm = 148;              // multiprocessors of this context's device (persistent grids)
  int n_clusters_bound = 0;    // upper bound of the number of atom clusters (kernels_nlist.cu) for grid sizing; the count itself lives on the device
  // cuFFT
  std::map<int, cufftHandle> plan_fwd, plan_inv;   // keyed by (rounded) batch size
  char* fft_work = nullptr; size_t fft_work_bytes = 0;   // work area shared by every plan
  // EVB
  EvbDev e;                    // device pointers of the EVB working set
  EvbHost eh;                  // pinned host read-back area + the last evaluated step's results
  void* evb_scratch = nullptr; // EvbScratch (kernels_evb.cu): device scratch of the MS-EVB build, owned by this context
  bool evb_overlap_solver = false;    // the branches of evb_build were joined on aux[0] (not on the main stream): evb_mix runs the solver there
  bool evb_assemble_pending = false;  // evb_build left the Hamiltonian assembly to the solver kernel
  bool throughput_mode = false;       // replica ensembles (rpb_ensemble_step): other replicas fill the SMs, so the pair kernel takes a full wave
  int evb_s_bound_fixed = 0;          // while a step graph is captured: the diabat-count bound its grids are sized for (0: from the last count)
  bool evb_h_exchange_in_solver = false;   // sharded step, tree solver: assembly + Hamiltonian all-reduce run in the solver kernel's prologue
  bool evb_join_pending = false;      // evb_commit queued its read-back copies on aux[3]; the main stream has not joined them yet
  bool evb_quirk_types_present = true; // ... and such a type (or the conjugate of a present type) occurs in the current molecule table (rpb_upload_state)
  std::vector<int> conj_pairs_host;    // evb_conjugate_pairs, 0-based (-1 none)
  bool evb_any_multi_basic = false;   // some molecule type has more than one atom that can be protonated (reference re-ordering quirk possible)
  bool mirror_stale = false;          // a committed hop changed the molecule table on the device: the host mirror is refreshed before use
  int grid_capacity = 0;       // number of K^3 grids usable in d.Q / d.theta (4 spare ones follow for the rounded FFT batch)
  int evb_solver = 0;          // 0: tree-structured ground-state solver (default)  1: block Jacobi (RPB_EVB_SOLVER=jacobi)
  std::vector<char> mt_multi_basic;   // per molecule type: more than one atom that can be protonated
  double evb_rcand = 0.0;      // candidate-list radius of the diabat real-space deltas
  double evb_rep_reach = 0.0;  // largest cutoff of the EVB proton-acceptor repulsion
  // pinned scratch
  double* h_en = nullptr;      // [E_NSLOT]
  int* h_flags = nullptr;      // [8]
  void* staging = nullptr;     // pinned staging area of rpb_upload_state / rpb_download_state (also caches the last uploaded tables)
  void* staging_up[2] = {nullptr, nullptr};   // double-buffered pinned images of {xq, vel} for uploads: no wait for the previous upload's copy
  cudaEvent_t ev_up[2] = {nullptr, nullptr}; int up_parity = 0;
  char* state_block = nullptr; // device: xq | vel | force in ONE allocation, so that a state transfer is one copy each way
  std::vector<int> raw_type, raw_mol; std::vector<double> raw_mass;   // the caller's tables as last uploaded (change detection by memcmp)
  bool image_valid = false;    // the pinned staging area holds the device state {xq, vel, force} as of the end of the last rpb_step
  int download_streak = 0;     // > 1: the caller downloads the full state after every call
  bool ke_valid = false;       // last_en.kinetic_energy belongs to the current velocities (computed by the step's last kernel)
  bool serial_streams = false;
  StepGraph graph[8];          // [0] non-reactive step, [1..3] MS-EVB step with grids sized for <= 32 / 56 / evb_max_states diabats;
                               // [4 + k]: the same for the odd arena parity of a state-sharded step (kernels_peer.cu)
  bool graph_failed = false;   // stream capture of a step did not work on this context: plain launches
  bool state_cache_valid = false;   // the staging area mirrors the per-atom / per-molecule tables on the device
  // measurement
  long long n_launch = 0, n_fft = 0;
  bool timers_on = false;
  std::vector<cudaEvent_t> ev_pool;   // 2 events per recorded interval
  std::vector<int> ev_id;
  int ev_used = 0;
  double t_ms[T_NTIMER];
  long long t_calls[T_NTIMER];
  rpb_energies last_en;
  PeerExchange peer;
};

template <typename T>
int dev_alloc(rpb_ctx* ctx, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  ctx->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}

// Launchers always use ctx->stream; a StreamScope redirects them to a side stream for the lifetime of the scope.
struct StreamScope {
  rpb_ctx* c; cudaStream_t saved;
  StreamScope(rpb_ctx* c_, cudaStream_t s) : c(c_), saved(c_->stream) { c->stream = s; }
  ~StreamScope() { c->stream = saved; }
};
// `to` waits for everything queued on `from` so far (fork or join, depending on the direction)
inline void stream_depend(rpb_ctx* c, int ev, cudaStream_t from, cudaStream_t to) {
  cudaEventRecord(c->ev_sync[ev], from);
  cudaStreamWaitEvent(to, c->ev_sync[ev], 0);
}

// Per-phase CUDA-event timers.  Events are only RECORDED on the launching stream while a step runs (no
// synchronisation, so enabling them does not serialise the step); rpb_timers_get resolves them afterwards.
struct ScopedTimer {
  rpb_ctx* c; int id; int slot;
  ScopedTimer(rpb_ctx* c_, int id_) : c(c_), id(id_), slot(-1) {
    if (c->timers_on && c->ev_used < (int)c->ev_pool.size() / 2) {
      slot = c->ev_used++;
      c->ev_id[slot] = id;
      cudaEventRecord(c->ev_pool[2 * slot], c->stream);
    }
  }
  ~ScopedTimer() { if (slot >= 0) cudaEventRecord(c->ev_pool[2 * slot + 1], c->stream); }
};

// ---- rpb_api.cu
int calculate_total_force_energy(rpb_ctx*, bool evb_principal);   // total_energy_forces.f90:19-99
// ---- kernels_md.cu
void launch_integrate_first(rpb_ctx*);      // md_integration.f90:469-491
void launch_integrate_second(rpb_ctx*);     // md_integration.f90:507-532
void launch_update_com_shift(rpb_ctx*, bool shift);
// ---- kernels_nlist.cu
int verlet_setup(rpb_ctx*);                 // per-context sizing of the cooperative rebuild kernels
int launch_verlet_update(rpb_ctx*);         // total_energy_forces.f90:30-39
int launch_verlet_force_rebuild(rpb_ctx*);  // construct_verlet_list + displacement init
struct CommitArgs;
int launch_commit_and_rebuild(rpb_ctx*, const CommitArgs*);   // hop commit + the forced rebuild of ms_evb.f90:223-225 in ONE cooperative launch; acts only if the device-side hop flag is set
int launch_verlet_reference_list(rpb_ctx*); // parity accessor: the reference's half list in its row order
void launch_zero_forces(rpb_ctx*);
void launch_kinetic_energy(rpb_ctx*);
int measure_fp64_peak(rpb_ctx*, double* tflops);
// ---- kernels_pair.cu
void launch_pair_verlet(rpb_ctx*, bool shard_by_rank);   // pair_int_real_space.f90:135-371; shard: rank r takes clusters [NC r/R, NC (r+1)/R)
void launch_molecule_terms(rpb_ctx*);       // pair_int_real_space.f90:386-588 + intra_bonded_interactions.f90:17-552
// ---- kernels_pme.cu
int pme_round_batch(int batch);
int pme_get_plans(rpb_ctx*, int batch, cufftHandle* fwd, cufftHandle* inv);
void launch_scaled_coords(rpb_ctx*);
void launch_spread_principal(rpb_ctx*);     // pme.f90:184-264
int launch_convolve(rpb_ctx*, int first_grid, int n_grids, double* e_recip_dev, bool inverse);  // pme.f90:73-129
void launch_gather(rpb_ctx*, const double* theta, double* out_force, bool add_to_force);       // pme.f90:346-498
// ---- kernels_fft.cu
int fft_conv_supported(rpb_ctx*);   // 1: hand-written batched FFT convolution handles this grid size, 0: cuFFT path
int fft_conv_batched(rpb_ctx*, int first_grid, int n_grids, double* e_recip_dev, bool inverse);
void fft_conv_free(rpb_ctx*);
// ---- kernels_peer.cu
void peer_begin(rpb_ctx*, int kind);      // producers of partial `kind` write into this step's parity of the arena
int peer_allreduce(rpb_ctx*, int kind);   // one kernel: signal, wait, pull + add in rank order
void peer_args_h(rpb_ctx*, void* peer_args_out);   // the Hamiltonian exchange as an argument block for the solver kernel (rpb_peer.cuh)
void peer_free(rpb_ctx*);
// ---- kernels_evb.cu
int evb_alloc(rpb_ctx*);
void evb_free(rpb_ctx*);
int evb_enumerate_async(rpb_ctx*, int part);   // launched early in the principal evaluation (part 0: the kernel, part 1: read-back on aux[3], images); evb_build waits for its read-back
void evb_clear_early(rpb_ctx*);      // accumulators of the build for the previous S + margin (launch on a stream with slack)
int evb_build(rpb_ctx*);
int evb_mix(rpb_ctx*, const double* coeff_override_host, double* force_out_host);
int evb_commit(rpb_ctx*);
void evb_join_readback(rpb_ctx*);   // main stream waits for the read-back copies evb_commit queued on aux[3]
int evb_readback(rpb_ctx*);          // synchronising read of the last step's results + sticky error flags
