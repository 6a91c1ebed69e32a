// Peer-memory exchange of the state-sharded MS-EVB step: the two all-reduce(sum) steps of
// ms_evb_calculate_total_force_energy -- Hamiltonian elements after the per-diabat loops (ms_evb.f90:657-687,
// 2026-2086) and the Hellmann-Feynman partial forces (ms_evb.f90:292-309) -- done by ONE kernel each over NVLink /
// NVSwitch peer memory instead of a library collective:
//
//   * every rank owns an "arena" (cudaMalloc, exported with cudaIpcGetMemHandle, opened by the peers): two parities
//     of its partial Hamiltonian block and partial force, plus one arrival flag per (kind, peer);
//   * the kernel that produces the partial (k_evb_assemble / k_evb_mix_forces + k_evb_gather_mix) writes straight into
//     the arena; k_peer_allreduce then (1) releases "partial #seq is complete" into every peer's flag slot with a
//     system-scope store, (2) waits for the peers' flags in its own arena, (3) pulls the peers' partials through
//     NVLink (ld.relaxed.sys, 16 B per lane) and adds them in RANK ORDER, so every rank obtains the bit-identical sum
//     -- the replicated parts of the step (solver, integrator, enumeration) stay in lock-step without any broadcast;
//   * double buffering by the parity of the sequence number makes a trailing barrier unnecessary: a rank can only
//     overwrite parity p after it has passed collective seq+1, which needs every peer's flag seq+1, which a peer sets
//     only after its own collective seq (the reads of parity p) has finished in stream order.
//
// The host never takes part: the whole sharded step runs inside rpb_step() like the single-GPU step.  A peer that
// never arrives (crashed rank) is reported after PEER_TIMEOUT_NS instead of hanging the device.
#include <cstring>
#include "rpb_host.h"
#include "rpb_peer.cuh"

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerArgs a) {
  const unsigned long long seq = *a.seq_ptr + 1ull;   // (updated by the last block of the previous collective of this kind)
  // (1) announce: everything the earlier kernels of this stream wrote into the local arena is complete
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flag_at[threadIdx.x], seq);
  }
  // (2) wait for every rank's partial #seq
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < a.world; r++)
      while (ld_acquire_sys(&a.my_flags[r]) < seq) {
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) { atomicMax(&a.err_flag[3], 30 + r); break; }
        __nanosleep(64);
      }
  }
  __syncthreads();
  // (3) pull and add in rank order
  const int n2 = a.n >> 1;
  if ((a.n & 1) && blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {   // odd tail (3N of an odd atom count)
    double s = ld_relaxed_sys_f64(a.part[0] + a.n - 1);
    for (int r = 1; r < a.world; r++) s += ld_relaxed_sys_f64(a.part[r] + a.n - 1);
    a.out[a.n - 1] = s;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
    double2 s = ld_relaxed_sys_f64x2(a.part[0] + 2 * (size_t)i);
    for (int r = 1; r < a.world; r++) {
      const double2 v = ld_relaxed_sys_f64x2(a.part[r] + 2 * (size_t)i);
      s.x += v.x; s.y += v.y;
    }
    reinterpret_cast<double2*>(a.out)[i] = s;
  }
  // the last block to finish (every block has read the old number by then) publishes the new sequence number
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.done, 1u) == gridDim.x - 1) { *a.seq_ptr = seq; *a.done = 0u; __threadfence(); }
  }
}

static size_t pad2(size_t n) { return (n + 1) & ~(size_t)1; }

static int peer_layout(rpb_ctx* c) {
  PeerExchange& p = c->peer;
  p.n_act[PEER_H] = 3 * RPB_MAXS + E_NSLOT;
  p.n_act[PEER_F] = 3 * c->d.N;
  p.n[PEER_H] = (int)pad2(p.n_act[PEER_H]);       // strides keep every parity 16-byte aligned
  p.n[PEER_F] = (int)pad2(p.n_act[PEER_F]);
  p.off_flags = 0;
  p.off[PEER_H] = pad2(2 * RPB_MAX_RANKS);                 // flags: [2 kinds][RPB_MAX_RANKS] x 8 bytes
  p.off[PEER_F] = p.off[PEER_H] + 2 * (size_t)p.n[PEER_H];
  p.arena_doubles = p.off[PEER_F] + 2 * (size_t)p.n[PEER_F];
  return 0;
}

static int peer_alloc(rpb_ctx* c) {
  PeerExchange& p = c->peer;
  if (p.arena) return 0;
  peer_layout(c);
  // a dedicated cudaMalloc (not a sub-allocation): the IPC handle exports the whole allocation
  cudaError_t e = cudaMalloc((void**)&p.arena, p.arena_doubles * sizeof(double));
  if (e != cudaSuccess) { c->err = std::string("peer arena cudaMalloc: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  e = cudaMemset(p.arena, 0, p.arena_doubles * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p.h_total, p.n[PEER_H] * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p.seq_dev, 4 * sizeof(unsigned long long));     // [2] numbers, [2] block counters
  if (e == cudaSuccess) e = cudaMemset(p.seq_dev, 0, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { c->err = std::string("peer arena init: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  p.h_local = c->e.h_diag; p.f_local = c->e.f_mix;
  return 0;
}

static int peer_enable(rpb_ctx* c, int world) {
  PeerExchange& p = c->peer;
  p.world = world;
  p.peer[c->d.rank] = p.arena;
  p.seq[PEER_H] = p.seq[PEER_F] = 0;
  cudaMemset(p.seq_dev, 0, 4 * sizeof(unsigned long long));
  p.on = true;
  return 0;
}

void peer_free(rpb_ctx* c) {
  PeerExchange& p = c->peer;
  for (int r = 0; r < RPB_MAX_RANKS; r++)
    if (p.opened[r] && p.peer[r]) { cudaIpcCloseMemHandle(p.peer[r]); p.peer[r] = nullptr; p.opened[r] = false; }
  if (p.arena) { cudaFree(p.arena); p.arena = nullptr; }
  if (p.h_total) { cudaFree(p.h_total); p.h_total = nullptr; }
  if (p.seq_dev) { cudaFree(p.seq_dev); p.seq_dev = nullptr; }
  p.on = false;
}

// Called before the phase that produces partial `kind`: the producers write into this step's parity of the arena.
void peer_begin(rpb_ctx* c, int kind) {
  PeerExchange& p = c->peer;
  p.seq[kind]++;                                                  // host mirror of the device-side number: selects the parity
  double* part = p.arena + p.off[kind] + (size_t)(p.seq[kind] & 1) * p.n[kind];
  if (kind == PEER_H) c->e.h_diag = part; else c->e.f_mix = part;
}

// The collective itself, on the main stream.  PEER_H: the sum lands in a local block that the solver reads through
// e.h_diag; PEER_F: the sum lands in d.force (the principal diabat's partial force was saved to dF slot 0 by evb_build).
static void peer_fill_args(rpb_ctx* c, int kind, PeerArgs& a) {
  PeerExchange& p = c->peer;
  memset(&a, 0, sizeof(a));
  const size_t part_off = p.off[kind] + (size_t)(p.seq[kind] & 1) * p.n[kind];
  for (int r = 0; r < p.world; r++) {
    a.part[r] = p.peer[r] + part_off;
    a.flag_at[r] = reinterpret_cast<unsigned long long*>(p.peer[r] + p.off_flags) + kind * RPB_MAX_RANKS + c->d.rank;
  }
  a.my_flags = reinterpret_cast<const unsigned long long*>(p.arena + p.off_flags) + kind * RPB_MAX_RANKS;
  a.out = (kind == PEER_H) ? p.h_total : c->d.force;
  a.err_flag = c->d.err_flag;
  a.seq_ptr = p.seq_dev + kind;
  a.done = reinterpret_cast<unsigned int*>(p.seq_dev + 2 + kind);
  a.n = p.n_act[kind];
  a.world = p.world;
}

// argument block of the Hamiltonian exchange for the solver kernel, which runs the exchange in its prologue
void peer_args_h(rpb_ctx* c, void* out) { peer_fill_args(c, PEER_H, *static_cast<PeerArgs*>(out)); c->e.h_diag = c->peer.h_total; }

int peer_allreduce(rpb_ctx* c, int kind) {
  PeerExchange& p = c->peer;
  PeerArgs a;
  peer_fill_args(c, kind, a);
  const int blocks = std::max(1, std::min((a.n / 2 + 255) / 256, 148 * 2));
  k_peer_allreduce<<<blocks, 256, 0, c->main_stream>>>(a);
  c->n_launch++;
  if (kind == PEER_H) c->e.h_diag = p.h_total; else p.f_reduced_in_place = true;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { c->err = std::string("k_peer_allreduce: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  return 0;
}

extern "C" {

int rpb_peer_export(rpb_ctx* c, void* handle_out) {
  if (!c || !handle_out) return RPB_ERR_ARG;
  if (!c->have_evb) { c->err = "rpb_set_evb must precede rpb_peer_export"; return RPB_ERR_STATE; }
  static_assert(sizeof(cudaIpcMemHandle_t) == RPB_PEER_HANDLE_BYTES, "handle size");
  int rc = peer_alloc(c);
  if (rc) return rc;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->peer.arena);
  if (e != cudaSuccess) { c->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int rpb_peer_import(rpb_ctx* c, const void* handles, int world_size) {
  if (!c || !handles) return RPB_ERR_ARG;
  if (world_size != c->d.world || world_size > RPB_MAX_RANKS) { c->err = "rpb_peer_import: world size mismatch (or more than RPB_MAX_RANKS ranks)"; return RPB_ERR_ARG; }
  if (!c->peer.arena) { c->err = "rpb_peer_export must precede rpb_peer_import"; return RPB_ERR_STATE; }
  PeerExchange& p = c->peer;
  for (int r = 0; r < world_size; r++) {
    if (r == c->d.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * RPB_PEER_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      c->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
      return RPB_ERR_CUDA;
    }
    p.peer[r] = (double*)ptr; p.opened[r] = true;
  }
  return peer_enable(c, world_size);
}

int rpb_peer_attach_local(rpb_ctx** ranks, int world_size) {
  if (!ranks || world_size < 1 || world_size > RPB_MAX_RANKS) return RPB_ERR_ARG;
  for (int r = 0; r < world_size; r++) {
    rpb_ctx* c = ranks[r];
    if (!c || c->d.rank != r || c->d.world != world_size) { if (c) c->err = "rpb_peer_attach_local: contexts must be passed in rank order"; return RPB_ERR_ARG; }
    if (!c->have_evb) { c->err = "rpb_set_evb must precede rpb_peer_attach_local"; return RPB_ERR_STATE; }
    int rc = peer_alloc(c);
    if (rc) return rc;
  }
  for (int r = 0; r < world_size; r++) {
    rpb_ctx* c = ranks[r];
    for (int q = 0; q < world_size; q++) {
      if (q != r && ranks[q]->cfg.device != c->cfg.device) {
        cudaSetDevice(c->cfg.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(ranks[q]->cfg.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
        cudaGetLastError();
      }
      c->peer.peer[q] = ranks[q]->peer.arena;
    }
    peer_enable(c, world_size);
  }
  return 0;
}

int rpb_peer_enabled(rpb_ctx* c) { return (c && c->peer.on) ? 1 : 0; }

}  // extern "C"
