// Peer-memory exchange: argument block and device helpers shared by k_peer_allreduce (kernels_peer.cu) and the solver
// kernel, which performs the Hamiltonian exchange in its own prologue (kernels_evb.cu).
#pragma once
#include "rpb_dev.cuh"

#define PEER_TIMEOUT_NS 20000000000ull

struct PeerArgs {
  const double* part[RPB_MAX_RANKS];          // partial of rank r (this parity), as mapped into this process
  unsigned long long* flag_at[RPB_MAX_RANKS]; // flag slot [kind][my rank] inside rank r's arena
  const unsigned long long* my_flags;         // [kind][0..world) in the local arena
  double* out;
  int* err_flag;
  unsigned long long* seq_ptr;                // device-side sequence number of this kind's LAST collective: a captured step graph replays with fresh numbers
  unsigned int* done;                         // blocks of this launch that have finished (the last one publishes the new number)
  int n;                                      // number of doubles
  int world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}


// All-reduce(sum) by ONE CTA inside another kernel (small blocks: the Hamiltonian elements).  Every thread of the CTA must
// call it; on return a.out holds the rank-ordered sum and the CTA is synchronised.
__device__ __forceinline__ void peer_allreduce_cta(const PeerArgs& a) {
  const unsigned long long seq = *a.seq_ptr + 1ull;
  __syncthreads();                                  // the partial was written by this CTA
  if (threadIdx.x < a.world) { __threadfence_system(); st_release_sys(a.flag_at[threadIdx.x], seq); }
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < a.world; r++)
      while (ld_acquire_sys(&a.my_flags[r]) < seq) {
        if (global_timer_ns() - t0 > PEER_TIMEOUT_NS) { atomicMax(&a.err_flag[3], 30 + r); break; }
        __nanosleep(64);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    double s = ld_relaxed_sys_f64(a.part[0] + i);
    for (int r = 1; r < a.world; r++) s += ld_relaxed_sys_f64(a.part[r] + i);
    a.out[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) { *a.seq_ptr = seq; __threadfence(); }
  __syncthreads();
}
