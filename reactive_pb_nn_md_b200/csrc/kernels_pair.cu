// Real-space pair forces over the cluster-pair (tile) list, and per-molecule bonded / intramolecular terms.
// Replaces:
//   pairwise_real_space_verlet            src/pair_int_real_space.f90:135-371
//   pairwise_real_space_{ewald,LJ,sapt}   src/pair_int_real_space.f90:621-759
//   intra_molecular_pairwise_energy_force src/pair_int_real_space.f90:386-588
//   intra_molecular_energy_force          src/intra_bonded_interactions.f90:17-552
//
// Pair kernel (k_pair_tiles): works on the HALF list of cluster-pair tiles (kernels_nlist.cu: a tile = cluster I x cluster J,
// <= 3 consecutive atoms of one molecule each, with a 9-bit mask of exactly the reference's listed atom pairs; one list
// word serves nine atom pairs, every listed pair is stored and evaluated once).  Per listed pair: minimum image with a
// reciprocal box (the shift can only differ from the reference's division for |dr| ~ L/2, far outside the cutoff) and
// the cutoff test dr^2 < r_c^2.  Per in-cutoff pair: ONE rsqrt replaces the sqrt and the six divisions of
// pair_int_real_space.f90:621-645,698-759 (relative differences ~1e-16, the interpolated tables are continuous across
// bins), and the erfc / ewaldscale tables are read interleaved (one 32-byte load per pair).  FP64 work -- see DESIGN.md.
#include <cmath>
#include <cstdlib>
#include "rpb_host.h"
#include "rpb_bonded.cuh"

// cold path of the pair kernel: a SAPT row with non-zero coefficients (pairwise_real_space_sapt :651-690; the example
// force field has none).  Out of line and fed with scalars so that it costs the hot loop no registers.
// (returns {energy, force factor} BY VALUE: reference parameters would pin the caller's loop-carried accumulators to the stack)
__device__ __noinline__ double2 sapt_pair(const double* __restrict__ tt_t, const double* __restrict__ dtt_t, double tt_max, int tt_grid,
                                          double dr2, const double* par) {
  const double A = par[0], B = par[1], C6 = par[2], C8 = par[3], C10 = par[4], C12 = par[5];
  const double r = sqrt(dr2);
  const double dr6 = dr2 * dr2 * dr2, dr8 = dr6 * dr2, dr10 = dr8 * dr2, dr12 = dr10 * dr2;
  const int idx = (int)ceil(B * r / tt_max * (double)tt_grid);
  const double* tt = &tt_t[4 * (idx - 1)];
  const double* dt = &dtt_t[4 * (idx - 1)];
  const double ex = exp(-1 * B * r);
  const double e_vdw = A * ex - tt[0] * C6 / dr6 - tt[1] * C8 / dr8 - tt[2] * C10 / dr10 - tt[3] * C12 / dr12;
  const double fac = r * A * B * ex + r * (B * dt[0]) * C6 / dr6 - tt[0] * 6.0 * C6 / dr6 + r * (B * dt[1]) * C8 / dr8 -
                     tt[1] * 8.0 * C8 / dr8 + r * (B * dt[2]) * C10 / dr10 - tt[2] * 10.0 * C10 / dr10 +
                     r * (B * dt[3]) * C12 / dr12 - tt[3] * 12.0 * C12 / dr12;
  return make_double2(e_vdw, fac / dr2);
}

// 1/sqrt(x) for x in the range of squared pair distances (normal, far from the exponent limits): the hardware's fp64
// reciprocal-square-root seed (MUFU.RSQ64H, ONE conversion-unit instruction, ~2^-22) and two Newton-Raphson steps in fp64
// -- no special-case branches, no float conversions, ~2 ulp (the library routine: 1 ulp)
__device__ __forceinline__ double rsqrt_pair(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// sum of nine per-lane values over the warp; on return lane k (k < 9) holds the total of v[k].  Transposing butterfly:
// at each of the first three levels a lane keeps half of its values and hands the other half to its partner, so the
// eight values v[0..7] cost 4 + 2 + 1 + 2 = 9 exchanges instead of 40 (v[8]: a plain butterfly).
__device__ __forceinline__ double warp_sum9(const double (&v)[9], const int lane) {
  double w4[4], w2[2], w1;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int k = 0; k < 4; k++) { const double keep = up ? v[k + 4] : v[k], give = up ? v[k] : v[k + 4]; w4[k] = keep + __shfl_xor_sync(0xffffffffu, give, 16); }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int k = 0; k < 2; k++) { const double keep = up ? w4[k + 2] : w4[k], give = up ? w4[k] : w4[k + 2]; w2[k] = keep + __shfl_xor_sync(0xffffffffu, give, 8); }
  }
  {
    const bool up = lane & 4;
    const double keep = up ? w2[1] : w2[0], give = up ? w2[0] : w2[1];
    w1 = keep + __shfl_xor_sync(0xffffffffu, give, 4);
  }
  w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  // lane l now holds the total of value 4*(l>>4 & 1) + 2*(l>>3 & 1) + (l>>2 & 1)
  double x = v[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  // bring value k to lane k
  const int src = ((lane & 4) ? 16 : 0) | ((lane & 2) ? 8 : 0) | ((lane & 1) ? 4 : 0);
  const double t = __shfl_sync(0xffffffffu, w1, src);
  return lane == 8 ? x : t;
}

// Ampere-style asynchronous copies global -> shared (LDGSTS): no register staging, no scoreboard stall at the issue
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The grid is one resident wave; warp g takes the g-th of (warps of the grid) equal, contiguous ranges of the tile list
// (balanced in tiles whatever the row lengths; a work counter would cost one same-address atomic per piece, which the L2
// serialises).  A range is cut into pieces at the cluster boundaries: within a piece the cluster I (<= 3 atoms) is
// broadcast from shared memory.  Every listed pair is stored ONCE (the list is a half list of tiles, kernels_nlist.cu),
// so a pair is evaluated once: F_I accumulates in registers over the piece, F_J of a lane's atom goes to d.force with
// three atomic adds per visit.
//   1. tile cull -- a lane takes one tile (cluster J): minimum-image distance of the two first atoms against
//      r_c + 2 x (largest cluster extent now).  ~37 % of the listed tiles lie wholly between the cutoff and the list
//      radius and are dropped here at the cost of one 32-byte gather and ~20 fp64 operations per NINE pair slots; the
//      survivors' list words are compacted (ballot + prefix count) into a per-warp ring in shared memory, and the
//      coordinates / types of their (up to) three atoms are fetched into the ring by asynchronous copies;
//   2. interaction, one cull iteration behind (the copies have landed) -- 32 J atoms per round: a lane takes one atom of
//      one tile and its three pairs with the atoms of I (three independent dependency chains per lane).  A pair that is not listed or outside the cutoff runs the
//      same Coulomb arithmetic on harmless operands (r^2 = 1, q_i q_j = 0) instead of diverging: ~77 % of the slots of a
//      surviving tile are live.
// One minimum-image shift per (I, J atom) from I's first atom instead of one per pair: identical results whenever
// r_cutoff + (largest cluster extent) < L/2 (a pair inside the cutoff then has that very shift, and a pair that would
// need another one is outside the cutoff with either); checked per launch on the device, per-pair shifts otherwise.
#define PAIR_QT 80               // ring capacity in tiles: < 11 left over + 2 x 32 from two cull iterations
struct PairRing {
  double4 xq[PAIR_QT][3]; int ty[PAIR_QT][4]; unsigned ent[PAIR_QT];
  double4 nxt_pi[3]; int nxt_ty[4]; double4 nxt_first[32];    // next piece: its cluster's atoms / types, first atoms of its first cull chunk
};

template <int TPB_, int MINB, bool SPA>
__global__ void __launch_bounds__(TPB_, MINB) k_pair_tiles(Dev d, int rank, int world) {
  extern __shared__ __align__(16) unsigned char sh_dyn[];   // the warps' rings, then [nT*nT][6] vdw parameters
  __shared__ int sh_vt[RPB_MAXT * RPB_MAXT];   // atype_vdw_type; 2 = SAPT row with all-zero coefficients (contributes exactly 0)
  __shared__ double sh_red[32];
  __shared__ double4 sh_pi[TPB_ / 32][3];      // the warp's cluster I (uniform over the lanes: broadcast reads instead of registers)
  PairRing& Q = reinterpret_cast<PairRing*>(sh_dyn)[threadIdx.x >> 5];
  double* sh_par = reinterpret_cast<double*>(sh_dyn + (TPB_ / 32) * sizeof(PairRing));
  // the force-field rows travel into shared memory by asynchronous copies while the warps look up their ranges (every
  // dependent global load at the head of this kernel is ~0.5 us of a ~50 us kernel)
  const int npar = d.nT * d.nT * 6;
  for (int k = threadIdx.x; k < npar; k += blockDim.x) cp_async_8(&sh_par[k], &d.vdw_param[k]);
  for (int k = threadIdx.x; k < d.nT * d.nT; k += blockDim.x) cp_async_4(&sh_vt[k], &d.vdw_type[k]);
  cp_async_commit();
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double4* pi = sh_pi[threadIdx.x >> 5];
  const int N = d.N;
  const int NC = *d.n_clusters;                // on the device: a committed hop can change it
  const int c_begin = (int)((long long)NC * rank / world), c_end = (int)((long long)NC * (rank + 1) / world);
  const double bx = d.box[0], by = d.box[1], bz = d.box[2];
  const double ibx = d.inv_box[0], iby = d.inv_box[1], ibz = d.inv_box[2];
  const double rc2 = d.rc2, inv_dx = d.inv_erfc_dx;
  const int nT = d.nT;
  // largest cluster extent NOW (k_verlet_disp, every evaluation; bonds stretch between list builds)
  const double ext = __longlong_as_double((long long)d.vstat[4]);
  const double rc = sqrt(d.rc2);
  // SPA (one minimum-image shift per (I, J atom)) is chosen by the host from the box and the cutoff; should a cluster be
  // larger than that choice assumed, the step fails loudly instead of computing with a wrong image
  constexpr bool shift_per_atom = SPA;
  if (SPA && !(rc + ext < 0.5 * fmin(bx, fmin(by, bz)))) { if (threadIdx.x == 0) atomicMax(&d.err_flag[1], 2); return; }
  const double cull2 = (rc + 2.0 * ext) * (rc + 2.0 * ext) * (1.0 + 1e-12);
  const unsigned* __restrict__ L = d.tile_list;
  double e_el = 0.0, e_vdw = 0.0;
  const int gwarp = blockIdx.x * (TPB_ / 32) + (threadIdx.x >> 5), nwarp = gridDim.x * (TPB_ / 32);
  const int* __restrict__ tp = d.tile_point;
  // this warp's range of the rank's tiles; the cluster holding its first tile by a 32-ary search over the row pointers
  const int T0 = tp[RPB_TILE_PARTS * c_begin], T1 = tp[RPB_TILE_PARTS * c_end];
  const int per = max(64, (T1 - T0 + nwarp - 1) / nwarp);
  const long long tb64 = (long long)T0 + (long long)gwarp * per;
  const int t_begin = (int)min(tb64, (long long)T1), t_end = min(T1, t_begin + per);
  int I0 = c_begin;
  if (t_begin < t_end) {
    int lo = c_begin, hi = c_end;              // tp[PARTS * lo] <= t_begin < tp[PARTS * hi]
    {
      // rows of a liquid have similar lengths: 32 probes around the proportional guess usually bracket the answer at once
      const int guess = c_begin + (int)(((long long)(t_begin - T0) * (c_end - c_begin)) / max(T1 - T0, 1));
      const int probe = min(max(guess - 15 + lane, c_begin), c_end);
      const unsigned le = __ballot_sync(0xffffffffu, tp[RPB_TILE_PARTS * probe] <= t_begin);
      const int nle = __popc(le);
      if (nle > 0) lo = __shfl_sync(0xffffffffu, probe, nle - 1);
      if (nle < 32) hi = __shfl_sync(0xffffffffu, probe, nle);
    }
    while (hi - lo > 1) {
      const int probe = lo + (int)(((long long)(hi - lo) * (lane + 1)) / 33);
      const unsigned le = __ballot_sync(0xffffffffu, tp[RPB_TILE_PARTS * probe] <= t_begin);   // monotone: a prefix of the lanes
      const int nle = __popc(le);
      const int nlo = __shfl_sync(0xffffffffu, probe, max(nle - 1, 0)), nhi = __shfl_sync(0xffffffffu, probe, min(nle, 31));
      if (nle) lo = nlo;
      if (nle < 32) hi = nhi;
    }
    I0 = lo;
  }
  // force-field rows have landed: a SAPT row with all-zero coefficients contributes exactly 0 -> type 2 (skipped)
  cp_async_wait<0>();
  __syncthreads();
  for (int k = threadIdx.x; k < d.nT * d.nT; k += blockDim.x)
    if (sh_vt[k] == 1) {
      const double* P = &sh_par[6 * k];
      if (P[0] == 0.0 && P[2] == 0.0 && P[3] == 0.0 && P[4] == 0.0 && P[5] == 0.0) sh_vt[k] = 2;
    }
  __syncthreads();
  int pf_I = -1, pf_info = 0, pf_re = 0, prev_re = 0;     // header of the NEXT piece, fetched while the current one is processed
  unsigned pf_ent = 0u;
  for (int I = I0; I < c_end && t_begin < t_end; I++) {
    const bool pf = pf_I == I;                 // (uniform over the warp)
    const unsigned pf_ent_cur = pf_ent;
    int rs, re, info;
    if (pf) { rs = prev_re; re = pf_re; info = pf_info; }
    else { rs = tp[RPB_TILE_PARTS * I]; re = tp[RPB_TILE_PARTS * (I + 1)]; info = d.cl_info[I]; }
    pf_I = -1;
    if (rs >= t_end) break;
    const int vs = max(rs, t_begin), vf = min(re, t_end);
    if (vs >= vf) continue;
    const int fi = info & 0xffffff, ni = info >> 24;
    int ti[3];
    {
      double4 hp = make_double4(0.0, 0.0, 0.0, 0.0);
      int ht = 0;
      if (lane < 3) {
        if (pf) { hp = Q.nxt_pi[lane]; ht = Q.nxt_ty[lane]; }
        else { const int ia = fi + (lane < ni ? lane : 0); hp = d.xq[ia]; ht = d.type[ia]; }
      }
#pragma unroll
      for (int a = 0; a < 3; a++) ti[a] = __shfl_sync(0xffffffffu, ht, a) * nT;
      __syncwarp();
      if (lane < 3) sh_pi[threadIdx.x >> 5][lane] = hp;
      __syncwarp();
    }
    // the next piece (cluster I + 1, if this warp's range goes on): row end, cluster word and the list words of its first
    // chunk are requested now; the copies that depend on them are issued from the first cull iteration below
    const bool want_next = I + 1 < c_end && re < t_end;
    if (want_next) {
      pf_info = d.cl_info[I + 1];
      pf_re = tp[RPB_TILE_PARTS * (I + 2)];
      pf_ent = re + lane < t_end ? L[re + lane] : 0u;
    }
    bool next_issued = false;
    double f[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int qh = 0, qn = 0;                        // ring head and fill in SLOTS (a tile = three slots, one per atom of J); uniform over the warp

    // phase 2: (up to) 32 slots of the ring, one J atom per lane (a tile may straddle two rounds: its atoms are independent)
    auto interact = [&](const int avail) {
      const int nt = min(avail, 32);
      const int sl = (qh + lane) % (3 * PAIR_QT);
      const int slot = sl / 3, b = sl - 3 * slot;
      const unsigned ent = lane < nt ? Q.ent[slot] : 0u;
      const unsigned mb = (ent >> (23 + b)) & 0x49u;          // bit 3a: pair (a, b) is listed
      double4 pj = make_double4(0.0, 0.0, 0.0, 0.0);
      int tj = 0;
      const int g = (int)(ent & 0x7fffffu) + b;
      if (mb) { pj = Q.xq[slot][b]; tj = Q.ty[slot][b]; }
      double sx = 0.0, sy = 0.0, sz = 0.0;
      if (shift_per_atom) {
        const double4 p0 = pi[0];
        sx = bx * floor_fp64pipe(fma(p0.x - pj.x, ibx, 0.5));
        sy = by * floor_fp64pipe(fma(p0.y - pj.y, iby, 0.5));
        sz = bz * floor_fp64pipe(fma(p0.z - pj.z, ibz, 0.5));
      }
      double dx[3], dy[3], dz[3], inv_r[3], c2[3], qq[3];
      bool live[3];
      double4 tb[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double4 pa = pi[a];
        dx[a] = pa.x - pj.x; dy[a] = pa.y - pj.y; dz[a] = pa.z - pj.z;
        if (shift_per_atom) { dx[a] -= sx; dy[a] -= sy; dz[a] -= sz; }
        else {
          dx[a] = fma(-bx, floor_fp64pipe(fma(dx[a], ibx, 0.5)), dx[a]);
          dy[a] = fma(-by, floor_fp64pipe(fma(dy[a], iby, 0.5)), dy[a]);
          dz[a] = fma(-bz, floor_fp64pipe(fma(dz[a], ibz, 0.5)), dz[a]);
        }
        double r2 = fma(dz[a], dz[a], fma(dy[a], dy[a], dx[a] * dx[a]));
        live[a] = ((mb >> (3 * a)) & 1u) && r2 < rc2;
        qq[a] = live[a] ? pa.w * pj.w : 0.0;
        r2 = live[a] ? r2 : 1.0;
        inv_r[a] = rsqrt_pair(r2);
        // linear_interpolation_ewald_tables  pair_int_real_space.f90:740-759
        const double x1 = (r2 * inv_r[a]) * inv_dx;
        int ii;
        const double ci = ceil_fp64pipe(x1, ii);
        tb[a] = ldg256(&d.es2_t[ii]);
        c2[a] = (x1 + 1.0) - ci;
      }
      double fjx = 0.0, fjy = 0.0, fjz = 0.0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double ir = inv_r[a], inv_r2 = ir * ir, c1 = 1.0 - c2[a];
        const double qr = qq[a] * ir;
        e_el = fma(qr, fma(c2[a], tb[a].z, c1 * tb[a].x), e_el);
        double fs = (qr * inv_r2) * fma(c2[a], tb[a].w, c1 * tb[a].y);
        if (live[a]) {
          const int pidx = ti[a] + tj;
          const int vt = sh_vt[pidx];
          if (vt == 0) {                         // pairwise_real_space_LJ :621-645
            const double c12 = sh_par[6 * pidx], c6 = sh_par[6 * pidx + 1];
            const double r6 = inv_r2 * inv_r2 * inv_r2, c12r6 = c12 * r6;
            e_vdw = fma(r6, c12r6 - c6, e_vdw);
            fs = fma(inv_r2 * r6, 12.0 * c12r6 - 6.0 * c6, fs);
          } else if (vt == 1) {                  // pairwise_real_space_sapt :651-690 (generic path)
            const double2 sp = sapt_pair(d.tt, d.dtt, d.tt_max, d.tt_grid, 1.0 / inv_r2, &sh_par[6 * pidx]);
            e_vdw += sp.x; fs += sp.y;
          }
        }
        const double gx = dx[a] * fs, gy = dy[a] * fs, gz = dz[a] * fs;      // (a dead slot has fs = 0)
        f[3 * a] += gx; f[3 * a + 1] += gy; f[3 * a + 2] += gz;
        fjx -= gx; fjy -= gy; fjz -= gz;
      }
      if (live[0] || live[1] || live[2]) {
        atomicAdd(&d.force[3 * g], fjx); atomicAdd(&d.force[3 * g + 1], fjy); atomicAdd(&d.force[3 * g + 2], fjz);
      }
      qh = (qh + nt) % (3 * PAIR_QT); qn -= nt;
    };

    // phase 1, software-pipelined: list words two iterations ahead, first-atom gathers one ahead
    unsigned ent_c = 0u, ent_n = 0u;
    double4 p_c = make_double4(0.0, 0.0, 0.0, 0.0);
    if (pf) {                                  // (a prefetched piece starts at its row start: vs == rs)
      ent_c = vs + lane < vf ? pf_ent_cur : 0u;
      if (vs + lane < vf) p_c = Q.nxt_first[lane];
    } else {
      if (vs + lane < vf) ent_c = L[vs + lane];
      if (vs + lane < vf) p_c = ldg256(&d.xq[ent_c & 0x7fffffu]);
    }
    if (vs + 32 + lane < vf) ent_n = L[vs + 32 + lane];
    int ready = 0;                             // slots of the ring whose asynchronous copies have been waited for
    for (int k0 = vs; k0 < vf; k0 += 32) {
      unsigned ent_nn = 0u;
      double4 p_n = make_double4(0.0, 0.0, 0.0, 0.0);
      if (k0 + 64 + lane < vf) ent_nn = L[k0 + 64 + lane];
      if (k0 + 32 + lane < vf) p_n = ldg256(&d.xq[ent_n & 0x7fffffu]);
      const double4 p0 = pi[0];
      double r0 = p0.x - p_c.x, r1 = p0.y - p_c.y, r2 = p0.z - p_c.z;
      r0 = fma(-bx, floor_fp64pipe(fma(r0, ibx, 0.5)), r0);
      r1 = fma(-by, floor_fp64pipe(fma(r1, iby, 0.5)), r1);
      r2 = fma(-bz, floor_fp64pipe(fma(r2, ibz, 0.5)), r2);
      const bool keep = (k0 + lane < vf) && fma(r2, r2, fma(r1, r1, r0 * r0)) < cull2;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int slot = ((qh + qn) / 3 + __popc(m & lt)) % PAIR_QT;       // (qh + qn is a multiple of 3: tiles are appended whole)
        const int fj = (int)(ent_c & 0x7fffffu);
        Q.ent[slot] = ent_c;
#pragma unroll
        for (int a = 0; a < 3; a++)
          if (fj + a < N) {                      // (a cluster of fewer atoms: the slot is never read, its mask bits are 0)
            cp_async_16(&Q.xq[slot][a], &d.xq[fj + a]);
            cp_async_16(reinterpret_cast<char*>(&Q.xq[slot][a]) + 16, reinterpret_cast<const char*>(&d.xq[fj + a]) + 16);
            cp_async_4(&Q.ty[slot][a], &d.type[fj + a]);
          }
      }
      if (want_next && !next_issued) {         // (once per piece, in the group of its first iteration: landed long before the switch)
        next_issued = true;
        const int nfi = pf_info & 0xffffff, nni = pf_info >> 24;
        __syncwarp();                            // the previous contents of the slots were consumed at this piece's start
        if (lane < 3) {
          const int ia = nfi + (lane < nni ? lane : 0);
          cp_async_16(&Q.nxt_pi[lane], &d.xq[ia]);
          cp_async_16(reinterpret_cast<char*>(&Q.nxt_pi[lane]) + 16, reinterpret_cast<const char*>(&d.xq[ia]) + 16);
          cp_async_4(&Q.nxt_ty[lane], &d.type[ia]);
        }
        if (re + lane < min(pf_re, t_end)) {
          const double4* src = &d.xq[pf_ent & 0x7fffffu];
          cp_async_16(&Q.nxt_first[lane], src);
          cp_async_16(reinterpret_cast<char*>(&Q.nxt_first[lane]) + 16, reinterpret_cast<const char*>(src) + 16);
        }
      }
      cp_async_commit();
      // the copies of the PREVIOUS iterations have landed once all but the newest group are complete
      cp_async_wait<1>();
      __syncwarp();
      while (ready >= 32) { interact(ready); ready -= 32; }
      qn += 3 * __popc(m);
      ready = qn;                                // (this iteration's tiles become usable after the next wait)
      ent_c = ent_n; ent_n = ent_nn; p_c = p_n;
    }
    cp_async_wait<0>();
    __syncwarp();
    while (qn > 0) interact(qn);               // what is left at the end of the piece
    __syncwarp();
    // F_I of this piece: one ADD per component (the other pieces of the row, the F_J of other rows, the bonded branch and
    // the PME gather add to d.force concurrently)
    const double mine = warp_sum9(f, lane);
    if (lane < 3 * ni) atomicAdd(&d.force[3 * fi + lane], mine);
    if (want_next) { pf_I = I + 1; prev_re = re; }
  }
  e_el = block_sum(e_el, sh_red);
  e_vdw = block_sum(e_vdw, sh_red);
  if (threadIdx.x == 0) { atomicAdd(&d.en[E_ELEC], e_el); atomicAdd(&d.en[E_VDW], e_vdw); }
}

// one thread per molecule: intramolecular non-bonded (exclusion correction, 1-4) + bonds/angles/dihedrals
__global__ void k_molecule_terms(Dev d) {
  __shared__ double sh_red[32];
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  MolEnergies E = {0, 0, 0, 0, 0};
  if (m < d.M) {
    int f0 = d.mol_first[m], n = d.mol_natom[m];
    const MolTypeDev& T = d.mt[d.mol_type[m]];
    double x[RPB_MA][3], q[RPB_MA], f[RPB_MA][3];
    int ty[RPB_MA];
    for (int a = 0; a < n; a++) {
      double4 p = d.xq[f0 + a];
      x[a][0] = p.x; x[a][1] = p.y; x[a][2] = p.z; q[a] = p.w;
      ty[a] = d.type[f0 + a];
      f[a][0] = f[a][1] = f[a][2] = 0.0;
    }
    molecule_terms(d, T, n, x, ty, q, f, E, true, true);
    for (int a = 0; a < n; a++)
      for (int k = 0; k < 3; k++) atomicAdd(&d.force[3 * (f0 + a) + k], f[a][k]);
  }
  double e;
  e = block_sum(E.e_elec, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_ELEC], e);
  e = block_sum(E.e_vdw, sh_red);  if (threadIdx.x == 0) atomicAdd(&d.en[E_VDW], e);
  e = block_sum(E.e_bond, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_BOND], e);
  e = block_sum(E.e_angle, sh_red); if (threadIdx.x == 0) atomicAdd(&d.en[E_ANGLE], e);
  e = block_sum(E.e_dih, sh_red);  if (threadIdx.x == 0) atomicAdd(&d.en[E_DIH], e);
}

template <int T, int MINB>
static int launch_pair_variant(rpb_ctx* c, bool shard, int waves_x4 = 4) {
  // state-sharded runs also shard the principal diabat's pair forces: rank r takes the clusters [NC r / R, NC (r+1) / R); the
  // partial forces and energies ride the two all-reduces the sharded step has anyway.
  const int R = shard ? c->d.world : 1, r = shard ? c->d.rank : 0;
  const int wpb = T / 32;
  const int blocks = std::max(1, c->n_sm * MINB * waves_x4 / 4);     // one resident wave
  const size_t shmem = (size_t)wpb * sizeof(PairRing) + (size_t)((c->d.nT * c->d.nT * 6 + 1) & ~1) * sizeof(double);
  // one minimum-image shift per (I, J atom) needs r_cutoff + (cluster extent) < L/2: taken when the box leaves 4 A for the
  // extent (three bonded atoms), verified against the actual extents on the device
  const double half_box = 0.5 * std::min(c->d.box[0], std::min(c->d.box[1], c->d.box[2]));
  const bool spa = half_box - std::sqrt(c->d.rc2) > 4.0;
  static bool attr_set = false;     // (same for every context: a function attribute)
  if (!attr_set) {
    cudaFuncSetAttribute(k_pair_tiles<T, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_pair_tiles<T, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  if (spa) k_pair_tiles<T, MINB, true><<<blocks, T, shmem, c->stream>>>(c->d, r, R);
  else k_pair_tiles<T, MINB, false><<<blocks, T, shmem, c->stream>>>(c->d, r, R);
  return 0;
}

void launch_pair_verlet(rpb_ctx* c, bool shard) {
  ScopedTimer t(c, T_PAIR);
  static const int variant = getenv("RPB_PAIR_VARIANT") ? atoi(getenv("RPB_PAIR_VARIANT")) : 0;
  switch (variant) {
    case 1: launch_pair_variant<128, 3>(c, shard); break;
    case 2: launch_pair_variant<128, 4>(c, shard); break;
    case 3: launch_pair_variant<64, 6>(c, shard); break;
    case 4: launch_pair_variant<64, 7>(c, shard); break;
    case 5: launch_pair_variant<128, 3>(c, shard, 8); break;        // two waves of half-length ranges
    case 6: launch_pair_variant<128, 3>(c, shard, 3); break;        // leaves a quarter of the CTA slots to the side streams
    case 7: launch_pair_variant<256, 1>(c, shard); break;
    // MS-EVB evaluations (shard == true): the Hamiltonian chain on the side streams bounds the step, and its short kernels
    // need CTA slots while the pair kernel runs -- three quarters of a wave
    default: launch_pair_variant<128, 3>(c, shard, (shard && !c->throughput_mode) ? 3 : 4); break;
  }
  c->n_launch += 1;
}

void launch_molecule_terms(rpb_ctx* c) {
  ScopedTimer t(c, T_INTRA);
  k_molecule_terms<<<(c->d.M + 127) / 128, 128, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}
