// Real-space pair forces over the cluster-pair (tile) list, and per-molecule bonded / intramolecular terms.
// Replaces:
//   pairwise_real_space_verlet            src/pair_int_real_space.f90:135-371
//   pairwise_real_space_{ewald,LJ,sapt}   src/pair_int_real_space.f90:621-759
//   intra_molecular_pairwise_energy_force src/pair_int_real_space.f90:386-588
//   intra_molecular_energy_force          src/intra_bonded_interactions.f90:17-552
//
// Pair kernel (k_pair_tiles): a warp per row part of a cluster I (<= 3 consecutive atoms of one molecule, held in
// registers); a lane takes one atom of a cluster J per iteration -- its (up to) three pairs with the atoms of I, whose
// listed subset are three bits of the tile's 9-bit mask (kernels_nlist.cu: exactly the reference's listed pairs).  A list
// word serves nine atom pairs.  Both directions of a tile are stored, so F_I is reduced in registers + shuffles and added
// to d.force once per row part: no j-scatter (bonded terms and the PME gather add to d.force too).  Energies are halved.  Per listed pair: minimum image with a reciprocal box (the shift can only differ from the reference's
// division for |dr| ~ L/2, far outside the cutoff) and the cutoff test dr^2 < r_c^2.  Per in-cutoff pair: ONE rsqrt
// replaces the sqrt and the six divisions of pair_int_real_space.f90:621-645,698-759 (relative differences ~1e-16,
// the interpolated tables are continuous across bins), and the erfc / ewaldscale tables are read interleaved
// (one 32-byte load per pair).  Pairs outside the mask or the cutoff run the same Coulomb arithmetic on harmless
// operands (r^2 = 1, q_i q_j = 0, table entry 1) instead of diverging.  FP64-pipe bound -- see DESIGN.md.
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

// A warp works on a few consecutive row parts (cluster I, part) of the tile list; rank r of R works on the clusters
// [NC r / R, NC (r+1) / R).  Two phases per row part:
//   1. candidate test -- a lane takes one ATOM of a cluster J per iteration (three consecutive lanes share a list word) and
//      tests its three pairs with the atoms of I: mask bit, minimum image, r^2 < r_c^2.  42 % of the listed pairs lie
//      between the cutoff and the list radius and 18 % of a tile's slots are not listed at all, so the pairs that pass
//      are COMPACTED (ballot + prefix count) into a per-warp queue in shared memory;
//   2. interaction -- whenever the queue holds NB x 32 pairs, every lane pops NB of them: rsqrt, table index, ONE 32-byte
//      gather of the interleaved erfc / ewaldscale entries per pair, Coulomb, LJ / SAPT, force.  All lanes carry real
//      pairs, and a lane has NB independent dependency chains (and NB table gathers) in flight: the loop is bound by the
//      latency of that chain, not by a pipe.
// One minimum-image shift per (I, J atom) from I's first atom instead of one per pair: identical results whenever
// r_cutoff + (largest cluster extent) < L/2 (a pair inside the cutoff then has that very shift, and a pair that would
// need another one is outside the cutoff with either); checked per launch on the device, per-pair shifts otherwise.
#define PAIR_QCAP 256            // pairs a warp's queue can hold: < NB x 32 left over + 96 from one chunk
struct PairQueue { double dx[PAIR_QCAP], dy[PAIR_QCAP], dz[PAIR_QCAP], r2[PAIR_QCAP], qq[PAIR_QCAP]; int meta[PAIR_QCAP]; };

template <int TPB_, int MINB, int NB>
__global__ void __launch_bounds__(TPB_, MINB) k_pair_tiles(Dev d, int rank, int world, int ppw, unsigned int* __restrict__ counters) {
  extern __shared__ double sh_par[];           // [nT*nT][6] vdw parameters, then the warps' queues
  __shared__ int sh_vt[RPB_MAXT * RPB_MAXT];   // atype_vdw_type; 2 = SAPT row with all-zero coefficients (contributes exactly 0)
  __shared__ double sh_red[32];
  __shared__ double4 sh_pi[TPB_ / 32][3];      // the warp's cluster I (uniform over the lanes: broadcast reads instead of registers)
  const int npar = d.nT * d.nT * 6;
  PairQueue& Q = reinterpret_cast<PairQueue*>(sh_par + ((npar + 1) & ~1))[threadIdx.x >> 5];
  for (int k = threadIdx.x; k < npar; k += blockDim.x) sh_par[k] = d.vdw_param[k];
  for (int k = threadIdx.x; k < d.nT * d.nT; k += blockDim.x) {
    int vt = d.vdw_type[k];
    if (vt == 1) {
      const double* P = &d.vdw_param[6 * k];
      if (P[0] == 0.0 && P[2] == 0.0 && P[3] == 0.0 && P[4] == 0.0 && P[5] == 0.0) vt = 2;
    }
    sh_vt[k] = vt;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const double4* pi = sh_pi[threadIdx.x >> 5];
  const int NC = *d.n_clusters;                // on the device: a committed hop can change it
  const int c_begin = (int)((long long)NC * rank / world), c_end = (int)((long long)NC * (rank + 1) / world);
  const int n_work = RPB_TILE_PARTS * (c_end - c_begin);
  const double bx = d.box[0], by = d.box[1], bz = d.box[2];
  const double ibx = d.inv_box[0], iby = d.inv_box[1], ibz = d.inv_box[2];
  const double rc2 = d.rc2, inv_dx = d.inv_erfc_dx;
  const int nT = d.nT;
  const double ext = __longlong_as_double((long long)d.vstat[0]);
  const bool shift_per_atom = sqrt(d.rc2) + ext < 0.5 * fmin(bx, fmin(by, bz));
  const unsigned* __restrict__ L = d.tile_list;
  double e_el = 0.0, e_vdw = 0.0;
  // A warp works on PPW consecutive pieces (row parts); the header of the next piece -- a chain of dependent loads
  // (cluster -> first atom -> types / coordinates; row pointers) -- is fetched while the current piece is processed.
  // The grid is NOT persistent: CTAs live ~15 us, so the short kernels of the high-priority side streams (enumeration,
  // images, PME, bonded terms) get SM resources as CTAs retire instead of waiting for the whole pair kernel.
  struct Header { int info, vs, vf, tpack; double4 p; };
  auto load_header = [&](int wk, Header& h) {
    h.info = 0; h.vs = 0; h.vf = 0; h.tpack = 0; h.p = make_double4(0.0, 0.0, 0.0, 0.0);
    if (wk < n_work) {
      const int rp = RPB_TILE_PARTS * c_begin + wk;
      h.info = d.cl_info[rp / RPB_TILE_PARTS];
      h.vs = d.tile_point[rp]; h.vf = d.tile_point[rp + 1];
      const int fi = h.info & 0xffffff, ni = h.info >> 24;
      if (lane < 3) { const int ia = fi + (lane < ni ? lane : 0); h.p = d.xq[ia]; h.tpack = d.type[ia]; }
    }
  };
  // pieces are handed out by a counter (balance: the last CTAs of the grid find nothing left and exit), ppw per warp
  int w = 0, w1 = 0;
  if (lane == 0) { w = (int)atomicAdd(&counters[0], 1u); if (ppw > 1) w1 = (int)atomicAdd(&counters[0], 1u); }
  w = __shfl_sync(0xffffffffu, w, 0); w1 = ppw > 1 ? __shfl_sync(0xffffffffu, w1, 0) : n_work;
  Header H, H1;
  load_header(w, H);
  for (int it = 0; it < ppw && w < n_work; it++) {
    int w2 = n_work;
    if (it + 2 < ppw && lane == 0) w2 = (int)atomicAdd(&counters[0], 1u);   // consumed two pieces from now
    load_header(w1, H1);                                                    // consumed at the next switch
    const int fi = H.info & 0xffffff, ni = H.info >> 24;
    const int vs = H.vs, vf = H.vf;
    int ti[3];
#pragma unroll
    for (int a = 0; a < 3; a++) ti[a] = __shfl_sync(0xffffffffu, H.tpack, a) * nT;
    __syncwarp();
    if (lane < 3) sh_pi[threadIdx.x >> 5][lane] = H.p;
    __syncwarp();
    const int nslot = 3 * (vf - vs);
    double f[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    int head = 0, cnt = 0;                     // queue state (uniform over the warp)

    // phase 2: NB pairs per lane (fewer in the last round of a row part)
    auto interact = [&]() {
      const int n = min(cnt, 32 * NB);
      double dx[NB], dy[NB], dz[NB], inv_r[NB], c2[NB], qq[NB];
      int meta[NB];
      bool on[NB];
      double4 tb[NB];
#pragma unroll
      for (int u = 0; u < NB; u++) {
        on[u] = 32 * u + lane < n;
        const int q = (head + 32 * u + lane) & (PAIR_QCAP - 1);
        double r2 = 1.0;
        dx[u] = dy[u] = dz[u] = 0.0; qq[u] = 0.0; meta[u] = 0;
        if (on[u]) { dx[u] = Q.dx[q]; dy[u] = Q.dy[q]; dz[u] = Q.dz[q]; r2 = Q.r2[q]; qq[u] = Q.qq[q]; meta[u] = Q.meta[q]; }
        inv_r[u] = rsqrt_pair(r2);
        // linear_interpolation_ewald_tables  pair_int_real_space.f90:740-759
        const double x1 = (r2 * inv_r[u]) * inv_dx;
        int ii;
        const double ci = ceil_fp64pipe(x1, ii);
        tb[u] = ldg256(&d.es2_t[ii]);
        c2[u] = (x1 + 1.0) - ci;
      }
#pragma unroll
      for (int u = 0; u < NB; u++) {
        const double ir = inv_r[u], inv_r2 = ir * ir, c1 = 1.0 - c2[u];
        const double qr = qq[u] * ir;
        e_el = fma(qr, fma(c2[u], tb[u].z, c1 * tb[u].x), e_el);
        double fs = (qr * inv_r2) * fma(c2[u], tb[u].w, c1 * tb[u].y);
        const int pidx = meta[u] & 0xffff, a = meta[u] >> 16;
        if (on[u]) {
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
        const double gx = dx[u] * fs, gy = dy[u] * fs, gz = dz[u] * fs;      // (an empty slot has d = 0)
        if (a == 0) { f[0][0] += gx; f[0][1] += gy; f[0][2] += gz; }
        else if (a == 1) { f[1][0] += gx; f[1][1] += gy; f[1][2] += gz; }
        else { f[2][0] += gx; f[2][1] += gy; f[2][2] += gz; }
      }
      head = (head + n) & (PAIR_QCAP - 1); cnt -= n;
    };

    // software pipeline of the list: word of iteration +2, gathers of iteration +1
    unsigned ent_c = 0u, ent_n = 0u;
    double4 pj_c = make_double4(0.0, 0.0, 0.0, 0.0);
    int tj_c = 0;
    { const int k = lane; if (k < nslot) ent_c = L[vs + k / 3]; }
    { const int k = 32 + lane; if (k < nslot) ent_n = L[vs + k / 3]; }
    { const int k = lane; if (k < nslot) { const int g = (ent_c & 0x7fffff) + k % 3; pj_c = ldg256(&d.xq[g]); tj_c = __ldg(&d.type[g]); } }
    for (int k0 = 0; k0 < nslot; k0 += 32) {
      const int k = k0 + lane, b = k % 3;
      double4 pj_n = make_double4(0.0, 0.0, 0.0, 0.0);
      int tj_n = 0;
      unsigned ent_nn = 0u;
      if (k + 32 < nslot) { const int g = (ent_n & 0x7fffff) + (k + 32) % 3; pj_n = ldg256(&d.xq[g]); tj_n = __ldg(&d.type[g]); }
      if (k + 64 < nslot) ent_nn = L[vs + (k + 64) / 3];
      const unsigned mbits = (k < nslot) ? (ent_c >> (23 + b)) : 0u;     // bit 3a: pair (a, b) is listed
      // ---- phase 1: the three candidate pairs of this J atom, compacted into the queue
      double sx = 0.0, sy = 0.0, sz = 0.0;
      if (shift_per_atom) {
        const double4 p0 = pi[0];
        sx = bx * floor_fp64pipe(fma(p0.x - pj_c.x, ibx, 0.5));
        sy = by * floor_fp64pipe(fma(p0.y - pj_c.y, iby, 0.5));
        sz = bz * floor_fp64pipe(fma(p0.z - pj_c.z, ibz, 0.5));
      }
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const double4 pa = pi[a];
        double dx = pa.x - pj_c.x, dy = pa.y - pj_c.y, dz = pa.z - pj_c.z;
        if (shift_per_atom) { dx -= sx; dy -= sy; dz -= sz; }
        else {
          dx = fma(-bx, floor_fp64pipe(fma(dx, ibx, 0.5)), dx);
          dy = fma(-by, floor_fp64pipe(fma(dy, iby, 0.5)), dy);
          dz = fma(-bz, floor_fp64pipe(fma(dz, ibz, 0.5)), dz);
        }
        const double dr2 = fma(dz, dz, fma(dy, dy, dx * dx));
        const bool live = ((mbits >> (3 * a)) & 1u) && dr2 < rc2;
        const unsigned m = __ballot_sync(0xffffffffu, live);
        if (live) {
          const int q = (head + cnt + __popc(m & lt)) & (PAIR_QCAP - 1);
          Q.dx[q] = dx; Q.dy[q] = dy; Q.dz[q] = dz; Q.r2[q] = dr2; Q.qq[q] = pa.w * pj_c.w;
          Q.meta[q] = (ti[a] + tj_c) | (a << 16);
        }
        cnt += __popc(m);
      }
      __syncwarp();
      // ---- phase 2 on full rounds
      while (cnt >= 32 * NB) interact();
      ent_c = ent_n; ent_n = ent_nn; pj_c = pj_n; tj_c = tj_n;
    }
    while (cnt > 0) interact();                // what is left at the end of the row part
    // F_I of this row part: lanes -> lane 0..8 by xor shuffles; one ADD per component and part (the other parts of the
    // row, the bonded branch and the PME gather add to d.force concurrently)
    double mine = 0.0;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double x = f[a][c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 3 * a + c) mine = x;
      }
    if (lane < 3 * ni) atomicAdd(&d.force[3 * fi + lane], mine);
    H = H1; w = w1;
    w1 = (it + 2 < ppw) ? __shfl_sync(0xffffffffu, w2, 0) : n_work;
  }
  e_el = block_sum(e_el, sh_red);
  e_vdw = block_sum(e_vdw, sh_red);
  if (threadIdx.x == 0) {
    atomicAdd(&d.en[E_ELEC], 0.5 * e_el); atomicAdd(&d.en[E_VDW], 0.5 * e_vdw);
    // the last CTA to finish re-arms the work counter for the next launch
    __threadfence();
    if (atomicAdd(&counters[1], 1u) == gridDim.x - 1) { counters[0] = 0u; counters[1] = 0u; __threadfence(); }
  }
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

template <int T, int MINB, int NB>
static int launch_pair_variant(rpb_ctx* c, bool shard, int ppw, int pad_kb = 0) {
  // state-sharded runs also shard the principal diabat's pair forces: rank r takes the clusters [NC r / R, NC (r+1) / R); the
  // partial forces and energies ride the two all-reduces the sharded step has anyway.
  const int R = shard ? c->d.world : 1, r = shard ? c->d.rank : 0;
  const long long pieces = (long long)RPB_TILE_PARTS * ((c->n_clusters_bound + R - 1) / R + 1);
  const int wpb = T / 32;
  const int blocks = (int)std::max(1LL, (pieces + (long long)wpb * ppw - 1) / ((long long)wpb * ppw));
  // pad_kb: shared memory requested beyond what the kernel uses, to cap the CTAs resident per SM -- at full occupancy the
  // pair kernel holds every register of an SM, and each short kernel of the MS-EVB chain then waits for a pair CTA to retire
  const size_t shmem = std::max((size_t)((c->d.nT * c->d.nT * 6 + 1) & ~1) * sizeof(double) + wpb * sizeof(PairQueue), (size_t)pad_kb * 1024);
  static bool attr_set = false;     // (same for every context: a function attribute)
  if (!attr_set) { cudaFuncSetAttribute(k_pair_tiles<T, MINB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr_set = true; }
  k_pair_tiles<T, MINB, NB><<<blocks, T, shmem, c->stream>>>(c->d, r, R, ppw, reinterpret_cast<unsigned int*>(c->d.vstat + 3));
  return 0;
}

void launch_pair_verlet(rpb_ctx* c, bool shard) {
  ScopedTimer t(c, T_PAIR);
  static const int variant = getenv("RPB_PAIR_VARIANT") ? atoi(getenv("RPB_PAIR_VARIANT")) : 0;
  switch (variant) {
    case 1: launch_pair_variant<128, 3, 3>(c, shard, 2); break;      // 168 registers, three pairs per lane in flight
    case 2: launch_pair_variant<128, 3, 3>(c, shard, 4); break;
    case 3: launch_pair_variant<128, 3, 3>(c, shard, 4, 80); break;      // two CTAs per SM
    case 4: launch_pair_variant<128, 3, 3>(c, shard, 2, 80); break;
    case 5: launch_pair_variant<128, 3, 3>(c, shard, 4, 120); break;     // one CTA per SM
    default: launch_pair_variant<128, 3, 3>(c, shard, 4); break;
  }
  c->n_launch += 1;
}

void launch_molecule_terms(rpb_ctx* c) {
  ScopedTimer t(c, T_INTRA);
  k_molecule_terms<<<(c->d.M + 127) / 128, 128, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}
