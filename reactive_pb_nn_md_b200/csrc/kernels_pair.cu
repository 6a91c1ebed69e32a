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
__device__ __noinline__ void sapt_pair(const double* __restrict__ tt_t, const double* __restrict__ dtt_t, double tt_max, int tt_grid,
                                       double dr2, const double* par, double& e_vdw, double& fs) {
  const double A = par[0], B = par[1], C6 = par[2], C8 = par[3], C10 = par[4], C12 = par[5];
  const double r = sqrt(dr2);
  const double dr6 = dr2 * dr2 * dr2, dr8 = dr6 * dr2, dr10 = dr8 * dr2, dr12 = dr10 * dr2;
  const int idx = (int)ceil(B * r / tt_max * (double)tt_grid);
  const double* tt = &tt_t[4 * (idx - 1)];
  const double* dt = &dtt_t[4 * (idx - 1)];
  const double ex = exp(-1 * B * r);
  e_vdw += A * ex - tt[0] * C6 / dr6 - tt[1] * C8 / dr8 - tt[2] * C10 / dr10 - tt[3] * C12 / dr12;
  const double fac = r * A * B * ex + r * (B * dt[0]) * C6 / dr6 - tt[0] * 6.0 * C6 / dr6 + r * (B * dt[1]) * C8 / dr8 -
                     tt[1] * 8.0 * C8 / dr8 + r * (B * dt[2]) * C10 / dr10 - tt[2] * 10.0 * C10 / dr10 +
                     r * (B * dt[3]) * C12 / dr12 - tt[3] * 12.0 * C12 / dr12;
  fs += fac / dr2;
}

// 1/sqrt(x) for x in the range of squared pair distances (normal, far from the exponent limits): single-precision seed
// and two Newton-Raphson steps in fp64 -- no special-case branches, ~2 ulp (the library routine: 1 ulp)
__device__ __forceinline__ double rsqrt_pair(double x) {
  double y = (double)rsqrtf((float)x);
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// Persistent warps: every warp takes row parts (cluster I, part) of the tile list in a grid-stride loop; rank r of R takes
// the clusters [NC r / R, NC (r+1) / R).  Inside a row part one lane handles one ATOM of a J cluster per iteration (three
// consecutive lanes share a list word): its three pairs with the atoms of I.  The list word is fetched two iterations
// ahead and the coordinate / type gathers one iteration ahead, so the only latency a warp waits for inside an iteration
// is that of its three table gathers, which are issued together.
// SHIFT_PER_ATOM: one minimum-image shift per (I, J atom) from I's first atom instead of one per pair -- identical
// results whenever r_cutoff + (largest cluster extent) < L/2 (a pair inside the cutoff then has that very shift, and a
// pair that would need another one is outside the cutoff with either); checked per launch on the device.
template <int TPB_, int MINB>
__global__ void __launch_bounds__(TPB_, MINB) k_pair_tiles(Dev d, int rank, int world) {
  extern __shared__ double sh_par[];           // [nT*nT][6] vdw parameters
  __shared__ int sh_vt[RPB_MAXT * RPB_MAXT];   // atype_vdw_type; 2 = SAPT row with all-zero coefficients (contributes exactly 0)
  __shared__ double sh_red[32];
  for (int k = threadIdx.x; k < d.nT * d.nT * 6; k += blockDim.x) sh_par[k] = d.vdw_param[k];
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
  const int NC = *d.n_clusters;                // on the device: a committed hop can change it
  const int c_begin = (int)((long long)NC * rank / world), c_end = (int)((long long)NC * (rank + 1) / world);
  const double ibx = d.inv_box[0], iby = d.inv_box[1], ibz = d.inv_box[2];
  const double bx = d.box[0], by = d.box[1], bz = d.box[2];
  const int nT = d.nT;
  const double ext = __longlong_as_double((long long)d.vstat[0]);
  const bool shift_per_atom = sqrt(d.rc2) + ext < 0.5 * fmin(bx, fmin(by, bz));
  const unsigned* __restrict__ L = d.tile_list;
  double e_el = 0.0, e_vdw = 0.0;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int rp = RPB_TILE_PARTS * c_begin + gw; rp < RPB_TILE_PARTS * c_end; rp += nw) {
    const int I = rp / RPB_TILE_PARTS;
    const int info = d.cl_info[I];
    const int fi = info & 0xffffff, ni = info >> 24;
    double4 pi[3];
    int ti[3];
#pragma unroll
    for (int a = 0; a < 3; a++) { const int ia = fi + (a < ni ? a : 0); pi[a] = d.xq[ia]; ti[a] = d.type[ia] * nT; }
    const int vs = d.tile_point[rp], vf = d.tile_point[rp + 1];
    const int nslot = 3 * (vf - vs);
    double f[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
    // software pipeline: word of iteration +2, gathers of iteration +1
    unsigned ent_c = 0u, ent_n = 0u;
    double4 pj_c = make_double4(0.0, 0.0, 0.0, 0.0);
    int tj_c = 0;
    { const int k = lane; if (k < nslot) ent_c = L[vs + k / 3]; }
    { const int k = 32 + lane; if (k < nslot) ent_n = L[vs + k / 3]; }
    { const int k = lane, b = k % 3; if (k < nslot) { const int g = (ent_c & 0x7fffff) + b; pj_c = ldg256(&d.xq[g]); tj_c = __ldg(&d.type[g]); } }
    for (int k0 = 0; k0 < nslot; k0 += 32) {
      const int k = k0 + lane, b = k % 3;
      // issue the next iteration's gathers and the list word after that
      double4 pj_n = make_double4(0.0, 0.0, 0.0, 0.0);
      int tj_n = 0;
      unsigned ent_nn = 0u;
      if (k + 32 < nslot) { const int g = (ent_n & 0x7fffff) + (k + 32) % 3; pj_n = ldg256(&d.xq[g]); tj_n = __ldg(&d.type[g]); }
      if (k + 64 < nslot) ent_nn = L[vs + (k + 64) / 3];
      const unsigned mbits = (k < nslot) ? (ent_c >> (23 + b)) : 0u;     // bit 3a of mbits: pair (a, b) is listed
      // minimum-image shift of this J atom relative to the cluster's first atom
      double sx = 0.0, sy = 0.0, sz = 0.0;
      if (shift_per_atom) {
        sx = bx * floor(fma(pi[0].x - pj_c.x, ibx, 0.5));
        sy = by * floor(fma(pi[0].y - pj_c.y, iby, 0.5));
        sz = bz * floor(fma(pi[0].z - pj_c.z, ibz, 0.5));
      }
      double sdx[3], sdy[3], sdz[3], sinv[3], sc2[3], sqq[3];
      double4 tb[3];
      bool live[3];
#pragma unroll
      for (int a = 0; a < 3; a++) {   // minimum image, cutoff, table index, table load
        double dx = pi[a].x - pj_c.x, dy = pi[a].y - pj_c.y, dz = pi[a].z - pj_c.z;
        if (shift_per_atom) { dx -= sx; dy -= sy; dz -= sz; }
        else {
          dx = fma(-bx, floor(fma(dx, ibx, 0.5)), dx);
          dy = fma(-by, floor(fma(dy, iby, 0.5)), dy);
          dz = fma(-bz, floor(fma(dz, ibz, 0.5)), dz);
        }
        const double dr2 = fma(dz, dz, fma(dy, dy, dx * dx));
        live[a] = ((mbits >> (3 * a)) & 1u) && dr2 < d.rc2;
        const double d2 = live[a] ? dr2 : 1.0;
        const double inv_r = rsqrt_pair(d2);
        // linear_interpolation_ewald_tables  pair_int_real_space.f90:740-759
        const double x1 = (d2 * inv_r) * d.inv_erfc_dx;
        const double ci = ceil(x1);
        tb[a] = ldg256(&d.es2_t[live[a] ? (int)ci : 1]);
        sc2[a] = (x1 + 1.0) - ci;
        sinv[a] = inv_r;
        sdx[a] = dx; sdy[a] = dy; sdz[a] = dz;
        sqq[a] = live[a] ? pi[a].w * pj_c.w : 0.0;
      }
#pragma unroll
      for (int a = 0; a < 3; a++) {   // energies and force
        const double inv_r = sinv[a], inv_r2 = inv_r * inv_r, c1 = 1.0 - sc2[a];
        const double qr = sqq[a] * inv_r;
        e_el = fma(qr, fma(sc2[a], tb[a].z, c1 * tb[a].x), e_el);
        double fs = (qr * inv_r2) * fma(sc2[a], tb[a].w, c1 * tb[a].y);
        if (live[a]) {
          const int pidx = ti[a] + tj_c;
          const int vt = sh_vt[pidx];
          if (vt == 0) {                       // pairwise_real_space_LJ :621-645
            const double c12 = sh_par[6 * pidx], c6 = sh_par[6 * pidx + 1];
            const double r6 = inv_r2 * inv_r2 * inv_r2, c12r6 = c12 * r6;
            e_vdw = fma(r6, c12r6 - c6, e_vdw);
            fs = fma(inv_r2 * r6, 12.0 * c12r6 - 6.0 * c6, fs);
          } else if (vt == 1) {                // pairwise_real_space_sapt :651-690 (generic path)
            sapt_pair(d.tt, d.dtt, d.tt_max, d.tt_grid, 1.0 / inv_r2, &sh_par[6 * pidx], e_vdw, fs);
          }
        }
        f[a][0] = fma(sdx[a], fs, f[a][0]); f[a][1] = fma(sdy[a], fs, f[a][1]); f[a][2] = fma(sdz[a], fs, f[a][2]);
      }
      ent_c = ent_n; ent_n = ent_nn; pj_c = pj_n; tj_c = tj_n;
    }
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
  }
  e_el = block_sum(e_el, sh_red);
  e_vdw = block_sum(e_vdw, sh_red);
  if (threadIdx.x == 0) { atomicAdd(&d.en[E_ELEC], 0.5 * e_el); atomicAdd(&d.en[E_VDW], 0.5 * e_vdw); }
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
static void launch_pair_variant(rpb_ctx* c, bool shard, int ctas_per_sm) {
  // state-sharded runs also shard the principal diabat's pair forces: rank r takes the clusters [NC r / R, NC (r+1) / R); the
  // partial forces and energies ride the two all-reduces the sharded step has anyway.  Persistent warps: a fixed grid.
  const int R = shard ? c->d.world : 1, r = shard ? c->d.rank : 0;
  const long long parts = (long long)RPB_TILE_PARTS * ((c->n_clusters_bound + R - 1) / R + 1);
  const int blocks = (int)std::max(1LL, std::min((long long)c->n_sm * ctas_per_sm, (parts * 32 + T - 1) / T));
  const size_t shmem = (size_t)c->d.nT * c->d.nT * 6 * sizeof(double);
  k_pair_tiles<T, MINB><<<blocks, T, shmem, c->stream>>>(c->d, r, R);
}

void launch_pair_verlet(rpb_ctx* c, bool shard) {
  ScopedTimer t(c, T_PAIR);
  static const int variant = getenv("RPB_PAIR_VARIANT") ? atoi(getenv("RPB_PAIR_VARIANT")) : 0;
  switch (variant) {
    case 1: launch_pair_variant<128, 4>(c, shard, 4); break;      // 128 registers
    case 2: launch_pair_variant<128, 3>(c, shard, 3); break;      // 168
    case 3: launch_pair_variant<256, 2>(c, shard, 2); break;      // 128
    case 4: launch_pair_variant<128, 5>(c, shard, 5); break;      // 96
    case 5: launch_pair_variant<128, 4>(c, shard, 8); break;      // 128 registers, two waves of CTAs
    default: launch_pair_variant<128, 4>(c, shard, 4); break;
  }
  c->n_launch += 1;
}

void launch_molecule_terms(rpb_ctx* c) {
  ScopedTimer t(c, T_INTRA);
  k_molecule_terms<<<(c->d.M + 127) / 128, 128, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}
