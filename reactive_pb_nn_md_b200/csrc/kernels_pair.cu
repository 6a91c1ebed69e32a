// Real-space pair forces over the half Verlet list, and per-molecule bonded / intramolecular terms.
// Replaces:
//   pairwise_real_space_verlet            src/pair_int_real_space.f90:135-371
//   pairwise_real_space_{ewald,LJ,sapt}   src/pair_int_real_space.f90:621-759
//   intra_molecular_pairwise_energy_force src/pair_int_real_space.f90:386-588
//   intra_molecular_energy_force          src/intra_bonded_interactions.f90:17-552
//
// Pair kernel: one warp per i-atom over its row of the SYMMETRIC list (both directions of every listed pair, built
// next to the reference-ordered half list at every rebuild): each pair is evaluated from both of its atoms, F_i is
// reduced in registers + shuffles and stored once -- no atomics, so the forces are reproducible run to run -- and the
// energies are halved.  Per listed pair: minimum image with a reciprocal box (the shift can only differ from the
// reference's division for |dr| ~ L/2, far outside the cutoff) and the cutoff test.  Per in-cutoff pair: ONE rsqrt
// replaces the sqrt and the six divisions of pair_int_real_space.f90:621-645,698-759 (relative differences ~1e-16,
// the interpolated tables are continuous across bins), and the erfc / ewaldscale tables are read interleaved.
// Bound by the FP64 pipe -- see DESIGN.md.
#include <cstdlib>
#include "rpb_host.h"
#include "rpb_bonded.cuh"

#define PAIR_TPB 128

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

// BF: branch-free evaluation -- out-of-cutoff (and padding) lanes run the same arithmetic on harmless operands (r^2 = 1, q_i q_j = 0,
// table entry 1) instead of sitting out behind divergent branches; the in-cutoff lanes execute exactly the same operations.
// 1/sqrt(x) for x in the range of squared pair distances (normal, far from the exponent limits): single-precision seed
// and two Newton-Raphson steps in fp64 -- no special-case branches, ~2 ulp (the library routine: 1 ulp)
__device__ __forceinline__ double rsqrt_pair(double x) {
  double y = (double)rsqrtf((float)x);
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  return y;
}

// MODE bit 0 (BF), bit 1: rsqrt_pair instead of the library rsqrt
template <int PAIR_B, int TPB_, int MINB, int MODE = 0>
__global__ void __launch_bounds__(TPB_, MINB) k_pair_verlet(Dev d, int i_begin, int i_end) {
  constexpr bool BF = (MODE & 1) != 0, FR = (MODE & 2) != 0;
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
  const int nwarp_total = (gridDim.x * blockDim.x) >> 5;
  const double ibx = d.inv_box[0], iby = d.inv_box[1], ibz = d.inv_box[2];
  const double bx = d.box[0], by = d.box[1], bz = d.box[2];
  const int* __restrict__ L = d.full_list;
  double e_el = 0.0, e_vdw = 0.0;
  for (int i = i_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < i_end; i += nwarp_total) {
    const int vs = d.full_point[i], vf = d.full_point[i + 1];
    const double4 pi = d.xq[i];
    const int ti = d.type[i] * d.nT;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    // The loop is bound by memory latency, not arithmetic: a warp issues in order and stalls at the first USE of a
    // pending load, so every lane works on PAIR_B neighbours at once -- PAIR_B list entries, then PAIR_B 32-byte
    // coordinate gathers, then PAIR_B 32-byte table gathers are issued back to back (one exposed latency per batch and
    // stage instead of one per neighbour), and the list entries of the next batch are fetched a full batch ahead.
    // (Measured on B200, C3, first version: 107 us one neighbour at a time -> 59 us with PAIR_B = 4; staging the gathers through
    // shared memory with cp.async was slower, 142 us -- twice the L1 wavefronts for 16-byte copies.)
    int jn[PAIR_B];
#pragma unroll
    for (int k = 0; k < PAIR_B; k++) { const int v = vs + 32 * k + lane; jn[k] = v < vf ? L[v] : -1; }
    for (int base = vs; base < vf; base += 32 * PAIR_B) {
      int j[PAIR_B];
      double4 p[PAIR_B];
#pragma unroll
      for (int k = 0; k < PAIR_B; k++) { j[k] = jn[k]; p[k] = make_double4(0.0, 0.0, 0.0, 0.0); if (j[k] >= 0) p[k] = ldg256(&d.xq[j[k] & 0xffffff]); }
#pragma unroll
      for (int k = 0; k < PAIR_B; k++) { const int v = base + 32 * (PAIR_B + k) + lane; jn[k] = v < vf ? L[v] : -1; }
      double sdx[PAIR_B], sdy[PAIR_B], sdz[PAIR_B], sinv[PAIR_B], sc2[PAIR_B], sqq[PAIR_B];
      double4 tb[PAIR_B];       // {erfc[i-1], scale[i-1], erfc[i], scale[i]}
      bool in[PAIR_B];
#pragma unroll
      for (int k = 0; k < PAIR_B; k++) {   // minimum image, cutoff, table index, table load
        double dx = pi.x - p[k].x, dy = pi.y - p[k].y, dz = pi.z - p[k].z;
        dx = fma(-bx, floor(fma(dx, ibx, 0.5)), dx);
        dy = fma(-by, floor(fma(dy, iby, 0.5)), dy);
        dz = fma(-bz, floor(fma(dz, ibz, 0.5)), dz);
        const double dr2 = fma(dz, dz, fma(dy, dy, dx * dx));
        in[k] = j[k] >= 0 && dr2 < d.rc2;
        if (BF) {
          const double d2 = in[k] ? dr2 : 1.0;
          const double inv_r = FR ? rsqrt_pair(d2) : rsqrt(d2);
          const double x1 = (d2 * inv_r) * d.inv_erfc_dx;
          const double ci = ceil(x1);
          tb[k] = ldg256(&d.es2_t[in[k] ? (int)ci : 1]);
          sc2[k] = (x1 + 1.0) - ci;
          sinv[k] = inv_r;
          sdx[k] = dx; sdy[k] = dy; sdz[k] = dz; sqq[k] = in[k] ? pi.w * p[k].w : 0.0;
          continue;
        }
        tb[k] = make_double4(0.0, 0.0, 0.0, 0.0);
        sdx[k] = dx; sdy[k] = dy; sdz[k] = dz; sqq[k] = pi.w * p[k].w;
        sinv[k] = 1.0; sc2[k] = 0.0;
        if (in[k]) {
          const double inv_r = FR ? rsqrt_pair(dr2) : rsqrt(dr2);
          // linear_interpolation_ewald_tables  pair_int_real_space.f90:740-759
          const double x1 = (dr2 * inv_r) * d.inv_erfc_dx;
          const double ci = ceil(x1);
          tb[k] = ldg256(&d.es2_t[(int)ci]);
          sc2[k] = (x1 + 1.0) - ci;
          sinv[k] = inv_r;
        }
      }
#pragma unroll
      for (int k = 0; k < PAIR_B; k++) {   // energies and force of the in-cutoff pairs
        if (!BF && !in[k]) continue;
        const int pidx = BF ? (in[k] ? ti + (j[k] >> 24) : 0) : ti + (j[k] >> 24);
        const double inv_r = sinv[k], inv_r2 = inv_r * inv_r, c1 = 1.0 - sc2[k];
        const double qr = sqq[k] * inv_r;
        e_el = fma(qr, fma(sc2[k], tb[k].z, c1 * tb[k].x), e_el);
        double fs = (qr * inv_r2) * fma(sc2[k], tb[k].w, c1 * tb[k].y);
        const int vt = (BF && !in[k]) ? -1 : sh_vt[pidx];
        if (BF) {                            // LJ with the coefficients masked to zero for every lane that has no LJ term
          const bool lj = vt == 0;
          const double c12 = lj ? sh_par[6 * pidx] : 0.0, c6 = lj ? sh_par[6 * pidx + 1] : 0.0;
          const double r6 = inv_r2 * inv_r2 * inv_r2, c12r6 = c12 * r6;
          e_vdw = fma(r6, c12r6 - c6, e_vdw);
          fs = fma(inv_r2 * r6, 12.0 * c12r6 - 6.0 * c6, fs);
          if (vt == 1) sapt_pair(d.tt, d.dtt, d.tt_max, d.tt_grid, 1.0 / inv_r2, &sh_par[6 * pidx], e_vdw, fs);
        } else if (vt == 0) {                // pairwise_real_space_LJ :621-645
          const double c12 = sh_par[6 * pidx], c6 = sh_par[6 * pidx + 1];
          const double r6 = inv_r2 * inv_r2 * inv_r2, c12r6 = c12 * r6;
          e_vdw = fma(r6, c12r6 - c6, e_vdw);
          fs = fma(inv_r2 * r6, 12.0 * c12r6 - 6.0 * c6, fs);
        } else if (vt == 1) {                // pairwise_real_space_sapt :651-690 (generic path)
          sapt_pair(d.tt, d.dtt, d.tt_max, d.tt_grid, 1.0 / inv_r2, &sh_par[6 * pidx], e_vdw, fs);
        }
        fx = fma(sdx[k], fs, fx); fy = fma(sdy[k], fs, fy); fz = fma(sdz[k], fs, fz);
      }
    }
    fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
    if (lane == 0) { atomicAdd(&d.force[3 * i], fx); atomicAdd(&d.force[3 * i + 1], fy); atomicAdd(&d.force[3 * i + 2], fz); }   // one RED per component: the bonded branch adds to d.force concurrently
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

template <int B, int T, int M, int MODE = 0>
static void launch_pair_variant(rpb_ctx* c, bool shard) {
  // state-sharded runs also shard the principal diabat's pair forces: rank r takes the atoms [N r / R, N (r+1) / R); the
  // partial forces and energies ride the two all-reduces the sharded step has anyway
  const int R = shard ? c->d.world : 1, r = shard ? c->d.rank : 0, N = c->d.N;
  const int i0 = (int)((long long)N * r / R), i1 = (int)((long long)N * (r + 1) / R);
  const int wpb = T / 32, blocks = std::max(1, (i1 - i0 + wpb - 1) / wpb);
  const size_t shmem = (size_t)c->d.nT * c->d.nT * 6 * sizeof(double);
  k_pair_verlet<B, T, M, MODE><<<blocks, T, shmem, c->stream>>>(c->d, i0, i1);
}

void launch_pair_verlet(rpb_ctx* c, bool shard) {
  ScopedTimer t(c, T_PAIR);
  static const int variant = getenv("RPB_PAIR_VARIANT") ? atoi(getenv("RPB_PAIR_VARIANT")) : 0;
  switch (variant) {
    case 1: launch_pair_variant<2, 256, 3>(c, shard); break;
    case 2: launch_pair_variant<2, 256, 4>(c, shard); break;
    case 3: launch_pair_variant<3, 256, 2>(c, shard); break;
    case 4: launch_pair_variant<3, 128, 5>(c, shard); break;
    case 5: launch_pair_variant<4, 256, 2>(c, shard); break;
    case 6: launch_pair_variant<2, 128, 6>(c, shard); break;
    case 7: launch_pair_variant<4, 128, 3>(c, shard); break;
    case 8: launch_pair_variant<3, 128, 3>(c, shard); break;
    case 9: launch_pair_variant<3, 192, 2>(c, shard); break;
    case 10: launch_pair_variant<2, 256, 2>(c, shard); break;
    case 11: launch_pair_variant<2, 192, 3>(c, shard); break;
    case 12: launch_pair_variant<2, 128, 4>(c, shard); break;
    case 13: launch_pair_variant<2, 512, 1>(c, shard); break;
    case 14: launch_pair_variant<2, 256, 2, 1>(c, shard); break;
    case 15: launch_pair_variant<3, 256, 2, 1>(c, shard); break;
    case 16: launch_pair_variant<2, 256, 2, 2>(c, shard); break;
    case 17: launch_pair_variant<2, 256, 2, 3>(c, shard); break;
    case 18: launch_pair_variant<3, 256, 2, 3>(c, shard); break;
    default: launch_pair_variant<2, 256, 2, 2>(c, shard); break;   // best of the sweep on B200: batches of 2 x 32 neighbours fit 128 registers without
                                                                // spills (3 x 32 spills loaded coordinates to local memory in the hot loop: ncu source
                                                                // page, profiles/README.md); C2 107 -> 97 us, C4 260 -> 234 us; with the branch-free rsqrt_pair 94 / 228 us
  }
  c->n_launch += 1;
}

void launch_molecule_terms(rpb_ctx* c) {
  ScopedTimer t(c, T_INTRA);
  k_molecule_terms<<<(c->d.M + 127) / 128, 128, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}
