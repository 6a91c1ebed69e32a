// Device-side data model of the B200-native MS-EVB force path (sm_100a, fp64).
//
// Everything the step needs lives in HBM for the whole run; `Dev` is the flat table of device
// pointers + scalars that every kernel receives by value (constant bank).  Layouts:
//   xq[N]        double4 {x,y,z,charge}   32-byte aligned -> two LDG.128 per neighbour
//   vel/force    double[3N]               (3,N) column-major == reference layout
//   Q / theta    double[S][K][K][K]       n1 fastest (pme.f90 Q(n1+1,n2+1,n3+1)), batch stride K^3
//   FQ           complex[S][K][K][K/2+1]  cuFFT D2Z output, m1 half-spectrum fastest
//   snapshots    per (state, level) images of the <=4 molecules of a proton-hop chain (evb.cuh)
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include "../../include/rpbmd.h"

#define RPB_MAXT RPB_MAX_N_ATOM_TYPE
#define RPB_MAXM RPB_MAX_N_MOLE_TYPE
#define RPB_MAXI RPB_MAX_INTERACTION_TYPE
#define RPB_MAXS RPB_EVB_MAX_STATES
#define RPB_MAXC RPB_EVB_MAX_CHAIN
#define RPB_MA RPB_MAX_MOLE_ATOMS
#define RPB_MAXB 8    // bonds / angles / dihedrals per molecule type
#define RPB_CHAIN_MOLS (RPB_MAXC + 1)
#define RPB_TILE_PARTS 4   // a cluster's row of the pair list is built and consumed in this many independent parts
#define RPB_TILE_TMPCAP 160 // entries of a row part kept in the rebuild's scratch (longer rows are swept twice; water at r_v = 12 A: ~70)

// energy accumulator slots (device doubles)
enum { E_ELEC = 0, E_VDW, E_BOND, E_ANGLE, E_DIH, E_RECIP, E_KE, E_REP, E_NSLOT };

// molecule_type_data (glob_v.f90:299-317) with bonded parameters resolved per term at upload
// (parameters are looked up by atom type in the reference; a molecule's atom types always equal its
// type's template after evb retyping + reorder, ms_evb.f90:888-927).
struct MolTypeDev {
  int n_atom;
  int atom_type[RPB_MA];
  int n_bond, n_angle, n_dih;
  int bond[RPB_MAXB][2]; int bond_kind[RPB_MAXB]; double bond_par[RPB_MAXB][3];
  int angle[RPB_MAXB][3]; int angle_kind[RPB_MAXB]; double angle_par[RPB_MAXB][2];
  int dih[RPB_MAXB][4]; int dih_kind[RPB_MAXB]; double dih_par[RPB_MAXB][6];
  int pair_excl[RPB_MA][RPB_MA];
  int reactive_proton[RPB_MA], reactive_basic[RPB_MA];
  int heavy_acid_atom;   // get_heavy_atom_transfer_acid for this type (-1 if not an acid)
  int heavy_base_atom;   // get_heavy_atom_transfer_base for this type (-1 if not a base)
  int bonded_heavy[RPB_MA];  // find_bonded_atom_hydrogen per atom (-1 if not exactly one bond)
};

struct EvbTables {  // glob_v.f90:77-120, 0-based types, -1 = empty row
  int da_int[RPB_MAXI][3]; double da_par[RPB_MAXI][6];
  int pa_int[RPB_MAXI][2]; double pa_par[RPB_MAXI][5];
  int dc_int[RPB_MAXI][3]; double dc_par[RPB_MAXI][10]; int dc_type[RPB_MAXI];
  double exch_atomic[RPB_MAXT]; double exch_proton[RPB_MAXM][RPB_MAXM];
  int conj_pairs[RPB_MAXM], conj_atom[RPB_MAXT];
  double ref_energy[RPB_MAXM]; int proton_index[RPB_MAXM], heavy_acid_index[RPB_MAXM];
  double atype_chg[RPB_MAXT];
};

struct Dev {
  int N, M, K, nT, nMT;
  int rank, world;
  double box[3], inv_box[3];
  double kk[3];          // diagonal of construct_reciprocal_lattice_vector (REAL*4 volume)
  double rc2, rv2, verlet_skin;
  double alpha, erf_factor, conv, conv_kin, dt, pi;
  double erfc_dx, tt_max; int tt_grid; double spline_grid;
  double ewald_self;
  double cut_solv2, cut_pair2;
  int max_chain, max_states;
  // atoms
  double4* xq; double* vel; double* force; double* mass; int* type; int* mol_of_atom;
  // molecules
  int* mol_first; int* mol_natom; int* mol_type; double* r_com;
  int* hydronium;        // device scalar: 0-based hydronium molecule
  // force field
  const double* vdw_param;    // [nT*nT][6]
  const double* vdw_param14;
  const int* vdw_type;        // [nT*nT]
  const int* freeze;          // [nT]
  const MolTypeDev* mt;       // [nMT]
  const EvbTables* evb;
  // tables
  const double *B6, *B5, *erfc_t, *scale_t, *tt, *dtt, *CBh;
  const double4* es2_t;  // es2_t[i] = {erfc_t[i-1], scale_t[i-1], erfc_t[i], scale_t[i]}: both table points of an interpolation in ONE 32-byte load
  double inv_erfc_dx;
  // verlet
  // The reference's half list (verlet_point / neighbor_list, reference row order, 1-based) is a parity ACCESSOR: it is
  // generated on demand from the positions of the last rebuild (vbuild_xq).  The step itself consumes the cluster-pair
  // ("tile") list: clusters are runs of <= 3 consecutive atoms of one molecule (a water is one cluster), a tile is an
  // ordered pair of clusters (I, J) with a 9-bit mask of the atom pairs that the reference's list holds
  // (different molecule, |r_ij| < r_verlet at build time); both directions of every tile are stored.
  int* verlet_point; int* neighbor_list; int verlet_cap;
  double4* vbuild_xq;                       // positions at the last rebuild
  int* n_clusters;                          // device scalar
  int* cl_info;                             // [cluster] first atom | n_atom << 24
  int* mol_cl_first; int* mol_ncl;          // [M+1] first cluster of every molecule / [M] clusters per molecule
  unsigned* tile_tmp;                       // rebuild scratch: [row part][RPB_TILE_TMPCAP]
  int* tile_point; unsigned* tile_list;     // CSR over (cluster, part) rows [RPB_TILE_PARTS * I + part]; entry = first atom of cluster J | mask << 23 (bit 3a+b: atom a of I with atom b of J)
  long long tile_cap;
  double4* csort_xq; int* csort_mol; int* csort_info;   // cell-sorted copies of the clusters for the sweep: [3 slot + b], [slot], [slot]
  double4* vsort_xq; int* vsort_mol; int* vsort_entry;  // cell-sorted atom copies (reference-list accessor)
  unsigned long long* vstat;                // [0] bits of the largest cluster extent at the build  [1] listed atom pairs  [2] rebuild counter  [3] pair-kernel work counters  [4] bits of the largest cluster extent now
  double* vstore; double* vdisp; int* flag_verlet; int* rebuild_now;
  const int* commit_hop;                    // device flag: the MS-EVB solver selected a new hydronium molecule this step (null without MS-EVB)
  int* err_flag;  // [0] atom with |F|>1e5 (1-based, 0 none)  [1] verlet overflow  [2] too many diabats  [3] evb lookup failure
  int ncx, ncy, ncz, dia, dib, dic;
  int coop_blocks;                          // grid of the cooperative rebuild kernels on this context's device
  int* cell_count; int* cell_start; int* cell_atoms; int* atom_cell; int* row_count;
  double* maxd;  // two largest displacements
  int* vdone;    // arrival counter of the displacement kernel's blocks
  // PME
  double* uscale; double* Q; double* theta; cufftDoubleComplex* FQ; double* force_recip;
  // energies
  double* en;            // [E_NSLOT]
};

#define CUDA_OK(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
      return RPB_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#ifdef __CUDACC__
// one 256-bit read-only load (LDG.E.ENL2.256): a 32-byte-aligned double4 costs one L1 wavefront per distinct line
// instead of the two of a pair of LDG.128 -- the scattered gathers of the pair kernel are bound by exactly that
__device__ __forceinline__ double4 ldg256(const double4* p) {
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
// ---- strict fp64 helpers (the library is compiled with --fmad=false; fma() is used explicitly
//      only where the summation order is free) ----
__device__ __forceinline__ double min_image(double d, double L) { return d - L * floor(d / L + 0.5); }

// floor / ceil of |t| < 2^51 on the FP64 pipe (two DADDs with directed rounding around 1.5 * 2^52) instead of the
// conversion unit's FRND.F64, which runs at a fraction of the FP64 rate on sm_100 (ncu: the XU pipe, not the FP64 pipe,
// was the busiest unit of the pair kernel).  Exact: t + M lies in [2^52, 2^53), where the spacing of doubles is 1.
#define RPB_MAGIC_2P52 6755399441055744.0
__device__ __forceinline__ double floor_fp64pipe(double t) { return __dadd_rn(__dadd_rd(t, RPB_MAGIC_2P52), -RPB_MAGIC_2P52); }
// ceil(t) as a double and as an int (the low word of the biased sum), 0 <= t < 2^31
__device__ __forceinline__ double ceil_fp64pipe(double t, int& i) {
  const double b = __dadd_ru(t, RPB_MAGIC_2P52);
  i = __double2loint(b);
  return __dadd_rn(b, -RPB_MAGIC_2P52);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0; `sh` holds >= 32 doubles
__device__ __forceinline__ double block_sum(double v, double* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? sh[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

// linear_interpolation_ewald_tables pair_int_real_space.f90:740-759
__device__ __forceinline__ void ewald_tables(const Dev& d, double r, double& ev, double& sv) {
  double x1 = r / d.erfc_dx;
  double ci = ceil(x1);
  int i = (int)ci;
  double c2 = (x1 + 1.0) - ci;
  double c1 = 1.0 - c2;
  ev = c1 * __ldg(&d.erfc_t[i - 1]) + c2 * __ldg(&d.erfc_t[i]);
  sv = c1 * __ldg(&d.scale_t[i - 1]) + c2 * __ldg(&d.scale_t[i]);
}

// one non-bonded pair: table-erfc Coulomb + LJ / SAPT (pair_int_real_space.f90:621-759).
// dr = r_i - r_j (already minimum-imaged), par -> 6 vdw parameters, vt = atype_vdw_type.
// Returns energies; f = force on i (force on j is -f).
__device__ __forceinline__ void pair_terms(const Dev& d, const double dr[3], double dr2, double qq, int vt,
                                           const double* par, bool do_coulomb, double& e_el, double& e_vdw, double f[3]) {
  double fs = 0.0;
  e_el = 0.0; e_vdw = 0.0;
  double r = sqrt(dr2);
  if (do_coulomb) {
    double ev, sv;
    ewald_tables(d, r, ev, sv);
    e_el = qq / r * ev;
    fs = qq / (dr2 * r) * sv;
  }
  if (vt == 0) {
    double dr6 = dr2 * dr2 * dr2, dr12 = dr6 * dr6;
    double c12 = par[0], c6 = par[1];
    e_vdw = c12 / dr12 - c6 / dr6;
    fs += (12.0 * c12 / dr12 - 6.0 * c6 / dr6) / dr2;
  } else if (vt == 1) {
    double A = par[0], B = par[1], C6 = par[2], C8 = par[3], C10 = par[4], C12 = par[5];
    if (A != 0.0 || C6 != 0.0 || C8 != 0.0 || C10 != 0.0 || C12 != 0.0) {  // all-zero rows contribute exactly 0
      double dr6 = dr2 * dr2 * dr2, dr8 = dr6 * dr2, dr10 = dr8 * dr2, dr12 = dr10 * dr2;
      int idx = (int)ceil(B * r / d.tt_max * (double)d.tt_grid);
      const double* tt = &d.tt[4 * (idx - 1)];
      const double* dt = &d.dtt[4 * (idx - 1)];
      double ex = exp(-1 * B * r);
      e_vdw = A * ex - tt[0] * C6 / dr6 - tt[1] * C8 / dr8 - tt[2] * C10 / dr10 - tt[3] * C12 / dr12;
      double fac = r * A * B * ex + r * (B * dt[0]) * C6 / dr6 - tt[0] * 6.0 * C6 / dr6 + r * (B * dt[1]) * C8 / dr8 -
                   tt[1] * 8.0 * C8 / dr8 + r * (B * dt[2]) * C10 / dr10 - tt[2] * 10.0 * C10 / dr10 +
                   r * (B * dt[3]) * C12 / dr12 - tt[3] * 12.0 * C12 / dr12;
      fs += fac / dr2;
    }
  }
  f[0] = dr[0] * fs; f[1] = dr[1] * fs; f[2] = dr[2] * fs;
}

// intra_pme_exclusion pair_int_real_space.f90:781-816 (exact erfc, no table)
__device__ __forceinline__ void excl_terms(const Dev& d, const double dr[3], double dr2, double qq, double& e_el, double f[3]) {
  double rm = sqrt(dr2);
  if (rm < 1e-8) {
    e_el = -d.erf_factor * qq * d.conv;
    f[0] = f[1] = f[2] = 0.0;
  } else {
    double ec = erfc(rm * d.alpha) - 1.0;
    e_el = qq * ec / rm * d.conv;
    double ar = rm * d.alpha;
    double g = ec / (dr2 * rm) + d.erf_factor * exp(-(ar * ar)) / dr2;
    double s = qq * g * d.conv;
    f[0] = dr[0] * s; f[1] = dr[1] * s; f[2] = dr[2] * s;
  }
}
#endif
