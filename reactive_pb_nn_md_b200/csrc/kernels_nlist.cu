// Verlet displacement tracker and on-device neighbour-list rebuild.  Replaces (reference file:line):
//   update_verlet_displacements          src/general_routines.f90:1259-1337
//   construct_verlet_list_grid           src/general_routines.f90:1408-1595  (+ allocate_verlet_list :1206-1247)
//
// What the step consumes is a CLUSTER-PAIR ("tile") list, not the reference's atom list: atoms are grouped into clusters
// of <= 3 consecutive atoms of one molecule (a water is one cluster), and a tile (I, J) carries a 9-bit mask of exactly
// those atom pairs the reference's list holds -- different molecules, |r_ij|^2 < r_verlet^2 at build time with the
// orthorhombic minimum image of :1554-1560 -- so the SET of listed pairs, and with it which in-cutoff pairs a stale list
// misses (rebuild threshold 1.2 x skin > skin), is the reference's.  Both directions of a tile are stored: the pair kernel
// then needs no atomics and no scatter for the j-forces, one 96-byte gather serves nine pairs, and the list shrinks from
// 8 bytes per listed pair to ~1 (SURVEY 8d counts 4 P_v + 56 N algorithmic bytes).
//
// The reference's own half list in its row order (cells scanned ia, ib, ic nested relative to the atom's cell, ascending
// atom index inside a cell, :1523-1531, :1486-1493) is a parity accessor (rpb_get_neighbor_list): it is generated on
// demand by k_verlet_reference_list from the positions saved at the last rebuild.  Its length must also not exceed the
// reference's capacity ("please increase size of verlet neighbor list"): the rebuild checks that on the tile masks.
//
// The rebuild decision stays on the device: the cooperative rebuild kernel is launched every step and exits at once
// unless *flag_verlet says rebuild (or the caller forces it: init, hop commit).
#include <algorithm>
#include <cooperative_groups.h>
#include "rpb_host.h"
#include "rpb_commit.cuh"
namespace cg = cooperative_groups;

#define TPB 256

// per-block two largest |accumulated displacement|; the last block merges them and sets the flag (:1293-1326)
__global__ void k_verlet_disp(Dev d, double* blk_top2, int* done_counter) {
  __shared__ double s1[TPB], s2[TPB];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int rebuild = *d.rebuild_now;
  double nrm = 0.0;
  if (i < d.N) {
    double4 p = d.xq[i];
    double xn[3] = {p.x, p.y, p.z};
    {
      // current extent of the atom's cluster (distance to the cluster's first atom; molecules are never split by the
      // periodic wrap): the pair kernel culls whole tiles with the largest one.  d.vstat[4] is reset by k_verlet_rebuild, which precedes every launch of this kernel.
      const int f = d.mol_first[d.mol_of_atom[i]], k = (i - f) % 3;
      if (k) {
        const double4 p0 = d.xq[i - k];
        const double e2 = (p.x - p0.x) * (p.x - p0.x) + (p.y - p0.y) * (p.y - p0.y) + (p.z - p0.z) * (p.z - p0.z);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(sqrt(e2));
        if (bits > d.vstat[4]) atomicMax(&d.vstat[4], bits);      // non-negative doubles order like their bit patterns
      }
    }
    if (rebuild) {
      for (int k = 0; k < 3; k++) { d.vstore[3 * i + k] = xn[k]; d.vdisp[3 * i + k] = 0.0; }
    } else {
      double acc[3];
      for (int k = 0; k < 3; k++) {
        double xo = d.vstore[3 * i + k];
        double dr = xn[k] - xo;                       // pbc_shift(old,new)
        double sh = floor(d.inv_box[k] * dr + 0.5) * d.box[k];
        double dd = xn[k] - xo - sh;                  // pbc_dr
        acc[k] = d.vdisp[3 * i + k] + dd;
        d.vdisp[3 * i + k] = acc[k];
        d.vstore[3 * i + k] = xn[k];
      }
      nrm = sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2]);
    }
  }
  s1[threadIdx.x] = nrm; s2[threadIdx.x] = 0.0;
  __syncthreads();
  for (int o = TPB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double a1 = s1[threadIdx.x], a2 = s2[threadIdx.x], b1 = s1[threadIdx.x + o], b2 = s2[threadIdx.x + o];
      double m1 = fmax(a1, b1);
      double m2 = fmax(fmin(a1, b1), fmax(a2, b2));
      s1[threadIdx.x] = m1; s2[threadIdx.x] = m2;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { blk_top2[2 * blockIdx.x] = s1[0]; blk_top2[2 * blockIdx.x + 1] = s2[0]; }
  __shared__ int is_last;
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(done_counter, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int nblk = gridDim.x;
  double m1 = 0.0, m2 = 0.0;
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    double b1 = blk_top2[2 * b], b2 = blk_top2[2 * b + 1];
    double n1 = fmax(m1, b1), n2 = fmax(fmin(m1, b1), fmax(m2, b2));
    m1 = n1; m2 = n2;
  }
  s1[threadIdx.x] = m1; s2[threadIdx.x] = m2;
  __syncthreads();
  for (int o = TPB / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      double a1 = s1[threadIdx.x], a2 = s2[threadIdx.x], b1 = s1[threadIdx.x + o], b2 = s2[threadIdx.x + o];
      s1[threadIdx.x] = fmax(a1, b1);
      s2[threadIdx.x] = fmax(fmin(a1, b1), fmax(a2, b2));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    d.maxd[0] = s1[0]; d.maxd[1] = s2[0];
    int rb = *d.rebuild_now;
    if (rb == 1) *d.flag_verlet = 0;
    else if (rb == 0) *d.flag_verlet = ((s1[0] + s2[0]) > d.verlet_skin) ? 1 : 0;
    *done_counter = 0;
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int wrap_cell(int ig, int n) { return ig - (int)floor((double)(ig - 1) / (double)n) * n; }

// exclusive scan by ONE block: out[i] = base + sum_{j<i} in[j], out[n] = base + total.  Returns the total to every thread.
__device__ int block_scan_exclusive(const int* __restrict__ in, int* __restrict__ out, int n, int base) {
  __shared__ int part[TPB];
  __shared__ int total;
  const int tid = threadIdx.x, chunk = (n + TPB - 1) / TPB;
  const int s0 = min(tid * chunk, n), s1 = min(s0 + chunk, n);
  int sum = 0;
  for (int k = s0; k < s1; k++) sum += in[k];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) { int run = 0; for (int k = 0; k < TPB; k++) { int v = part[k]; part[k] = run; run += v; } out[n] = run + base; total = run; }
  __syncthreads();
  int run = part[tid] + base;
  for (int k = s0; k < s1; k++) { int v = in[k]; out[k] = run; run += v; }
  __syncthreads();
  return total;
}

__device__ __forceinline__ int cell_of(const Dev& d, const double4& p) {
  int ix = (int)floor((d.inv_box[0] * p.x) * d.ncx) + 1;
  int iy = (int)floor((d.inv_box[1] * p.y) * d.ncy) + 1;
  int iz = (int)floor((d.inv_box[2] * p.z) * d.ncz) + 1;
  ix = wrap_cell(ix, d.ncx); iy = wrap_cell(iy, d.ncy); iz = wrap_cell(iz, d.ncz);
  return (iz - 1) + d.ncz * ((iy - 1) + d.ncy * (ix - 1));   // z fastest: a column of the cell walk is contiguous
}

// ascending index inside each cell (deterministic order; for atoms == the reference's append-at-tail linked list :1486-1493)
__device__ __forceinline__ void sort_cells(const Dev& d, int ncell, int gtid, int gthreads) {
  for (int c = gtid; c < ncell; c += gthreads) {
    int s = d.cell_start[c], e = d.cell_start[c + 1];
    for (int a = s + 1; a < e; a++) {
      int v = d.cell_atoms[a], b = a - 1;
      while (b >= s && d.cell_atoms[b] > v) { d.cell_atoms[b + 1] = d.cell_atoms[b]; b--; }
      d.cell_atoms[b + 1] = v;
    }
  }
}

// range of cell offsets visited along one dimension: `reach` cells either way, never more than the n cells there are
__device__ __forceinline__ void offset_range(int reach, int n, int& lo, int& hi, bool& prune) {
  if (2 * reach + 1 <= n) { lo = -reach; hi = reach; prune = true; }
  else { lo = -((n - 1) / 2); hi = lo + n - 1; prune = false; }     // every cell once; offsets are no longer minimum images
}

// ------------------------------------------------------------------------------------------------
// tile sweep of one cluster I by one warp.  WRITE = false: count tiles and listed atom pairs; true: store the entries.
// Candidates come from the cell-sorted cluster copies; cells that cannot hold a cluster within reach
// R = r_verlet + 2 x (largest cluster extent) of I's first atom are skipped.
// A cluster's row is built (and later consumed) in RPB_TILE_PARTS independent parts: part p takes every RPB_TILE_PARTS-th
// (x, y) column of the cell walk, so four warps share the sweep of one cluster and four warps share its pair forces.
template <bool WRITE>
__device__ __forceinline__ void tile_sweep(const Dev& d, const int I, const int part, const int lane, const double R, int* __restrict__ queue) {
  const int info = d.cl_info[I], fi = info & 0xffffff, ni = info >> 24;
  const int mi = d.mol_of_atom[fi];
  double4 pi[3];
#pragma unroll
  for (int a = 0; a < 3; a++) pi[a] = d.xq[fi + (a < ni ? a : 0)];
  const int c = cell_of(d, pi[0]);
  const int iz = c % d.ncz + 1, iy = (c / d.ncz) % d.ncy + 1, ix = c / (d.ncz * d.ncy) + 1;
  const double wx = d.box[0] / d.ncx, wy = d.box[1] / d.ncy, wz = d.box[2] / d.ncz;
  double fx = (d.inv_box[0] * pi[0].x) * d.ncx; fx -= floor(fx);
  double fy = (d.inv_box[1] * pi[0].y) * d.ncy; fy -= floor(fy);
  double fz = (d.inv_box[2] * pi[0].z) * d.ncz; fz -= floor(fz);
  int xlo, xhi, ylo, yhi, zlo, zhi;
  bool px, py, pz;
  offset_range((int)floor(R / wx) + 1, d.ncx, xlo, xhi, px);
  offset_range((int)floor(R / wy) + 1, d.ncy, ylo, yhi, py);
  offset_range((int)floor(R / wz) + 1, d.ncz, zlo, zhi, pz);
  const double slack = 1e-6, R2 = R * R;
  // WRITE: straight into the row of the final list; otherwise into this row's fixed-capacity scratch (the entries beyond
  // the capacity are only counted: such a row is swept a second time once its place in the list is known)
  unsigned* __restrict__ out = WRITE ? d.tile_list + d.tile_point[RPB_TILE_PARTS * I + part] : d.tile_tmp + (size_t)(RPB_TILE_PARTS * I + part) * RPB_TILE_TMPCAP;
  int n_tile = 0;
  unsigned long long n_pair = 0;
  const int ny = yhi - ylo + 1, ncol = (xhi - xlo + 1) * ny;
  int qh = 0, qn = 0;                       // the warp's queue of near candidates (cell-sorted slots), uniform over the warp
  // the (up to) nine atom pairs of cluster I with the cluster in `slot` (-1: idle lane) -> one list word
  auto pair_tests = [&](const int slot) {
    unsigned mask = 0;
    int fj = 0;
    if (slot >= 0) {
      const int jinfo = d.csort_info[slot], nj = jinfo >> 24;
      fj = jinfo & 0xffffff;
#pragma unroll
      for (int b = 0; b < 3; b++) {
        if (b < nj) {
          const double4 pj = ldg256(&d.csort_xq[3 * slot + b]);
#pragma unroll
          for (int a = 0; a < 3; a++) {
            double r0 = pi[a].x - pj.x, r1 = pi[a].y - pj.y, r2 = pi[a].z - pj.z;
            r0 = r0 - d.box[0] * floor_fp64pipe(r0 * d.inv_box[0] + 0.5);
            r1 = r1 - d.box[1] * floor_fp64pipe(r1 * d.inv_box[1] + 0.5);
            r2 = r2 - d.box[2] * floor_fp64pipe(r2 * d.inv_box[2] + 0.5);
            if (a < ni && (r0 * r0 + r1 * r1 + r2 * r2) < d.rv2) mask |= 1u << (3 * a + b);
          }
        }
      }
    }
    const unsigned hit = __ballot_sync(0xffffffffu, mask != 0);
    if (mask) { const int pos = n_tile + __popc(hit & ((1u << lane) - 1u)); if (WRITE || pos < RPB_TILE_TMPCAP) out[pos] = (unsigned)fj | (mask << 23); }
    if (!WRITE) n_pair += __popc(mask);
    n_tile += __popc(hit);
    __syncwarp();
  };
  {
    // The (x, y) columns of this part, 32 at a time: every lane works out the slot range(s) of ONE column (a column's cells
    // within reach are contiguous in the cell-sorted arrays, in two pieces when the z range wraps), then the warp walks the
    // concatenation of the 32 ranges with all lanes busy (a lane finds its column by a search over the ranges' offsets).
    const int ncol_mine = (ncol - part + RPB_TILE_PARTS - 1) / RPB_TILE_PARTS;
    for (int kb = 0; kb < ncol_mine; kb += 32) {
      int s0 = 0, e0 = 0, s1 = 0, e1 = 0;
      if (kb + lane < ncol_mine) {
        const int col_id = part + RPB_TILE_PARTS * (kb + lane);
        const int ox = xlo + col_id / ny, oy = ylo + col_id % ny;
        const int g1 = wrap_cell(ix + ox, d.ncx);
        double gx = 0.0;
        if (px && ox != 0) gx = fmax(0.0, (ox > 0 ? (double)ox - fx : fx - (double)(ox + 1)) - slack) * wx;
        double gy = 0.0;
        if (py && oy != 0) gy = fmax(0.0, (oy > 0 ? (double)oy - fy : fy - (double)(oy + 1)) - slack) * wy;
        const double rem = R2 - (gx * gx + gy * gy);
        if (rem >= 0.0) {
          int z0 = zlo, z1 = zhi;
          if (pz) {
            const double zc = sqrt(rem) / wz + slack;
            z1 = min(zhi, (int)floor(fz + zc));
            z0 = -min(-zlo, (int)floor(1.0 - fz + zc));
          }
          const int g2 = wrap_cell(iy + oy, d.ncy);
          const int col = d.ncz * ((g2 - 1) + d.ncy * (g1 - 1));
          // [iz+z0, iz+z1] wrapped into 1..ncz: one or two contiguous ranges of the cell-sorted arrays
          const int len = z1 - z0 + 1, start = wrap_cell(iz + z0, d.ncz);
          s0 = d.cell_start[col + start - 1]; e0 = d.cell_start[col + min(start + len - 1, d.ncz)];
          if (start + len - 1 > d.ncz) { s1 = d.cell_start[col]; e1 = d.cell_start[col + start + len - 1 - d.ncz]; }
        }
      }
      const int n0 = e0 - s0, len = n0 + (e1 - s1);
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - len;
      for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + lane;
        int owner = 0;                      // the last lane whose range starts at or before t (ranges of length 0 never win)
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int ex = __shfl_sync(0xffffffffu, excl, (owner + step) & 31);
          if (ex <= t) owner += step;
        }
        const int off = t - __shfl_sync(0xffffffffu, excl, owner);
        const int cs0 = __shfl_sync(0xffffffffu, s0, owner), cn0 = __shfl_sync(0xffffffffu, n0, owner), cs1 = __shfl_sync(0xffffffffu, s1, owner);
        const int slot = off < cn0 ? cs0 + off : cs1 + (off - cn0);
        // HALF list of tiles: the unordered cluster pair {I, J} is stored in the row of I when I + J is odd and I < J, or
        // I + J is even and I > J (balanced rows without any ordering of the clusters in space)
        bool mine = false;
        if (t < total && d.csort_mol[slot] != mi) { const int J = d.cell_atoms[slot]; mine = ((I + J) & 1) ? (I < J) : (I > J); }
        bool near = false;
        if (mine) {     // first atoms farther apart than R: no atom pair can be listed
          const double4 p0 = ldg256(&d.csort_xq[3 * slot]);
          double r0 = pi[0].x - p0.x, r1 = pi[0].y - p0.y, r2 = pi[0].z - p0.z;
          r0 = r0 - d.box[0] * floor_fp64pipe(r0 * d.inv_box[0] + 0.5);
          r1 = r1 - d.box[1] * floor_fp64pipe(r1 * d.inv_box[1] + 0.5);
          r2 = r2 - d.box[2] * floor_fp64pipe(r2 * d.inv_box[2] + 0.5);
          near = (r0 * r0 + r1 * r1 + r2 * r2) < R2;
        }
        // ~1 candidate in 8 survives: the survivors are compacted into the warp's queue and the nine atom-pair tests run
        // on full warps
        const unsigned nm = __ballot_sync(0xffffffffu, near);
        if (near) queue[(qn + __popc(nm & ((1u << lane) - 1u))) & 63] = slot;
        qn += __popc(nm);
        __syncwarp();
        if (qn - qh >= 32) { pair_tests(queue[(qh + lane) & 63]); qh += 32; }
      }
    }
    if (qn > qh) pair_tests(lane < qn - qh ? queue[(qh + lane) & 63] : -1);
  }
  if (!WRITE) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_pair += __shfl_down_sync(0xffffffffu, n_pair, o);
    if (lane == 0) { d.row_count[RPB_TILE_PARTS * I + part] = n_tile; if (n_pair) atomicAdd(&d.vstat[1], n_pair); }
  }
}

// The whole rebuild as ONE cooperative kernel (grid-wide barriers between the phases): on the ~95 % of steps without a
// rebuild it costs a single launch that exits at once.
__device__ __forceinline__ void rebuild_phases(const Dev& d, cg::grid_group& grid, const int ncell) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const int gwarp = gtid >> 5, nwarps = gthreads >> 5, lane = threadIdx.x & 31;
  int* cursor = d.cell_count + (ncell + 1);
  // ---- clusters: runs of <= 3 consecutive atoms of one molecule
  for (int c = gtid; c < 2 * ncell + 2; c += gthreads) d.cell_count[c] = 0;
  for (int m = gtid; m < d.M; m += gthreads) d.mol_ncl[m] = (d.mol_natom[m] + 2) / 3;
  if (gtid == 0) { d.vstat[0] = 0ull; d.vstat[1] = 0ull; d.vstat[2] += 1ull; }
  grid.sync();
  if (blockIdx.x == 0) { const int nc = block_scan_exclusive(d.mol_ncl, d.mol_cl_first, d.M, 0); if (threadIdx.x == 0) *d.n_clusters = nc; }
  grid.sync();
  const int NC = *d.n_clusters;
  for (int m = gtid; m < d.M; m += gthreads) {
    const int f = d.mol_first[m], n = d.mol_natom[m], c0 = d.mol_cl_first[m];
    for (int k = 0; 3 * k < n; k++) {
      const int fa = f + 3 * k, na = min(3, n - 3 * k);
      d.cl_info[c0 + k] = fa | (na << 24);
      const double4 p0 = d.xq[fa];
      double ext2 = 0.0;
      for (int a = 1; a < na; a++) { const double4 p = d.xq[fa + a]; ext2 = fmax(ext2, (p.x - p0.x) * (p.x - p0.x) + (p.y - p0.y) * (p.y - p0.y) + (p.z - p0.z) * (p.z - p0.z)); }
      if (ext2 > 0.0) atomicMax(&d.vstat[0], (unsigned long long)__double_as_longlong(sqrt(ext2)));   // non-negative doubles order like their bit patterns
      const int c = cell_of(d, p0);
      d.atom_cell[c0 + k] = c;                      // (indexed by cluster in this kernel)
      atomicAdd(&d.cell_count[c], 1);
    }
  }
  for (int i = gtid; i < d.N; i += gthreads) d.vbuild_xq[i] = d.xq[i];     // what the reference-list accessor works from
  grid.sync();
  if (blockIdx.x == 0) block_scan_exclusive(d.cell_count, d.cell_start, ncell, 0);
  grid.sync();
  for (int I = gtid; I < NC; I += gthreads) {
    const int c = d.atom_cell[I];
    d.cell_atoms[d.cell_start[c] + atomicAdd(&cursor[c], 1)] = I;
  }
  grid.sync();
  sort_cells(d, ncell, gtid, gthreads);
  grid.sync();
  for (int s = gtid; s < NC; s += gthreads) {       // cell-sorted copies for the sweep
    const int I = d.cell_atoms[s], info = d.cl_info[I], f = info & 0xffffff, n = info >> 24;
    d.csort_info[s] = info; d.csort_mol[s] = d.mol_of_atom[f];
    for (int b = 0; b < 3; b++) d.csort_xq[3 * s + b] = d.xq[f + (b < n ? b : 0)];
  }
  grid.sync();
  const double R = sqrt(d.rv2) + 2.0 * __longlong_as_double((long long)d.vstat[0]) + 1e-9;
  __shared__ int sweep_queue[TPB / 32][64];       // per warp: near candidates waiting for their pair tests
  int* queue = sweep_queue[threadIdx.x >> 5];
  for (int w = gwarp; w < RPB_TILE_PARTS * NC; w += nwarps) tile_sweep<false>(d, w / RPB_TILE_PARTS, w % RPB_TILE_PARTS, lane, R, queue);
  grid.sync();
  if (blockIdx.x == 0) {
    block_scan_exclusive(d.row_count, d.tile_point, RPB_TILE_PARTS * NC, 0);
    // allocate_verlet_list / the overflow stop of :1562-1565: like the reference's half list, the tiles hold each listed pair once
    if (threadIdx.x == 0 && (long long)d.vstat[1] > (long long)d.verlet_cap) atomicMax(&d.err_flag[1], 1);
  }
  grid.sync();
  if (d.err_flag[1]) return;
  for (int w = gwarp; w < RPB_TILE_PARTS * NC; w += nwarps) {
    const int n = d.row_count[w];
    if (n <= RPB_TILE_TMPCAP) {                      // the usual case: copy the row out of its scratch
      const unsigned* __restrict__ src = d.tile_tmp + (size_t)w * RPB_TILE_TMPCAP;
      unsigned* __restrict__ dst = d.tile_list + d.tile_point[w];
      for (int k = lane; k < n; k += 32) dst[k] = src[k];
    } else tile_sweep<true>(d, w / RPB_TILE_PARTS, w % RPB_TILE_PARTS, lane, R, queue);
  }
}

__global__ void __launch_bounds__(TPB) k_verlet_rebuild(Dev d, int force_rebuild, int ncell) {
  cg::grid_group grid = cg::this_grid();
  // forced (init): flag_verlet_list untouched (flag_junk, ms_evb.f90:223-225)
  const int rb = force_rebuild ? 2 : ((*d.flag_verlet == 1) ? 1 : 0);
  // (d.vstat[4]: the largest cluster extent, re-accumulated by the k_verlet_disp that follows on this stream)
  if (blockIdx.x == 0 && threadIdx.x == 0) { *d.rebuild_now = rb; d.maxd[0] = 0.0; d.maxd[1] = 0.0; d.vstat[4] = 0ull; }
  if (!rb) return;
  rebuild_phases(d, grid, ncell);
}

// Hop commit (rpb_commit.cuh) + the forced rebuild + update_verlet_displacements(init) of ms_evb.f90:218-227 as ONE
// cooperative kernel at the end of every MS-EVB force evaluation: exits at once unless the solver selected a hop.
__global__ void __launch_bounds__(TPB) k_commit_and_rebuild(Dev d, CommitArgs a, int ncell) {
  if (!*d.commit_hop) return;
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  if (gtid == 0) commit_prepare(d, a.e, a.ci);
  grid.sync();
  for (int i = gtid; i < max(d.N, d.M); i += gthreads) commit_permute(d, a, i);
  grid.sync();
  for (int i = gtid; i < d.N; i += gthreads) commit_finish(d, a, i);
  grid.sync();
  rebuild_phases(d, grid, ncell);
  for (int i = gtid; i < d.N; i += gthreads) {       // update_verlet_displacements(init): flag_verlet_list untouched (flag_junk)
    const double4 p = d.xq[i];
    d.vstore[3 * i] = p.x; d.vstore[3 * i + 1] = p.y; d.vstore[3 * i + 2] = p.z;
    d.vdisp[3 * i] = 0.0; d.vdisp[3 * i + 1] = 0.0; d.vdisp[3 * i + 2] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Parity accessor: the reference's half list in the reference's row order, from the positions of the last rebuild.
// One warp per atom.  The reference visits the cells ia, ib, ic nested (:1523-1531) and, inside a cell, ascending atom
// index; cells are stored z-fastest here, so the innermost ic loop of one (ia, ib) column is ONE contiguous range of
// the cell-sorted arrays (two when the column wraps around the box), swept 64 atoms at a time with ballot-ordered
// appends.  Columns / cells that cannot hold an atom inside r_v are skipped (skipping empty-handed cells does not
// change the order of what is found).  Two passes (count, scan, write): no per-row capacity.
template <bool WRITE>
__device__ __forceinline__ void verlet_rows_atom(const Dev& d, const int i, const int lane) {
  const double4 pi = d.vbuild_xq[i];
  const int mi = d.mol_of_atom[i];
  const int c = d.atom_cell[i];
  const int iz = c % d.ncz + 1, iy = (c / d.ncz) % d.ncy + 1, ix = c / (d.ncz * d.ncy) + 1;
  int* __restrict__ out = WRITE ? d.neighbor_list + (d.verlet_point[i] - 1) : nullptr;
  int n_half = 0;
  const double wx = d.box[0] / d.ncx, wy = d.box[1] / d.ncy, wz = d.box[2] / d.ncz;
  double fx = (d.inv_box[0] * pi.x) * d.ncx; fx -= floor(fx);
  double fy = (d.inv_box[1] * pi.y) * d.ncy; fy -= floor(fy);
  double fz = (d.inv_box[2] * pi.z) * d.ncz; fz -= floor(fz);
  const bool prune_x = (d.dia + 1) * wx <= 0.5 * d.box[0], prune_y = (d.dib + 1) * wy <= 0.5 * d.box[1],
             prune_z = (d.dic + 1) * wz <= 0.5 * d.box[2];
  const double slack = 1e-6;     // cell units; the bounds only have to be conservative
  for (int ia = -d.dia; ia <= d.dia; ia++) {
    const int g1 = wrap_cell(ix + ia, d.ncx);
    double gx = 0.0;
    if (prune_x && ia != 0) gx = fmax(0.0, (ia > 0 ? (double)ia - fx : fx - (double)(ia + 1)) - slack) * wx;
    for (int ib = -d.dib; ib <= d.dib; ib++) {
      double gy = 0.0;
      if (prune_y && ib != 0) gy = fmax(0.0, (ib > 0 ? (double)ib - fy : fy - (double)(ib + 1)) - slack) * wy;
      const double rem = d.rv2 - (gx * gx + gy * gy);
      if (rem < 0.0) continue;                         // the whole column is out of reach
      int ic_lo = -d.dic, ic_hi = d.dic;
      if (prune_z) {
        const double zc = sqrt(rem) / wz + slack;      // reach along z in cell units
        ic_hi = min(d.dic, (int)floor(fz + zc));
        ic_lo = -min(d.dic, (int)floor(1.0 - fz + zc));
      }
      const int g2 = wrap_cell(iy + ib, d.ncy);
      const int col = d.ncz * ((g2 - 1) + d.ncy * (g1 - 1));
      // z segments of the column: [iz+ic_lo, iz+ic_hi] wrapped into 1..ncz, in the order the reference meets them
      int seg0[2], seg1[2], nseg = 1;
      {
        const int zl = iz + ic_lo, zh = iz + ic_hi;
        if (zh < 1) { seg0[0] = zl + d.ncz; seg1[0] = zh + d.ncz; }
        else if (zl > d.ncz) { seg0[0] = zl - d.ncz; seg1[0] = zh - d.ncz; }
        else if (zl < 1) { seg0[0] = zl + d.ncz; seg1[0] = d.ncz; seg0[1] = 1; seg1[1] = zh; nseg = 2; }
        else if (zh > d.ncz) { seg0[0] = zl; seg1[0] = d.ncz; seg0[1] = 1; seg1[1] = zh - d.ncz; nseg = 2; }
        else { seg0[0] = zl; seg1[0] = zh; }
      }
      for (int sg = 0; sg < nseg; sg++) {
        const int s = d.cell_start[col + seg0[sg] - 1], e = d.cell_start[col + seg1[sg]];
        for (int b = s; b < e; b += 64) {
          const int a0 = b + lane, a1 = b + 32 + lane;
          int j0 = -1, j1 = -1, m0 = mi, m1 = mi;
          double4 p0 = pi, p1 = pi;
          if (a0 < e) { j0 = d.vsort_entry[a0]; m0 = d.vsort_mol[a0]; p0 = ldg256(&d.vsort_xq[a0]); }
          if (a1 < e) { j1 = d.vsort_entry[a1]; m1 = d.vsort_mol[a1]; p1 = ldg256(&d.vsort_xq[a1]); }
          bool hit0 = false, hit1 = false;
          if (m0 != mi && i < j0) {
            double r0 = pi.x - p0.x, r1 = pi.y - p0.y, r2 = pi.z - p0.z;
            r0 = r0 - d.box[0] * floor_fp64pipe(r0 * d.inv_box[0] + 0.5);
            r1 = r1 - d.box[1] * floor_fp64pipe(r1 * d.inv_box[1] + 0.5);
            r2 = r2 - d.box[2] * floor_fp64pipe(r2 * d.inv_box[2] + 0.5);
            hit0 = (r0 * r0 + r1 * r1 + r2 * r2) < d.rv2;
          }
          if (m1 != mi && i < j1) {
            double r0 = pi.x - p1.x, r1 = pi.y - p1.y, r2 = pi.z - p1.z;
            r0 = r0 - d.box[0] * floor_fp64pipe(r0 * d.inv_box[0] + 0.5);
            r1 = r1 - d.box[1] * floor_fp64pipe(r1 * d.inv_box[1] + 0.5);
            r2 = r2 - d.box[2] * floor_fp64pipe(r2 * d.inv_box[2] + 0.5);
            hit1 = (r0 * r0 + r1 * r1 + r2 * r2) < d.rv2;
          }
          const unsigned below = (1u << lane) - 1u;
          const unsigned b0 = __ballot_sync(0xffffffffu, hit0), b1 = __ballot_sync(0xffffffffu, hit1);
          if (WRITE) {
            if (hit0) out[n_half + __popc(b0 & below)] = j0 + 1;
            if (hit1) out[n_half + __popc(b0) + __popc(b1 & below)] = j1 + 1;
          }
          n_half += __popc(b0) + __popc(b1);
        }
      }
    }
  }
  if (!WRITE && lane == 0) d.row_count[i] = n_half;
}

__global__ void __launch_bounds__(TPB) k_verlet_reference_list(Dev d, int ncell) {
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const int gwarp = gtid >> 5, nwarps = gthreads >> 5, lane = threadIdx.x & 31;
  int* cursor = d.cell_count + (ncell + 1);
  for (int c = gtid; c < 2 * ncell + 2; c += gthreads) d.cell_count[c] = 0;
  grid.sync();
  for (int i = gtid; i < d.N; i += gthreads) {       // cell of every atom (general_routines.f90:1458-1493)
    const int c = cell_of(d, d.vbuild_xq[i]);
    d.atom_cell[i] = c;
    atomicAdd(&d.cell_count[c], 1);
  }
  grid.sync();
  if (blockIdx.x == 0) block_scan_exclusive(d.cell_count, d.cell_start, ncell, 0);
  grid.sync();
  for (int i = gtid; i < d.N; i += gthreads) {
    const int c = d.atom_cell[i];
    d.cell_atoms[d.cell_start[c] + atomicAdd(&cursor[c], 1)] = i;
  }
  grid.sync();
  sort_cells(d, ncell, gtid, gthreads);
  grid.sync();
  for (int a = gtid; a < d.N; a += gthreads) {
    const int j = d.cell_atoms[a];
    d.vsort_xq[a] = d.vbuild_xq[j]; d.vsort_mol[a] = d.mol_of_atom[j]; d.vsort_entry[a] = j;
  }
  grid.sync();
  for (int i = gwarp; i < d.N; i += nwarps) verlet_rows_atom<false>(d, i, lane);
  grid.sync();
  if (blockIdx.x == 0) {
    const int total = block_scan_exclusive(d.row_count, d.verlet_point, d.N, 1);
    if (threadIdx.x == 0 && total > d.verlet_cap) atomicMax(&d.err_flag[1], 1);
  }
  grid.sync();
  if (d.err_flag[1]) return;
  for (int i = gwarp; i < d.N; i += nwarps) verlet_rows_atom<true>(d, i, lane);
}

// ------------------------------------------------------------------------------------------------
int verlet_setup(rpb_ctx* c) {     // per context: the cooperative grid of THIS context's device
  int per_sm = 0, per_sm2 = 0, per_sm3 = 0, sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_verlet_rebuild, TPB, 0) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_verlet_reference_list, TPB, 0) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm3, k_commit_and_rebuild, TPB, 0) != cudaSuccess || per_sm < 1 || per_sm2 < 1 || per_sm3 < 1) {
    c->err = "cannot size the cooperative neighbour-list kernels on this device";
    return RPB_ERR_CUDA;
  }
  // ONE block per SM: a cooperative grid starts only when all of its blocks can be resident, and the single-CTA enumeration
  // kernel (512 threads, 117 KB of shared memory) that runs next to it must not make it wait for an SM
  c->d.coop_blocks = std::max(2, sms);
  c->n_sm = sms;
  return 0;
}

static int verlet_common(rpb_ctx* c, int force_rebuild) {
  ScopedTimer t(c, T_VERLET);
  Dev& d = c->d;
  int ncell = d.ncx * d.ncy * d.ncz;
  double* blk_top2 = d.maxd + 8;
  const int nb = (d.N + TPB - 1) / TPB;
  void* args[] = {(void*)&d, (void*)&force_rebuild, (void*)&ncell};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_verlet_rebuild, dim3(d.coop_blocks), dim3(TPB), args, 0, c->stream);
  if (e != cudaSuccess) { c->err = std::string("neighbour-list rebuild launch: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  k_verlet_disp<<<nb, TPB, 0, c->stream>>>(d, blk_top2, d.vdone);
  e = cudaGetLastError();
  if (e != cudaSuccess) { c->err = std::string("k_verlet_disp launch: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  c->n_launch += 2;
  return 0;
}

int launch_verlet_update(rpb_ctx* c) { const int f = c->rebuild_forced ? 1 : 0; c->rebuild_forced = false; return verlet_common(c, f); }
int launch_verlet_force_rebuild(rpb_ctx* c) { c->rebuild_forced = false; return verlet_common(c, 1); }

int launch_commit_and_rebuild(rpb_ctx* c, const CommitArgs* a) {
  Dev& d = c->d;
  int ncell = d.ncx * d.ncy * d.ncz;
  void* args[] = {(void*)&d, (void*)a, (void*)&ncell};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_commit_and_rebuild, dim3(d.coop_blocks), dim3(TPB), args, 0, c->stream);
  if (e != cudaSuccess) { c->err = std::string("hop-commit launch: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  c->n_launch += 1;
  return 0;
}

// the reference-ordered half list of the last rebuild -> d.verlet_point / d.neighbor_list (accessor only)
int launch_verlet_reference_list(rpb_ctx* c) {
  Dev& d = c->d;
  int ncell = d.ncx * d.ncy * d.ncz;
  void* args[] = {(void*)&d, (void*)&ncell};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)k_verlet_reference_list, dim3(d.coop_blocks), dim3(TPB), args, 0, c->stream);
  if (e != cudaSuccess) { c->err = std::string("reference-list launch: ") + cudaGetErrorString(e); return RPB_ERR_CUDA; }
  c->n_launch += 1;
  return 0;
}
