// Batched PME reciprocal-space convolution  theta = IDFT( CB * DFT(Q) ),  E_rec = 1/2 sum Q theta,  hand-written for
// sm_100a in fp64.  Replaces the MKL DFTI sequence of pme.f90:73-129 / ms_evb.f90:2026-2050 (forward 3-D DFT, multiply
// by CB, backward 3-D DFT, energy) for every grid of the batch (principal diabat + all owned diabats) with THREE
// kernels instead of cuFFT's 3 + 3 passes and a separate CB pass:
//   A  (grid, z-slab)   real slab -> 2-D DFT over (x, y) in shared memory -> half spectrum            [reads 8 B, writes ~8.3 B / point]
//   B  (grid, ky-plane) DFT along z, x CB (fused E_rec, Parseval in k-space), inverse DFT along z, in place  [16.7 B / point]
//   C  (grid, z-slab)   inverse 2-D DFT over (ky, kx) -> real theta slab                                [reads ~8.3 B, writes 8 B]
// ~50 B of global traffic per grid point where the library path moves ~125 B.  The x direction uses the two-for-one
// trick (rows y, y+1 packed as one complex line), every 1-D transform is a mixed-radix (4, 3, 2) Stockham autosort FFT
// run by a group of 16 threads on a shared-memory line.  Unnormalised transform pair, forward exp(-i..), like DFTI
// with its default scale (ms_evb.f90:1999) and like cuFFT, which remains the path for grid sizes with other factors.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include "rpb_host.h"

#define FFT_GROUP 16                 // threads per 1-D transform
#define FFT_TPB 256                  // 16 transforms in flight per CTA
#define FFT_NG (FFT_TPB / FFT_GROUP)

struct FftPlan { int K, Kh, npass; int radix[8]; };

// scratch lines are stored with one padding slot after every 8 elements: the stride-4 stores of a radix-4 pass then
// fall into distinct 16-byte bank groups (6-way conflicts without it, measured with ncu)
__host__ __device__ __forceinline__ int fpad(int e) { return e + (e >> 3); }
#define FFT_LINE_PITCH(K) ((K) + ((K) >> 3) + 1)

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }

// ---- one Stockham pass, radix R, with everything but the data known at compile time:
//   N = length still to be transformed at this pass, S = stride (product of the radices already done), K = line length
template <int R, int N, int S, int K, bool INV>
__device__ __forceinline__ void fft_pass(const double2* __restrict__ src, double2* __restrict__ dst, const double2* __restrict__ W, int t) {
  constexpr int M = N / R, TW = K / N, NB = K / R;     // butterflies of the pass: NB, FFT_GROUP threads share them
#pragma unroll
  for (int i0 = 0; i0 < NB; i0 += FFT_GROUP) {
    const int idx = i0 + t;
    if (NB % FFT_GROUP != 0 && idx >= NB) break;
    const int pp = idx / S, q = idx % S;               // S is a compile-time constant: shifts / multiplies
    const int ib = q + S * pp, ob = q + S * R * pp;      // element indices; lines are stored padded (fpad)
    constexpr int istep = S * M;
    double2 w1 = W[pp * TW];
    if (INV) w1.y = -w1.y;
#define IN(k) src[fpad(ib + (k) * istep)]
#define OUT(j) dst[fpad(ob + (j) * S)]
    if (R == 4) {
      const double2 a0 = IN(0), a1 = IN(1), a2 = IN(2), a3 = IN(3);
      const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
      const double2 mi = INV ? make_double2(-d13.y, d13.x) : make_double2(d13.y, -d13.x);   // -/+ i (a1 - a3)
      double2 w2 = W[2 * pp * TW], w3 = W[3 * pp * TW];
      if (INV) { w2.y = -w2.y; w3.y = -w3.y; }
      OUT(0) = cadd(s02, s13);
      if (M == 1) { OUT(1) = cadd(d02, mi); OUT(2) = csub(s02, s13); OUT(3) = csub(d02, mi); }   // last pass: twiddles are 1
      else { OUT(1) = cmul(cadd(d02, mi), w1); OUT(2) = cmul(csub(s02, s13), w2); OUT(3) = cmul(csub(d02, mi), w3); }
    } else if (R == 3) {
      const double h = 0.86602540378443864676;   // sqrt(3)/2
      const double2 a0 = IN(0), a1 = IN(1), a2 = IN(2);
      const double2 s12 = cadd(a1, a2), d12 = csub(a1, a2);
      const double2 cc = make_double2(fma(-0.5, s12.x, a0.x), fma(-0.5, s12.y, a0.y));
      const double2 e = INV ? make_double2(-h * d12.y, h * d12.x) : make_double2(h * d12.y, -h * d12.x);
      double2 w2 = W[2 * pp * TW];
      if (INV) w2.y = -w2.y;
      OUT(0) = cadd(a0, s12);
      if (M == 1) { OUT(1) = cadd(cc, e); OUT(2) = csub(cc, e); }
      else { OUT(1) = cmul(cadd(cc, e), w1); OUT(2) = cmul(csub(cc, e), w2); }
    } else {
      const double2 a0 = IN(0), a1 = IN(1);
      OUT(0) = cadd(a0, a1);
      OUT(1) = (M == 1) ? csub(a0, a1) : cmul(csub(a0, a1), w1);
    }
#undef IN
#undef OUT
  }
}

// compile-time radix schedules of the supported line lengths (4s first, then 3s, then 2s -- as fft_make_plan)
template <int K, bool INV>
__device__ __forceinline__ double2* fft_line_fixed(double2* a, double2* b, const double2* __restrict__ W, int t, unsigned gmask) {
  if (K == 64) {
    fft_pass<4, 64, 1, 64, INV>(a, b, W, t); __syncwarp(gmask);
    fft_pass<4, 16, 4, 64, INV>(b, a, W, t); __syncwarp(gmask);
    fft_pass<4, 4, 16, 64, INV>(a, b, W, t); __syncwarp(gmask);
    return b;
  } else if (K == 48) {
    fft_pass<4, 48, 1, 48, INV>(a, b, W, t); __syncwarp(gmask);
    fft_pass<4, 12, 4, 48, INV>(b, a, W, t); __syncwarp(gmask);
    fft_pass<3, 3, 16, 48, INV>(a, b, W, t); __syncwarp(gmask);
    return b;
  } else if (K == 36) {
    fft_pass<4, 36, 1, 36, INV>(a, b, W, t); __syncwarp(gmask);
    fft_pass<3, 9, 4, 36, INV>(b, a, W, t); __syncwarp(gmask);
    fft_pass<3, 3, 12, 36, INV>(a, b, W, t); __syncwarp(gmask);
    return b;
  } else {   // K == 32
    fft_pass<4, 32, 1, 32, INV>(a, b, W, t); __syncwarp(gmask);
    fft_pass<4, 8, 4, 32, INV>(b, a, W, t); __syncwarp(gmask);
    fft_pass<2, 2, 16, 32, INV>(a, b, W, t); __syncwarp(gmask);
    return b;
  }
}

// One 1-D DFT of length K on a contiguous shared-memory line, by the FFT_GROUP threads of a group (t = 0..15).
// Stockham autosort: pass p reads src, writes dst, then the roles swap; returns the buffer that holds the result.
// W[j] = exp(-2 pi i j / K); INV conjugates twiddles and butterflies (unnormalised backward transform).
// KT > 0: line length known at compile time (32, 36, 48, 64); KT == 0: generic run-time schedule from the plan.
template <int KT, bool INV>
__device__ __forceinline__ double2* fft_line(double2* src, double2* dst, const double2* __restrict__ W, const FftPlan& P, int t, unsigned gmask) {
  if (KT > 0) return fft_line_fixed<KT, INV>(src, dst, W, t, gmask);
  int n = P.K, s = 1;
  for (int ps = 0; ps < P.npass; ps++) {
    const int r = P.radix[ps], m = n / r, tw = P.K / n;
    for (int idx = t; idx < P.K / r; idx += FFT_GROUP) {
      const int pp = idx / s, q = idx - pp * s;
      const int ib = q + s * pp, ob = q + s * r * pp;
      const int istep = s * m;
      double2 w1 = W[pp * tw];
      if (INV) w1.y = -w1.y;
#define IN(k) src[fpad(ib + (k) * istep)]
#define OUT(j) dst[fpad(ob + (j) * s)]
      if (r == 4) {
        const double2 a0 = IN(0), a1 = IN(1), a2 = IN(2), a3 = IN(3);
        const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
        const double2 mi = INV ? make_double2(-d13.y, d13.x) : make_double2(d13.y, -d13.x);   // -/+ i (a1 - a3)
        double2 w2 = W[2 * pp * tw], w3 = W[3 * pp * tw];
        if (INV) { w2.y = -w2.y; w3.y = -w3.y; }
        OUT(0) = cadd(s02, s13);
        OUT(1) = cmul(cadd(d02, mi), w1);
        OUT(2) = cmul(csub(s02, s13), w2);
        OUT(3) = cmul(csub(d02, mi), w3);
      } else if (r == 3) {
        const double h = 0.86602540378443864676;   // sqrt(3)/2
        const double2 a0 = IN(0), a1 = IN(1), a2 = IN(2);
        const double2 s12 = cadd(a1, a2), d12 = csub(a1, a2);
        const double2 cc = make_double2(fma(-0.5, s12.x, a0.x), fma(-0.5, s12.y, a0.y));
        const double2 e = INV ? make_double2(-h * d12.y, h * d12.x) : make_double2(h * d12.y, -h * d12.x);
        double2 w2 = W[2 * pp * tw];
        if (INV) w2.y = -w2.y;
        OUT(0) = cadd(a0, s12);
        OUT(1) = cmul(cadd(cc, e), w1);
        OUT(2) = cmul(csub(cc, e), w2);
      } else {
        const double2 a0 = IN(0), a1 = IN(1);
        OUT(0) = cadd(a0, a1);
        OUT(1) = cmul(csub(a0, a1), w1);
      }
#undef IN
#undef OUT
    }
    __syncwarp(gmask);
    double2* tmp = src; src = dst; dst = tmp;
    n = m; s *= r;
  }
  return src;
}

// shared memory: [slab K x Kh complex][FFT_NG x 2 lines of K complex][K twiddles]
struct FftSmem {
  double2 *slab, *scr, *W;
  __device__ FftSmem(unsigned char* base, const FftPlan& P) {
    slab = reinterpret_cast<double2*>(base);
    scr = slab + (size_t)P.K * P.Kh;
    W = scr + (size_t)FFT_NG * 2 * FFT_LINE_PITCH(P.K);
  }
};
static size_t fft_smem_bytes(const FftPlan& P) { return ((size_t)P.K * P.Kh + (size_t)FFT_NG * 2 * FFT_LINE_PITCH(P.K) + P.K) * sizeof(double2); }

__device__ __forceinline__ void load_twiddles(double2* W, const double2* __restrict__ Wg, int K) {
  for (int j = threadIdx.x; j < K; j += blockDim.x) W[j] = Wg[j];
}

// transform every column of the slab (lines along the slow index, stride = Kh) in place
template <int KT, bool INV>
__device__ __forceinline__ void fft_columns(const FftSmem& S, const FftPlan& P, int g, int t, unsigned gmask) {
  double2* l0 = S.scr + (size_t)g * 2 * FFT_LINE_PITCH(P.K);
  double2* l1 = l0 + FFT_LINE_PITCH(P.K);
  for (int col = g; col < P.Kh; col += FFT_NG) {
    for (int i = t; i < P.K; i += FFT_GROUP) l0[fpad(i)] = S.slab[(size_t)i * P.Kh + col];
    __syncwarp(gmask);
    const double2* res = fft_line<KT, INV>(l0, l1, S.W, P, t, gmask);
    for (int i = t; i < P.K; i += FFT_GROUP) S.slab[(size_t)i * P.Kh + col] = res[fpad(i)];
    __syncwarp(gmask);
  }
}

// ---- A: real slab (z fixed) -> half spectrum of the 2-D DFT over (x, y):  FQ[g][z][ky][kx], kx = 0..K/2
template <int KT>
__global__ void __launch_bounds__(FFT_TPB) k_fft_fwd_xy(FftPlan P, const double* __restrict__ Q, double2* __restrict__ FQ, const double2* __restrict__ Wg) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  FftSmem S(fft_smem, P);
  const int K = KT > 0 ? KT : P.K, Kh = K / 2 + 1, z = blockIdx.x;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)Kh * K * K;
  const double* q = Q + blockIdx.y * K3 + (size_t)z * K * K;
  double2* out = FQ + blockIdx.y * Kh3 + (size_t)z * K * Kh;
  const int g = threadIdx.x / FFT_GROUP, t = threadIdx.x % FFT_GROUP;
  const unsigned gmask = 0xffffu << (16 * ((threadIdx.x & 31) / 16));
  load_twiddles(S.W, Wg, K);
  __syncthreads();
  double2* l0 = S.scr + (size_t)g * 2 * FFT_LINE_PITCH(K);
  double2* l1 = l0 + FFT_LINE_PITCH(K);
  // x direction, two real rows per complex transform
  for (int y = 2 * g; y < K; y += 2 * FFT_NG) {
    for (int x = t; x < K; x += FFT_GROUP) l0[fpad(x)] = make_double2(q[(size_t)y * K + x], q[(size_t)(y + 1) * K + x]);
    __syncwarp(gmask);
    const double2* C = fft_line<KT, false>(l0, l1, S.W, P, t, gmask);
    for (int k = t; k < Kh; k += FFT_GROUP) {
      const double2 a = C[fpad(k)], b = C[fpad((K - k) % K)];
      // A_y = (C[k] + conj C[K-k]) / 2 ,  A_{y+1} = (C[k] - conj C[K-k]) / (2i)
      S.slab[(size_t)y * Kh + k] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
      S.slab[(size_t)(y + 1) * Kh + k] = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));
    }
    __syncwarp(gmask);
  }
  __syncthreads();
  fft_columns<KT, false>(S, P, g, t, gmask);
  __syncthreads();
  for (int e = threadIdx.x; e < K * Kh; e += blockDim.x) out[e] = S.slab[e];
}

// ---- B: plane ky fixed: DFT along z for every kx, x CB, E_rec, inverse DFT along z, back in place
template <int KT>
__global__ void __launch_bounds__(FFT_TPB) k_fft_z_conv(FftPlan P, double2* __restrict__ FQ, const double* __restrict__ CBh, const double2* __restrict__ Wg,
                                                        double conv, double* __restrict__ e_out) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  __shared__ double red[32];
  FftSmem S(fft_smem, P);
  const int K = KT > 0 ? KT : P.K, Kh = K / 2 + 1, ky = blockIdx.x;
  const size_t Kh3 = (size_t)Kh * K * K;
  double2* base = FQ + blockIdx.y * Kh3 + (size_t)ky * Kh;      // element (z, kx) at base[z*K*Kh + kx]
  const int g = threadIdx.x / FFT_GROUP, t = threadIdx.x % FFT_GROUP;
  const unsigned gmask = 0xffffu << (16 * ((threadIdx.x & 31) / 16));
  load_twiddles(S.W, Wg, K);
  for (int e = threadIdx.x; e < K * Kh; e += blockDim.x) { const int z = e / Kh, kx = e - z * Kh; S.slab[e] = base[(size_t)z * K * Kh + kx]; }
  __syncthreads();
  fft_columns<KT, false>(S, P, g, t, gmask);
  __syncthreads();
  double acc = 0.0;
  for (int e = threadIdx.x; e < K * Kh; e += blockDim.x) {
    const int kz = e / Kh, kx = e - kz * Kh;
    const double cb = __ldg(&CBh[(size_t)kx + (size_t)Kh * (ky + (size_t)K * kz)]);
    double2 v = S.slab[e];
    const double wgt = (kx == 0 || 2 * kx == K) ? 1.0 : 2.0;
    acc = fma(wgt * cb, fma(v.x, v.x, v.y * v.y), acc);
    v.x *= cb; v.y *= cb;
    S.slab[e] = v;
  }
  __syncthreads();
  fft_columns<KT, true>(S, P, g, t, gmask);
  __syncthreads();
  for (int e = threadIdx.x; e < K * Kh; e += blockDim.x) { const int z = e / Kh, kx = e - z * Kh; base[(size_t)z * K * Kh + kx] = S.slab[e]; }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&e_out[blockIdx.y], 0.5 * acc * conv);
}

// ---- C: slab z fixed: inverse DFT along ky, then along kx (Hermitian half -> two real rows per complex transform)
template <int KT>
__global__ void __launch_bounds__(FFT_TPB) k_fft_inv_xy(FftPlan P, const double2* __restrict__ FQ, double* __restrict__ theta, const double2* __restrict__ Wg) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  FftSmem S(fft_smem, P);
  const int K = KT > 0 ? KT : P.K, Kh = K / 2 + 1, z = blockIdx.x;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)Kh * K * K;
  const double2* in = FQ + blockIdx.y * Kh3 + (size_t)z * K * Kh;
  double* out = theta + blockIdx.y * K3 + (size_t)z * K * K;
  const int g = threadIdx.x / FFT_GROUP, t = threadIdx.x % FFT_GROUP;
  const unsigned gmask = 0xffffu << (16 * ((threadIdx.x & 31) / 16));
  load_twiddles(S.W, Wg, K);
  for (int e = threadIdx.x; e < K * Kh; e += blockDim.x) S.slab[e] = in[e];
  __syncthreads();
  fft_columns<KT, true>(S, P, g, t, gmask);
  __syncthreads();
  double2* l0 = S.scr + (size_t)g * 2 * FFT_LINE_PITCH(K);
  double2* l1 = l0 + FFT_LINE_PITCH(K);
  for (int y = 2 * g; y < K; y += 2 * FFT_NG) {
    // Z[kx] = A_y[kx] + i A_{y+1}[kx] over the full kx range; kx > K/2 from the Hermitian symmetry A[K-kx] = conj A[kx]
    for (int kx = t; kx < K; kx += FFT_GROUP) {
      double2 a, b;
      if (kx < Kh) { a = S.slab[(size_t)y * Kh + kx]; b = S.slab[(size_t)(y + 1) * Kh + kx]; }
      else { a = S.slab[(size_t)y * Kh + (K - kx)]; b = S.slab[(size_t)(y + 1) * Kh + (K - kx)]; a.y = -a.y; b.y = -b.y; }
      l0[fpad(kx)] = make_double2(a.x - b.y, a.y + b.x);
    }
    __syncwarp(gmask);
    const double2* R = fft_line<KT, true>(l0, l1, S.W, P, t, gmask);
    for (int x = t; x < K; x += FFT_GROUP) { const double2 v = R[fpad(x)]; out[(size_t)y * K + x] = v.x; out[(size_t)(y + 1) * K + x] = v.y; }
    __syncwarp(gmask);
  }
}

// ================================================================================================
// Register-blocked variant for K = 16 * R2 (R2 = 3: K = 48, R2 = 4: K = 64) -- the grid sizes of the benchmark configs.
// A line is transformed by R2 threads: thread r loads the decimated sub-sequence x[R2 j + r], j = 0..15, runs a
// 16-point FFT entirely in registers (two radix-4 stages, compile-time twiddles, 16-way instruction-level
// parallelism, no synchronisation), applies W_K^(r k) and parks the result in shared memory; after ONE warp-level
// sync thread u combines the R2 partial spectra into the 16 outputs X[k + 16 u].  10 (K = 48) or 8 (K = 64) lines per
// warp are in flight, against 2 for the 16-threads-per-line scheme above, and a line costs one sync instead of three.
// Shared-memory pitches (17 per partial spectrum, R2*17 per line, 27 / 34 per slab row) make the strided accesses of a
// quarter-warp fall into distinct 16-byte bank groups.
// ================================================================================================
template <int R2> struct L16 {
  static constexpr int K = 16 * R2, Kh = K / 2 + 1;
  static constexpr int LPW = 32 / R2;                       // lines per warp
  static constexpr int NW = (Kh + LPW - 1) / LPW;           // warps per CTA: one round covers the Kh columns
  static constexpr int TPB = 32 * NW, NL = LPW * NW;
  static constexpr int LP = R2 * 17;                        // line-buffer pitch (complex elements)
  static constexpr int PS = (R2 == 3) ? 27 : 34;            // slab row pitch  (>= Kh)
  static constexpr size_t SMEM = ((size_t)K * PS + (size_t)NL * LP + K) * sizeof(double2);
};

template <bool INV>
__device__ __forceinline__ void dft4(double2 a0, double2 a1, double2 a2, double2 a3, double2& b0, double2& b1, double2& b2, double2& b3) {
  const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
  const double2 mi = INV ? make_double2(-d13.y, d13.x) : make_double2(d13.y, -d13.x);   // -/+ i (a1 - a3)
  b0 = cadd(s02, s13); b1 = cadd(d02, mi); b2 = csub(s02, s13); b3 = csub(d02, mi);
}

// 16-point DFT in registers, natural order in and out:  n = 4a + b,  k = k1 + 4 k2,
//   X[k1 + 4 k2] = sum_b W4^(b k2) [ W16^(b k1) sum_a W4^(a k1) x[4a + b] ]
template <bool INV>
__device__ __forceinline__ void fft16(double2 (&v)[16]) {
  constexpr double c1 = 0.92387953251128673848, s1 = 0.38268343236508978178, c2 = 0.70710678118654752440;
  double2 t[4][4];
#pragma unroll
  for (int b = 0; b < 4; b++) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b], t[b][0], t[b][1], t[b][2], t[b][3]);
  // W16^m = (cos(pi m / 8), -/+ sin(pi m / 8)) for m = b k1
  const double sg = INV ? -1.0 : 1.0;
  t[1][1] = cmul(t[1][1], make_double2(c1, -sg * s1));
  t[1][2] = cmul(t[1][2], make_double2(c2, -sg * c2));
  t[1][3] = cmul(t[1][3], make_double2(s1, -sg * c1));
  t[2][1] = cmul(t[2][1], make_double2(c2, -sg * c2));
  t[2][2] = INV ? make_double2(-t[2][2].y, t[2][2].x) : make_double2(t[2][2].y, -t[2][2].x);   // W16^4 = -/+ i
  t[2][3] = cmul(t[2][3], make_double2(-c2, -sg * c2));
  t[3][1] = cmul(t[3][1], make_double2(s1, -sg * c1));
  t[3][2] = cmul(t[3][2], make_double2(-c2, -sg * c2));
  t[3][3] = cmul(t[3][3], make_double2(-c1, sg * s1));
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) dft4<INV>(t[0][k1], t[1][k1], t[2][k1], t[3][k1], v[k1], v[k1 + 4], v[k1 + 8], v[k1 + 12]);
}

// One line of length K = 16 R2 by the R2 threads (r = 0..R2-1) of a line slot; every lane of the warp calls this the
// same number of times (inactive slots pass active = false).  ld(e) returns element e, st(e, v) stores output e.
template <int R2, bool INV, typename LD, typename ST>
__device__ __forceinline__ void line_fft16(LD ld, ST st, double2* Z, const double2* __restrict__ W, int r, bool active) {
  constexpr int K = 16 * R2;
  double2 v[16];
  if (active) {
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = ld(R2 * j + r);
    fft16<INV>(v);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      double2 w = W[r * k];
      if (INV) w.y = -w.y;
      Z[r * 17 + k] = (r == 0) ? v[k] : cmul(v[k], w);
    }
  }
  __syncwarp();
  if (active) {
    double2 cf[R2];            // exp(-/+ 2 pi i r' u / R2), u = r
#pragma unroll
    for (int q = 0; q < R2; q++) { cf[q] = W[(K / R2) * ((q * r) % R2)]; if (INV) cf[q].y = -cf[q].y; }
#pragma unroll
    for (int k = 0; k < 16; k++) {
      double2 acc = Z[k];
#pragma unroll
      for (int q = 1; q < R2; q++) { const double2 z = Z[q * 17 + k]; acc.x = fma(cf[q].x, z.x, fma(-cf[q].y, z.y, acc.x)); acc.y = fma(cf[q].x, z.y, fma(cf[q].y, z.x, acc.y)); }
      v[k] = acc;
    }
  }
  __syncwarp();                // every partial spectrum has been read: st may overwrite the line buffer or the source
  if (active) {
#pragma unroll
    for (int k = 0; k < 16; k++) st(k + 16 * r, v[k]);
  }
}

template <int R2>
__global__ void __launch_bounds__(L16<R2>::TPB) k_fft16_fwd_xy(const double* __restrict__ Q, double2* __restrict__ FQ, const double2* __restrict__ Wg) {
  typedef L16<R2> C;
  constexpr int K = C::K, Kh = C::Kh, PS = C::PS;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* slab = reinterpret_cast<double2*>(fft_smem);
  double2* Zall = slab + (size_t)K * PS;
  double2* W = Zall + (size_t)C::NL * C::LP;
  const int z = blockIdx.x;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)Kh * K * K;
  const double* q = Q + blockIdx.y * K3 + (size_t)z * K * K;
  double2* out = FQ + blockIdx.y * Kh3 + (size_t)z * K * Kh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / R2, r = lane % R2, line = warp * C::LPW + slot;
  const bool lane_ok = lane < C::LPW * R2;
  double2* Z = Zall + (size_t)line * C::LP;
  load_twiddles(W, Wg, K);
  __syncthreads();
  {   // x direction: rows 2 line, 2 line + 1 packed into one complex transform, then untangled into the half spectrum
    const bool act = lane_ok && line < K / 2;
    const double* r0 = q + (size_t)(2 * line) * K;
    line_fft16<R2, false>([&](int e) { return make_double2(r0[e], r0[K + e]); },
                          [&](int e, double2 v) { Z[(e >> 4) * 17 + (e & 15)] = v; }, Z, W, r, act);
    __syncwarp();
    if (act)
      for (int k = r; k < Kh; k += R2) {
        const int km = (K - k) % K;
        const double2 a = Z[(k >> 4) * 17 + (k & 15)], b = Z[(km >> 4) * 17 + (km & 15)];
        slab[(size_t)(2 * line) * PS + k] = make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        slab[(size_t)(2 * line + 1) * PS + k] = make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x));
      }
  }
  __syncthreads();
  {   // y direction: column `line` of the slab, in place
    const bool act = lane_ok && line < Kh;
    line_fft16<R2, false>([&](int e) { return slab[(size_t)e * PS + line]; },
                          [&](int e, double2 v) { slab[(size_t)e * PS + line] = v; }, Z, W, r, act);
  }
  __syncthreads();
#pragma unroll
  for (int e0 = 0; e0 < K * Kh; e0 += C::TPB) { const int e = e0 + threadIdx.x; if (e < K * Kh) { const int ky = e / Kh, kx = e - ky * Kh; out[e] = slab[(size_t)ky * PS + kx]; } }
}

template <int R2>
__global__ void __launch_bounds__(L16<R2>::TPB) k_fft16_z_conv(double2* __restrict__ FQ, const double* __restrict__ CBh, const double2* __restrict__ Wg,
                                                              double conv, double* __restrict__ e_out) {
  typedef L16<R2> C;
  constexpr int K = C::K, Kh = C::Kh, PS = C::PS;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  __shared__ double red[32];
  double2* slab = reinterpret_cast<double2*>(fft_smem);
  double2* Zall = slab + (size_t)K * PS;
  double2* W = Zall + (size_t)C::NL * C::LP;
  const int ky = blockIdx.x;
  const size_t Kh3 = (size_t)Kh * K * K;
  double2* base = FQ + blockIdx.y * Kh3 + (size_t)ky * Kh;      // element (z, kx) at base[z*K*Kh + kx]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / R2, r = lane % R2, line = warp * C::LPW + slot;
  const bool act = lane < C::LPW * R2 && line < Kh;
  double2* Z = Zall + (size_t)line * C::LP;
  load_twiddles(W, Wg, K);
  {   // all loads of the plane in flight before the first is consumed (a plain loop serialises one L2 round trip per element)
    constexpr int NIT = (K * Kh + C::TPB - 1) / C::TPB;
    double2 tmp[NIT];
#pragma unroll
    for (int i = 0; i < NIT; i++) { const int e = i * C::TPB + threadIdx.x; if (e < K * Kh) { const int zz = e / Kh, kx = e - zz * Kh; tmp[i] = base[(size_t)zz * K * Kh + kx]; } }
#pragma unroll
    for (int i = 0; i < NIT; i++) { const int e = i * C::TPB + threadIdx.x; if (e < K * Kh) { const int zz = e / Kh, kx = e - zz * Kh; slab[(size_t)zz * PS + kx] = tmp[i]; } }
  }
  __syncthreads();
  line_fft16<R2, false>([&](int e) { return slab[(size_t)e * PS + line]; },
                        [&](int e, double2 v) { slab[(size_t)e * PS + line] = v; }, Z, W, r, act);
  __syncthreads();
  double acc = 0.0;
  {
    constexpr int NIT = (K * Kh + C::TPB - 1) / C::TPB;
    double cbv[NIT];
#pragma unroll
    for (int i = 0; i < NIT; i++) { const int e = i * C::TPB + threadIdx.x; cbv[i] = 0.0; if (e < K * Kh) { const int kz = e / Kh, kx = e - kz * Kh; cbv[i] = __ldg(&CBh[(size_t)kx + (size_t)Kh * (ky + (size_t)K * kz)]); } }
#pragma unroll
    for (int i = 0; i < NIT; i++) {
      const int e = i * C::TPB + threadIdx.x;
      if (e < K * Kh) {
        const int kz = e / Kh, kx = e - kz * Kh;
        const double cb = cbv[i];
        double2 v = slab[(size_t)kz * PS + kx];
        const double wgt = (kx == 0 || 2 * kx == K) ? 1.0 : 2.0;
        acc = fma(wgt * cb, fma(v.x, v.x, v.y * v.y), acc);
        v.x *= cb; v.y *= cb;
        slab[(size_t)kz * PS + kx] = v;
      }
    }
  }
  __syncthreads();
  line_fft16<R2, true>([&](int e) { return slab[(size_t)e * PS + line]; },
                       [&](int e, double2 v) { slab[(size_t)e * PS + line] = v; }, Z, W, r, act);
  __syncthreads();
#pragma unroll
  for (int e0 = 0; e0 < K * Kh; e0 += C::TPB) { const int e = e0 + threadIdx.x; if (e < K * Kh) { const int zz = e / Kh, kx = e - zz * Kh; base[(size_t)zz * K * Kh + kx] = slab[(size_t)zz * PS + kx]; } }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&e_out[blockIdx.y], 0.5 * acc * conv);
}

template <int R2>
__global__ void __launch_bounds__(L16<R2>::TPB) k_fft16_inv_xy(const double2* __restrict__ FQ, double* __restrict__ theta, const double2* __restrict__ Wg) {
  typedef L16<R2> C;
  constexpr int K = C::K, Kh = C::Kh, PS = C::PS;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* slab = reinterpret_cast<double2*>(fft_smem);
  double2* Zall = slab + (size_t)K * PS;
  double2* W = Zall + (size_t)C::NL * C::LP;
  const int z = blockIdx.x;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)Kh * K * K;
  const double2* in = FQ + blockIdx.y * Kh3 + (size_t)z * K * Kh;
  double* out = theta + blockIdx.y * K3 + (size_t)z * K * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = lane / R2, r = lane % R2, line = warp * C::LPW + slot;
  const bool lane_ok = lane < C::LPW * R2;
  double2* Z = Zall + (size_t)line * C::LP;
  load_twiddles(W, Wg, K);
  {
    constexpr int NIT = (K * Kh + C::TPB - 1) / C::TPB;
    double2 tmp[NIT];
#pragma unroll
    for (int i = 0; i < NIT; i++) { const int e = i * C::TPB + threadIdx.x; if (e < K * Kh) tmp[i] = in[e]; }
#pragma unroll
    for (int i = 0; i < NIT; i++) { const int e = i * C::TPB + threadIdx.x; if (e < K * Kh) { const int ky = e / Kh, kx = e - ky * Kh; slab[(size_t)ky * PS + kx] = tmp[i]; } }
  }
  __syncthreads();
  {
    const bool act = lane_ok && line < Kh;
    line_fft16<R2, true>([&](int e) { return slab[(size_t)e * PS + line]; },
                         [&](int e, double2 v) { slab[(size_t)e * PS + line] = v; }, Z, W, r, act);
  }
  __syncthreads();
  {   // x direction: Z[kx] = A_y[kx] + i A_{y+1}[kx] over the full kx range (Hermitian extension), two real rows out
    const bool act = lane_ok && line < K / 2;
    const double2* ra = slab + (size_t)(2 * line) * PS;
    double* o0 = out + (size_t)(2 * line) * K;
    line_fft16<R2, true>([&](int e) {
                           double2 a, b;
                           if (e < Kh) { a = ra[e]; b = ra[PS + e]; }
                           else { a = ra[K - e]; b = ra[PS + K - e]; a.y = -a.y; b.y = -b.y; }
                           return make_double2(a.x - b.y, a.y + b.x);
                         },
                         [&](int e, double2 v) { o0[e] = v.x; o0[K + e] = v.y; }, Z, W, r, act);
  }
}

// ------------------------------------------------------------------------------------------------
static bool fft_make_plan(int K, FftPlan& P) {
  P.K = K; P.Kh = K / 2 + 1; P.npass = 0;
  if (K % 2 || K < 8 || K > 128) return false;      // the two-for-one packing needs an even K; one line per 16 threads
  int n = K;
  while (n % 4 == 0) { P.radix[P.npass++] = 4; n /= 4; }
  while (n % 3 == 0) { P.radix[P.npass++] = 3; n /= 3; }
  while (n % 2 == 0) { P.radix[P.npass++] = 2; n /= 2; }
  return n == 1 && P.npass <= 8;
}

// calls f with std::integral_constant<int, KT>: the compile-time specialisation for this grid size, or 0 (generic)
template <typename F>
static void fft_dispatch(int K, F&& f) {
  switch (K) {
    case 32: f(std::integral_constant<int, 32>()); break;
    case 36: f(std::integral_constant<int, 36>()); break;
    case 48: f(std::integral_constant<int, 48>()); break;
    case 64: f(std::integral_constant<int, 64>()); break;
    default: f(std::integral_constant<int, 0>()); break;
  }
}

struct FftState { bool tried = false, ok = false, group16 = false; FftPlan plan; double2* W = nullptr; };
static std::map<rpb_ctx*, FftState> g_fft;

void fft_conv_free(rpb_ctx* c) { g_fft.erase(c); }

// one-time set-up; < 0 on error, else 1 if this context uses the hand-written path, 0 if it needs cuFFT
static int fft_conv_init(rpb_ctx* c) {
  FftState& st = g_fft[c];
  if (!st.tried) {
    st.tried = true;
    const char* mode = getenv("RPB_FFT");                 // RPB_FFT=cufft forces the library path (cross-check)
    st.ok = fft_make_plan(c->d.K, st.plan) && !(mode && std::string(mode) == "cufft");
    if (st.ok) {
      const int K = c->d.K;
      std::vector<double2> w(K);
      for (int j = 0; j < K; j++) { const double a = -2.0 * 3.14159265358979323846 * (double)j / (double)K; w[j] = make_double2(cos(a), sin(a)); }
      int rc = dev_alloc(c, &st.W, (size_t)K);
      if (rc) return rc;
      if (cudaMemcpy(st.W, w.data(), K * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess) { c->err = "twiddle upload failed"; return RPB_ERR_CUDA; }
      const int smem = (int)fft_smem_bytes(st.plan);
      st.group16 = mode && std::string(mode) == "group16";     // RPB_FFT=group16: the 16-threads-per-line kernels for 48 / 64 too
      cudaFuncSetAttribute(k_fft16_fwd_xy<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<3>::SMEM);
      cudaFuncSetAttribute(k_fft16_z_conv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<3>::SMEM);
      cudaFuncSetAttribute(k_fft16_inv_xy<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<3>::SMEM);
      cudaFuncSetAttribute(k_fft16_fwd_xy<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<4>::SMEM);
      cudaFuncSetAttribute(k_fft16_z_conv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<4>::SMEM);
      cudaFuncSetAttribute(k_fft16_inv_xy<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L16<4>::SMEM);
      fft_dispatch(st.plan.K, [&](auto kt) {
        constexpr int KT = decltype(kt)::value;
        cudaFuncSetAttribute(k_fft_fwd_xy<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_fft_z_conv<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_fft_inv_xy<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      });
    }
  }
  return st.ok ? 1 : 0;
}
int fft_conv_supported(rpb_ctx* c) { return fft_conv_init(c); }

// returns 1 if the hand-written path handled the batch, 0 if the grid size needs the cuFFT path, < 0 on error
int fft_conv_batched(rpb_ctx* c, int first_grid, int n_grids, double* e_recip_dev, bool inverse) {
  const int ini = fft_conv_init(c);
  if (ini <= 0) return ini;
  FftState& st = g_fft[c];
  const FftPlan& P = st.plan;
  const int K = P.K;
  const size_t K3 = (size_t)K * K * K, Kh3 = (size_t)P.Kh * K * K, smem = fft_smem_bytes(P);
  double2* FQ = reinterpret_cast<double2*>(c->d.FQ) + Kh3 * first_grid;
  const dim3 grid(K, n_grids);
  if ((K == 48 || K == 64) && !st.group16) {
    auto run = [&](auto r2) {
      constexpr int R2 = decltype(r2)::value;
      typedef L16<R2> C;
      {
        ScopedTimer t(c, T_FFT);
        k_fft16_fwd_xy<R2><<<grid, C::TPB, C::SMEM, c->stream>>>(c->d.Q + K3 * first_grid, FQ, st.W);
        c->n_launch += 1;
      }
      {
        ScopedTimer t(c, T_CONV);
        cudaMemsetAsync(e_recip_dev + first_grid, 0, n_grids * sizeof(double), c->stream);
        k_fft16_z_conv<R2><<<grid, C::TPB, C::SMEM, c->stream>>>(FQ, c->d.CBh, st.W, c->d.conv, e_recip_dev + first_grid);
        c->n_launch += 1;
      }
      if (inverse) {
        ScopedTimer t(c, T_FFT);
        k_fft16_inv_xy<R2><<<grid, C::TPB, C::SMEM, c->stream>>>(FQ, c->d.theta + K3 * first_grid, st.W);
        c->n_launch += 1;
      }
    };
    if (K == 48) run(std::integral_constant<int, 3>()); else run(std::integral_constant<int, 4>());
    return 1;
  }
  fft_dispatch(K, [&](auto kt) {
    constexpr int KT = decltype(kt)::value;
    {
      ScopedTimer t(c, T_FFT);
      k_fft_fwd_xy<KT><<<grid, FFT_TPB, smem, c->stream>>>(P, c->d.Q + K3 * first_grid, FQ, st.W);
      c->n_launch += 1;
    }
    {
      ScopedTimer t(c, T_CONV);
      cudaMemsetAsync(e_recip_dev + first_grid, 0, n_grids * sizeof(double), c->stream);
      k_fft_z_conv<KT><<<grid, FFT_TPB, smem, c->stream>>>(P, FQ, c->d.CBh, st.W, c->d.conv, e_recip_dev + first_grid);
      c->n_launch += 1;
    }
    if (inverse) {
      ScopedTimer t(c, T_FFT);
      k_fft_inv_xy<KT><<<grid, FFT_TPB, smem, c->stream>>>(P, FQ, c->d.theta + K3 * first_grid, st.W);
      c->n_launch += 1;
    }
  });
  return 1;
}
