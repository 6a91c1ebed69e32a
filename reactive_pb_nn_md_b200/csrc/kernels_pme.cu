// Smooth-PME reciprocal space on the device.  Replaces:
//   create_scaled_direct_coordinates   src/general_routines.f90:497-524
//   grid_Q                             src/pme.f90:184-264
//   FFT * CB * FFT^-1, E = 1/2 sum Q theta   src/pme.f90:73-129  (MKL DFTI -> cuFFT D2Z/Z2D)
//   derivative_grid_Q                  src/pme.f90:346-498
// Table look-ups keep the reference's nearest-above index arithmetic ceil(arg/6*1e5) bit for bit
// (division, then multiply, no contraction), so the weights are the reference's weights.
//
// The input grid is real and CB is real and even for an orthorhombic box, so the half-spectrum
// D2Z/Z2D pair is exact; E_rec is accumulated in k-space in the same pass that applies CB
// (Parseval: sum_x Q theta = sum_k CB |F|^2 for the unnormalised pair) -- no extra pass over Q, theta.
#include "rpb_host.h"
#include "rpb_pme.cuh"

#define TPB 256

__global__ void k_scaled_coords(Dev d) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.N) return;
  double4 p = d.xq[i];
  double x[3] = {p.x, p.y, p.z}, u[3];
  scaled_coords(d, x, u);
  d.uscale[3 * i] = u[0]; d.uscale[3 * i + 1] = u[1]; d.uscale[3 * i + 2] = u[2];
}

__global__ void k_spread(Dev d, double* Q) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= d.N) return;
  double u[3] = {d.uscale[3 * w], d.uscale[3 * w + 1], d.uscale[3 * w + 2]};
  spread_atom_warp(d, Q, u, d.xq[w].w, 1.0, lane);
}

// FQ *= CB (half spectrum) and E_s = 1/2 conv sum_k w_k CB |F|^2, for n_grids grids
__global__ void k_conv_energy(Dev d, cufftDoubleComplex* FQ, int n_grids, double* e_out) {
  __shared__ double sh[32];
  int Kh = d.K / 2 + 1;
  size_t per = (size_t)Kh * d.K * d.K;
  int g = blockIdx.y;
  double acc = 0.0;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < per; e += (size_t)gridDim.x * blockDim.x) {
    int m1 = (int)(e % Kh);
    double cb = __ldg(&d.CBh[e]);
    cufftDoubleComplex v = FQ[(size_t)g * per + e];
    double wgt = (m1 == 0 || (2 * m1 == d.K)) ? 1.0 : 2.0;
    acc += wgt * cb * (v.x * v.x + v.y * v.y);
    v.x = v.x * cb; v.y = v.y * cb;
    FQ[(size_t)g * per + e] = v;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(&e_out[g], 0.5 * acc * d.conv);
}

__global__ void k_gather(Dev d, const double* __restrict__ theta, double* out, int add_to_force) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= d.N) return;
  double u[3] = {d.uscale[3 * w], d.uscale[3 * w + 1], d.uscale[3 * w + 2]};
  double F[3];
  gather_atom_warp(d, theta, u, d.xq[w].w, lane, F);
  if (lane < 3) {
    double v = lane == 0 ? F[0] : (lane == 1 ? F[1] : F[2]);
    if (out) out[3 * w + lane] = v;
    if (add_to_force) atomicAdd(&d.force[3 * w + lane], v);
  }
}

// ------------------------------------------------------------------------------------------------
// cuFFT plans are cached per batch size; batch sizes are rounded up to a multiple of 4 (the spare grids behind the
// last owned diabat are transformed too and ignored) so that a run whose number of diabats changes from step to step
// never builds a plan inside the step: rpb_set_evb pre-creates every size.  All plans share one work area (they run
// on one stream).
int pme_round_batch(int batch) { return batch <= 1 ? 1 : (batch + 3) / 4 * 4; }

int pme_get_plans(rpb_ctx* ctx, int batch, cufftHandle* fwd, cufftHandle* inv) {
  batch = pme_round_batch(batch);
  auto it = ctx->plan_fwd.find(batch);
  if (it == ctx->plan_fwd.end()) {
    int K = ctx->d.K;
    int n[3] = {K, K, K};
    cufftHandle pf, pi;
    size_t wf = 0, wi = 0;
    if (cufftCreate(&pf) != CUFFT_SUCCESS || cufftCreate(&pi) != CUFFT_SUCCESS || cufftSetAutoAllocation(pf, 0) != CUFFT_SUCCESS ||
        cufftSetAutoAllocation(pi, 0) != CUFFT_SUCCESS ||
        cufftMakePlanMany(pf, 3, n, nullptr, 1, K * K * K, nullptr, 1, K * K * (K / 2 + 1), CUFFT_D2Z, batch, &wf) != CUFFT_SUCCESS ||
        cufftMakePlanMany(pi, 3, n, nullptr, 1, K * K * (K / 2 + 1), nullptr, 1, K * K * K, CUFFT_Z2D, batch, &wi) != CUFFT_SUCCESS) {
      ctx->err = "cufftMakePlanMany failed";
      return RPB_ERR_CUDA;
    }
    cufftSetStream(pf, ctx->stream);
    cufftSetStream(pi, ctx->stream);
    ctx->plan_fwd[batch] = pf;
    ctx->plan_inv[batch] = pi;
    size_t need = std::max(wf, wi);
    if (need > ctx->fft_work_bytes) {
      cudaStreamSynchronize(ctx->stream);
      char* w = nullptr;
      int rc = dev_alloc(ctx, &w, need);
      if (rc) return rc;
      ctx->fft_work = w; ctx->fft_work_bytes = need;
      for (auto& kv : ctx->plan_fwd) cufftSetWorkArea(kv.second, w);
      for (auto& kv : ctx->plan_inv) cufftSetWorkArea(kv.second, w);
    } else {
      cufftSetWorkArea(pf, ctx->fft_work);
      cufftSetWorkArea(pi, ctx->fft_work);
    }
  }
  *fwd = ctx->plan_fwd[batch];
  *inv = ctx->plan_inv[batch];
  return 0;
}

void launch_scaled_coords(rpb_ctx* c) {
  k_scaled_coords<<<(c->d.N + TPB - 1) / TPB, TPB, 0, c->stream>>>(c->d);
  c->n_launch += 1;
}

void launch_spread_principal(rpb_ctx* c) {
  ScopedTimer t(c, T_SPREAD);
  size_t K3 = (size_t)c->d.K * c->d.K * c->d.K;
  cudaMemsetAsync(c->d.Q, 0, K3 * sizeof(double), c->stream);
  launch_scaled_coords(c);
  k_spread<<<(c->d.N * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(c->d, c->d.Q);
  c->n_launch += 1;
}

int launch_convolve(rpb_ctx* c, int first_grid, int n_grids, double* e_recip_dev, bool inverse) {
  if (n_grids <= 0) return 0;
  {   // hand-written fused path (kernels_fft.cu); cuFFT below only for grid sizes with factors other than 2 and 3
    const int r = fft_conv_batched(c, first_grid, n_grids, e_recip_dev, inverse);
    if (r != 0) return r < 0 ? r : 0;
  }
  cufftHandle pf, pi;
  int rc = pme_get_plans(c, n_grids, &pf, &pi);
  if (rc) return rc;
  int K = c->d.K;
  size_t K3 = (size_t)K * K * K, Kh3 = (size_t)K * K * (K / 2 + 1);
  {
    ScopedTimer t(c, T_FFT);
    cufftSetStream(pf, c->stream);
    if (cufftExecD2Z(pf, c->d.Q + K3 * first_grid, c->d.FQ + Kh3 * first_grid) != CUFFT_SUCCESS) { c->err = "cufftExecD2Z failed"; return RPB_ERR_CUDA; }
    c->n_fft += 1;
  }
  {
    ScopedTimer t(c, T_CONV);
    cudaMemsetAsync(e_recip_dev + first_grid, 0, n_grids * sizeof(double), c->stream);
    dim3 grid((unsigned)std::min<size_t>((Kh3 + TPB - 1) / TPB, 148 * 4), n_grids);
    k_conv_energy<<<grid, TPB, 0, c->stream>>>(c->d, c->d.FQ + Kh3 * first_grid, n_grids, e_recip_dev + first_grid);
    c->n_launch += 1;
  }
  if (inverse) {
    ScopedTimer t(c, T_FFT);
    cufftSetStream(pi, c->stream);
    if (cufftExecZ2D(pi, c->d.FQ + Kh3 * first_grid, c->d.theta + K3 * first_grid) != CUFFT_SUCCESS) { c->err = "cufftExecZ2D failed"; return RPB_ERR_CUDA; }
    c->n_fft += 1;
  }
  return 0;
}

void launch_gather(rpb_ctx* c, const double* theta, double* out_force, bool add_to_force) {
  ScopedTimer t(c, T_GATHER);
  k_gather<<<(c->d.N * 32 + TPB - 1) / TPB, TPB, 0, c->stream>>>(c->d, theta, out_force, add_to_force ? 1 : 0);
  c->n_launch += 1;
}
