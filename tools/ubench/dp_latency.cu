// Dependent-issue latency and throughput of the fp64 pipe on this GPU (cycles per instruction, one warp / many warps).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat(double* out, long long* cyc, int n, int mode) {
  double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 1e-9;
  double a2 = a + 1, a3 = a + 2, a4 = a + 3;
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < n; i++) { a = fma(a, b, c); }
  if (mode == 1) for (int i = 0; i < n; i++) { a = a * b; }
  if (mode == 2) for (int i = 0; i < n; i++) { a = a + c; }
  if (mode == 3) for (int i = 0; i < n; i++) { a = fma(a, b, c); a2 = fma(a2, b, c); }
  if (mode == 4) for (int i = 0; i < n; i++) { a = fma(a, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); a4 = fma(a4, b, c); }
  if (mode == 5) for (int i = 0; i < n; i++) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); a = y + 1.5; }
  if (mode == 6) for (int i = 0; i < n; i++) { a = __dadd_rd(a, 6755399441055744.0) - 6755399441055744.0 + 0.3; }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + a2 + a3 + a4;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const char* names[] = {"DFMA chain", "DMUL chain", "DADD chain", "2 DFMA chains", "4 DFMA chains", "RSQ64H+DADD chain", "magic floor (3 dep DADD)"};
  const int n = 4096;
  for (int mode = 0; mode < 7; mode++)
    for (int warps : {1, 4, 8, 16, 32}) {      // warps per SM (one CTA per SM, 148 CTAs)
      k_lat<<<148, 32 * warps>>>(out, cyc, n, mode);
      k_lat<<<148, 32 * warps>>>(out, cyc, n, mode);
      long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-26s warps/SM %2d : %.2f cycles per loop iteration\n", names[mode], warps, (double)h / n);
    }
  return 0;
}
