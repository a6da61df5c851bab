// Micro-benchmarks (B200): dependent DFMA latency, DFMA throughput vs warps,
// L1-hit __ldg latency (warp-uniform), shared-memory load latency.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_chain(double *out, double a, double b, int n, long long *cyc) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i += 8) {
    x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
    x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dfma_indep(double *out, double a, double b, int n, long long *cyc) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i += 8) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void ldg_chain(const int *tab, int n, int *out, long long *cyc) {
  int idx = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) idx = __ldg(tab + idx);
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void lds_chain(int n, int *out, long long *cyc) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 32) & 1023;
  __syncthreads();
  int idx = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; ++i) idx = s[idx];
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out; long long *cyc; int *tab, *iout;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8); cudaMalloc(&tab, 4096 * 4); cudaMalloc(&iout, 4096);
  int h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i + 8) & 1023;
  cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
  long long c; const int n = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    dfma_chain<<<1, 32 * warps>>>(out, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, %2d warps/SM: %.2f cycles per DFMA (warp 0)\n", warps, (double)c / n);
  }
  for (int warps : {1, 4, 8, 16, 32}) {
    dfma_indep<<<1, 32 * warps>>>(out, 1.0000001, 1e-9, n, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 independent chains, %2d warps/SM: %.2f cycles per DFMA per warp -> %.1f DFMA lanes/clk/SM\n", warps, (double)c / n,
           32.0 * warps * n / c);
  }
  ldg_chain<<<1, 32>>>(tab, n, iout, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("__ldg pointer chase (L1 hit, warp-uniform): %.1f cycles\n", (double)c / n);
  lds_chain<<<1, 32>>>(n, iout, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS pointer chase: %.1f cycles\n", (double)c / n);
  return 0;
}
