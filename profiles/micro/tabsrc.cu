// Micro-benchmark (B200): where should the per-row factor tables of the
// partitioned solve come from?  Every thread runs the chunk forward + backward
// recurrences (M = 32 rows in registers, 5 table planes, same arithmetic as
// chunk_core.cuh) `iters` times; 256-thread blocks (16 lines x 16 chunks), two
// per SM, like the y-sweep.  Table source:
//   0  shared memory, 16-byte loads (what the strided kernels do today)
//   1  kernel-parameter constant bank, per-thread (warp-uniform) index -> LDC
//   2  global memory through L1 (__ldg), 16-byte loads
//   3  shared memory, but the chunk index is blockIdx-uniform (LDS broadcast of
//      one address per warp) - isolates the 2-chunks-per-warp effect
// Output: cycles per block-iteration and row-updates per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int M = 32, P = 16, L = M * P;
struct Tab { double v[5 * L]; };

template <int MODE, int W>
__global__ void __launch_bounds__(256, 2) run(const __grid_constant__ Tab tab, const double *__restrict__ gtab, double *out,
                                               int iters, long long *cyc) {
  __shared__ __align__(16) double s_tab[5 * L];
  const int w = threadIdx.x, p = threadIdx.y;
  for (int e = p * W + w; e < 5 * L; e += 256) s_tab[e] = gtab[e];
  __syncthreads();
  double v[M];
#pragma unroll
  for (int t = 0; t < M; ++t) v[t] = 1e-3 * (w + t) + p;
  const int r0 = (p % P) * M;
  double acc = 0.0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    auto ld2 = [&](int plane, int t) -> double2 {
      if (MODE == 0) return *reinterpret_cast<const double2 *>(s_tab + plane * L + r0 + t);
      if (MODE == 1) return make_double2(tab.v[plane * L + r0 + t], tab.v[plane * L + r0 + t + 1]);
      if (MODE == 2) return __ldg(reinterpret_cast<const double2 *>(gtab + plane * L + r0 + t));
      if (MODE == 4) {
        const volatile double *q = s_tab + plane * L + r0 + t;   // two 8-byte shared loads
        return make_double2(q[0], q[1]);
      }
      return make_double2(__ldg(gtab + plane * L + r0 + t), __ldg(gtab + plane * L + r0 + t + 1));
    };
#pragma unroll
    for (int t = 0; t < M; t += 2) { const double2 c = ld2(0, t); v[t] *= c.x; v[t + 1] *= c.y; }
    double prev = 0.0;
#pragma unroll
    for (int t = 0; t < M; t += 2) {
      const double2 c = ld2(1, t);
      prev = fma(-c.x, prev, v[t]); v[t] = prev;
      prev = fma(-c.y, prev, v[t + 1]); v[t + 1] = prev;
    }
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int t = 0; t < M; t += 2) { const double2 c = ld2(2, t); a0 = fma(c.x, v[t], a0); a1 = fma(c.y, v[t + 1], a1); }
    const double alpha = a0 + a1;
#pragma unroll
    for (int t = 0; t < M; t += 2) { const double2 c = ld2(3, t); v[t] = fma(-alpha, c.x, v[t]); v[t + 1] = fma(-alpha, c.y, v[t + 1]); }
    double nxt = v[M - 1];
#pragma unroll
    for (int t = M - 2; t >= 0; t -= 2) {
      const double2 c = ld2(4, t);
      if (t + 1 < M - 1) { nxt = fma(-c.y, nxt, v[t + 1]); v[t + 1] = nxt; }
      nxt = fma(-c.x, nxt, v[t]); v[t] = nxt;
    }
    acc += nxt;
  }
  const long long t1 = clock64();
  double s = acc;
#pragma unroll
  for (int t = 0; t < M; ++t) s += v[t];
  out[blockIdx.x * 256 + p * W + w] = s;
  if (w == 0 && p == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int W>
void go(const Tab &h, const double *g, double *out, long long *cyc, int blocks, int iters, bool uniform_chunk) {
  dim3 block(W, 256 / W);
  run<MODE, W><<<blocks, block>>>(h, g, out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  run<MODE, W><<<blocks, block>>>(h, g, out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double rows = (double)blocks * 256 * M * iters;
  printf("mode %d W=%d: %.3f ms, %lld cycles/block for %d iters (%.1f cycles per chunk pass), %.2f row-updates/cycle/SM, err=%s\n", MODE, W, ms,
         c, iters, (double)c / iters, rows / 148.0 / (double)c, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  static Tab h;
  for (int i = 0; i < 5 * L; ++i) h.v[i] = 0.3 + 1e-4 * (i % 97);
  double *g, *out; long long *cyc;
  cudaMalloc(&g, sizeof(Tab)); cudaMalloc(&out, 296 * 256 * sizeof(double)); cudaMalloc(&cyc, 8);
  cudaMemcpy(g, h.v, sizeof(Tab), cudaMemcpyHostToDevice);
  const int iters = 2000;
  go<0, 16>(h, g, out, cyc, 296, iters, false);
  go<0, 32>(h, g, out, cyc, 296, iters, false);
  go<4, 16>(h, g, out, cyc, 296, iters, false);
  go<4, 32>(h, g, out, cyc, 296, iters, false);
  go<2, 16>(h, g, out, cyc, 296, iters, false);
  go<2, 32>(h, g, out, cyc, 296, iters, false);
  go<6, 16>(h, g, out, cyc, 296, iters, false);
  go<6, 32>(h, g, out, cyc, 296, iters, false);
  go<1, 32>(h, g, out, cyc, 296, iters, false);
  return 0;
}
