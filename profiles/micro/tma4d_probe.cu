// tma4d_probe.cu - round-2 probe for the TMA-staged x-sweep (kernels_xt.cu).
// Questions answered on a B200 before the kernel was written:
//  1. does cuTensorMapEncodeTiled accept a rank-4 float64 map whose strides are NOT
//     ascending: dims (16, ny, nz, nx/16), strides (nx*8, ny*nx*8, 128)?
//  2. where does SWIZZLE_128B put element (c, row, plane, seg) of a box (16, 6, 2, S)?
//     expected: line = seg*12 + plane*6 + row; byte = line*128 + (((c>>1) ^ (line&7))<<4) + (c&1)*8
//  3. are out-of-range rows/planes (coordinate -1 / >= dim) zero-filled?
//  4. does a swizzled store box (16, 4, 1, S) land where expected?
//  5. how fast can 2 CTAs/SM stream all (2 planes x 4 rows) patches of a 512^3 field through
//     the three load boxes (20 rows per 8 lines) - the load-side ceiling of the kernel.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma4d_probe tma4d_probe.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static PFN_cuTensorMapEncodeTiled get_encode() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// one block: load box at (j0-1, k0), dump the raw shared memory
__global__ void probe_load(const __grid_constant__ CUtensorMap tm, double *dump, int n_doubles, int j0m1, int k0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + n_doubles * 8);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, n_doubles * 8);
    tma_load_4d(sm, &tm, bar, 0, j0m1, k0, 0);
  }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < n_doubles; i += blockDim.x) dump[i] = reinterpret_cast<double *>(sm)[i];
}

// one block: fill smem (swizzled, logical value = 1000*row + 100000*seg + c) and store box (16,4,1,S) at (j0,k0)
__global__ void probe_store(const __grid_constant__ CUtensorMap tm, int S, int j0, int k0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  for (int e = threadIdx.x; e < S * 4 * 16; e += blockDim.x) {
    const int c = e & 15, row = (e >> 4) & 3, seg = e >> 6;
    const int line = seg * 4 + row;
    const int byte = line * 128 + ((((c >> 1) ^ (line & 7))) << 4) + (c & 1) * 8;
    *reinterpret_cast<double *>(sm + byte) = 1000.0 * row + 100000.0 * seg + c;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_4d(&tm, sm, 0, j0, k0, 0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// streaming ceiling: persistent CTAs, each patch = boxes C (16,6,2,S), ZL (16,4,1,S), ZH (16,4,1,S); optionally the
// consumers read everything back (5 LDS.128 per 2 cells) and write d = sum to W through a swizzled store
template <int MODE>
__global__ void __launch_bounds__(256, 2)
stream_patches(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmZ,
               const __grid_constant__ CUtensorMap tmO, int S, int ny, int nz, int n_tiles, int tiles_y, double *sink) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int szC = S * 12 * 128, szZ = S * 4 * 128;
  unsigned char *C = sm, *ZL = sm + szC, *ZH = ZL + szZ;
  uint64_t *bar = reinterpret_cast<uint64_t *>(ZH + szZ);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int rw = lane & 3, pc = (lane >> 2) & 1, pz = (lane >> 3) & 1, pp = lane >> 4;
  const int p = 4 * w + 2 * pp + pc;
  uint32_t parity = 0;
  double acc = 0.0;
  int t = blockIdx.x;
  if (t < n_tiles && threadIdx.x == 0) {
    const int k0 = 2 * (t / tiles_y), j0 = 4 * (t % tiles_y);
    mbar_expect_tx(bar, szC);
    tma_load_4d(C, &tmC, bar, 0, j0 - 1, k0, 0);
    mbar_expect_tx(bar + 1, 2 * szZ);
    tma_load_4d(ZL, &tmZ, bar + 1, 0, j0, k0 - 1, 0);
    tma_load_4d(ZH, &tmZ, bar + 1, 0, j0, k0 + 2, 0);
  }
  for (; t < n_tiles; t += gridDim.x) {
    const int k0 = 2 * (t / tiles_y), j0 = 4 * (t % tiles_y);
    mbar_wait(bar, parity);
    mbar_wait(bar + 1, parity);
    parity ^= 1;
    double v[16];
    if (MODE >= 1 && p < S) {
      // 5 stencil streams of this thread's chunk (seg p), chunk layout, swizzled
      const int lc = p * 12 + pz * 6 + rw + 1;           // centre line
      const int lzl = p * 4 + rw, lzh = p * 4 + rw;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double2 c = *reinterpret_cast<const double2 *>(C + lc * 128 + ((u ^ (lc & 7)) << 4));
        const double2 ym = *reinterpret_cast<const double2 *>(C + (lc - 1) * 128 + ((u ^ ((lc - 1) & 7)) << 4));
        const double2 yp = *reinterpret_cast<const double2 *>(C + (lc + 1) * 128 + ((u ^ ((lc + 1) & 7)) << 4));
        double2 zm, zp;
        if (pz == 0) {
          zm = *reinterpret_cast<const double2 *>(ZL + lzl * 128 + ((u ^ (lzl & 7)) << 4));
          zp = *reinterpret_cast<const double2 *>(C + (lc + 6) * 128 + ((u ^ ((lc + 6) & 7)) << 4));
        } else {
          zm = *reinterpret_cast<const double2 *>(C + (lc - 6) * 128 + ((u ^ ((lc - 6) & 7)) << 4));
          zp = *reinterpret_cast<const double2 *>(ZH + lzh * 128 + ((u ^ (lzh & 7)) << 4));
        }
        v[2 * u] = c.x + 0.25 * (ym.x + yp.x + zm.x + zp.x);
        v[2 * u + 1] = c.y + 0.25 * (ym.y + yp.y + zm.y + zp.y);
      }
    }
    __syncthreads();   // stage consumed
    if (threadIdx.x == 0 && t + (int)gridDim.x < n_tiles) {
      const int tn = t + gridDim.x;
      const int kn = 2 * (tn / tiles_y), jn = 4 * (tn % tiles_y);
      mbar_expect_tx(bar, szC);
      tma_load_4d(C, &tmC, bar, 0, jn - 1, kn, 0);
    }
    if (MODE >= 2) {
      // result into the ZL / ZH buffers (plane 0 / plane 1), swizzled, then TMA store
      if (p < S) {
        unsigned char *O = pz == 0 ? ZL : ZH;
        const int lo = p * 4 + rw;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<double2 *>(O + lo * 128 + ((u ^ (lo & 7)) << 4)) = make_double2(v[2 * u], v[2 * u + 1]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_4d(&tmO, ZL, 0, j0, k0, 0);
        tma_store_4d(&tmO, ZH, 0, j0, k0 + 1, 0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += v[u];
    }
    if (threadIdx.x == 0 && t + (int)gridDim.x < n_tiles) {
      const int tn = t + gridDim.x;
      const int kn = 2 * (tn / tiles_y), jn = 4 * (tn % tiles_y);
      mbar_expect_tx(bar + 1, 2 * szZ);
      tma_load_4d(ZL, &tmZ, bar + 1, 0, jn, kn - 1, 0);
      tma_load_4d(ZH, &tmZ, bar + 1, 0, jn, kn + 2, 0);
    }
  }
  if (MODE == 1 && acc == 12345.678) sink[0] = acc;
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static bool encode4(PFN_cuTensorMapEncodeTiled enc, CUtensorMap *m, void *base, int nx, int ny, int nz, int rows, int planes,
                    CUtensorMapSwizzle sw) {
  const int S = nx / 16;
  cuuint64_t dims[4] = {16, (cuuint64_t)ny, (cuuint64_t)nz, (cuuint64_t)S};
  cuuint64_t strides[3] = {(cuuint64_t)nx * 8, (cuuint64_t)ny * nx * 8, 128};
  cuuint32_t box[4] = {16, (cuuint32_t)rows, (cuuint32_t)planes, (cuuint32_t)S};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) printf("encode (rows %d planes %d) failed: CUresult %d\n", rows, planes, (int)rc);
  return rc == CUDA_SUCCESS;
}

int main() {
  auto enc = get_encode();
  {
    // ---- layout probe on a small field
    const int nx = 64, ny = 12, nz = 6, S = nx / 16;
    const size_t n = (size_t)nx * ny * nz;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i + 1.0;
    double *dT, *dW, *dump;
    CK(cudaMalloc(&dT, n * 8));
    CK(cudaMalloc(&dW, n * 8));
    CK(cudaMemset(dW, 0, n * 8));
    CK(cudaMemcpy(dT, h.data(), n * 8, cudaMemcpyHostToDevice));
    CUtensorMap tmC, tmO;
    const bool okC = encode4(enc, &tmC, dT, nx, ny, nz, 6, 2, CU_TENSOR_MAP_SWIZZLE_128B);
    const bool okO = encode4(enc, &tmO, dW, nx, ny, nz, 4, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    printf("Q1 encode non-ascending strides: C %s, O %s\n", okC ? "ok" : "FAILED", okO ? "ok" : "FAILED");
    if (!okC || !okO) return 2;
    const int nd = S * 12 * 16;
    CK(cudaMalloc(&dump, nd * 8));
    for (int trial = 0; trial < 2; ++trial) {
      const int j0m1 = trial == 0 ? 3 : -1, k0 = trial == 0 ? 2 : 5;   // trial 1: row -1 and plane 6 out of range
      CK(cudaFuncSetAttribute(probe_load, cudaFuncAttributeMaxDynamicSharedMemorySize, nd * 8 + 64));
      probe_load<<<1, 128, nd * 8 + 64>>>(tmC, dump, nd, j0m1, k0);
      CK(cudaDeviceSynchronize());
      std::vector<double> d(nd);
      CK(cudaMemcpy(d.data(), dump, nd * 8, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int seg = 0; seg < S; ++seg)
        for (int pl = 0; pl < 2; ++pl)
          for (int row = 0; row < 6; ++row)
            for (int c = 0; c < 16; ++c) {
              const int line = seg * 12 + pl * 6 + row;
              const int byte = line * 128 + ((((c >> 1) ^ (line & 7))) << 4) + (c & 1) * 8;
              const int j = j0m1 + row, k = k0 + pl, i = seg * 16 + c;
              const double want = (j < 0 || j >= ny || k < 0 || k >= nz) ? 0.0 : h[((size_t)k * ny + j) * nx + i];
              if (d[byte / 8] != want) {
                if (bad < 5) printf("  mismatch seg %d pl %d row %d c %d: got %g want %g\n", seg, pl, row, c, d[byte / 8], want);
                ++bad;
              }
            }
      printf("Q%d load layout [seg][plane][row][16] + 128B swizzle%s: %s (%d mismatches)\n", trial == 0 ? 2 : 3,
             trial ? " with OOB zero fill" : "", bad ? "WRONG" : "as expected", bad);
    }
    CK(cudaFuncSetAttribute(probe_store, cudaFuncAttributeMaxDynamicSharedMemorySize, S * 4 * 128));
    probe_store<<<1, 128, S * 4 * 128>>>(tmO, S, 10, 3);   // rows 10, 11 in range, 12, 13 clipped
    CK(cudaDeviceSynchronize());
    std::vector<double> w(n);
    CK(cudaMemcpy(w.data(), dW, n * 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
          double want = 0.0;
          if (k == 3 && j >= 10 && j < 14) want = 1000.0 * (j - 10) + 100000.0 * (i / 16) + (i % 16);
          if (w[((size_t)k * ny + j) * nx + i] != want) ++bad;
        }
    printf("Q4 swizzled store box (16,4,1,S) with clipped rows: %s (%d mismatches)\n", bad ? "WRONG" : "as expected", bad);
  }
  {
    // ---- streaming ceiling at 512^3
    const int nx = 512, ny = 512, nz = 512, S = nx / 16;
    const size_t n = (size_t)nx * ny * nz;
    double *dT, *dW, *sink;
    CK(cudaMalloc(&dT, n * 8));
    CK(cudaMalloc(&dW, n * 8));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(dT, 0, n * 8));
    CUtensorMap tmC, tmZ, tmO;
    encode4(enc, &tmC, dT, nx, ny, nz, 6, 2, CU_TENSOR_MAP_SWIZZLE_128B);
    encode4(enc, &tmZ, dT, nx, ny, nz, 4, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    encode4(enc, &tmO, dW, nx, ny, nz, 4, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    const int tiles_y = ny / 4, n_tiles = tiles_y * (nz / 2);
    const size_t smem = (size_t)S * 20 * 128 + 64;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 3; ++mode) {
      auto kern = mode == 0 ? stream_patches<0> : (mode == 1 ? stream_patches<1> : stream_patches<2>);
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int per_sm = 1; per_sm <= 2; ++per_sm) {
        const int grid = prop.multiProcessorCount * per_sm;
        for (int it = 0; it < 3; ++it) kern<<<grid, 256, smem>>>(tmC, tmZ, tmO, S, ny, nz, n_tiles, tiles_y, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 10;
        for (int it = 0; it < reps; ++it) kern<<<grid, 256, smem>>>(tmC, tmZ, tmO, S, ny, nz, n_tiles, tiles_y, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= reps;
        printf("Q5 mode %d (%s) %d CTA/SM: %.3f ms per 512^3 sweep -> %.0f GB/s of field read%s\n", mode,
               mode == 0 ? "TMA loads only" : (mode == 1 ? "loads + 5 LDS.128 streams" : "loads + LDS + swizzled STS + TMA store"),
               per_sm, ms, n * 8 / (ms * 1e-3) / 1e9, mode == 2 ? " (+ same written)" : "");
      }
    }
  }
  return 0;
}
