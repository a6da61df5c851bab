// peer.cu - NVLink peer-memory plumbing of the multi-GPU z-slab path: device
// buffers that other ranks of the node can map (CUDA IPC), and flags in those
// buffers for stream-ordered signalling between GPUs.  The data itself is
// written into the neighbours' memory by the kernels that produce it
// (z_forward pushes the chunk interface values, kernels_strided.cu) or by
// cudaMemcpyAsync over NVLink (halo planes); nothing here touches the host.
// There is no reference counterpart: heatsim2 is single-process.
#include "hs2_common.cuh"

namespace {

constexpr int MAX_FLAGS = 16;
struct FlagList {
  int n;
  unsigned long long *p[MAX_FLAGS];
};

// The grid that produced the data has completed before this one starts (same
// stream), which makes its peer writes visible system-wide; the release store
// publishes the flag after them.
__global__ void flag_signal_kernel(FlagList fl, unsigned long long value) {
  __threadfence_system();
  const int i = threadIdx.x;
  if (i < fl.n) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(fl.p[i]), "l"(value) : "memory");
}

// Spin (bounded) until every flag has reached `value`.  Flags live in this
// GPU's own memory; the peers write them over NVLink.  On timeout the status
// word is set and the kernel returns, so a lost peer can never hang the GPU.
__global__ void flag_wait_kernel(FlagList fl, unsigned long long value, long long max_cycles, int *status) {
  const int i = threadIdx.x;
  // once a wait has timed out the run is lost: later waits return at once, so
  // the process reaches its status check quickly instead of timing out per step
  if (i < fl.n && *reinterpret_cast<volatile int *>(status) == 0) {
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(fl.p[i]) : "memory");
      if (v >= value) break;
      if (clock64() - t0 > max_cycles) {
        atomicExch(status, 1);
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

int make_list(const uint64_t *flag_ptrs, int n, FlagList *fl) {
  HS2_REQUIRE(n >= 0 && n <= MAX_FLAGS && (n == 0 || flag_ptrs), "flag list: %d entries (max %d)", n, MAX_FLAGS);
  fl->n = n;
  for (int i = 0; i < n; ++i) fl->p[i] = reinterpret_cast<unsigned long long *>(flag_ptrs[i]);
  return HS2_OK;
}

}  // namespace

extern "C" {

int hs2_peer_alloc(int64_t bytes, void **d_ptr, void *handle64) {
  HS2_REQUIRE(bytes > 0 && d_ptr && handle64, "hs2_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void *p = nullptr;
  HS2_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  HS2_CUDA_CHECK(cudaMemset(p, 0, (size_t)bytes));
  HS2_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    hs2_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return HS2_E_CUDA;
  }
  memcpy(handle64, &h, 64);
  *d_ptr = p;
  return HS2_OK;
}

int hs2_peer_open(const void *handle64, void **d_ptr) {
  HS2_REQUIRE(handle64 && d_ptr, "hs2_peer_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  HS2_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HS2_OK;
}

int hs2_peer_close(void *d_ptr) {
  if (d_ptr) HS2_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
  return HS2_OK;
}

int hs2_peer_free(void *d_ptr) {
  if (d_ptr) HS2_CUDA_CHECK(cudaFree(d_ptr));
  return HS2_OK;
}

int hs2_flag_signal(const uint64_t *flag_ptrs, int n, uint64_t value, void *stream) {
  FlagList fl;
  int rc = make_list(flag_ptrs, n, &fl);
  if (rc) return rc;
  if (n == 0) return HS2_OK;
  flag_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fl, (unsigned long long)value);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

int hs2_flag_wait(const uint64_t *flag_ptrs, int n, uint64_t value, double timeout_s, int *d_status, void *stream) {
  FlagList fl;
  int rc = make_list(flag_ptrs, n, &fl);
  if (rc) return rc;
  HS2_REQUIRE(d_status, "hs2_flag_wait: NULL status word");
  if (n == 0) return HS2_OK;
  const long long cycles = (long long)(timeout_s * 1.9e9);
  flag_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(fl, (unsigned long long)value, cycles, d_status);
  HS2_CUDA_CHECK(cudaGetLastError());
  return HS2_OK;
}

int hs2_copy_async(void *d_dst, const void *d_src, int64_t bytes, void *stream) {
  HS2_REQUIRE(d_dst && d_src && bytes >= 0, "hs2_copy_async: bad argument");
  HS2_CUDA_CHECK(cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return HS2_OK;
}

}  // extern "C"
