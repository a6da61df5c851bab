// Shared declarations of the hs2_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hs2_b200.h"

// the most common chunk table of an axis, handed to the kernels by value (chunk_core.cuh)
struct UTab {
  double v[HS2_T_PLANES][32];   // [plane][row in chunk], rows >= chunk unused
};

struct hs2_owned;                     // plan_build.cu: device buffers and host tables of a built plan
void hs2_owned_free(hs2_owned *o);

struct hs2_plan {
  hs2_plan_desc d;
  hs2_owned *owned;    // non-NULL for plans made by hs2_plan_build
  int64_t n;           // nz*ny*nx
  int sm_count;
  int max_smem_optin;
  int last_kernel[3];  // HS2_K_* of the last sweep per axis (hs2_plan_last_kernel)
  UTab utab[3];        // copy of axis[a].h_utab (passed to the kernels by value)
  bool has_utab[3];
  void *graph_cache;   // run_steps.cu: replay graph of the last hs2_run_steps configuration
};
void hs2_graph_cache_free(hs2_plan *plan);

void hs2_set_error(const char *fmt, ...);

// Optional phase timing (build with -DHS2_PHASE_TIMING): thread 0 of every
// block adds the cycles spent between marks to g_hs2_phase[slot].
#ifdef HS2_PHASE_TIMING
static __device__ unsigned long long g_hs2_phase[16];   // one copy per translation unit
#define HS2_MARK_DECL long long hs2_t0__ = clock64()
#define HS2_MARK(slot)                                                          \
  do {                                                                          \
    if (threadIdx.x == 0 && threadIdx.y == 0) {                                 \
      const long long t__ = clock64();                                          \
      atomicAdd(&g_hs2_phase[slot], (unsigned long long)(t__ - hs2_t0__));      \
      hs2_t0__ = t__;                                                           \
    }                                                                           \
  } while (0)
#else
#define HS2_MARK_DECL
#define HS2_MARK(slot)
#endif

#define HS2_CUDA_CHECK(call)                                                   \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      hs2_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),   \
                    __FILE__, __LINE__);                                       \
      return HS2_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

#define HS2_REQUIRE(cond, ...)                                                 \
  do {                                                                         \
    if (!(cond)) {                                                             \
      hs2_set_error(__VA_ARGS__);                                              \
      return HS2_E_INVALID;                                                    \
    }                                                                          \
  } while (0)

// kernels_v1.cu - global-memory reference-quality path (any size)
int hs2_v1_sweep_x(hs2_plan *p, const double *T, double *W, const hs2_source *src,
                   const double *halo_lo, const double *halo_hi, cudaStream_t st);
int hs2_v1_sweep_y(hs2_plan *p, double *W, cudaStream_t st);
int hs2_v1_sweep_z(hs2_plan *p, const double *T, double *Tout, double *W, cudaStream_t st);

// kernels_strided.cu - register/shared-memory tile kernels of the strided axes
bool hs2_tile_supported(const hs2_plan *p, int axis);
int hs2_tile_sweep_y(hs2_plan *p, double *W, cudaStream_t st);
int hs2_tile_sweep_z(hs2_plan *p, const double *T, double *Tout, double *W, cudaStream_t st);
int hs2_zfused(hs2_plan *pl, double *data, const double *Tin, double *Tout, double *Yall, int n_peers, const uint64_t *peer_y,
               double timeout_s, int *status, int *tile_lines, cudaStream_t st);
int hs2_fill_empty(void *d_ptr, int64_t n_doubles, cudaStream_t st);
int hs2_zdist(hs2_plan *pl, int phase, double *data, const double *Tin, double *Tout, double *Y, int64_t line0,
              int64_t n_lines, int n_peers, const uint64_t *peer_y, bool full_cols, cudaStream_t st);
// kernels_xt.cu - x sweep on TMA-staged patches (default where it applies); *done = false: fall through
bool hs2_tile_xt_supported(const hs2_plan *p);
int hs2_tile_sweep_xt(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                      const double *halo_hi, int part, cudaStream_t st, bool *done);
// kernels_xw.cu - x sweep, one warp per line (nx = 512, source-free steps); *done = false: fall through
bool hs2_tile_xw_supported(const hs2_plan *p);
int hs2_tile_sweep_xw(hs2_plan *p, const double *T, double *W, const double *halo_lo, const double *halo_hi, int part,
                      cudaStream_t st, bool *done);
// kernels_xf.cu - x sweep with the explicit x-term folded into the solve
bool hs2_tile_xf_supported(const hs2_plan *p);
int hs2_tile_sweep_xf(hs2_plan *p, const double *T, double *W, const hs2_source *src, const double *halo_lo,
                      const double *halo_hi, int part, cudaStream_t st);
