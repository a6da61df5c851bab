// run_steps.cu - many time steps per call with the field resident in HBM, and observation on the device.
//
// The reference's demos step and observe from Python: one run_adi_steps call per step, the whole field copied
// back, then indexed on the host (demos/steelonfoam.py:132-143) or reduced to the z-min surface temperature
// (heatsim2/surface_temperature.py:4-36, demos/ktest.py).  Here the loop is native: hs2_run_steps ping-pongs the
// field between two device buffers, records probe cells and the surface estimate with ONE small kernel per
// recorded instant, and - where the grid is small enough for launch latency to matter - replays the steps as a
// CUDA graph, so a device-resident user pays neither Python nor launch overhead per step.
#include "hs2_common.cuh"

namespace {

// probes: out[row][q] = T[cell[q]];  surface: out[row][j][i] = estimate from planes 0 and 1.
// `row` comes from a device counter so that the same launch can be replayed from a CUDA graph.
// Arithmetic of the surface estimate in the reference's order (surface_temperature.py:30-36), no contraction:
//   a = (T1 - T0) / (2 dz^2); c = T0 - 0.25 a dz^2; lin = T0 - (T1 - T0) / 2; (lin + c) / 2
__global__ void observe_kernel(const double *__restrict__ T, const int64_t *__restrict__ cells, int n_probes,
                               double *__restrict__ probe_rec, double *__restrict__ surf_rec, int64_t plane,
                               double two_dz2, double dz2, const unsigned long long *__restrict__ counter) {
  const unsigned long long row = counter ? *counter : 0ull;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (probe_rec && tid < n_probes) probe_rec[row * (unsigned long long)n_probes + tid] = T[cells[tid]];
  if (surf_rec && tid < plane) {
    const double t0 = T[tid], t1 = T[plane + tid];
    const double dT = __dsub_rn(t1, t0);
    const double a = __ddiv_rn(dT, two_dz2);
    const double c = __dsub_rn(t0, __dmul_rn(__dmul_rn(0.25, a), dz2));
    const double lin = __dsub_rn(t0, __ddiv_rn(dT, 2.0));
    surf_rec[row * (unsigned long long)plane + tid] = __ddiv_rn(__dadd_rn(lin, c), 2.0);
  }
}

__global__ void bump_kernel(unsigned long long *counter) { *counter += 1ull; }

struct Observer {
  const int64_t *cells;
  int n_probes;
  double *probe_rec, *surf_rec;
  int64_t plane;
  double dz;
  unsigned long long *counter;
  bool active() const { return (cells && n_probes > 0 && probe_rec) || surf_rec; }
  int launch(const double *T, cudaStream_t st) const {
    const int64_t n = surf_rec ? (plane > n_probes ? plane : n_probes) : n_probes;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    const double dz2 = dz * dz;
    observe_kernel<<<blocks, threads, 0, st>>>(T, cells, (cells && probe_rec) ? n_probes : 0, probe_rec, surf_rec, plane,
                                               2.0 * dz2, dz2, counter);
    HS2_CUDA_CHECK(cudaGetLastError());
    if (counter) {
      bump_kernel<<<1, 1, 0, st>>>(counter);
      HS2_CUDA_CHECK(cudaGetLastError());
    }
    return HS2_OK;
  }
};

// steps [first, first + count) of the run: step n reads buf[n & 1] and writes buf[(n + 1) & 1]
int run_range(hs2_plan *plan, double *const buf[2], double *work, int64_t first, int64_t count, int every,
              const Observer &obs, cudaStream_t st) {
  for (int64_t n = first; n < first + count; ++n) {
    int rc = hs2_step(plan, buf[n & 1], buf[(n + 1) & 1], work, nullptr, nullptr, nullptr, st);
    if (rc) return rc;
    if (obs.active() && (n + 1) % every == 0) {
      rc = obs.launch(buf[(n + 1) & 1], st);
      if (rc) return rc;
    }
  }
  return HS2_OK;
}

// the replay graph of the last hs2_run_steps configuration, owned by the plan (released by hs2_plan_destroy)
struct GraphKey {
  double *a, *b, *work;
  int64_t unit;
  int every;
  const int64_t *cells;
  int n_probes;
  double *probe_rec, *surf_rec;
  double dz;
  unsigned long long *counter;
  bool operator==(const GraphKey &o) const {
    return a == o.a && b == o.b && work == o.work && unit == o.unit && every == o.every && cells == o.cells &&
           n_probes == o.n_probes && probe_rec == o.probe_rec && surf_rec == o.surf_rec && dz == o.dz && counter == o.counter;
  }
};
struct GraphCache {
  GraphKey key;
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};

}  // namespace

void hs2_graph_cache_free(hs2_plan *plan) {
  GraphCache *c = static_cast<GraphCache *>(plan->graph_cache);
  if (!c) return;
  if (c->exec) cudaGraphExecDestroy(c->exec);
  if (c->graph) cudaGraphDestroy(c->graph);
  delete c;
  plan->graph_cache = nullptr;
}

extern "C" {

int hs2_observe(hs2_plan *plan, const double *d_T, const int64_t *d_probe_cells, int n_probes, double *d_probe_out,
                double *d_surface_out, double dz, void *stream) {
  HS2_REQUIRE(plan && d_T, "hs2_observe: NULL argument");
  HS2_REQUIRE(n_probes >= 0 && (n_probes == 0 || (d_probe_cells && d_probe_out)), "hs2_observe: probe arrays missing");
  HS2_REQUIRE(!d_surface_out || (plan->d.nz >= 2 && dz != 0.0), "hs2_observe: the surface estimate needs two planes and dz != 0");
  Observer obs{d_probe_cells, n_probes, d_probe_out, d_surface_out, plan->d.ny * plan->d.nx, dz, nullptr};
  if (!obs.active()) return HS2_OK;
  return obs.launch(d_T, (cudaStream_t)stream);
}

int hs2_run_steps(hs2_plan *plan, double *d_T_a, double *d_T_b, double *d_work, int64_t first_step, int64_t nsteps, int every,
                  const int64_t *d_probe_cells, int n_probes, double *d_probe_rec, double *d_surface_rec, double dz,
                  uint64_t *d_counter, int use_graph, void *stream) {
  HS2_REQUIRE(plan && d_T_a && d_T_b && d_work, "hs2_run_steps: NULL argument");
  HS2_REQUIRE(d_T_a != d_T_b && d_T_a != d_work && d_T_b != d_work, "hs2_run_steps: the three field buffers must be distinct");
  HS2_REQUIRE(first_step >= 0 && nsteps >= 0 && every >= 1, "hs2_run_steps: first_step >= 0, nsteps >= 0 and every >= 1 required");
  HS2_REQUIRE(plan->d.z_chunks_global == 0, "hs2_run_steps: slab plans are stepped by the multi-GPU driver");
  HS2_REQUIRE(n_probes >= 0 && (n_probes == 0 || !d_probe_rec || d_probe_cells), "hs2_run_steps: probe cells missing");
  HS2_REQUIRE(!d_surface_rec || (plan->d.nz >= 2 && dz != 0.0), "hs2_run_steps: the surface estimate needs two planes and dz != 0");
  Observer obs{d_probe_cells, n_probes, d_probe_rec, d_surface_rec, plan->d.ny * plan->d.nx, dz,
               reinterpret_cast<unsigned long long *>(d_counter)};
  HS2_REQUIRE(!obs.active() || d_counter, "hs2_run_steps: recording needs d_counter (a 64-bit device word holding the first free record row)");
  cudaStream_t st = (cudaStream_t)stream;
  double *const buf[2] = {d_T_a, d_T_b};
  // replay unit: a whole number of recording periods that also brings the field back to the buffer it started in
  const int64_t period = obs.active() ? every : 1;
  const int64_t unit = (period & 1) ? 2 * period : period;
  const int64_t end = first_step + nsteps;
  const int64_t aligned = (first_step + unit - 1) / unit * unit;      // first step number the unit can start at
  const int64_t reps = aligned < end ? (end - aligned) / unit : 0;
  if (!use_graph || reps < 4 || unit > 256) return run_range(plan, buf, d_work, first_step, nsteps, every, obs, st);

  const GraphKey key{d_T_a, d_T_b, d_work, unit, every, obs.cells, obs.n_probes, obs.probe_rec, obs.surf_rec, dz, obs.counter};
  GraphCache *cache = static_cast<GraphCache *>(plan->graph_cache);
  if (cache && !(cache->key == key)) {
    // another configuration: the old graph may still be executing on the caller's stream
    HS2_CUDA_CHECK(cudaStreamSynchronize(st));
    hs2_graph_cache_free(plan);
    cache = nullptr;
  }
  if (!cache) {
    // capture one unit on a private stream (the legacy default stream cannot be captured); it is replayed on the caller's
    cudaStream_t cs = nullptr;
    HS2_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int rc = HS2_OK;
    cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = run_range(plan, buf, d_work, 0, unit, every, obs, cs);
      e = cudaStreamEndCapture(cs, &graph);
      if (rc == HS2_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    }
    cudaStreamDestroy(cs);
    if (rc != HS2_OK || e != cudaSuccess || !exec) {
      // not capturable on this path: plain launches (nothing of the unit has run; a genuine error shows up again there)
      (void)cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      return run_range(plan, buf, d_work, first_step, nsteps, every, obs, st);
    }
    cache = new GraphCache{key, graph, exec};
    plan->graph_cache = cache;
  }
  {
    const int rc = run_range(plan, buf, d_work, first_step, aligned - first_step, every, obs, st);
    if (rc) return rc;
  }
  for (int64_t r = 0; r < reps; ++r) {
    cudaError_t e = cudaGraphLaunch(cache->exec, st);
    if (e != cudaSuccess) {
      hs2_set_error("hs2_run_steps: cudaGraphLaunch failed: %s", cudaGetErrorString(e));
      return HS2_E_CUDA;
    }
  }
  return run_range(plan, buf, d_work, aligned + reps * unit, end - (aligned + reps * unit), every, obs, st);
}

}  // extern "C"
