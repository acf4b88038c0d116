// C-ABI: tode_solve_fused -- whole solve of a built-in analytic field in one launch.
#include <climits>

#include "api_common.cuh"
#include "erk_fused.cuh"

namespace tode {

__global__ void summary_init_kernel(int* summary, int n_launches) {
  summary[0] = 0;
  summary[1] = INT_MAX;
  summary[2] = 0;
  summary[3] = n_launches;  // kernels this tode_solve_fused call launches (incl. this one)
  summary[4] = 0;  // [4..5]: 64-bit work queue of the persistent f2 kernel
  summary[5] = 0;
  summary[6] = 0;
  summary[7] = 0;
}

// Multi-GPU epilogue (tode_solution.peer_global): one thread publishes this launch's iteration count and
// its first failing iteration to every replica with system-scope atomics (replicas are zeroed before the
// launches of a step, so both travel as maxima: [0] = max iterations, [2] = max (INT_MAX - first failure)).
// A replay launch (iter_cap > 0) publishes nothing: the caller already knows the batch-wide cap.
struct PeerGlobals {
  int n;
  int* g[TODE_MAX_PEERS];
};
__global__ void peer_epilogue_kernel(const int* summary, PeerGlobals pg, int after_replay) {
  if (after_replay) return;
  const int iters = summary[0], first_fail = summary[1];
  for (int p = 0; p < pg.n; ++p) {
    atomicMax_system(pg.g[p] + 0, iters);
    if (first_fail != INT_MAX) atomicMax_system(pg.g[p] + 2, INT_MAX - first_fail);
  }
  __threadfence_system();
}

// fused_impl.cuh, instantiated in fused_f32f32.cu / fused_f64f64.cu / fused_f32f64.cu / fused_f64f32.cu
template <typename D, typename T>
int launch_fused(int field, const double* fp, const tode_tableau* tab, const tode_controller* ctrl,
                 const tode_problem* prob, const tode_solution* sol, int64_t iter_cap, cudaStream_t stream);

}  // namespace tode

using namespace tode;

extern "C" int tode_solve_fused(int field, const double* field_params, const tode_tableau* tab,
                                const tode_controller* ctrl, const tode_problem* prob,
                                const tode_solution* sol, int64_t iter_cap, void* stream) {
  if (!field_params || !tab || !ctrl || !prob || !sol) return TODE_EINVAL;
  if (!prob->y0 || !prob->t_start || !prob->t_end || (prob->T > 0 && !prob->t_eval)) return TODE_EINVAL;
  if (!sol->ys || !sol->n_steps || !sol->n_accepted || !sol->n_initialized || !sol->status || !sol->summary)
    return TODE_EINVAL;
#define CALL(D, T) launch_fused<D, T>(field, field_params, tab, ctrl, prob, sol, iter_cap, static_cast<cudaStream_t>(stream))
  const int rc = [&]() -> int { TODE_DISPATCH_DT(prob->data_dtype, prob->time_dtype, CALL); }();
#undef CALL
  if (rc != 0 || sol->n_peers <= 0 || prob->B == 0) return rc;
  PeerGlobals pg{};
  pg.n = sol->n_peers;
  for (int p = 0; p < sol->n_peers; ++p) pg.g[p] = sol->peer_global[p];
  peer_epilogue_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(sol->summary, pg, iter_cap > 0 ? 1 : 0);
  return launch_status();
}
