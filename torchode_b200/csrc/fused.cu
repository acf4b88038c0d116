// C-ABI: tode_solve_fused -- whole solve of a built-in analytic field in one launch.
#include <climits>

#include "api_common.cuh"
#include "erk_fused.cuh"

namespace tode {

__global__ void summary_init_kernel(int* summary) {
  summary[0] = 0;
  summary[1] = INT_MAX;
  summary[2] = 0;
  summary[3] = 0;
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
  TODE_DISPATCH_DT(prob->data_dtype, prob->time_dtype, CALL);
#undef CALL
}
