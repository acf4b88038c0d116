// C-ABI: tode_solve_fused -- whole solve of a built-in analytic field in one launch.
#include <climits>
#include <cstdlib>

#include "api_common.cuh"
#include "erk_fused.cuh"

namespace tode {

__global__ void summary_init_kernel(int* summary) {
  summary[0] = 0;
  summary[1] = INT_MAX;
  summary[2] = 0;
  summary[3] = 0;
}

// resident 128-thread CTAs per SM the fused kernel is compiled for; measured on C2 / C3
// (profiles/r01_fused_occupancy.txt): fp64 state prefers 6 (80 registers, a few spilled bytes,
// -5 % time), fp32 state 4
#ifndef TODE_FUSED_MINB
#define TODE_FUSED_MINB (sizeof(D) == 8 ? 6 : 4)
#endif

template <typename D, typename T, int F, int FIELD>
static int launch_fused_f(const FusedArgs<D, T>& a, cudaStream_t stream) {
  constexpr int kThreads = 128;
  summary_init_kernel<<<1, 1, 0, stream>>>(a.summary);
  const unsigned grid = (unsigned)((a.B + kThreads - 1) / kThreads);
#ifdef TODE_FUSED_TUNE
  // tuning build only: pick the occupancy variant at run time
  const char* env = std::getenv("TODE_FUSED_MINB");
  const int minb = env ? std::atoi(env) : TODE_FUSED_MINB;
  switch (minb) {
    case 3: solve_fused_kernel<D, T, F, FIELD, 3><<<grid, kThreads, 0, stream>>>(a); break;
    case 5: solve_fused_kernel<D, T, F, FIELD, 5><<<grid, kThreads, 0, stream>>>(a); break;
    case 6: solve_fused_kernel<D, T, F, FIELD, 6><<<grid, kThreads, 0, stream>>>(a); break;
    case 8: solve_fused_kernel<D, T, F, FIELD, 8><<<grid, kThreads, 0, stream>>>(a); break;
    default: solve_fused_kernel<D, T, F, FIELD, 4><<<grid, kThreads, 0, stream>>>(a); break;
  }
#else
  solve_fused_kernel<D, T, F, FIELD, TODE_FUSED_MINB><<<grid, kThreads, 0, stream>>>(a);
#endif
  return launch_status();
}

template <typename D, typename T>
static int launch_fused(int field, const double* fp, const tode_tableau* tab, const tode_controller* ctrl,
                        const tode_problem* prob, const tode_solution* sol, int64_t iter_cap,
                        cudaStream_t stream) {
  if (tab->n_stages != kStagesFused) return TODE_ENOSUP;
  // load_row / store_row use one vector access per row for F == 2 (and F == 4 in fp32)
  const size_t al = prob->F == 2 ? 2 * sizeof(D) : ((prob->F == 4 && sizeof(D) == 4) ? 16 : sizeof(D));
  if (!aligned_to(prob->y0, al) || !aligned_to(sol->ys, al)) return TODE_EALIGN;
  FusedArgs<D, T> a{};
  a.tab = make_tab<D, T>(tab);
  a.ctrl = make_ctrl<D, T>(ctrl);
  for (int i = 0; i < TODE_MAX_FIELD_PARAMS; ++i) a.fp[i] = fp[i];
  a.B = prob->B;
  a.Tn = prob->T;
  a.y0 = static_cast<const D*>(prob->y0);
  a.t_start = static_cast<const T*>(prob->t_start);
  a.t_end = static_cast<const T*>(prob->t_end);
  a.t_eval = static_cast<const T*>(prob->t_eval);
  a.te_stride = prob->t_eval_stride_b;
  a.dt0 = static_cast<const T*>(prob->dt0);
  a.ys = static_cast<D*>(sol->ys);
  a.n_steps = reinterpret_cast<long long*>(sol->n_steps);
  a.n_accepted = reinterpret_cast<long long*>(sol->n_accepted);
  a.n_initialized = reinterpret_cast<long long*>(sol->n_initialized);
  a.status = reinterpret_cast<long long*>(sol->status);
  a.t_final = static_cast<T*>(sol->t_final);
  a.dt_final = static_cast<T*>(sol->dt_final);
  a.summary = sol->summary;
  a.iter_cap = iter_cap;
  a.e_init = round_exp<D>(1.0 / (double)tab->order);
  if (a.B == 0) return 0;
  switch (field) {
    case TODE_FIELD_LINEAR:
      switch (prob->F) {
        case 1: return launch_fused_f<D, T, 1, TODE_FIELD_LINEAR>(a, stream);
        case 2: return launch_fused_f<D, T, 2, TODE_FIELD_LINEAR>(a, stream);
        case 3: return launch_fused_f<D, T, 3, TODE_FIELD_LINEAR>(a, stream);
        case 4: return launch_fused_f<D, T, 4, TODE_FIELD_LINEAR>(a, stream);
        default: return TODE_ENOSUP;
      }
    case TODE_FIELD_VAN_DER_POL:
      if (prob->F != 2) return TODE_EINVAL;
      return launch_fused_f<D, T, 2, TODE_FIELD_VAN_DER_POL>(a, stream);
    case TODE_FIELD_LOTKA_VOLTERRA:
      if (prob->F != 2) return TODE_EINVAL;
      return launch_fused_f<D, T, 2, TODE_FIELD_LOTKA_VOLTERRA>(a, stream);
    default:
      return TODE_EINVAL;
  }
}

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
