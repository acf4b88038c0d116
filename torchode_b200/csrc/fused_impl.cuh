// Launcher of solve_fused_kernel, explicitly instantiated per (data, time) dtype pair in
// fused_f32f32.cu / fused_f64f64.cu / fused_f32f64.cu / fused_f64f32.cu (parallel build).
#pragma once
#include <climits>
#include <cstdlib>

#include "api_common.cuh"
#include "erk_fused.cuh"
#include "erk_fused_f2.cuh"

namespace tode {

// resident 128-thread CTAs per SM the fused kernel is compiled for, measured on C2 / C3: fp64
// state 5 (96 registers, ~100 B of spills; 6 = 80 registers spills ~200 B and is 4 % slower since
// the table-driven pow, 4 and 7 are slower too), fp32 state 4
#ifndef TODE_FUSED_MINB
#define TODE_FUSED_MINB (sizeof(D) == 8 ? 5 : 4)
#endif

__global__ void summary_init_kernel(int* summary, int n_launches);

template <typename D, typename T, int F, int FIELD, int CK, bool TE>
static int launch_fused_k(const FusedArgs<D, T>& a, cudaStream_t stream) {
  constexpr int kThreads = kFusedThreads;
  summary_init_kernel<<<1, 1, 0, stream>>>(a.summary, 2 + (a.n_peers > 0 ? 1 : 0));
  const unsigned grid = (unsigned)((a.B + kThreads - 1) / kThreads);
  solve_fused_kernel<D, T, F, FIELD, TODE_FUSED_MINB, CK, TE><<<grid, kThreads, 0, stream>>>(a);
  return launch_status();
}

// fp32 state and time, two features, t_eval rows (BASELINE configs[2]): the packed-fp32 persistent
// kernel of erk_fused_f2.cuh -- pre-pass (initial step, monotonicity) + solve with lane refill.
// The pre-pass lends the low words of the int64 n_steps output as its (B) float scratch.
#ifndef TODE_F2_MINB
#define TODE_F2_MINB 4
#endif
template <int FIELD, int CK>
static int launch_fused_f2_k(const FusedArgs<float, float>& a, cudaStream_t stream) {
  summary_init_kernel<<<1, 1, 0, stream>>>(a.summary, 3 + (a.n_peers > 0 ? 1 : 0));
  float* dt_scratch = reinterpret_cast<float*>(a.n_steps);
  fused_f2_init_kernel<FIELD><<<(unsigned)((a.B + 255) / 256), 256, 0, stream>>>(a, dt_scratch, 2);
  int ctas_per_sm = TODE_F2_MINB;
  if (const char* e = getenv("TODE_F2_CTAS")) {  // tuning knob: resident CTAs per SM the grid is sized for
    const int v = atoi(e);
    if (v >= 1 && v <= TODE_F2_MINB) ctas_per_sm = v;
  }
  const unsigned grid = grid_for(a.B, kF2Threads, ctas_per_sm);
  solve_fused_f2_kernel<FIELD, TODE_F2_MINB, CK><<<grid, kF2Threads, 0, stream>>>(
      a, dt_scratch, 2, reinterpret_cast<unsigned long long*>(a.summary + 4));
  return launch_status();
}
template <int FIELD>
static int launch_fused_f2(const FusedArgs<float, float>& a, int ck, cudaStream_t stream) {
  switch (ck) {
    case 0: return launch_fused_f2_k<FIELD, 0>(a, stream);
    case 1: return launch_fused_f2_k<FIELD, 1>(a, stream);
    default: return launch_fused_f2_k<FIELD, 2>(a, stream);
  }
}
// does the problem fit the f2 kernel?  (ys replicas written by the kernel keep the general kernel)
template <typename D, typename T>
static bool f2_route(const FusedArgs<D, T>& a, long long F) {
  if (sizeof(D) != 4 || sizeof(T) != 4 || F != 2 || a.Tn <= 0 || a.Tn > 0x3fffffff) return false;
  if (!aligned_to(a.summary, 8) || getenv("TODE_NO_F2") != nullptr) return false;
  for (int p = 0; p < a.n_peers; ++p)
    if (a.p_ys[p] != nullptr) return false;
  return true;
}

// SPEC: instantiate the specialised variants (same data / time dtype, the built-in 2-feature
// fields and 1- / 2-feature linear decay); everything else runs the general <2, true> kernel
template <typename D, typename T, int F, int FIELD, bool SPEC>
static int launch_fused_f(const FusedArgs<D, T>& a, int ck, cudaStream_t stream) {
  if constexpr (SPEC) {
    const bool te = a.Tn > 0;
    switch (ck * 2 + (te ? 1 : 0)) {
      case 0: return launch_fused_k<D, T, F, FIELD, 0, false>(a, stream);
      case 1: return launch_fused_k<D, T, F, FIELD, 0, true>(a, stream);
      case 2: return launch_fused_k<D, T, F, FIELD, 1, false>(a, stream);
      case 3: return launch_fused_k<D, T, F, FIELD, 1, true>(a, stream);
      case 4: return launch_fused_k<D, T, F, FIELD, 2, false>(a, stream);
      default: break;
    }
  }
  return launch_fused_k<D, T, F, FIELD, 2, true>(a, stream);
}

template <typename D, typename T>
int launch_fused(int field, const double* fp, const tode_tableau* tab, const tode_controller* ctrl,
                 const tode_problem* prob, const tode_solution* sol, int64_t iter_cap, cudaStream_t stream) {
  if (tab->n_stages != kStagesFused) return TODE_ENOSUP;
  // load_row / store_row use one vector access per row for F == 2 (and F == 4 in fp32)
  const size_t al = prob->F == 2 ? 2 * sizeof(D) : ((prob->F == 4 && sizeof(D) == 4) ? 16 : sizeof(D));
  if (!aligned_to(prob->y0, al) || !aligned_to(sol->ys, al)) return TODE_EALIGN;
  FusedArgs<D, T> a{};
  a.tab = make_tab<D, T>(tab);
  a.ctrl = make_ctrl<D, T>(ctrl);
  a.pow = make_powtab();
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
  if (sol->n_peers < 0 || sol->n_peers > TODE_MAX_PEERS) return TODE_EINVAL;
  a.n_peers = sol->n_peers;
  a.peer_row0 = sol->peer_row0;
  for (int p = 0; p < sol->n_peers; ++p) {
    const int n_stats = (sol->peer_n_steps[p] != nullptr) + (sol->peer_n_accepted[p] != nullptr) +
                        (sol->peer_n_initialized[p] != nullptr) + (sol->peer_status[p] != nullptr);
    if ((n_stats != 0 && n_stats != 4) || !sol->peer_global[p]) return TODE_EINVAL;
    if (sol->peer_ys[p] && !aligned_to(sol->peer_ys[p], al)) return TODE_EALIGN;
    a.p_ys[p] = static_cast<D*>(sol->peer_ys[p]);
    a.p_n_steps[p] = reinterpret_cast<long long*>(sol->peer_n_steps[p]);
    a.p_n_accepted[p] = reinterpret_cast<long long*>(sol->peer_n_accepted[p]);
    a.p_n_initialized[p] = reinterpret_cast<long long*>(sol->peer_n_initialized[p]);
    a.p_status[p] = reinterpret_cast<long long*>(sol->peer_status[p]);
  }
  if (a.B == 0) return 0;
  // what the controller needs: 0 = no history, 1 = r1 only, 2 = r1 and r2
  const int ck = !a.ctrl.pid ? 0 : (a.ctrl.e_prev2 == 0.0 ? 1 : 2);
  constexpr bool kSame = sizeof(D) == sizeof(T);
  if constexpr (sizeof(D) == 4 && sizeof(T) == 4) {
    if (f2_route(a, prob->F)) {
      switch (field) {
        case TODE_FIELD_LINEAR: return launch_fused_f2<TODE_FIELD_LINEAR>(a, ck, stream);
        case TODE_FIELD_VAN_DER_POL: return launch_fused_f2<TODE_FIELD_VAN_DER_POL>(a, ck, stream);
        case TODE_FIELD_LOTKA_VOLTERRA: return launch_fused_f2<TODE_FIELD_LOTKA_VOLTERRA>(a, ck, stream);
        default: return TODE_EINVAL;
      }
    }
  }
  switch (field) {
    case TODE_FIELD_LINEAR:
      switch (prob->F) {
        case 1: return launch_fused_f<D, T, 1, TODE_FIELD_LINEAR, kSame>(a, ck, stream);
        case 2: return launch_fused_f<D, T, 2, TODE_FIELD_LINEAR, kSame>(a, ck, stream);
        case 3: return launch_fused_f<D, T, 3, TODE_FIELD_LINEAR, false>(a, ck, stream);
        case 4: return launch_fused_f<D, T, 4, TODE_FIELD_LINEAR, false>(a, ck, stream);
        default: return TODE_ENOSUP;
      }
    case TODE_FIELD_VAN_DER_POL:
      if (prob->F != 2) return TODE_EINVAL;
      return launch_fused_f<D, T, 2, TODE_FIELD_VAN_DER_POL, kSame>(a, ck, stream);
    case TODE_FIELD_LOTKA_VOLTERRA:
      if (prob->F != 2) return TODE_EINVAL;
      return launch_fused_f<D, T, 2, TODE_FIELD_LOTKA_VOLTERRA, kSame>(a, ck, stream);
    default:
      return TODE_EINVAL;
  }
}

}  // namespace tode
