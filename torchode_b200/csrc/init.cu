// C-ABI: initial step size (Hairer heuristic around the user's second f evaluation) and
// solver-state initialisation.
#include "api_common.cuh"
#include "erk_init_split.cuh"
#include "erk_kernels.cuh"

namespace tode {

template <typename D, typename T>
static int fill_init_args(InitArgs<D, T>& a, const tode_tableau* tab, const tode_controller* ctrl,
                          const tode_state* st) {
  a.tab = make_tab<D, T>(tab);
  a.ctrl = make_ctrl<D, T>(ctrl);
  a.B = st->B;
  a.F = st->F;
  a.Tn = st->T;
  a.t_start = static_cast<const T*>(st->t_start);
  a.t_end = static_cast<const T*>(st->t_end);
  a.t_eval = static_cast<const T*>(st->t_eval);
  a.te_stride = st->t_eval_stride_b;
  a.t = static_cast<T*>(st->t);
  a.dt = static_cast<T*>(st->dt);
  a.y = static_cast<const D*>(st->y);
  a.f0 = static_cast<const D*>(st->f0);
  a.r1 = static_cast<D*>(st->r1);
  a.r2 = static_cast<D*>(st->r2);
  a.running = st->running;
  a.n_steps = st->n_steps;
  a.n_accepted = st->n_accepted;
  a.status = st->status;
  a.cursor = st->cursor;
  a.y_eval = static_cast<D*>(st->y_eval);
  a.t_nodes = static_cast<T*>(st->t_nodes);
  a.ctl = st->ctl;
  a.scratch = static_cast<D*>(st->scratch);
  a.scratch_elems = st->scratch_elems;
  a.e_init = round_exp<D>(1.0 / (double)tab->order);
  a.sqrt_f = (D)std::sqrt((double)st->F);
  return 0;
}

// few samples x huge rows: one warp per (sample, chunk), two launches per part (erk_init_split.cuh)
template <typename D, typename T, int VEC, bool PART_B>
static int launch_init_split(const InitArgs<D, T>& a, long long cpr, cudaStream_t stream) {
  const unsigned grid = grid_for(a.B * cpr, kBlock / 32, 8);
  if (PART_B) {
    init_split_b_partial_kernel<D, T, VEC><<<grid, kBlock, 0, stream>>>(a);
    init_split_b_finish_kernel<D, T><<<grid_for(a.B, kBlock, 1), kBlock, 0, stream>>>(a, cpr);
  } else {
    init_split_a_partial_kernel<D, T, VEC><<<grid, kBlock, 0, stream>>>(a);
    init_split_a_finish_kernel<D, T, VEC><<<grid, kBlock, 0, stream>>>(a);
  }
  return launch_status();
}

template <typename D, typename T, int VEC, bool PART_B>
static int launch_init_vec(const InitArgs<D, T>& a, cudaStream_t stream) {
  // same condition as the split finish (finish_impl.cuh); needs room for the chunk partials
  const long long n = a.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  if (n >= 2 * kChunkVec && a.B < 16LL * sm_count() && a.scratch != nullptr &&
      a.scratch_elems >= init_split_scratch_need(a.B, cpr))
    return launch_init_split<D, T, VEC, PART_B>(a, cpr, stream);
  const bool thread_per_sample = geom_lanes(a.F / VEC) == 1;
  const unsigned grid = grid_for(a.B, thread_per_sample ? kBlock : kBlock / 32, 8);
  if (thread_per_sample) {
    if (PART_B) init_step_b_kernel<D, T, 1, VEC><<<grid, kBlock, 0, stream>>>(a);
    else init_step_a_kernel<D, T, 1, VEC><<<grid, kBlock, 0, stream>>>(a);
  } else {
    if (PART_B) init_step_b_kernel<D, T, 32, VEC><<<grid, kBlock, 0, stream>>>(a);
    else init_step_a_kernel<D, T, 32, VEC><<<grid, kBlock, 0, stream>>>(a);
  }
  return launch_status();
}

template <typename D, typename T, bool PART_B>
static int launch_init(const InitArgs<D, T>& a, cudaStream_t stream) {
  if (a.B == 0) return 0;
  const int vec = geom_vec<D>(a.F);
  if (sizeof(D) == 4 && vec == 4) return launch_init_vec<D, T, (sizeof(D) == 4 ? 4 : 2), PART_B>(a, stream);
  if (vec == 2) return launch_init_vec<D, T, 2, PART_B>(a, stream);
  return launch_init_vec<D, T, 1, PART_B>(a, stream);
}

template <typename D, typename T>
static int init_a(const tode_tableau* tab, const tode_controller* ctrl, const tode_state* st,
                  void* y1_out, void* t1_out, cudaStream_t stream) {
  const size_t al = sizeof(D) * geom_vec<D>(st->F);
  if (!aligned_to(st->y, al) || !aligned_to(st->f0, al) || !aligned_to(y1_out, al)) return TODE_EALIGN;
  InitArgs<D, T> a{};
  fill_init_args(a, tab, ctrl, st);
  a.y1_out = static_cast<D*>(y1_out);
  a.t1_out = static_cast<T*>(t1_out);
  return launch_init<D, T, false>(a, stream);
}

template <typename D, typename T>
static int init_b(const tode_tableau* tab, const tode_controller* ctrl, const tode_state* st,
                  const void* f1, const void* dt0, cudaStream_t stream) {
  const size_t al = sizeof(D) * geom_vec<D>(st->F);
  if (!aligned_to(st->y, al) || !aligned_to(st->f0, al) || !aligned_to(st->y_eval, al) ||
      (f1 && !aligned_to(f1, al)))
    return TODE_EALIGN;
  const cudaError_t e = cudaMemsetAsync(st->ctl, 0, sizeof(int32_t) * TODE_CTL_WORDS, stream);
  if (e != cudaSuccess) return (int)e;
  InitArgs<D, T> a{};
  fill_init_args(a, tab, ctrl, st);
  a.f1 = static_cast<const D*>(f1);
  a.dt0 = static_cast<const T*>(dt0);
  return launch_init<D, T, true>(a, stream);
}

static int check_state(const tode_state* st, const tode_controller* ctrl) {
  if (!st->t || !st->dt || !st->y || !st->f0 || !st->running || !st->n_steps || !st->n_accepted ||
      !st->status || !st->y_eval || !st->ctl || !st->t_start || !st->t_end)
    return TODE_EINVAL;
  if (ctrl->pid && (!st->r1 || !st->r2)) return TODE_EINVAL;
  if (st->T > 0 && (!st->t_eval || !st->cursor)) return TODE_EINVAL;
  return 0;
}

}  // namespace tode

using namespace tode;

extern "C" int tode_init_step_a(const tode_tableau* tab, const tode_controller* ctrl,
                                const tode_state* st, void* y1_out, void* t1_out, void* stream) {
  if (!tab || !ctrl || !st || !y1_out || !t1_out) return TODE_EINVAL;
  if (!st->y || !st->f0 || !st->t_start || !st->t_end || !st->scratch || st->scratch_elems < 2 * st->B)
    return TODE_EINVAL;
#define CALL(D, T) init_a<D, T>(tab, ctrl, st, y1_out, t1_out, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

extern "C" int tode_init_step_b(const tode_tableau* tab, const tode_controller* ctrl,
                                const tode_state* st, const void* f1, void* stream) {
  if (!tab || !ctrl || !st || !f1) return TODE_EINVAL;
  if (!st->scratch || st->scratch_elems < 2 * st->B) return TODE_EINVAL;
  if (int rc = check_state(st, ctrl)) return rc;
#define CALL(D, T) init_b<D, T>(tab, ctrl, st, f1, nullptr, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

extern "C" int tode_init_with_dt0(const tode_tableau* tab, const tode_controller* ctrl,
                                  const tode_state* st, const void* dt0, void* stream) {
  if (!tab || !ctrl || !st || !dt0) return TODE_EINVAL;
  if (int rc = check_state(st, ctrl)) return rc;
#define CALL(D, T) init_b<D, T>(tab, ctrl, st, nullptr, dt0, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}
