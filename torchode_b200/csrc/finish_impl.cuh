// Launcher of erk_finish_kernel, explicitly instantiated per (data, time) dtype pair in
// finish_f32f32.cu / finish_f64f64.cu / finish_f32f64.cu / finish_f64f32.cu (parallel build).
#pragma once
#include "api_common.cuh"
#include "erk_kernels.cuh"
#include "erk_finish_split.cuh"

namespace tode {

template <typename D, typename T, int G, int VEC, int CI>
static int launch_finish_cfg(const FinishArgs<D, T>& a, cudaStream_t stream) {
  const long long gpb = kBlock / G;
  // persistent-style grid: few CTAs -> few termination atomics; 8 resident CTAs / SM
  const unsigned grid = grid_for(a.B, gpb, 8);
  erk_finish_kernel<D, T, G, VEC, CI><<<grid, kBlock, 0, stream>>>(a);
  return launch_status();
}

// few samples x huge rows: one warp per (sample, chunk), three launches (erk_finish_split.cuh)
template <typename D, typename T, int VEC>
static int launch_finish_split(const FinishArgs<D, T>& a, cudaStream_t stream) {
  const long long n = a.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long need = a.B * cpr + (a.B * (long long)sizeof(SplitAux<T>) + 32) / (long long)sizeof(D) + 8;
  if (a.scratch == nullptr || a.scratch_elems < need) return TODE_EINVAL;
  const unsigned grid = grid_for(a.B * cpr, kBlock / 32, 8);
  finish_split_partial_kernel<D, T, VEC><<<grid, kBlock, 0, stream>>>(a);
  finish_split_control_kernel<D, T><<<grid_for(a.B, kBlock / 32, 1), kBlock, 0, stream>>>(a, cpr);
  finish_split_commit_kernel<D, T, VEC><<<grid, kBlock, 0, stream>>>(a);
  return launch_status();
}

template <typename D, typename T, int VEC>
static int launch_finish_vec(const FinishArgs<D, T>& a, cudaStream_t stream) {
  // warp-per-sample cannot fill the machine with few, very long rows
  if (a.not_yet == nullptr && a.F / VEC >= 2 * kChunkVec && a.B < 16LL * sm_count())
    return launch_finish_split<D, T, VEC>(a, stream);
  // lane-group size of the canonical geometry; groups of 1..16 lanes keep the 9 rows in
  // registers (one chunk per lane), warp-per-sample streams (re-reads hit L1/L2)
  int g = geom_lanes(a.F / VEC);
  if (a.not_yet != nullptr && g > 1) g = 32;  // mask mode needs one sample per warp
  switch (g) {
    case 1: return launch_finish_cfg<D, T, 1, VEC, 1>(a, stream);
    case 2: return launch_finish_cfg<D, T, 2, VEC, 1>(a, stream);
    case 4: return launch_finish_cfg<D, T, 4, VEC, 1>(a, stream);
    case 8: return launch_finish_cfg<D, T, 8, VEC, 1>(a, stream);
    case 16: return launch_finish_cfg<D, T, 16, VEC, 1>(a, stream);
    default: return launch_finish_cfg<D, T, 32, VEC, 0>(a, stream);
  }
}

template <typename D, typename T>
int launch_finish(const tode_tableau* tab, const tode_controller* ctrl, const tode_state* st,
                         const void* const* k, const void* y1, cudaStream_t stream) {
  if (tab->n_stages != kStages) return TODE_ENOSUP;
  const int vec = geom_vec<D>(st->F);
  const size_t al = sizeof(D) * vec;
  if (!aligned_to(st->y, al) || !aligned_to(st->f0, al) || !aligned_to(y1, al) ||
      !aligned_to(st->y_eval, al))
    return TODE_EALIGN;
  FinishArgs<D, T> a{};
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
  a.y = static_cast<D*>(st->y);
  a.f0 = static_cast<D*>(st->f0);
  a.r1 = static_cast<D*>(st->r1);
  a.r2 = static_cast<D*>(st->r2);
  a.running = st->running;
  a.n_steps = st->n_steps;
  a.n_accepted = st->n_accepted;
  a.status = st->status;
  a.cursor = st->cursor;
  a.not_yet = st->not_yet;
  a.y_eval = static_cast<D*>(st->y_eval);
  a.t_nodes = static_cast<T*>(st->t_nodes);
  a.ctl = st->ctl;
  for (int s = 0; s < kStages; ++s) {
    if (!aligned_to(k[s], al)) return TODE_EALIGN;
    a.k[s] = static_cast<const D*>(k[s]);
  }
  a.y1 = static_cast<const D*>(y1);
  a.sqrt_f = (D)std::sqrt((double)st->F);
  a.scratch = static_cast<D*>(st->scratch);
  a.scratch_elems = st->scratch_elems;
  if (a.B == 0) return 0;
  if (sizeof(D) == 4 && vec == 4) return launch_finish_vec<D, T, (sizeof(D) == 4 ? 4 : 2)>(a, stream);
  if (vec == 2) return launch_finish_vec<D, T, 2>(a, stream);
  return launch_finish_vec<D, T, 1>(a, stream);
}

}  // namespace tode
