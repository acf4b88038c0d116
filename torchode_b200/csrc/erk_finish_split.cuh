// Split-mode finish for few samples with a huge feature dimension (config C5: 64 x 2^20):
// a row is cut into the chunks of the canonical reduction order (kChunkVec vectors), one WARP
// per (sample, chunk), three launches per iteration:
//
//   finish_split_partial_kernel  9 rows read once: per-chunk sum of squares / max   -> scratch
//   finish_split_control_kernel  one warp per sample: chunk partials added in ascending
//                                order, controller, commit decision, all per-sample scalars,
//                                dense-output bookkeeping, termination flag
//   finish_split_commit_kernel   per chunk: dense output of the crossed t_eval points (re-read),
//                                then y <- y1, f0 <- k[S-1] where accepted
//
// Same arithmetic, same canonical order -> same bits as the single-launch kernel; only the
// parallelisation differs (the norm must be complete before accept is known, and accept gates
// the commit).  Mask ("general") mode is not supported here (the launcher never selects it).
#pragma once
#include "erk_kernels.cuh"

namespace tode {

template <typename T>
struct SplitAux {
  T t0, dt;
  int cur_old, cur_new;
  int flags;  // bit0: sample was running, bit1: accepted (commit), bit2: evaluate t_end (T == 0)
  int pad;
};

template <typename D, typename T>
TODE_DEV D* split_partials(const FinishArgs<D, T>& A) {
  return A.scratch;
}
template <typename D, typename T>
TODE_DEV SplitAux<T>* split_aux(const FinishArgs<D, T>& A, long long chunks_per_row) {
  // after the partials, 16-byte aligned
  unsigned long long p = reinterpret_cast<unsigned long long>(A.scratch + A.B * chunks_per_row);
  p = (p + 15ull) & ~15ull;
  return reinterpret_cast<SplitAux<T>*>(p);
}

// Chunk partials of one sample combined in ascending chunk order (RowNorm::flush / result) by a whole
// warp: the lanes fetch 32 partials at a time (one coalesced load instead of 32 dependent round trips to
// L2 by a single thread: 21 us -> 3 us per launch at 256 chunks) and every lane adds them in the canonical
// order through shuffles.  All lanes return the norm.
template <typename D>
TODE_DEV D warp_combine_partials(const D* p, long long cpr, int kind, int lane) {
  D total = (D)0;
  bool first = true;
  for (long long base = 0; base < cpr; base += 32) {
    const long long i = base + lane;
    const D v = i < cpr ? p[i] : (D)0;
    const int cnt = cpr - base < 32 ? (int)(cpr - base) : 32;
    for (int k = 0; k < cnt; ++k) {
      const D x = __shfl_sync(0xffffffffu, v, k);
      if (first) total = x;
      else total = kind == TODE_NORM_MAX ? max_nan_nn(total, x) : add(total, x);
      first = false;
    }
  }
  return kind == TODE_NORM_MAX ? total : fsqrt(total);
}

template <typename D, typename T, int VEC>
__global__ void __launch_bounds__(kBlock) finish_split_partial_kernel(const __grid_constant__ FinishArgs<D, T> A) {
  if (A.ctl[TODE_CTL_STOP]) return;
  constexpr int S = kStages;
  const int lane = threadIdx.x & 31;
  const long long n = A.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long warps_total = A.B * cpr;
  const TabP<D, T>& tab = A.tab;
  const CtrlP<D, T>& c = A.ctrl;
  D* partials = split_partials(A);
  for (long long w = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5; w < warps_total;
       w += ((long long)gridDim.x * kBlock) >> 5) {
    const long long b = w / cpr, ch = w % cpr;
    if (!A.running[b]) continue;  // warp-uniform
    const D dtD = (D)A.dt[b];
    const long long row = b * A.F;
    D part = (D)0;
    bool first = true;
#pragma unroll 2
    for (int i = 0; i < kChunkVec / 32; ++i) {
      const long long j = ch * kChunkVec + lane + 32LL * i;
      if (j < n) {
        const long long e = row + j * VEC;
        D y0v[VEC], y1v[VEC], kv[S][VEC];
        VecIO<D, VEC>::ld(A.y + e, y0v);
        VecIO<D, VEC>::ld(A.y1 + e, y1v);
#pragma unroll
        for (int s = 0; s < S; ++s) VecIO<D, VEC>::ld(A.k[s] + e, kv[s]);
#pragma unroll
        for (int x = 0; x < VEC; ++x) {
          D ks[S];
#pragma unroll
          for (int s = 0; s < S; ++s) ks[s] = kv[s][x];
          const D err = weighted_sum<D, S>(dtD, tab.b_err, ks);
          const D bounds = ffma(c.rtol, max_nan_nn(fabs_(y0v[x]), fabs_(y1v[x])), c.atol);
          const D q = fdiv(fabs_(err), bounds);
          if (c.norm == TODE_NORM_MAX) {
            part = first ? fabs_(q) : max_nan_nn(part, fabs_(q));
            first = false;
          } else {
            sumsq_acc(part, first, fdiv(q, A.sqrt_f));
          }
        }
      }
    }
    const D v = c.norm == TODE_NORM_MAX ? group_max<D, 32>(part) : group_sum<D, 32>(part);
    if (lane == 0) partials[w] = v;
  }
}

// one WARP per sample: the lanes combine the chunk partials, lane 0 does the per-sample scalar work
template <typename D, typename T>
__global__ void __launch_bounds__(kBlock) finish_split_control_kernel(const __grid_constant__ FinishArgs<D, T> A,
                                                                        long long cpr) {
  SplitAux<T>* aux = split_aux(A, cpr);
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * kBlock) >> 5;
  if (A.ctl[TODE_CTL_STOP]) {
    // no-op iteration (launched speculatively after the stop): clear the step records so
    // that the commit kernel of this iteration does nothing
    for (long long b = warp0; b < A.B; b += n_warps)
      if (lane == 0) aux[b].flags = 0;
    return;
  }
  const CtrlP<D, T>& c = A.ctrl;
  const D* partials = split_partials(A);
  int my_running = 0, my_failed = 0;
  for (long long b = warp0; b < A.B; b += n_warps) {
    SplitAux<T> a;
    a.flags = 0;
    a.cur_old = a.cur_new = 0;
    a.pad = 0;
    a.t0 = (T)0;
    a.dt = (T)0;
    if (A.running[b]) {  // warp-uniform
      const D nrm = warp_combine_partials<D>(partials + b * cpr, cpr, c.norm, lane);
      if (lane == 0) {
        const T t0 = A.t[b], dt = A.dt[b], ts = A.t_start[b], te = A.t_end[b];
        const int ns = A.n_steps[b] + 1;
        const D r1 = c.pid ? A.r1[b] : (D)1, r2 = c.pid ? A.r2[b] : (D)1;
        const Decision<D, T> d = decide_step<D, T>(c, nrm, t0, dt, ts, te, r1, r2, ns, true);
        int cur = 0;
        a.flags = 1 | (d.upd ? 2 : 0);
        a.t0 = t0;
        a.dt = dt;
        if (A.Tn == 0) {
          if (!d.running_new || d.status != TODE_SUCCESS || (c.iter_cap > 0 && (long long)ns >= c.iter_cap))
            a.flags |= 4;
        } else {
          const T* tev = A.t_eval + b * A.te_stride;
          cur = A.cursor[b];
          a.cur_old = cur;
          while (cur < A.Tn && ffma(d.dir, d.t_new, mul(-d.dir, tev[cur])) >= (T)0) ++cur;  // :216-223
          a.cur_new = cur;
        }
        store_sample_scalars<D, T>(A, b, d, dt, ts, te, ns, cur, my_running, my_failed);
        if (A.flip != nullptr && d.upd) A.flip[b] ^= 1;  // commit by pointer flip (heat_step.cu)
      }
    }
    if (lane == 0) aux[b] = a;
  }
  publish_termination(A.ctl, my_running, my_failed);
}

// Runs after the control kernel of the same iteration (which may just have set the stop flag
// for the FOLLOWING iterations), so it does not test the stop flag: it is gated by the
// per-sample step records, which the control kernel clears in a no-op iteration.
template <typename D, typename T, int VEC>
__global__ void __launch_bounds__(kBlock) finish_split_commit_kernel(const __grid_constant__ FinishArgs<D, T> A) {
  constexpr int S = kStages;
  const int lane = threadIdx.x & 31;
  const long long n = A.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long warps_total = A.B * cpr;
  const TabP<D, T>& tab = A.tab;
  const SplitAux<T>* aux = split_aux(A, cpr);
  for (long long w = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5; w < warps_total;
       w += ((long long)gridDim.x * kBlock) >> 5) {
    const long long b = w / cpr, ch = w % cpr;
    const SplitAux<T> a = aux[b];
    const bool upd = (a.flags & 2) != 0;
    const bool eval_end = (a.flags & 4) != 0;
    const int n_pts = eval_end ? 1 : (a.cur_new - a.cur_old);
    if (!(a.flags & 1) || (!upd && n_pts == 0)) continue;  // warp-uniform
    const D dtD = (D)a.dt;
    const long long row = b * A.F;
    const T* tev = A.Tn > 0 ? A.t_eval + b * A.te_stride : nullptr;
    const T te = A.t_end[b];
    for (int i = 0; i < kChunkVec / 32; ++i) {
      const long long j = ch * kChunkVec + lane + 32LL * i;
      if (j >= n) continue;
      const long long e = row + j * VEC;
      D y1v[VEC], k6v[VEC];
      if (n_pts > 0) {
        D y0v[VEC], kv[S][VEC], co[VEC][5];
        VecIO<D, VEC>::ld(A.y + e, y0v);
        VecIO<D, VEC>::ld(A.y1 + e, y1v);
#pragma unroll
        for (int s = 0; s < S; ++s) VecIO<D, VEC>::ld(A.k[s] + e, kv[s]);
#pragma unroll
        for (int x = 0; x < VEC; ++x) {
          D ks[S];
#pragma unroll
          for (int s = 0; s < S; ++s) ks[s] = kv[s][x];
          interp_coeffs<D, T, S>(tab, dtD, y0v[x], y1v[x], ks, co[x]);
          k6v[x] = kv[S - 1][x];
        }
        for (int p = 0; p < n_pts; ++p) {
          const T tq = eval_end ? te : tev[a.cur_old + p];
          const D xq = interp_x<D, T>(tq, a.t0, a.dt);
          D out[VEC];
#pragma unroll
          for (int x = 0; x < VEC; ++x) out[x] = horner4<D>(co[x], xq);
          D* dst = eval_end ? A.y_eval + row : A.y_eval + (b * A.Tn + a.cur_old + p) * A.F;
          VecIO<D, VEC>::st(dst + j * VEC, out);
        }
      } else {
        VecIO<D, VEC>::ld(A.y1 + e, y1v);
        VecIO<D, VEC>::ld(A.k[S - 1] + e, k6v);
      }
      if (upd) {
        VecIO<D, VEC>::st(A.y + e, y1v);
        VecIO<D, VEC>::st(A.f0 + e, k6v);
      }
    }
  }
}

}  // namespace tode
