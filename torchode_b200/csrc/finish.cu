// C-ABI: tode_erk_finish (one launch per loop iteration) and the stand-alone
// tode_adapt_step_size controller op.
#include "api_common.cuh"
#include "erk_kernels.cuh"

namespace tode {

template <typename D, typename T>
int launch_finish(const tode_tableau* tab, const tode_controller* ctrl, const tode_state* st,
                  const void* const* k, const void* y1, cudaStream_t stream);

// ---- stand-alone controller op (step_size_controllers.py:393-429 / 738-774) --------------
template <typename D, typename T>
struct AdaptArgs {
  CtrlP<D, T> ctrl;
  long long B, F;
  const T* dt;
  const D* y0;
  const D* y1;
  const D* err;
  const D* r1;
  const D* r2;
  uint8_t* accept;
  T* dt_next;
  D* ratio;
  D* r1o;
  D* r2o;
  long long* status;
  D sqrt_f;
};

template <typename D, typename T, int G, int VEC>
__global__ void __launch_bounds__(kBlock) adapt_kernel(const __grid_constant__ AdaptArgs<D, T> A) {
  const int lane = threadIdx.x % G;
  const long long gpb = kBlock / G;
  const long long n = A.F / VEC;
  const long long n_it = (n + G - 1) / G;
  const CtrlP<D, T>& c = A.ctrl;
  for (long long base = (long long)blockIdx.x * gpb; base < A.B; base += (long long)gridDim.x * gpb) {
    const long long b = base + threadIdx.x / G;
    const bool act = b < A.B;
    const long long row = b * A.F;
    RowNorm<D, G> rn(c.norm, A.sqrt_f);
    for (long long it = 0; it < n_it; ++it) {
      const long long j = lane + it * G;
      if (act && j < n) {
        D y0v[VEC], y1v[VEC], ev[VEC];
        VecIO<D, VEC>::ld(A.y0 + row + j * VEC, y0v);
        VecIO<D, VEC>::ld(A.y1 + row + j * VEC, y1v);
        VecIO<D, VEC>::ld(A.err + row + j * VEC, ev);
#pragma unroll
        for (int x = 0; x < VEC; ++x) {
          const D bounds = ffma(c.rtol, max_nan_nn(fabs_(y0v[x]), fabs_(y1v[x])), c.atol);
          rn.add(fdiv(fabs_(ev[x]), bounds));
        }
      }
      rn.slot_done();
    }
    const D nrm = rn.result();
    if (act && lane == 0) {
      const D r1 = A.r1 != nullptr ? A.r1[b] : (D)1;
      const D r2 = A.r2 != nullptr ? A.r2[b] : (D)1;
      const CtrlOut<D, T> o = controller<D, T>(c, nrm, A.dt[b], r1, r2);
      A.accept[b] = (uint8_t)o.accept;
      A.dt_next[b] = o.dt_next;
      if (A.ratio != nullptr) A.ratio[b] = o.ratio;
      if (A.r1o != nullptr) A.r1o[b] = o.r1;
      if (A.r2o != nullptr) A.r2o[b] = o.r2;
      A.status[b] = o.status;
    }
  }
}

template <typename D, typename T, int VEC>
static int launch_adapt_vec(const AdaptArgs<D, T>& a, cudaStream_t stream) {
  if (geom_lanes(a.F / VEC) == 1) {
    adapt_kernel<D, T, 1, VEC><<<grid_for(a.B, kBlock, 8), kBlock, 0, stream>>>(a);
  } else {
    adapt_kernel<D, T, 32, VEC><<<grid_for(a.B, kBlock / 32, 8), kBlock, 0, stream>>>(a);
  }
  return launch_status();
}

template <typename D, typename T>
static int launch_adapt(const tode_controller* ctrl, int64_t B, int64_t F, const void* dt,
                        const void* y0, const void* y1, const void* err, const void* r1, const void* r2,
                        uint8_t* accept, void* dt_next, void* ratio, void* r1o, void* r2o,
                        int64_t* status, cudaStream_t stream) {
  const int vec = geom_vec<D>(F);
  const size_t al = sizeof(D) * vec;
  if (!aligned_to(y0, al) || !aligned_to(y1, al) || !aligned_to(err, al)) return TODE_EALIGN;
  AdaptArgs<D, T> a{};
  a.ctrl = make_ctrl<D, T>(ctrl);
  a.B = B;
  a.F = F;
  a.dt = static_cast<const T*>(dt);
  a.y0 = static_cast<const D*>(y0);
  a.y1 = static_cast<const D*>(y1);
  a.err = static_cast<const D*>(err);
  a.r1 = static_cast<const D*>(r1);
  a.r2 = static_cast<const D*>(r2);
  a.accept = accept;
  a.dt_next = static_cast<T*>(dt_next);
  a.ratio = static_cast<D*>(ratio);
  a.r1o = static_cast<D*>(r1o);
  a.r2o = static_cast<D*>(r2o);
  a.status = reinterpret_cast<long long*>(status);
  a.sqrt_f = (D)std::sqrt((double)F);
  if (B == 0) return 0;
  if (sizeof(D) == 4 && vec == 4) return launch_adapt_vec<D, T, (sizeof(D) == 4 ? 4 : 2)>(a, stream);
  if (vec == 2) return launch_adapt_vec<D, T, 2>(a, stream);
  return launch_adapt_vec<D, T, 1>(a, stream);
}

}  // namespace tode

using namespace tode;

extern "C" int tode_erk_finish(const tode_tableau* tab, const tode_controller* ctrl,
                               const tode_state* st, const void* const* k, const void* y1,
                               void* stream) {
  if (!tab || !ctrl || !st || !k || !y1) return TODE_EINVAL;
  if (!st->t || !st->dt || !st->y || !st->f0 || !st->running || !st->n_steps || !st->n_accepted ||
      !st->status || !st->y_eval || !st->ctl || !st->t_start || !st->t_end)
    return TODE_EINVAL;
  if (ctrl->pid && (!st->r1 || !st->r2)) return TODE_EINVAL;
  if (st->T > 0 && (!st->t_eval || (!st->cursor && !st->not_yet))) return TODE_EINVAL;
  for (int s = 0; s < tab->n_stages; ++s)
    if (!k[s]) return TODE_EINVAL;
#define CALL(D, T) launch_finish<D, T>(tab, ctrl, st, k, y1, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

extern "C" int tode_adapt_step_size(const tode_controller* ctrl, int32_t data_dtype, int32_t time_dtype,
                                    int64_t B, int64_t F, const void* dt, const void* y0,
                                    const void* y1, const void* err, const void* r1, const void* r2,
                                    uint8_t* accept_out, void* dt_next_out, void* ratio_out,
                                    void* r1_out, void* r2_out, int64_t* status_out, void* stream) {
  if (!ctrl || !dt || !y0 || !y1 || !err || !accept_out || !dt_next_out || !status_out) return TODE_EINVAL;
#define CALL(D, T)                                                                                   \
  launch_adapt<D, T>(ctrl, B, F, dt, y0, y1, err, r1, r2, accept_out, dt_next_out, ratio_out, r1_out, \
                     r2_out, status_out, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(data_dtype, time_dtype, CALL);
#undef CALL
}
