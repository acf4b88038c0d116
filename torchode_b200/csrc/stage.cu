// C-ABI: tode_erk_stage and the light stand-alone protocol ops
// (tode_erk_weighted_sum, tode_time_nodes, tode_interp_eval).
#include "api_common.cuh"
#include "erk_kernels.cuh"

namespace tode {

// vectors per thread: keep >= ~96 bytes of loads in flight per thread (early stages read few rows
// and narrow rows give narrow vectors: with 2 vectors per thread NK = 1, 2 on 8-byte rows reached
// only 78 % / 83 % of the copy bandwidth, profiles/r01_bench_c2.json roofline_kernels)
#ifndef TODE_STAGE_BYTES_IN_FLIGHT
#define TODE_STAGE_BYTES_IN_FLIGHT 96
#endif
template <typename D, int VEC, int NK>
constexpr int stage_unroll() {
  const int per_vec = (NK + 1) * VEC * (int)sizeof(D);
  const int u = (TODE_STAGE_BYTES_IN_FLIGHT + per_vec - 1) / per_vec;
  return u < 2 ? 2 : (u > 8 ? 8 : u);
}

template <typename D, typename T, int VEC, int NK>
static int launch_stage_nk(const tode_tableau* tab, int stage, const tode_state* st,
                           const void* const* k, void* y_out, cudaStream_t stream) {
  constexpr int UNROLL = stage_unroll<D, VEC, NK>();
  StageArgs<D, T, NK> a{};
  for (int j = 0; j < NK; ++j) {
    a.a[j] = (D)tab->a[stage][j];
    a.k[j] = static_cast<const D*>(k[j]);
  }
  a.y = static_cast<const D*>(st->y);
  a.out = static_cast<D*>(y_out);
  a.dt = static_cast<const T*>(st->dt);
  a.running = st->running;
  a.ctl = st->ctl;
  a.n_vec = st->B * st->F / VEC;
  a.F = st->F;
  if (a.n_vec == 0) return 0;
  const long long per_block = (long long)kBlock * UNROLL;
  const unsigned grid = (unsigned)((a.n_vec + per_block - 1) / per_block);
  erk_stage_kernel<D, T, VEC, NK, UNROLL><<<grid, kBlock, 0, stream>>>(a);
  return launch_status();
}

template <typename D, typename T, int VEC>
static int launch_stage_vec(const tode_tableau* tab, int stage, const tode_state* st,
                            const void* const* k, void* y_out, cudaStream_t stream) {
  switch (stage) {
    case 1: return launch_stage_nk<D, T, VEC, 1>(tab, stage, st, k, y_out, stream);
    case 2: return launch_stage_nk<D, T, VEC, 2>(tab, stage, st, k, y_out, stream);
    case 3: return launch_stage_nk<D, T, VEC, 3>(tab, stage, st, k, y_out, stream);
    case 4: return launch_stage_nk<D, T, VEC, 4>(tab, stage, st, k, y_out, stream);
    case 5: return launch_stage_nk<D, T, VEC, 5>(tab, stage, st, k, y_out, stream);
    case 6: return launch_stage_nk<D, T, VEC, 6>(tab, stage, st, k, y_out, stream);
    default: return TODE_EINVAL;
  }
}

template <typename D, typename T>
static int launch_stage(const tode_tableau* tab, int stage, const tode_state* st,
                        const void* const* k, void* y_out, cudaStream_t stream) {
  // widest vector that divides F and for which every operand is aligned
  int vec = geom_vec<D>(st->F);
  auto ok = [&](int v) {
    const size_t a = sizeof(D) * v;
    if (!aligned_to(st->y, a) || !aligned_to(y_out, a)) return false;
    for (int j = 0; j < stage; ++j)
      if (!aligned_to(k[j], a)) return false;
    return true;
  };
  while (vec > 1 && !ok(vec)) vec >>= 1;
  if (sizeof(D) == 4 && vec == 4) return launch_stage_vec<D, T, (sizeof(D) == 4 ? 4 : 2)>(tab, stage, st, k, y_out, stream);
  if (vec == 2) return launch_stage_vec<D, T, 2>(tab, stage, st, k, y_out, stream);
  return launch_stage_vec<D, T, 1>(tab, stage, st, k, y_out, stream);
}

// ---- light protocol kernels -------------------------------------------------------------
template <typename D, typename T>
struct WSumArgs {
  D w[TODE_MAX_STAGES];
  const D* k[TODE_MAX_STAGES];
  const T* dt;
  const D* base;
  D* out;
  long long N, F;
  int S;
};

template <typename D, typename T>
__global__ void __launch_bounds__(kBlock) weighted_sum_kernel(const __grid_constant__ WSumArgs<D, T> A) {
  for (long long e = (long long)blockIdx.x * kBlock + threadIdx.x; e < A.N;
       e += (long long)gridDim.x * kBlock) {
    const D dtD = (D)A.dt[e / A.F];
    D kv[TODE_MAX_STAGES];
    for (int s = 0; s < A.S; ++s) kv[s] = A.k[s][e];
    const D acc = weighted_sum_n<D>(dtD, A.w, kv, A.S);
    A.out[e] = A.base != nullptr ? add(A.base[e], acc) : acc;
  }
}

template <typename D, typename T>
static int launch_wsum(const tode_tableau* tab, int which, int64_t B, int64_t F, const void* dt,
                       const void* const* k, const void* base, void* out, cudaStream_t stream) {
  WSumArgs<D, T> a{};
  a.S = tab->n_stages;
  for (int s = 0; s < a.S; ++s) {
    a.w[s] = (D)(which == TODE_W_B ? tab->b[s] : tab->b_err[s]);
    a.k[s] = static_cast<const D*>(k[s]);
  }
  a.dt = static_cast<const T*>(dt);
  a.base = static_cast<const D*>(base);
  a.out = static_cast<D*>(out);
  a.N = B * F;
  a.F = F;
  if (a.N == 0) return 0;
  weighted_sum_kernel<D, T><<<grid_for(a.N, kBlock, 8), kBlock, 0, stream>>>(a);
  return launch_status();
}

template <typename T>
struct NodesArgs {
  T c[TODE_MAX_STAGES];
  const T* t0;
  const T* dt;
  T* out;
  long long B;
  int S;
};
template <typename T>
__global__ void __launch_bounds__(kBlock) time_nodes_kernel(const __grid_constant__ NodesArgs<T> A) {
  for (long long b = (long long)blockIdx.x * kBlock + threadIdx.x; b < A.B;
       b += (long long)gridDim.x * kBlock) {
    const T t0 = A.t0[b], dt = A.dt[b];
    for (int i = 0; i < A.S; ++i) A.out[(long long)i * A.B + b] = ffma(A.c[i], dt, t0);  // runge_kutta.py:259
  }
}
template <typename T>
static int launch_nodes(const tode_tableau* tab, int64_t B, const void* t0, const void* dt, void* out,
                        cudaStream_t stream) {
  NodesArgs<T> a{};
  a.S = tab->n_stages;
  for (int i = 0; i < a.S; ++i) a.c[i] = (T)tab->c[i];
  a.t0 = static_cast<const T*>(t0);
  a.dt = static_cast<const T*>(dt);
  a.out = static_cast<T*>(out);
  a.B = B;
  if (B == 0) return 0;
  time_nodes_kernel<T><<<grid_for(B, kBlock, 8), kBlock, 0, stream>>>(a);
  return launch_status();
}

template <typename D, typename T>
struct InterpArgs {
  TabP<D, T> tab;
  const T* t0;
  const T* dt;
  const T* t;
  const D* y0;
  const D* y1;
  const D* k[TODE_MAX_STAGES];
  const long long* idx;
  D* out;
  long long N, F;
};
template <typename D, typename T>
__global__ void __launch_bounds__(kBlock) interp_eval_kernel(const __grid_constant__ InterpArgs<D, T> A) {
  constexpr int S = TODE_MAX_STAGES;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < A.N * A.F;
       i += (long long)gridDim.x * kBlock) {
    const long long nq = i / A.F, f = i % A.F;
    const long long b = A.idx[nq];
    const long long e = b * A.F + f;
    const T dt = A.dt[b];
    const D x = interp_x<D, T>(A.t[nq], A.t0[b], dt);
    D ks[S], co[5];
#pragma unroll
    for (int s = 0; s < S; ++s) ks[s] = A.k[s][e];
    interp_coeffs<D, T, S>(A.tab, (D)dt, A.y0[e], A.y1[e], ks, co);
    A.out[i] = horner4<D>(co, x);
  }
}
template <typename D, typename T>
static int launch_interp(const tode_tableau* tab, int64_t F, int64_t N, const void* t0, const void* dt,
                         const void* y0, const void* y1, const void* const* k, const void* t,
                         const int64_t* idx, void* out, cudaStream_t stream) {
  if (tab->n_stages != TODE_MAX_STAGES) return TODE_ENOSUP;
  InterpArgs<D, T> a{};
  a.tab = make_tab<D, T>(tab);
  a.t0 = static_cast<const T*>(t0);
  a.dt = static_cast<const T*>(dt);
  a.t = static_cast<const T*>(t);
  a.y0 = static_cast<const D*>(y0);
  a.y1 = static_cast<const D*>(y1);
  for (int s = 0; s < TODE_MAX_STAGES; ++s) a.k[s] = static_cast<const D*>(k[s]);
  a.idx = reinterpret_cast<const long long*>(idx);
  a.out = static_cast<D*>(out);
  a.N = N;
  a.F = F;
  if (N * F == 0) return 0;
  interp_eval_kernel<D, T><<<grid_for(N * F, kBlock, 8), kBlock, 0, stream>>>(a);
  return launch_status();
}

}  // namespace tode

using namespace tode;

extern "C" int tode_erk_stage(const tode_tableau* tab, int stage, const tode_state* st,
                              const void* const* k, void* y_out, void* stream) {
  if (!tab || !st || !k || !y_out || !st->y || !st->dt) return TODE_EINVAL;
  if (stage < 1 || stage >= tab->n_stages || stage > 6) return TODE_EINVAL;
  for (int j = 0; j < stage; ++j)
    if (!k[j]) return TODE_EINVAL;
#define CALL(D, T) launch_stage<D, T>(tab, stage, st, k, y_out, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

extern "C" int tode_erk_weighted_sum(const tode_tableau* tab, int which, int32_t data_dtype,
                                     int32_t time_dtype, int64_t B, int64_t F, const void* dt,
                                     const void* const* k, const void* base, void* out, void* stream) {
  if (!tab || !dt || !k || !out || (which != TODE_W_B && which != TODE_W_BERR)) return TODE_EINVAL;
  if (tab->n_stages < 1 || tab->n_stages > TODE_MAX_STAGES) return TODE_EINVAL;
#define CALL(D, T) launch_wsum<D, T>(tab, which, B, F, dt, k, base, out, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(data_dtype, time_dtype, CALL);
#undef CALL
}

extern "C" int tode_time_nodes(const tode_tableau* tab, int32_t time_dtype, int64_t B, const void* t0,
                               const void* dt, void* out, void* stream) {
  if (!tab || !t0 || !dt || !out) return TODE_EINVAL;
  if (time_dtype == TODE_F32) return launch_nodes<float>(tab, B, t0, dt, out, static_cast<cudaStream_t>(stream));
  if (time_dtype == TODE_F64) return launch_nodes<double>(tab, B, t0, dt, out, static_cast<cudaStream_t>(stream));
  return TODE_EINVAL;
}

extern "C" int tode_interp_eval(const tode_tableau* tab, int32_t data_dtype, int32_t time_dtype,
                                int64_t B, int64_t F, int64_t N, const void* t0, const void* dt,
                                const void* y0, const void* y1, const void* const* k, const void* t,
                                const int64_t* idx, void* out, void* stream) {
  (void)B;
  if (!tab || !t0 || !dt || !y0 || !y1 || !k || (N > 0 && (!t || !idx || !out))) return TODE_EINVAL;
#define CALL(D, T) launch_interp<D, T>(tab, F, N, t0, dt, y0, y1, k, t, idx, out, static_cast<cudaStream_t>(stream))
  TODE_DISPATCH_DT(data_dtype, time_dtype, CALL);
#undef CALL
}
