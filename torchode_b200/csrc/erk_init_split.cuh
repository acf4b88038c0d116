// Split-mode initial step for few samples with a huge feature dimension (config C5: 64 x 2^20),
// the counterpart of erk_finish_split.cuh: with one warp per SAMPLE the two init kernels leave all
// but B warps of the machine idle (measured: 17.7 + 14.7 ms at 64 x 2^20).  Here a row is cut into the
// chunks of the canonical reduction order, one WARP per (sample, chunk):
//
//   part a   init_split_a_partial_kernel   per-chunk partials of d0 = |y0 * inv|, d1 = |f0 * inv|
//            init_split_a_finish_kernel    every warp re-adds its sample's partials in ascending chunk
//                                          order (dt0), then writes its chunk of y1 = y0 + dir dt0 f0
//   part b   init_split_b_partial_kernel   per-chunk partials of |(f1 - f0) * inv|, y_eval seed copy
//            init_split_b_finish_kernel    one thread per sample: partials in ascending order, first
//                                          step, per-sample state (adjoints.py:59-126)
//
// Same arithmetic, same canonical order -> same bits as init_step_a/b_kernel (G = 32).
#pragma once
#include "erk_kernels.cuh"

namespace tode {

// partials live behind the two per-sample slots of part a: [2B, 2B + B*cpr) and [2B + B*cpr, 2B + 2*B*cpr)
template <typename D, typename T>
TODE_DEV D* init_partials(const InitArgs<D, T>& A, int which, long long cpr) {
  return A.scratch + 2 * A.B + (long long)which * A.B * cpr;
}
inline long long init_split_scratch_need(long long B, long long cpr) { return 2 * B + 2 * B * cpr; }

// chunk partial of one norm in the canonical order: lane partials ascending, xor-butterfly
template <typename D>
struct ChunkNorm {
  D part;
  bool first;
  int kind;
  D sqrt_f;
  TODE_DEV ChunkNorm(int kind_, D sqrt_f_) : part((D)0), first(true), kind(kind_), sqrt_f(sqrt_f_) {}
  TODE_DEV void add(D q) {
    if (kind == TODE_NORM_MAX) {
      part = first ? fabs_(q) : max_nan_nn(part, fabs_(q));
      first = false;
    } else {
      sumsq_acc(part, first, fdiv(q, sqrt_f));
    }
  }
  TODE_DEV D reduce() const { return kind == TODE_NORM_MAX ? group_max<D, 32>(part) : group_sum<D, 32>(part); }
};

// chunk partials of a sample combined in ascending chunk order (RowNorm::flush / result)
template <typename D>
TODE_DEV D combine_partials(const D* p, long long cpr, int kind) {
  D total = p[0];
  if (kind == TODE_NORM_MAX) {
    for (long long ch = 1; ch < cpr; ++ch) total = max_nan_nn(total, p[ch]);
    return total;
  }
  for (long long ch = 1; ch < cpr; ++ch) total = add(total, p[ch]);
  return fsqrt(total);
}

template <typename D, typename T, int VEC>
__global__ void __launch_bounds__(kBlock) init_split_a_partial_kernel(const __grid_constant__ InitArgs<D, T> A) {
  const int lane = threadIdx.x & 31;
  const long long n = A.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long warps_total = A.B * cpr;
  const CtrlP<D, T>& c = A.ctrl;
  D* p0 = init_partials(A, 0, cpr);
  D* p1 = init_partials(A, 1, cpr);
  for (long long w = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5; w < warps_total;
       w += ((long long)gridDim.x * kBlock) >> 5) {
    const long long b = w / cpr, ch = w % cpr;
    const long long row = b * A.F;
    ChunkNorm<D> n0(c.norm, A.sqrt_f), n1(c.norm, A.sqrt_f);
#pragma unroll 4
    for (int i = 0; i < kChunkVec / 32; ++i) {
      const long long j = ch * kChunkVec + lane + 32LL * i;
      if (j < n) {
        D yv[VEC], fv[VEC];
        VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
        VecIO<D, VEC>::ld(A.f0 + row + j * VEC, fv);
#pragma unroll
        for (int x = 0; x < VEC; ++x) {
          const D inv = fdiv((D)1, ffma(c.rtol, fabs_(yv[x]), c.atol));  // :461-462
          n0.add(mul(yv[x], inv));                                       // :464
          n1.add(mul(fv[x], inv));                                       // :465
        }
      }
    }
    const D r0 = n0.reduce(), r1 = n1.reduce();
    if (lane == 0) {
      p0[w] = r0;
      p1[w] = r1;
    }
  }
}

template <typename D, typename T, int VEC>
__global__ void __launch_bounds__(kBlock) init_split_a_finish_kernel(const __grid_constant__ InitArgs<D, T> A) {
  const int lane = threadIdx.x & 31;
  const long long n = A.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long warps_total = A.B * cpr;
  const CtrlP<D, T>& c = A.ctrl;
  const D* p0 = init_partials(A, 0, cpr);
  const D* p1 = init_partials(A, 1, cpr);
  for (long long w = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5; w < warps_total;
       w += ((long long)gridDim.x * kBlock) >> 5) {
    const long long b = w / cpr, ch = w % cpr;
    const long long row = b * A.F;
    const D d0 = combine_partials<D>(p0 + b * cpr, cpr, c.norm);
    const D d1 = combine_partials<D>(p1 + b * cpr, cpr, c.norm);
    const T ts = A.t_start[b], te = A.t_end[b];
    const D dt0 = init_dt0<D, T>(d0, d1, ts, te);
    const T dir = dir_of(ts, te);
    const D sdt = mul((D)dir, dt0);
#pragma unroll 4
    for (int i = 0; i < kChunkVec / 32; ++i) {
      const long long j = ch * kChunkVec + lane + 32LL * i;
      if (j < n) {
        D yv[VEC], fv[VEC], r[VEC];
        VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
        VecIO<D, VEC>::ld(A.f0 + row + j * VEC, fv);
#pragma unroll
        for (int x = 0; x < VEC; ++x) r[x] = ffma(sdt, fv[x], yv[x]);  // :473
        VecIO<D, VEC>::st(A.y1_out + row + j * VEC, r);
      }
    }
    if (ch == 0 && lane == 0) {
      A.t1_out[b] = ffma(dir, (T)dt0, ts);  // :475
      A.scratch[b] = dt0;
      A.scratch[A.B + b] = d1;
    }
  }
}

// f1 == NULL (user-supplied dt0): only the y_eval seed copy
template <typename D, typename T, int VEC>
__global__ void __launch_bounds__(kBlock) init_split_b_partial_kernel(const __grid_constant__ InitArgs<D, T> A) {
  const int lane = threadIdx.x & 31;
  const long long n = A.F / VEC;
  const long long cpr = (n + kChunkVec - 1) / kChunkVec;
  const long long warps_total = A.B * cpr;
  const CtrlP<D, T>& c = A.ctrl;
  D* p2 = init_partials(A, 0, cpr);
  for (long long w = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5; w < warps_total;
       w += ((long long)gridDim.x * kBlock) >> 5) {
    const long long b = w / cpr, ch = w % cpr;
    const long long row = b * A.F;
    // adjoints.py:123-126 (first t_eval point at t_start) / the end value of a solve that never steps
    const bool seed = A.Tn == 0 || A.t_eval[b * A.te_stride] == A.t_start[b];
    D* dst = A.y_eval + (A.Tn == 0 ? row : b * A.Tn * A.F);
    ChunkNorm<D> n2(c.norm, A.sqrt_f);
#pragma unroll 4
    for (int i = 0; i < kChunkVec / 32; ++i) {
      const long long j = ch * kChunkVec + lane + 32LL * i;
      if (j < n) {
        D yv[VEC];
        VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
        if (A.f1 != nullptr) {
          D f0v[VEC], f1v[VEC];
          VecIO<D, VEC>::ld(A.f0 + row + j * VEC, f0v);
          VecIO<D, VEC>::ld(A.f1 + row + j * VEC, f1v);
#pragma unroll
          for (int x = 0; x < VEC; ++x) {
            const D inv = fdiv((D)1, ffma(c.rtol, fabs_(yv[x]), c.atol));
            n2.add(mul(sub(f1v[x], f0v[x]), inv));
          }
        }
        if (seed) VecIO<D, VEC>::st(dst + j * VEC, yv);
      }
    }
    if (A.f1 != nullptr) {
      const D r = n2.reduce();
      if (lane == 0) p2[w] = r;
    }
  }
}

template <typename D, typename T>
__global__ void __launch_bounds__(kBlock) init_split_b_finish_kernel(const __grid_constant__ InitArgs<D, T> A,
                                                                       long long cpr) {
  const CtrlP<D, T>& c = A.ctrl;
  const D* p2 = init_partials(A, 0, cpr);
  int nonmono = 0;
  for (long long b = (long long)blockIdx.x * kBlock + threadIdx.x; b < A.B; b += (long long)gridDim.x * kBlock) {
    const T ts = A.t_start[b], te = A.t_end[b];
    const T dir = dir_of(ts, te);
    T dt;
    if (A.f1 != nullptr) {
      const D nrm2 = combine_partials<D>(p2 + b * cpr, cpr, c.norm);
      dt = init_dt_from_norm2<D, T>(A, nrm2, A.scratch[b], A.scratch[A.B + b], dir);
    } else {
      dt = A.dt0[b];
    }
    const T t_min = ts < te ? ts : te;
    const T t_max = ts < te ? te : ts;
    dt = clamp_nan(dt, sub(t_min, ts), sub(t_max, ts));  // adjoints.py:109
    int cur = 0;
    if (A.Tn > 0) {
      const T* tev = A.t_eval + b * A.te_stride;
      if (tev[0] == ts) cur = 1;  // :123-126 (the row itself was copied by the partial kernel)
      for (long long j = 1; j < A.Tn; ++j)
        if (mul(dir, tev[j]) < mul(dir, tev[j - 1])) nonmono = 1;
    }
    init_store_sample<D, T>(A, b, ts, dt, cur);
  }
  if (__any_sync(0xffffffffu, nonmono) && (threadIdx.x & 31) == 0) atomicOr(&A.ctl[TODE_CTL_NONMONO], 1);
}

}  // namespace tode
