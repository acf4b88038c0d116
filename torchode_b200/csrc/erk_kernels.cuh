// Kernels of the stage-wise path ("path A": opaque user vector field).
//
//   erk_stage_kernel   runge_kutta.py:259-263   one launch per stage, 128-bit vectorised
//   erk_finish_kernel  everything after the last stage of an iteration (see header of
//                      include/torchode_b200.h), one launch, no host sync
//   init_a/b kernels   Hairer initial step + state initialisation
//
// Layout: all (B,F) operands are row-major contiguous and 16-byte aligned.  A sample is
// owned by a group of G lanes (G = 1 or 32 here; the canonical reduction geometry of
// erk_math.cuh makes the result independent of that choice), lane l holds the vectors
// l, l+G, ... of the row.  Finished samples (running == 0) are neither read nor written.
#pragma once
#include "erk_math.cuh"

namespace tode {

constexpr int kBlock = 256;
constexpr int kStages = TODE_MAX_STAGES;  // fused finish path is specialised for 7 stages

// Resident CTAs per SM the finish kernel is compiled for (register budget 65536 / (256 n)).
// Measured on B200 (profiles/r01_finish_occupancy.txt): the kernel is latency-bound at low
// occupancy; fp32 thread-per-sample variants want 4 CTAs/SM (64 regs, a few spilled bytes),
// fp64 ones 2 (more would spill the double-precision controller state), warp-per-sample
// streaming variants 3.
#ifndef TODE_FINISH_MINB_F32
#define TODE_FINISH_MINB_F32 4
#endif
#ifndef TODE_FINISH_MINB_F32_V4
#define TODE_FINISH_MINB_F32_V4 4
#endif
#ifndef TODE_FINISH_MINB_F64
#define TODE_FINISH_MINB_F64 2
#endif
#ifndef TODE_FINISH_MINB_G32_F32
#define TODE_FINISH_MINB_G32_F32 4
#endif
#ifndef TODE_FINISH_MINB_G32_F64
#define TODE_FINISH_MINB_G32_F64 3
#endif
template <typename D, int G, int VEC>
constexpr int finish_min_blocks() {
  if (G == 32) return sizeof(D) == 8 ? TODE_FINISH_MINB_G32_F64 : TODE_FINISH_MINB_G32_F32;
  if (sizeof(D) == 8) return TODE_FINISH_MINB_F64;
  return VEC == 4 ? TODE_FINISH_MINB_F32_V4 : TODE_FINISH_MINB_F32;
}

// ---- vector load / store ------------------------------------------------------------
template <typename D, int VEC>
struct VecIO;
template <>
struct VecIO<float, 1> {
  TODE_DEV static void ld(const float* p, float* r) { r[0] = *p; }
  TODE_DEV static void st(float* p, const float* r) { *p = r[0]; }
};
template <>
struct VecIO<float, 2> {
  TODE_DEV static void ld(const float* p, float* r) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    r[0] = v.x; r[1] = v.y;
  }
  TODE_DEV static void st(float* p, const float* r) {
    *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  }
};
template <>
struct VecIO<float, 4> {
  TODE_DEV static void ld(const float* p, float* r) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  TODE_DEV static void st(float* p, const float* r) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  }
};
template <>
struct VecIO<double, 1> {
  TODE_DEV static void ld(const double* p, double* r) { r[0] = *p; }
  TODE_DEV static void st(double* p, const double* r) { *p = r[0]; }
};
template <>
struct VecIO<double, 2> {
  TODE_DEV static void ld(const double* p, double* r) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    r[0] = v.x; r[1] = v.y;
  }
  TODE_DEV static void st(double* p, const double* r) {
    *reinterpret_cast<double2*>(p) = make_double2(r[0], r[1]);
  }
};

// =====================================================================================
// Stage kernel
// =====================================================================================
template <typename D, typename T, int NK>
struct StageArgs {
  D a[NK];
  const D* k[NK];
  const D* y;
  D* out;
  const T* dt;
  const uint8_t* running;  // may be NULL
  const int* ctl;          // may be NULL
  long long n_vec;         // B * F / VEC
  long long F;
};

// one thread = UNROLL vectors of VEC elements; VEC divides F so a vector never straddles rows
template <typename D, typename T, int VEC, int NK, int UNROLL>
__global__ void __launch_bounds__(kBlock) erk_stage_kernel(const __grid_constant__ StageArgs<D, T, NK> A) {
  if (A.ctl != nullptr && A.ctl[TODE_CTL_STOP]) return;
  const long long vec_per_row = A.F / VEC;
  const long long base = ((long long)blockIdx.x * UNROLL) * kBlock + threadIdx.x;
  D kv[UNROLL][NK][VEC];
  D yv[UNROLL][VEC];
  D dtD[UNROLL];
  bool on[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const long long v = base + (long long)u * kBlock;
    on[u] = v < A.n_vec;
    if (on[u]) {
      const long long b = v / vec_per_row;
      if (A.running != nullptr && !A.running[b]) on[u] = false;
      if (on[u]) dtD[u] = (D)A.dt[b];
    }
  }
  // issue every load before the first use (memory-level parallelism)
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (on[u]) {
      const long long e = (base + (long long)u * kBlock) * VEC;
      VecIO<D, VEC>::ld(A.y + e, yv[u]);
#pragma unroll
      for (int j = 0; j < NK; ++j) VecIO<D, VEC>::ld(A.k[j] + e, kv[u][j]);
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (on[u]) {
      D r[VEC];
#pragma unroll
      for (int x = 0; x < VEC; ++x) {
        // einsum("j,jbf->bf", a[i,:i], k[:i]): FMA chain in ascending j, then addcmul(y0, dt, acc)
        D acc = mul(A.a[0], kv[u][0][x]);
#pragma unroll
        for (int j = 1; j < NK; ++j) acc = ffma(A.a[j], kv[u][j][x], acc);
        r[x] = ffma(dtD[u], acc, yv[u][x]);
      }
      VecIO<D, VEC>::st(A.out + (base + (long long)u * kBlock) * VEC, r);
    }
  }
}

// =====================================================================================
// Finish kernel
// =====================================================================================
template <typename D, typename T>
struct FinishArgs {
  TabP<D, T> tab;
  CtrlP<D, T> ctrl;
  long long B, F, Tn;
  const T* t_start;
  const T* t_end;
  const T* t_eval;
  long long te_stride;
  T* t;
  T* dt;
  D* y;
  D* f0;
  D* r1;
  D* r2;
  uint8_t* running;
  int* n_steps;
  int* n_accepted;
  int* status;
  int* cursor;
  uint8_t* not_yet;
  D* y_eval;
  T* t_nodes;
  int* ctl;
  const D* k[kStages];
  const D* y1;
  D sqrt_f;
  D* scratch;  // split (multi-CTA per sample) mode: chunk partials + per-sample step records
  long long scratch_elems;
  uint8_t* flip;  // step-fused route (heat_step.cu): per-sample buffer selector, toggled where a step is accepted
};

// butterfly over the G lanes of a sample group (strides 1, 2, ..., G/2)
template <typename D, int G>
TODE_DEV D group_sum(D v) {
#pragma unroll
  for (int m = 1; m < G; m <<= 1) v = add(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
template <typename D, int G>
TODE_DEV D group_max(D v) {
#pragma unroll
  for (int m = 1; m < G; m <<= 1) v = max_nan_nn(v, __shfl_xor_sync(0xffffffffu, v, m));  // v >= 0 or NaN
  return v;
}

// Per-sample norm over the row in the CANONICAL ORDER:
// the row is cut into chunks of kChunkVec vectors (= 32 vectors per lane of a warp); inside a
// chunk lane l sums the squares of its vectors in ascending order (first a product, then FMAs)
// and the lane partials are combined by an xor-butterfly 1,2,4,...; chunk sums are added in
// ascending chunk order.  Rows of up to kChunkVec vectors are a single chunk.
// add() is called by the owning lane for each of its elements; slot_done() and result() must
// be called by ALL lanes of the warp at the same points (they shuffle).
constexpr int kChunkVec = 1024;
template <typename D, int G>
struct RowNorm {
  D part, total, sqrt_f;
  int cnt, kind;
  bool first, first_chunk;
  TODE_DEV RowNorm(int kind_, D sqrt_f_)
      : part((D)0), total((D)0), sqrt_f(sqrt_f_), cnt(0), kind(kind_), first(true), first_chunk(true) {}
  TODE_DEV void add(D q) {
    if (kind == TODE_NORM_MAX) {
      part = first ? fabs_(q) : max_nan_nn(part, fabs_(q));
      first = false;
    } else {
      sumsq_acc(part, first, fdiv(q, sqrt_f));
    }
  }
  TODE_DEV void flush() {
    const D c = group_sum<D, G>(part);
    total = first_chunk ? c : tode::add(total, c);
    first_chunk = false;
    part = (D)0;
    first = true;
    cnt = 0;
  }
  // one vector slot per lane consumed (uniform across the warp)
  TODE_DEV void slot_done() {
    if (G == 32 && kind != TODE_NORM_MAX && ++cnt == kChunkVec / 32) flush();
  }
  TODE_DEV D result() {
    if (kind == TODE_NORM_MAX) return group_max<D, G>(part);
    if (cnt > 0 || first_chunk) flush();
    return fsqrt(total);
  }
};

// Block-wide accounting of (running samples, any failure) into the control block; the last
// CTA to arrive evaluates `any(running) & all(status == 0)` (adjoints.py:186-190).
TODE_DEV void publish_termination(int* ctl, int my_running, int my_failed) {
  __shared__ int s_run[kBlock / 32];
  __shared__ int s_fail[kBlock / 32];
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    my_running += __shfl_xor_sync(0xffffffffu, my_running, m);
    my_failed |= __shfl_xor_sync(0xffffffffu, my_failed, m);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_run[warp] = my_running;
    s_fail[warp] = my_failed;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, fail = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) {
      run += s_run[w];
      fail |= s_fail[w];
    }
    if (run) atomicAdd(&ctl[TODE_CTL_RUNNING], run);
    if (fail) atomicOr(&ctl[TODE_CTL_FAILED], 1);
    __threadfence();
    const int ticket = atomicAdd(&ctl[TODE_CTL_TICKET], 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const int total_run = atomicAdd(&ctl[TODE_CTL_RUNNING], 0);
      const int total_fail = atomicOr(&ctl[TODE_CTL_FAILED], 0);
      ctl[TODE_CTL_ITERS] += 1;
      if (total_run == 0 || total_fail != 0) ctl[TODE_CTL_STOP] = 1;
      ctl[TODE_CTL_RUNNING] = 0;
      ctl[TODE_CTL_FAILED] = 0;
      ctl[TODE_CTL_TICKET] = 0;
      __threadfence();
    }
  }
}

// Per-sample decision of one loop iteration: controller + commit / running / status logic of
// adjoints.py:150-181.  Shared by the single-launch finish kernel and the split-mode kernels.
template <typename D, typename T>
struct Decision {
  CtrlOut<D, T> o;
  T t_new, dir;
  int status;
  bool upd, running_new;
};

template <typename D, typename T>
TODE_DEV Decision<D, T> decide_step(const CtrlP<D, T>& c, D nrm, T t0, T dt, T ts, T te, D r1, D r2,
                                    int ns, bool act) {
  Decision<D, T> d;
  d.o = controller<D, T>(c, nrm, dt, r1, r2);
  d.upd = d.o.accept && act;            // adjoints.py:150
  d.t_new = d.upd ? add(t0, dt) : t0;   // :151
  d.dir = dir_of(ts, te);
  d.running_new = act && (ffma(d.dir, d.t_new, mul(-d.dir, te)) < (T)0);  // :169
  d.status = d.o.status;                                                  // :171-181
  if (c.max_steps >= 0 && (long long)ns >= c.max_steps) d.status = TODE_REACHED_MAX_STEPS;
  return d;
}

// Writes the per-sample scalars of a running sample after its iteration (one thread per sample).
template <typename D, typename T>
TODE_DEV void store_sample_scalars(const FinishArgs<D, T>& A, long long b, const Decision<D, T>& d, T dt,
                                   T ts, T te, int ns, int cur, int& my_running, int& my_failed) {
  const T t_min = ts < te ? ts : te;  // adjoints.py:66-67
  const T t_max = ts < te ? te : ts;
  T dt_new = d.running_new ? d.o.dt_next : dt;                           // :247
  dt_new = clamp_nan(dt_new, sub(t_min, d.t_new), sub(t_max, d.t_new));  // :251
  A.t[b] = d.t_new;
  A.dt[b] = dt_new;
  A.n_steps[b] = ns;
  if (d.upd) A.n_accepted[b] += 1;
  A.status[b] = d.status;
  A.running[b] = (uint8_t)d.running_new;
  if (A.ctrl.pid && d.running_new) {  // PIDController.merge_states :639-647
    A.r1[b] = d.o.r1;
    A.r2[b] = d.o.r2;
  }
  if (A.Tn > 0 && A.not_yet == nullptr) A.cursor[b] = cur;
  if (A.t_nodes != nullptr) {
#pragma unroll
    for (int i = 1; i < kStages; ++i) A.t_nodes[(long long)i * A.B + b] = ffma(A.tab.c[i], dt_new, d.t_new);
  }
  my_running += d.running_new ? 1 : 0;
  my_failed |= (d.status != TODE_SUCCESS) ? 1 : 0;
}

// G lanes per sample, VEC elements per vector, CI = number of vector chunks per lane whose
// commit operands (y1, k[S-1]) are kept in registers between the reduction pass and the commit
// (chunks beyond CI are re-read; the cache is only ever indexed with compile-time constants).
// NOTE: general (not_yet mask) mode requires G == 1 or G == 32 (one sample per warp); the
// launcher routes 1 < G < 32 problems in mask mode to the G == 32 instantiation (same bits).
template <typename D, typename T, int G, int VEC, int CI>
__global__ void __launch_bounds__(kBlock, finish_min_blocks<D, G, VEC>()) erk_finish_kernel(const __grid_constant__ FinishArgs<D, T> A) {
  if (A.ctl[TODE_CTL_STOP]) return;
  constexpr int S = kStages;
  constexpr int CC = CI > 0 ? CI : 1;
  const int lane = threadIdx.x % G;
  const long long gpb = kBlock / G;  // sample groups per block
  const long long n = A.F / VEC;     // vectors per row
  const long long n_it = (n + G - 1) / G;
  const TabP<D, T>& tab = A.tab;
  const CtrlP<D, T>& c = A.ctrl;
  int my_running = 0, my_failed = 0;

  for (long long base = (long long)blockIdx.x * gpb; base < A.B; base += (long long)gridDim.x * gpb) {
    const long long b = base + threadIdx.x / G;
    const bool act = (b < A.B) && A.running[b];
    if (!__any_sync(0xffffffffu, act)) continue;  // warp-uniform

    // every per-sample scalar is read here, before the group butterfly, and written by
    // lane 0 only after the __syncwarp() at the bottom
    T t0 = (T)0, dt = (T)0, ts = (T)0, te = (T)0;
    D r1 = (D)1, r2 = (D)1;
    int ns = 0, cur = 0;
    if (act) {
      t0 = A.t[b];
      dt = A.dt[b];
      ts = A.t_start[b];
      te = A.t_end[b];
      ns = A.n_steps[b] + 1;  // adjoints.py:161
      if (c.pid) {
        r1 = A.r1[b];
        r2 = A.r2[b];
      }
      if (A.Tn > 0 && A.not_yet == nullptr) cur = A.cursor[b];
    }
    const D dtD = (D)dt;  // runge_kutta.py:247
    const long long row = b * A.F;

    // ---- pass 1: error estimate, bounds, per-sample norm (step_size_controllers.py:394-400)
    // Only the two rows the commit needs (y1, k[S-1]) stay in registers (CI chunks per lane);
    // the dense output re-reads its operands, which were just loaded by the same lane (L1/L2).
    D y1c[CC][VEC], k6c[CC][VEC];
    RowNorm<D, G> rn(c.norm, A.sqrt_f);
    auto load_chunk = [&](long long it, D(&y0v)[VEC], D(&y1v)[VEC], D(&kv)[S][VEC]) {
      const long long e = row + (lane + it * G) * VEC;
      VecIO<D, VEC>::ld(A.y + e, y0v);
      VecIO<D, VEC>::ld(A.y1 + e, y1v);
#pragma unroll
      for (int s = 0; s < S; ++s) VecIO<D, VEC>::ld(A.k[s] + e, kv[s]);
    };
    auto reduce_chunk = [&](const D(&y0v)[VEC], const D(&y1v)[VEC], const D(&kv)[S][VEC]) {
#pragma unroll
      for (int x = 0; x < VEC; ++x) {
        D ks[S];
#pragma unroll
        for (int s = 0; s < S; ++s) ks[s] = kv[s][x];
        const D err = weighted_sum<D, S>(dtD, tab.b_err, ks);  // runge_kutta.py:269
        const D bounds = ffma(c.rtol, max_nan_nn(fabs_(y0v[x]), fabs_(y1v[x])), c.atol);
        rn.add(fdiv(fabs_(err), bounds));
      }
    };
    if (CI > 0) {
#pragma unroll
      for (int ci = 0; ci < CC; ++ci) {
        if (act && lane + (long long)ci * G < n) {
          D y0v[VEC], kv[S][VEC];
          load_chunk(ci, y0v, y1c[ci], kv);
          reduce_chunk(y0v, y1c[ci], kv);
#pragma unroll
          for (int x = 0; x < VEC; ++x) k6c[ci][x] = kv[S - 1][x];
        }
        rn.slot_done();
      }
    }
    for (long long it = CI; it < n_it; ++it) {
      if (act && lane + it * G < n) {
        D y0v[VEC], y1v[VEC], kv[S][VEC];
        load_chunk(it, y0v, y1v, kv);
        reduce_chunk(y0v, y1v, kv);
      }
      rn.slot_done();
    }
    const D nrm = rn.result();

    // ---- controller, commit decision (all lanes of the group, redundantly) --------------
    const Decision<D, T> d = decide_step<D, T>(c, nrm, t0, dt, ts, te, r1, r2, ns, act);
    const bool upd = d.upd, running_new = d.running_new;
    const T t_new = d.t_new, dir = d.dir;
    const int status = d.status;

    // ---- dense output with the data of THIS step (adjoints.py:215-234, 298-301) ---------
    auto eval_chunk = [&](long long it, const D(&y0v)[VEC], const D(&y1v)[VEC],
                          const D(&kv)[S][VEC], D x, D* dst_row) {
      D out[VEC];
#pragma unroll
      for (int xx = 0; xx < VEC; ++xx) {
        D ks[S], co[5];
#pragma unroll
        for (int s = 0; s < S; ++s) ks[s] = kv[s][xx];
        interp_coeffs<D, T, S>(tab, dtD, y0v[xx], y1v[xx], ks, co);
        out[xx] = horner4<D>(co, x);
      }
      VecIO<D, VEC>::st(dst_row + (lane + it * G) * VEC, out);
    };
    auto eval_point = [&](T tq, D* dst_row) {
      const D x = interp_x<D, T>(tq, t0, dt);
      for (long long it = 0; it < n_it; ++it) {
        if (lane + it * G < n) {
          D y0v[VEC], y1v[VEC], kv[S][VEC];
          load_chunk(it, y0v, y1v, kv);
          eval_chunk(it, y0v, y1v, kv, x, dst_row);
        }
      }
    };

    if (act) {
      if (A.Tn == 0) {
        // last effective step of this sample: it finishes or reports a failure
        // ... or the batch is known to stop here (replay after another sample's failure, adjoints.py:298-301)
        if (!running_new || status != TODE_SUCCESS || (c.iter_cap > 0 && (long long)ns >= c.iter_cap))
          eval_point(te, A.y_eval + row);
      } else {
        const T* tev = A.t_eval + b * A.te_stride;
        if (A.not_yet == nullptr) {
          while (cur < A.Tn) {
            const T tq = tev[cur];
            if (!(ffma(dir, t_new, mul(-dir, tq)) >= (T)0)) break;  // :216-223
            eval_point(tq, A.y_eval + (b * A.Tn + cur) * A.F);
            ++cur;
          }
        } else {
          uint8_t* ny = A.not_yet + b * A.Tn;
          for (long long jq = 0; jq < A.Tn; ++jq) {
            const uint8_t pending = ny[jq];
            if (G > 1) __syncwarp();  // one sample per warp: `act` is warp-uniform here
            if (pending) {
              const T tq = tev[jq];
              if (ffma(dir, t_new, mul(-dir, tq)) >= (T)0) {
                eval_point(tq, A.y_eval + (b * A.Tn + jq) * A.F);
                if (lane == 0) ny[jq] = 0;  // :232
              }
            }
          }
        }
      }
    }

    // ---- commit y <- y1, f0 <- k[S-1] where accepted (adjoints.py:152-155, runge_kutta.py:216-224)
    // comes after the dense output, which still reads the old y; a lane only ever touches
    // its own elements
    if (upd) {
      if (CI > 0) {
#pragma unroll
        for (int ci = 0; ci < CC; ++ci)
          if (lane + (long long)ci * G < n) {
            const long long e = row + (lane + (long long)ci * G) * VEC;
            VecIO<D, VEC>::st(A.y + e, y1c[ci]);
            VecIO<D, VEC>::st(A.f0 + e, k6c[ci]);
          }
      }
      for (long long it = CI; it < n_it; ++it) {
        if (lane + it * G < n) {
          const long long e = row + (lane + it * G) * VEC;
          D y1v[VEC], k6v[VEC];
          VecIO<D, VEC>::ld(A.y1 + e, y1v);
          VecIO<D, VEC>::ld(A.k[S - 1] + e, k6v);
          VecIO<D, VEC>::st(A.y + e, y1v);
          VecIO<D, VEC>::st(A.f0 + e, k6v);
        }
      }
    }

    // ---- per-sample scalars (lane 0 of the group) ---------------------------------------
    if (G > 1) __syncwarp();
    if (act && lane == 0) store_sample_scalars<D, T>(A, b, d, dt, ts, te, ns, cur, my_running, my_failed);
  }
  publish_termination(A.ctl, my_running, my_failed);
}

// =====================================================================================
// Initial step size (step_size_controllers.py:431-490 / 776-835) + state initialisation
// =====================================================================================
template <typename D, typename T>
struct InitArgs {
  TabP<D, T> tab;
  CtrlP<D, T> ctrl;
  long long B, F, Tn;
  const T* t_start;
  const T* t_end;
  const T* t_eval;
  long long te_stride;
  T* t;
  T* dt;
  const D* y;
  const D* f0;
  D* r1;
  D* r2;
  uint8_t* running;
  int* n_steps;
  int* n_accepted;
  int* status;
  int* cursor;
  D* y_eval;
  T* t_nodes;
  int* ctl;
  D* scratch;      // [0,B): dt0, [B,2B): d1, then the chunk partials of the split mode
  long long scratch_elems;
  D* y1_out;       // part a
  T* t1_out;       // part a
  const D* f1;     // part b: f(t1, y1), or NULL when the user supplied dt0
  const T* dt0;    // user-supplied initial step (used when f1 == NULL)
  double e_init;   // 1 / order, pre-rounded to the data dtype
  D sqrt_f;
};

// dt0 of part a from the first two norms (:467-471)
template <typename D, typename T>
TODE_DEV D init_dt0(D d0, D d1, T ts, T te) {
  const D dt0 = (d0 < (D)1e-5 || d1 < (D)1e-5) ? (D)1e-6 : fdiv(mul((D)0.01, d0), d1);  // :467-468
  return min_nan(dt0, (D)fabs_(sub(te, ts)));                                             // :471
}

// signed first step from the third norm (:481-490)
template <typename D, typename T>
TODE_DEV T init_dt_from_norm2(const InitArgs<D, T>& A, D nrm2, D dt0, D d1, T dir) {
  D d2 = fdiv(nrm2, dt0);
  // only the Integral copy guards dt0 == 0 (:481 vs :826)
  if (!A.ctrl.pid && dt0 == (D)0) d2 = (D)__longlong_as_double(0x7ff0000000000000LL);
  const D m = max_nan_nn(d1, d2);
  D dt1;
  if (m <= (D)1e-15) {
    dt1 = max_nan_nn((D)1e-6, mul(dt0, (D)1e-3));
  } else {
    // `0.01 / m` is Tensor.__rtruediv__ = m.reciprocal() * 0.01      (:484-488)
    dt1 = det_pow_t(mul(fdiv((D)1, m), (D)0.01), A.e_init);
  }
  return (T)mul((D)dir, min_nan(mul((D)100, dt0), dt1));  // :490
}

// per-sample state of a fresh solve (adjoints.py:59-126), written by one thread per sample
template <typename D, typename T>
TODE_DEV void init_store_sample(const InitArgs<D, T>& A, long long b, T ts, T dt, int cur) {
  A.t[b] = ts;
  A.dt[b] = dt;
  A.running[b] = 1;
  A.n_steps[b] = 0;
  A.n_accepted[b] = 0;
  A.status[b] = 0;
  if (A.ctrl.pid) {  // PIDState.default :542
    A.r1[b] = (D)1;
    A.r2[b] = (D)1;
  }
  if (A.cursor != nullptr) A.cursor[b] = cur;
  if (A.t_nodes != nullptr) {
#pragma unroll
    for (int i = 1; i < kStages; ++i) A.t_nodes[(long long)i * A.B + b] = ffma(A.tab.c[i], dt, ts);
  }
}

// part a: d0, d1, dt0, y1 = y0 + dir*dt0*f0, t1 = t0 + dir*dt0   (:459-479)
template <typename D, typename T, int G, int VEC>
__global__ void __launch_bounds__(kBlock) init_step_a_kernel(const __grid_constant__ InitArgs<D, T> A) {
  const int lane = threadIdx.x % G;
  const long long gpb = kBlock / G;
  const long long n = A.F / VEC;
  const long long n_it = (n + G - 1) / G;
  const CtrlP<D, T>& c = A.ctrl;
  for (long long base = (long long)blockIdx.x * gpb; base < A.B; base += (long long)gridDim.x * gpb) {
    const long long b = base + threadIdx.x / G;
    const bool act = b < A.B;
    const long long row = b * A.F;
    RowNorm<D, G> n0(c.norm, A.sqrt_f), n1(c.norm, A.sqrt_f);
    for (long long it = 0; it < n_it; ++it) {
      const long long j = lane + it * G;
      if (act && j < n) {
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
      n0.slot_done();
      n1.slot_done();
    }
    const D d0 = n0.result();
    const D d1 = n1.result();
    T ts = (T)0, te = (T)0;
    if (act) {
      ts = A.t_start[b];
      te = A.t_end[b];
    }
    const D dt0 = init_dt0<D, T>(d0, d1, ts, te);
    const T dir = dir_of(ts, te);
    const D sdt = mul((D)dir, dt0);
    for (long long it = 0; it < n_it; ++it) {
      const long long j = lane + it * G;
      if (act && j < n) {
        D yv[VEC], fv[VEC], r[VEC];
        VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
        VecIO<D, VEC>::ld(A.f0 + row + j * VEC, fv);
#pragma unroll
        for (int x = 0; x < VEC; ++x) r[x] = ffma(sdt, fv[x], yv[x]);  // :473
        VecIO<D, VEC>::st(A.y1_out + row + j * VEC, r);
      }
    }
    if (act && lane == 0) {
      A.t1_out[b] = ffma(dir, (T)dt0, ts);  // :475
      A.scratch[b] = dt0;
      A.scratch[A.B + b] = d1;
    }
  }
}

// part b: d2, dt1, dt = dir * min(100 dt0, dt1) (:481-490), then adjoints.py:59-126.
// The launcher zeroes the control block before this kernel.
template <typename D, typename T, int G, int VEC>
__global__ void __launch_bounds__(kBlock) init_step_b_kernel(const __grid_constant__ InitArgs<D, T> A) {
  const int lane = threadIdx.x % G;
  const long long gpb = kBlock / G;
  const long long n = A.F / VEC;
  const long long n_it = (n + G - 1) / G;
  const CtrlP<D, T>& c = A.ctrl;
  int nonmono = 0;
  for (long long base = (long long)blockIdx.x * gpb; base < A.B; base += (long long)gridDim.x * gpb) {
    const long long b = base + threadIdx.x / G;
    const bool act = b < A.B;
    const long long row = b * A.F;
    T ts = (T)0, te = (T)0;
    if (act) {
      ts = A.t_start[b];
      te = A.t_end[b];
    }
    const T dir = dir_of(ts, te);
    T dt = (T)0;
    if (A.f1 != nullptr) {
      RowNorm<D, G> n2(c.norm, A.sqrt_f);
      for (long long it = 0; it < n_it; ++it) {
        const long long j = lane + it * G;
        if (act && j < n) {
          D yv[VEC], f0v[VEC], f1v[VEC];
          VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
          VecIO<D, VEC>::ld(A.f0 + row + j * VEC, f0v);
          VecIO<D, VEC>::ld(A.f1 + row + j * VEC, f1v);
#pragma unroll
          for (int x = 0; x < VEC; ++x) {
            const D inv = fdiv((D)1, ffma(c.rtol, fabs_(yv[x]), c.atol));
            n2.add(mul(sub(f1v[x], f0v[x]), inv));
          }
        }
        n2.slot_done();
      }
      const D nrm2 = n2.result();
      D dt0 = (D)0, d1 = (D)0;
      if (act) {
        dt0 = A.scratch[b];
        d1 = A.scratch[A.B + b];
      }
      dt = init_dt_from_norm2<D, T>(A, nrm2, dt0, d1, dir);
    } else if (act) {
      dt = A.dt0[b];
    }
    // ---- adjoints.py:59-126 -------------------------------------------------------------
    const T t_min = ts < te ? ts : te;
    const T t_max = ts < te ? te : ts;
    dt = clamp_nan(dt, sub(t_min, ts), sub(t_max, ts));  // :109
    int cur = 0;
    if (act) {
      if (A.Tn > 0) {
        const T* tev = A.t_eval + b * A.te_stride;
        if (tev[0] == ts) {  // :123-126
          cur = 1;
          for (long long j = lane; j < n; j += G) {
            D yv[VEC];
            VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
            VecIO<D, VEC>::st(A.y_eval + b * A.Tn * A.F + j * VEC, yv);
          }
        }
        for (long long j = 1 + lane; j < A.Tn; j += G)
          if (mul(dir, tev[j]) < mul(dir, tev[j - 1])) nonmono = 1;
      } else {
        for (long long j = lane; j < n; j += G) {
          D yv[VEC];
          VecIO<D, VEC>::ld(A.y + row + j * VEC, yv);
          VecIO<D, VEC>::st(A.y_eval + row + j * VEC, yv);
        }
      }
      if (lane == 0) init_store_sample<D, T>(A, b, ts, dt, cur);
    }
  }
  if (__any_sync(0xffffffffu, nonmono) && (threadIdx.x & 31) == 0) atomicOr(&A.ctl[TODE_CTL_NONMONO], 1);
}

}  // namespace tode
