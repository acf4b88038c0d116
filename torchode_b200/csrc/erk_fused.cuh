// Fused whole-solve kernel ("path B"): AutoDiffAdjoint.solve (adjoints.py:43-311) for a
// built-in analytic vector field in ONE launch.  One thread owns one sample: y, the FSAL
// derivative and the six stage derivatives live in registers for the whole integration,
// every sample advances with its own adaptive dt (no lock-step between samples), HBM is
// touched only for y0 / t_start / t_end / t_eval on the way in and ys / stats on the way out.
// All arithmetic goes through erk_math.cuh, so results are bit-identical to the stage-wise
// path A run with the same field evaluated by PyTorch ops.
#pragma once
#include "erk_math.cuh"

namespace tode {

constexpr int kStagesFused = TODE_MAX_STAGES;  // Dopri5 and Tsit5 both have 7 stages
#ifndef TODE_FUSED_THREADS
#define TODE_FUSED_THREADS 128
#endif
constexpr int kFusedThreads = TODE_FUSED_THREADS;

// ---- built-in fields: one IEEE rounding per op of torchode_b200/fields.py forward -----
template <int FIELD, typename D, int F>
struct Field;

template <typename D, int F>
struct Field<TODE_FIELD_LINEAR, D, F> {
  D rate;
  __device__ explicit Field(const double* p) : rate((D)p[0]) {}
  TODE_DEV void operator()(const D* y, D* out) const {
#pragma unroll
    for (int i = 0; i < F; ++i) out[i] = mul(rate, y[i]);
  }
};

template <typename D>
struct Field<TODE_FIELD_VAN_DER_POL, D, 2> {
  D mu;
  __device__ explicit Field(const double* p) : mu((D)p[0]) {}
  TODE_DEV void operator()(const D* y, D* out) const {
    const D x = y[0], v = y[1];
    // dv = mu * (1 - x * x) * v - x
    const D dv = sub(mul(mul(mu, sub((D)1, mul(x, x))), v), x);
    out[0] = v;
    out[1] = dv;
  }
};

template <typename D>
struct Field<TODE_FIELD_LOTKA_VOLTERRA, D, 2> {
  D alpha, beta, delta, gamma;
  __device__ explicit Field(const double* p)
      : alpha((D)p[0]), beta((D)p[1]), delta((D)p[2]), gamma((D)p[3]) {}
  TODE_DEV void operator()(const D* y, D* out) const {
    const D x = y[0], z = y[1];
    const D xz = mul(x, z);
    out[0] = sub(mul(alpha, x), mul(beta, xz));
    out[1] = sub(mul(delta, xz), mul(gamma, z));
  }
};

template <typename D, int F>
struct Row {
  D v[F];
};

// Error ratio + controller through the checked (branching) functions: the rarely taken second
// opinion of the fused kernel's branch-free step (zero error, non-finite values, ...).
template <typename D, typename T, int F>
__device__ __noinline__ CtrlOut<D, T> error_control_checked(const CtrlP<D, T>& c, Row<D, F> err, Row<D, F> bounds,
                                                            T dt, D r1, D r2, double L1, double L2) {
  D q[F];
#pragma unroll
  for (int f = 0; f < F; ++f) q[f] = fdiv(fabs_(err.v[f]), bounds.v[f]);
  double Lr;
  return controller_l<D, T>(c, row_norm_small<D, F>(q, c.norm), dt, r1, r2, L1, L2, &Lr);
}

template <typename D, typename T>
struct FusedArgs {
  TabP<D, T> tab;
  CtrlP<D, T> ctrl;
  PowTab pow;
  double fp[TODE_MAX_FIELD_PARAMS];
  long long B, Tn;
  const D* y0;
  const T* t_start;
  const T* t_end;
  const T* t_eval;
  long long te_stride;
  const T* dt0;  // NULL -> initial-step heuristic
  D* ys;
  long long* n_steps;
  long long* n_accepted;
  long long* n_initialized;
  long long* status;
  T* t_final;
  T* dt_final;
  int* summary;
  long long iter_cap;  // <= 0: unlimited
  double e_init;       // 1/order pre-rounded to D
  // replicas of the gathered result buffers (tode_solution.peer_*): row peer_row0 + b of each
  int n_peers;
  long long peer_row0;
  D* p_ys[TODE_MAX_PEERS];
  long long* p_n_steps[TODE_MAX_PEERS];
  long long* p_n_accepted[TODE_MAX_PEERS];
  long long* p_n_initialized[TODE_MAX_PEERS];
  long long* p_status[TODE_MAX_PEERS];
};

template <typename D, int F>
TODE_DEV void load_row(const D* p, D* r) {
  if (F == 2) {
    if (sizeof(D) == 4) {
      const float2 v = *reinterpret_cast<const float2*>(p);
      r[0] = (D)v.x; r[1] = (D)v.y;
    } else {
      const double2 v = *reinterpret_cast<const double2*>(p);
      r[0] = (D)v.x; r[1] = (D)v.y;
    }
  } else if (F == 4 && sizeof(D) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    r[0] = (D)v.x; r[1] = (D)v.y; r[2] = (D)v.z; r[3] = (D)v.w;
  } else {
#pragma unroll
    for (int i = 0; i < F; ++i) r[i] = p[i];
  }
}
template <typename D, int F>
TODE_DEV void store_row(D* p, const D* r) {
  if (F == 2) {
    if (sizeof(D) == 4) {
      *reinterpret_cast<float2*>(p) = make_float2((float)r[0], (float)r[1]);
    } else {
      *reinterpret_cast<double2*>(p) = make_double2((double)r[0], (double)r[1]);
    }
  } else if (F == 4 && sizeof(D) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
  } else {
#pragma unroll
    for (int i = 0; i < F; ++i) p[i] = r[i];
  }
}

// MINB = resident CTAs per SM the kernel is compiled for (register budget 65536 / (128 MINB)).
// CK / TE = what this instantiation knows about the problem at compile time (state that is never
// used costs registers, and at 6 CTAs / SM the fp64 kernel has none to spare): CK 0 = integral
// controller, 1 = PID without derivative term (no r2 / L2), 2 = any; TE false = no t_eval.
// <2, true> handles every problem; the launcher picks the tightest instantiation that exists.
template <typename D, typename T, int F, int FIELD, int MINB, int CK, bool TE>
__global__ void __launch_bounds__(kFusedThreads, MINB) solve_fused_kernel(const __grid_constant__ FusedArgs<D, T> A) {
  constexpr int S = kStagesFused;
  const long long Tn = TE ? A.Tn : 0;
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = b < A.B;
  const TabP<D, T>& tab = A.tab;
  const CtrlP<D, T>& c = A.ctrl;
  const Field<FIELD, D, F> field(A.fp);
  const DivBy<D> div_sqrt_f((D)sqrt((double)F));
  __shared__ __align__(16) double s_pow[kPowSharedDoubles];  // tables of det_log2 / det_exp2
  pow_tables_to_shared(s_pow, threadIdx.x, kFusedThreads);
  __syncthreads();
  const PowShared pow_src{A.pow, reinterpret_cast<const double2*>(s_pow)};

  int ns = 0, nacc = 0, status = TODE_SUCCESS, cur = 0, fail_iter = 0x7fffffff, nonmono = 0;
  if (valid) {
    D y[F], k[S][F];  // k[0] is the FSAL slot f(t, y)
    load_row<D, F>(A.y0 + b * F, y);
    const T ts = A.t_start[b], te = A.t_end[b];
    const T dir = dir_of(ts, te);
    const T t_min = ts < te ? ts : te, t_max = ts < te ? te : ts;
    const T* tev = Tn > 0 ? A.t_eval + b * A.te_stride : nullptr;
    const long long row_elems = (Tn > 0 ? Tn : 1) * F;
    D* ye = A.ys + b * row_elems;
    // result row `idx` of this sample: here and into every replica of the gathered buffers
    auto put_row = [&](long long idx, const D* r) {
      store_row<D, F>(ye + idx * F, r);
      for (int p = 0; p < A.n_peers; ++p)
        if (A.p_ys[p] != nullptr) store_row<D, F>(A.p_ys[p] + (A.peer_row0 + b) * row_elems + idx * F, r);
    };
    T t = ts, dt;
    field(y, k[0]);  // controller.init / ExplicitRungeKutta.init: f0 = f(t_start, y0)

    // ---- initial step (step_size_controllers.py:453-490 / :798-835) ---------------------
    if (A.dt0 != nullptr) {
      dt = A.dt0[b];
    } else {
      D inv[F], q[F];
#pragma unroll
      for (int i = 0; i < F; ++i) inv[i] = fdiv((D)1, ffma(c.rtol, fabs_(y[i]), c.atol));
#pragma unroll
      for (int i = 0; i < F; ++i) q[i] = mul(y[i], inv[i]);
      const D d0 = row_norm_small<D, F>(q, c.norm);
#pragma unroll
      for (int i = 0; i < F; ++i) q[i] = mul(k[0][i], inv[i]);
      const D d1 = row_norm_small<D, F>(q, c.norm);
      D dt0 = (d0 < (D)1e-5 || d1 < (D)1e-5) ? (D)1e-6 : fdiv(mul((D)0.01, d0), d1);
      dt0 = min_nan(dt0, (D)fabs_(sub(te, ts)));
      const D sdt = mul((D)dir, dt0);
      D y1[F], f1[F];
#pragma unroll
      for (int i = 0; i < F; ++i) y1[i] = ffma(sdt, k[0][i], y[i]);
      field(y1, f1);
#pragma unroll
      for (int i = 0; i < F; ++i) q[i] = mul(sub(f1[i], k[0][i]), inv[i]);
      D d2 = fdiv(row_norm_small<D, F>(q, c.norm), dt0);
      if (!c.pid && dt0 == (D)0) d2 = (D)__longlong_as_double(0x7ff0000000000000LL);
      const D m = max_nan_nn(d1, d2);
      D dt1;
      if (m <= (D)1e-15) {
        dt1 = max_nan_nn((D)1e-6, mul(dt0, (D)1e-3));
      } else {
        dt1 = det_pow_t(mul(fdiv((D)1, m), (D)0.01), A.e_init);
      }
      dt = (T)mul((D)dir, min_nan(mul((D)100, dt0), dt1));
    }
    dt = clamp_nan(dt, sub(t_min, t), sub(t_max, t));  // adjoints.py:109

    // ---- evaluation exactly at t_start, monotonicity of the t_eval row -------------------
    if (Tn > 0) {
      if (tev[0] == ts) {  // adjoints.py:123-126
        put_row(0, y);
        cur = 1;
      }
      for (long long j = 1; j < Tn; ++j)
        if (mul(dir, tev[j]) < mul(dir, tev[j - 1])) nonmono = 1;
    } else {
      put_row(0, y);  // never hand out uninitialised memory
    }

    D r1 = (D)1, r2 = (D)1;
    double L1 = 0.0, L2 = 0.0;  // log2 of the PID history, carried instead of recomputed (same bits)
    bool running = true;
    // the step counter is 32-bit: limits beyond INT_MAX can never be reached
    const int iter_cap = (A.iter_cap > 0 && A.iter_cap < 0x7fffffffLL) ? (int)A.iter_cap : 0x7fffffff;
    const int max_steps = (c.max_steps >= 0 && c.max_steps < 0x7fffffffLL) ? (int)c.max_steps : 0x7fffffff;
    // ---- the loop (adjoints.py:135-260) ---------------------------------------------------
    while (running && status == TODE_SUCCESS && ns < iter_cap) {
      const D dtD = (D)dt;  // runge_kutta.py:247
      D y1[F];
#pragma unroll
      for (int i = 1; i < S; ++i) {
        // runge_kutta.py:261-263 (FMA chain in ascending j, then addcmul)
        const D* arow = tab.a[i];
#pragma unroll
        for (int f = 0; f < F; ++f) {
          D acc = mul(arow[0], k[0][f]);
#pragma unroll
          for (int j = 1; j < i; ++j) acc = ffma(arow[j], k[j][f], acc);
          y1[f] = ffma(dtD, acc, y[f]);
        }
        field(y1, k[i]);
      }
      // error ratio (runge_kutta.py:269, step_size_controllers.py:394-400) and controller: first
      // without branches (erk_math.cuh "fast path"); if any of its range flags is cleared, once
      // more through the checked functions
      Row<D, F> err, bounds;
#pragma unroll
      for (int f = 0; f < F; ++f) {
        D ks[S];
#pragma unroll
        for (int s = 0; s < S; ++s) ks[s] = k[s][f];
        err.v[f] = weighted_sum<D, S>(dtD, tab.b_err, ks);
        bounds.v[f] = ffma(c.rtol, max_abs_nan(y[f], y1[f]), c.atol);
      }
      bool ok = true;
      D q[F];
      div_chk_n<F, true>(err.v, bounds.v, q, ok);  // |err| / bounds
      D nrm;
      if (c.norm == TODE_NORM_MAX) {
        nrm = fabs_(q[0]);
#pragma unroll
        for (int f = 1; f < F; ++f) nrm = max_nan_nn(nrm, fabs_(q[f]));
      } else {
        D v[F];
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = div_sqrt_f(q[f], ok);
        nrm = fsqrt(row_sumsq_canonical<D, F>(v));
      }
      CtrlOut<D, T> o = controller_fast<D, T, CK>(c, nrm, dt, r1, r2, L1, L2, ok, pow_src);
      if (!ok) o = error_control_checked<D, T, F>(c, err, bounds, dt, r1, r2, L1, L2);
      const bool upd = o.accept;                 // running is true inside the loop
      const T t_new = upd ? add(t, dt) : t;      // adjoints.py:151
      ns += 1;                                   // :161
      nacc += upd ? 1 : 0;                       // :162
      const bool running_new = ffma(dir, t_new, mul(-dir, te)) < (T)0;  // :169
      status = o.status;                         // :171-181
      if (ns >= max_steps) status = TODE_REACHED_MAX_STEPS;

      // ---- dense output (adjoints.py:215-234, 298-301) ------------------------------------
      bool have_co = false;
      D co[F][5];
      auto eval_at = [&](T tq, long long idx) {
        if (!have_co) {
#pragma unroll
          for (int f = 0; f < F; ++f) {
            D ks[S];
#pragma unroll
            for (int s = 0; s < S; ++s) ks[s] = k[s][f];
            interp_coeffs<D, T, S>(tab, dtD, y[f], y1[f], ks, co[f]);
          }
          have_co = true;
        }
        const D x = interp_x<D, T>(tq, t, dt);
        D out[F];
#pragma unroll
        for (int f = 0; f < F; ++f) out[f] = horner4<D>(co[f], x);
        put_row(idx, out);
      };
      if (Tn == 0) {
        // the interpolant of the sample's LAST loop iteration, evaluated at t_end: the
        // iteration in which it finishes, fails, or the batch is cut off (iter_cap)
        if (!running_new || status != TODE_SUCCESS || ns >= iter_cap)
          eval_at(te, 0);
      } else {
        while (cur < Tn) {
          const T tq = tev[cur];
          if (!(ffma(dir, t_new, mul(-dir, tq)) >= (T)0)) break;
          eval_at(tq, cur);
          ++cur;
        }
      }

      // ---- commit (adjoints.py:151-155, runge_kutta.py:216-224) -----------------------------
      if (upd) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          y[f] = y1[f];
          k[0][f] = k[S - 1][f];
        }
      }
      t = t_new;
      T dt_new = running_new ? o.dt_next : dt;                          // :247
      dt = clamp_nan(dt_new, sub(t_min, t_new), sub(t_max, t_new));     // :251
      if (CK >= 1 && c.pid && running_new) {                            // :253-255
        if (o.accept) {
          if (CK >= 2) L2 = L1;
          L1 = o.L_ratio;
        }
        r1 = o.r1;
        if (CK >= 2) r2 = o.r2;
      }
      running = running_new;
      if (status != TODE_SUCCESS) fail_iter = ns;
    }

    A.n_steps[b] = ns;
    A.n_accepted[b] = nacc;
    A.n_initialized[b] = Tn > 0 ? cur : 1;
    A.status[b] = status;
    for (int p = 0; p < A.n_peers; ++p) {
      if (A.p_n_steps[p] == nullptr) continue;
      const long long g = A.peer_row0 + b;
      A.p_n_steps[p][g] = ns;
      A.p_n_accepted[p][g] = nacc;
      A.p_n_initialized[p][g] = Tn > 0 ? cur : 1;
      A.p_status[p][g] = status;
    }
    if (A.t_final != nullptr) A.t_final[b] = t;
    if (A.dt_final != nullptr) A.dt_final[b] = dt;
  }
  // batch summary (pre-set to {0, INT32_MAX, 0} by the launcher): loop iterations of the
  // lock-step reference = max n_steps; first iteration with a failure; non-monotone t_eval
  const int wmax = __reduce_max_sync(0xffffffffu, ns);
  const int wmin = __reduce_min_sync(0xffffffffu, fail_iter);
  const int wnm = __any_sync(0xffffffffu, nonmono);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&A.summary[0], wmax);
    if (wmin != 0x7fffffff) atomicMin(&A.summary[1], wmin);
    if (wnm) atomicOr(&A.summary[2], 1);
  }
}

}  // namespace tode
