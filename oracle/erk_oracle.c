/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not shipped, not on the product path.
 *
 * Plain-C CPU restatement of torchode's batch-parallel adaptive explicit
 * Runge-Kutta solve loop (torchode v1.0.1).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * PARITY PINNING: this restatement is checked against outputs of the real
 * reference (imported in the build container, script tests/golden/make_golden.py)
 * stored under tests/golden/*.npz -- see tests/test_oracle_golden.py.
 *
 * Rounding order.  Every floating-point operation below is written in the
 * order the reference's PyTorch CPU ops were measured to execute (probes in
 * DESIGN.md "Rounding contract"): stage combination = FMA chain in ascending j
 * followed by one FMA with dt; error estimate / interpolation weights = (dt*b_s)
 * formed first, then an un-fused multiply-add chain in ascending s; error
 * bounds, Horner steps, `.add(x, alpha=)` and addcmul = one FMA each.  The file
 * is compiled with -ffp-contract=off so that only the explicit fma() calls fuse.
 * pow() is replaced by the deterministic det_pow below (pure IEEE arithmetic,
 * identical bits on every CPU and on the GPU); it is correctly rounded for
 * float in all but ~1e-8 of cases and within a few ulp for double, which is the
 * accuracy class of the reference's own pow (Sleef u10 on CPU, CUDA pow on GPU).
 *
 * The body is instantiated four times (data dtype x time dtype) from
 * erk_oracle_impl.h.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/torchode_b200.h"

/* ------------------------------------------------------------------ */
/* deterministic log2 / exp2 / pow (double arithmetic, explicit FMAs)  */
/* ------------------------------------------------------------------ */

static inline double orc_bits_to_double(uint64_t u) {
  double d;
  memcpy(&d, &u, 8);
  return d;
}
static inline uint64_t orc_double_to_bits(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
}

/* log2(x) for finite x > 0 */
static double det_log2(double x) {
  int k = 0;
  uint64_t ix = orc_double_to_bits(x);
  if ((ix >> 52) == 0) { /* subnormal: scale by 2^54 (exact) */
    x = x * 18014398509481984.0;
    ix = orc_double_to_bits(x);
    k = -54;
  }
  k += (int)((ix >> 52) & 0x7ff) - 1023;
  double m = orc_bits_to_double((ix & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
  if (m > 1.4142135623730951) {
    m = m * 0.5;
    k += 1;
  }
  const double f = m - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  /* atanh series: log(m) = 2s (1 + z/3 + z^2/5 + ... + z^11/23) */
  double p = 1.0 / 23.0;
  p = fma(p, z, 1.0 / 21.0);
  p = fma(p, z, 1.0 / 19.0);
  p = fma(p, z, 1.0 / 17.0);
  p = fma(p, z, 1.0 / 15.0);
  p = fma(p, z, 1.0 / 13.0);
  p = fma(p, z, 1.0 / 11.0);
  p = fma(p, z, 1.0 / 9.0);
  p = fma(p, z, 1.0 / 7.0);
  p = fma(p, z, 1.0 / 5.0);
  p = fma(p, z, 1.0 / 3.0);
  const double two_s = 2.0 * s;
  const double log_m = fma(two_s * z, p, two_s);
  return fma(log_m, 1.4426950408889634, (double)k);
}

/* 2^z */
static double det_exp2(double z) {
  if (z != z) return z;
  if (z >= 1024.0) return INFINITY;
  if (z <= -1100.0) return 0.0;
  const double n = floor(z + 0.5);
  const double f = z - n; /* exact, |f| <= 0.5 */
  const double u = f * 0.6931471805599453;
  /* e^u, Taylor to degree 14 (|u| <= 0.3466: remainder < 2^-60) */
  double p = 1.0 / 87178291200.0;
  p = fma(p, u, 1.0 / 6227020800.0);
  p = fma(p, u, 1.0 / 479001600.0);
  p = fma(p, u, 1.0 / 39916800.0);
  p = fma(p, u, 1.0 / 3628800.0);
  p = fma(p, u, 1.0 / 362880.0);
  p = fma(p, u, 1.0 / 40320.0);
  p = fma(p, u, 1.0 / 5040.0);
  p = fma(p, u, 1.0 / 720.0);
  p = fma(p, u, 1.0 / 120.0);
  p = fma(p, u, 1.0 / 24.0);
  p = fma(p, u, 1.0 / 6.0);
  p = fma(p, u, 0.5);
  p = fma(p, u, 1.0);
  p = fma(p, u, 1.0);
  int e = (int)n;
  /* scale by 2^e in (at most) two exact steps, last one may round to subnormal */
  if (e < -1000) {
    p = p * orc_bits_to_double((uint64_t)(1023 - 600) << 52);
    e += 600;
  }
  return p * orc_bits_to_double((uint64_t)(1023 + e) << 52);
}

/* x^e with IEEE pow special cases for the inputs that occur (x >= 0 or NaN) */
static double det_pow(double x, double e) {
  if (e == 0.0) return 1.0;
  if (x != x || e != e) return x + e;
  if (x == 1.0) return 1.0;
  if (x == 0.0) return e < 0.0 ? INFINITY : 0.0;
  if (x < 0.0) return NAN;
  if (isinf(x)) return e < 0.0 ? 0.0 : INFINITY;
  return det_exp2(e * det_log2(x));
}

static inline float det_pow_f32(float x, double e) {
  /* torch: float tensor ** python scalar -> exponent rounded to float first */
  return (float)det_pow((double)x, (double)(float)e);
}
static inline double det_pow_f64(double x, double e) { return det_pow(x, e); }

/* fma overloads */
static inline float orc_fma_f32(float a, float b, float c) { return fmaf(a, b, c); }
static inline double orc_fma_f64(double a, double b, double c) { return fma(a, b, c); }
static inline float orc_sqrt_f32(float a) { return sqrtf(a); }
static inline double orc_sqrt_f64(double a) { return sqrt(a); }
static inline float orc_abs_f32(float a) { return fabsf(a); }
static inline double orc_abs_f64(double a) { return fabs(a); }

int orc_abi_version(void) { return TODE_ABI_VERSION; }

/* exported for the unit tests of the deterministic math */
double orc_det_pow_f64(double x, double e) { return det_pow_f64(x, e); }
float orc_det_pow_f32(float x, double e) { return det_pow_f32(x, e); }
double orc_det_log2(double x) { return det_log2(x); }
double orc_det_exp2(double x) { return det_exp2(x); }

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)

/* ---- f32 data, f32 time ---- */
#define Dt float
#define Tt float
#define DSUF f32
#define TSUF f32
#define SUF _f32_f32
#include "erk_oracle_impl.h"
#undef Dt
#undef Tt
#undef DSUF
#undef TSUF
#undef SUF

/* ---- f64 data, f64 time ---- */
#define Dt double
#define Tt double
#define DSUF f64
#define TSUF f64
#define SUF _f64_f64
#include "erk_oracle_impl.h"
#undef Dt
#undef Tt
#undef DSUF
#undef TSUF
#undef SUF

/* ---- f32 data, f64 time ---- */
#define Dt float
#define Tt double
#define DSUF f32
#define TSUF f64
#define SUF _f32_f64
#include "erk_oracle_impl.h"
#undef Dt
#undef Tt
#undef DSUF
#undef TSUF
#undef SUF

/* ---- f64 data, f32 time ---- */
#define Dt double
#define Tt float
#define DSUF f64
#define TSUF f32
#define SUF _f64_f32
#include "erk_oracle_impl.h"
#undef Dt
#undef Tt
#undef DSUF
#undef TSUF
#undef SUF

/* ------------------------------------------------------------------ */
/* dtype dispatch (same signatures as the product's C-ABI, host memory) */
/* ------------------------------------------------------------------ */

#define ORC_DISPATCH(dd, td, call)                                     \
  do {                                                                 \
    if ((dd) == TODE_F32 && (td) == TODE_F32) return call(_f32_f32);   \
    if ((dd) == TODE_F64 && (td) == TODE_F64) return call(_f64_f64);   \
    if ((dd) == TODE_F32 && (td) == TODE_F64) return call(_f32_f64);   \
    if ((dd) == TODE_F64 && (td) == TODE_F32) return call(_f64_f32);   \
    return TODE_EINVAL;                                                \
  } while (0)

int orc_erk_stage(const tode_tableau* tab, int stage, const tode_state* st,
                  const void* const* k, void* y_out) {
#define CALL(S) ORC_CAT(orc_erk_stage, S)(tab, stage, st, k, y_out)
  ORC_DISPATCH(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

int orc_erk_finish(const tode_tableau* tab, const tode_controller* ctrl,
                   const tode_state* st, const void* const* k, const void* y1) {
#define CALL(S) ORC_CAT(orc_erk_finish, S)(tab, ctrl, st, k, y1)
  ORC_DISPATCH(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

int orc_init_step_a(const tode_tableau* tab, const tode_controller* ctrl,
                    const tode_state* st, void* y1_out, void* t1_out) {
#define CALL(S) ORC_CAT(orc_init_step_a, S)(tab, ctrl, st, y1_out, t1_out)
  ORC_DISPATCH(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

int orc_init_step_b(const tode_tableau* tab, const tode_controller* ctrl,
                    const tode_state* st, const void* f1) {
#define CALL(S) ORC_CAT(orc_init_step_b, S)(tab, ctrl, st, f1)
  ORC_DISPATCH(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

int orc_init_with_dt0(const tode_tableau* tab, const tode_controller* ctrl,
                      const tode_state* st, const void* dt0) {
#define CALL(S) ORC_CAT(orc_init_with_dt0, S)(tab, ctrl, st, dt0)
  ORC_DISPATCH(st->data_dtype, st->time_dtype, CALL);
#undef CALL
}

int orc_solve_builtin(int field, const double* field_params, const tode_tableau* tab,
                      const tode_controller* ctrl, const tode_problem* prob,
                      const tode_solution* sol, int64_t iter_cap) {
#define CALL(S) ORC_CAT(orc_solve_builtin, S)(field, field_params, tab, ctrl, prob, sol, iter_cap)
  ORC_DISPATCH(prob->data_dtype, prob->time_dtype, CALL);
#undef CALL
}

int orc_erk_weighted_sum(const tode_tableau* tab, int which, int32_t data_dtype, int32_t time_dtype,
                         int64_t B, int64_t F, const void* dt, const void* const* k,
                         const void* base, void* out) {
#define CALL(S) ORC_CAT(orc_erk_weighted_sum, S)(tab, which, B, F, dt, k, base, out)
  ORC_DISPATCH(data_dtype, time_dtype, CALL);
#undef CALL
}

int orc_adapt_step_size(const tode_controller* ctrl, int32_t data_dtype, int32_t time_dtype,
                        int64_t B, int64_t F, const void* dt, const void* y0, const void* y1,
                        const void* err, const void* r1, const void* r2, uint8_t* accept_out,
                        void* dt_next_out, void* ratio_out, void* r1_out, void* r2_out,
                        int64_t* status_out) {
#define CALL(S)                                                                              \
  ORC_CAT(orc_adapt_step_size, S)(ctrl, B, F, dt, y0, y1, err, r1, r2, accept_out,           \
                                  dt_next_out, ratio_out, r1_out, r2_out, status_out)
  ORC_DISPATCH(data_dtype, time_dtype, CALL);
#undef CALL
}

int orc_interp_eval(const tode_tableau* tab, int32_t data_dtype, int32_t time_dtype, int64_t B,
                    int64_t F, int64_t N, const void* t0, const void* dt, const void* y0,
                    const void* y1, const void* const* k, const void* t, const int64_t* idx,
                    void* out) {
#define CALL(S) ORC_CAT(orc_interp_eval, S)(tab, B, F, N, t0, dt, y0, y1, k, t, idx, out)
  ORC_DISPATCH(data_dtype, time_dtype, CALL);
#undef CALL
}
