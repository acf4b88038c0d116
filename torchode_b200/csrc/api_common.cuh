// Host-side helpers shared by the C-ABI translation units: parameter packing (double ->
// data/time dtype, the same roundings ButcherTableau.to / the reference's scalar casts do),
// dtype dispatch and launch geometry.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#include "erk_math.cuh"

namespace tode {

template <typename D, typename T>
inline TabP<D, T> make_tab(const tode_tableau* t) {
  TabP<D, T> p{};
  p.n_stages = t->n_stages;
  p.interp = t->interp;
  p.order = t->order;
  for (int i = 0; i < TODE_MAX_STAGES; ++i) {
    p.b_err[i] = (D)t->b_err[i];
    p.c[i] = (T)t->c[i];
    for (int j = 0; j < TODE_MAX_STAGES; ++j) p.a[i][j] = (D)t->a[i][j];
    for (int r = 0; r < 3; ++r) p.w[r][i] = (D)t->w[r][i];
  }
  return p;
}

// exponent of `tensor ** python_float`: the scalar is rounded to the tensor dtype first
template <typename D>
inline double round_exp(double e) {
  return (double)(D)e;
}

template <typename D, typename T>
inline CtrlP<D, T> make_ctrl(const tode_controller* c) {
  CtrlP<D, T> p{};
  p.atol = (D)c->atol;
  p.rtol = (D)c->rtol;
  p.safety = (D)c->safety;
  p.factor_min = (D)c->factor_min;
  p.factor_max = (D)c->factor_max;
  p.almost_zero = (D)c->almost_zero;
  p.e_ratio = round_exp<D>(c->exp_ratio);
  p.e_prev = round_exp<D>(c->exp_prev);
  p.e_prev2 = round_exp<D>(c->exp_prev2);
  p.dt_min = (T)c->dt_min;
  p.dt_max = (T)c->dt_max;
  p.max_steps = c->max_steps;
  p.iter_cap = c->iter_cap;
  p.norm = c->norm;
  p.pid = c->pid;
  p.has_dt_min = c->has_dt_min;
  p.has_dt_max = c->has_dt_max;
  return p;
}

// the polynomial coefficients of det_log2 / det_exp2 (pow_tables.h) for a kernel parameter block
inline PowTab make_powtab() {
  const unsigned long long la[6] = {TODE_POW_LOG_A1, TODE_POW_LOG_A2, TODE_POW_LOG_A3,
                                    TODE_POW_LOG_A4, TODE_POW_LOG_A5, TODE_POW_LOG_A6};
  const unsigned long long ee[5] = {TODE_POW_EXP_E1, TODE_POW_EXP_E2, TODE_POW_EXP_E3, TODE_POW_EXP_E4,
                                    TODE_POW_EXP_E5};
  PowTab t{};
  static_assert(sizeof(la) == sizeof(t.log_a) && sizeof(ee) == sizeof(t.exp_e), "bit patterns of doubles");
  memcpy(t.log_a, la, sizeof(la));
  memcpy(t.exp_e, ee, sizeof(ee));
  return t;
}

inline int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
    cached = n;
  }
  return cached;
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int launch_status() {
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// persistent-style grid: enough CTAs to fill the machine, never more than the work needs
inline unsigned grid_for(long long work_items, long long items_per_block, int ctas_per_sm) {
  long long need = (work_items + items_per_block - 1) / items_per_block;
  const long long cap = (long long)sm_count() * ctas_per_sm;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (unsigned)need;
}

#define TODE_DISPATCH_DT(dd, td, CALL)                                    \
  do {                                                                    \
    if ((dd) == TODE_F32 && (td) == TODE_F32) return CALL(float, float);  \
    if ((dd) == TODE_F64 && (td) == TODE_F64) return CALL(double, double); \
    if ((dd) == TODE_F32 && (td) == TODE_F64) return CALL(float, double); \
    if ((dd) == TODE_F64 && (td) == TODE_F32) return CALL(double, float); \
    return TODE_EINVAL;                                                   \
  } while (0)

}  // namespace tode
