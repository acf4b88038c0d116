/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see erk_oracle.c).
 *
 * Body of the CPU restatement, instantiated per (data dtype Dt, time dtype Tt).
 * Every function cites the reference lines it follows (torchode v1.0.1,
 * paths relative to torchode/).
 */

#define FN(name) ORC_CAT(name, SUF)
#define DFMA ORC_CAT(orc_fma_, DSUF)
#define TFMA ORC_CAT(orc_fma_, TSUF)
#define DSQRT ORC_CAT(orc_sqrt_, DSUF)
#define DABS ORC_CAT(orc_abs_, DSUF)
#define TABS ORC_CAT(orc_abs_, TSUF)
#define DPOW ORC_CAT(det_pow_, DSUF)

/* torch.maximum / torch.minimum: NaN-propagating (ATen BinaryOpsKernel) */
static inline Dt FN(dmax)(Dt a, Dt b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
static inline Dt FN(dmin)(Dt a, Dt b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
/* torch.clamp(x, lo, hi): NaN in x propagates, otherwise min(max(x, lo), hi) */
static inline Dt FN(dclamp)(Dt x, Dt lo, Dt hi) {
  if (x != x) return x;
  Dt r = x < lo ? lo : x;
  return r > hi ? hi : r;
}
static inline Tt FN(tclamp)(Tt x, Tt lo, Tt hi) {
  if (x != x) return x;
  Tt r = x < lo ? lo : x;
  return r > hi ? hi : r;
}

/* problems.py:42  time_direction = where(t_end > t_start, 1, -1) */
static inline Tt FN(dir_of)(Tt t_start, Tt t_end) { return t_end > t_start ? (Tt)1 : (Tt)-1; }

/* step_size_controllers.py:170-186: rms_norm / max_norm over the feature row `q` (n = F).
 * max: NaN-propagating max |x_i|.  rms: sqrt(sum_i (x_i / sqrt(F))^2) with the squares
 * summed in the CANONICAL ORDER that the kernels' lane geometry defines:
 *   VEC = widest <=16-byte vector (in elements) dividing F, N = F / VEC vectors,
 *   G = min(32, next_pow2(N)) lanes; the row is cut into chunks of 1024 vectors; inside a
 *   chunk lane l accumulates vectors l, l+G, ... in ascending order (first square a plain
 *   product, later ones FMAs) and the G partials are combined by an xor-butterfly with
 *   strides 1, 2, ..., G/2; chunk sums are added in ascending chunk order.
 * For F <= 2 (and F == 4 in fp32) this is the plain sequential FMA chain. */
static inline Dt FN(row_norm)(const Dt* q, int64_t n, int norm_kind) {
  if (norm_kind == TODE_NORM_MAX) {
    Dt m = DABS(q[0]);
    for (int64_t i = 1; i < n; ++i) m = FN(dmax)(m, DABS(q[i]));
    return m;
  }
  const Dt sqrt_f = (Dt)sqrt((double)n);
  const int VEC = sizeof(Dt) == 4 ? ((n % 4 == 0) ? 4 : ((n % 2 == 0) ? 2 : 1)) : ((n % 2 == 0) ? 2 : 1);
  const int64_t N = n / VEC;
  int G = 1;
  while (G < 32 && G < N) G <<= 1;
  const int64_t CHUNK = 1024; /* vectors per chunk = 32 per lane of a warp */
  Dt total = (Dt)0;
  for (int64_t c0 = 0; c0 < N; c0 += CHUNK) {
    const int64_t c1 = c0 + CHUNK < N ? c0 + CHUNK : N;
    Dt part[32], nxt[32];
    for (int l = 0; l < G; ++l) {
      Dt s = (Dt)0;
      int first = 1;
      for (int64_t j = c0 + l; j < c1; j += G) {
        for (int u = 0; u < VEC; ++u) {
          const Dt v = q[j * VEC + u] / sqrt_f;
          if (first) {
            s = v * v;
            first = 0;
          } else {
            s = DFMA(v, v, s);
          }
        }
      }
      part[l] = s;
    }
    for (int m = 1; m < G; m <<= 1) {
      for (int l = 0; l < G; ++l) nxt[l] = part[l] + part[l ^ m];
      for (int l = 0; l < G; ++l) part[l] = nxt[l];
    }
    total = c0 == 0 ? part[0] : total + part[0];
  }
  return DSQRT(total);
}

typedef struct FN(ctrl_out) {
  int accept;
  Tt dt_next;
  Dt ratio;
  int status;
  Dt r1, r2;
} FN(ctrl_out);

/* step_size_controllers.py:400-429 (Integral) / :745-774 (PID), dt_factor
 * :289-294 / :598-620, update_state :649-671.  `nrm` = norm(|err|/bounds). */
static inline FN(ctrl_out) FN(controller)(const tode_controller* c, Dt nrm, Tt dt, Dt r1, Dt r2) {
  FN(ctrl_out) o;
  const Dt ratio = FN(dmax)(nrm, (Dt)c->almost_zero); /* :400 */
  o.ratio = ratio;
  o.accept = ratio < (Dt)1; /* :401 */
  Dt factor = (Dt)c->safety * DPOW(ratio, c->exp_ratio);
  if (c->pid) { /* :615-618: safety * factor1 * factor2 * factor3, left to right */
    factor = factor * DPOW(r1, c->exp_prev);
    factor = factor * DPOW(r2, c->exp_prev2);
  }
  factor = FN(dclamp)(factor, (Dt)c->factor_min, (Dt)c->factor_max); /* :294 / :620 */
  Tt dt_next = dt * (Tt)factor;                                      /* :404 */
  /* :407-411  isfinite(ratio) ? SUCCESS : INFINITE_NORM */
  int status = (ratio - ratio == (Dt)0) ? TODE_SUCCESS : TODE_INFINITE_NORM;
  if (c->has_dt_min || c->has_dt_max) { /* :414-422 */
    const Tt a = TABS(dt_next);
    Tt cl = a;
    if (a == a) {
      if (c->has_dt_min && cl < (Tt)c->dt_min) cl = (Tt)c->dt_min;
      if (c->has_dt_max && cl > (Tt)c->dt_max) cl = (Tt)c->dt_max;
    }
    const Tt sign = (Tt)((dt_next > (Tt)0) - (dt_next < (Tt)0)); /* torch.sign: 0 for NaN */
    dt_next = sign * cl;
    if (c->has_dt_min && a < (Tt)c->dt_min) status = TODE_REACHED_DT_MIN;
  }
  o.dt_next = dt_next;
  o.status = status;
  /* :664-671: history shifts only where the step was accepted */
  o.r1 = o.accept ? ratio : r1;
  o.r2 = o.accept ? r1 : r2;
  return o;
}

/* runge_kutta.py:269  einsum("b,s,sbf->bf", dt, b_err, k): (dt*w_s) first,
 * then un-fused multiply-add chain in ascending s (bmm on (B,1,S)x(B,S,F)). */
static inline Dt FN(weighted_sum)(Dt dtD, const double* w, int n_stages, const Dt* const* k, int64_t e) {
  Dt acc = (dtD * (Dt)w[0]) * k[0][e];
  for (int s = 1; s < n_stages; ++s) acc = acc + (dtD * (Dt)w[s]) * k[s][e];
  return acc;
}

/* Quartic coefficients of the local interpolant for one element.
 * dopri5.py:54-60 + interpolation.py:139-170, tsit5.py:124-139. */
static inline void FN(interp_coeffs)(const tode_tableau* tab, Dt dtD, Dt y0, Dt y1,
                                     const Dt* const* k, int64_t e, Dt* co /* a,b,c,d,e */) {
  const int S = tab->n_stages;
  if (tab->interp == TODE_INTERP_DOPRI5) {
    const Dt f0 = dtD * k[0][e];     /* :148 */
    const Dt f1 = dtD * k[S - 1][e]; /* :149 */
    const Dt ymid = y0 + FN(weighted_sum)(dtD, tab->w[0], S, k, e); /* :150 */
    /* :152  (2*(f1-f0)).add(y1+y0, alpha=-8).add(y_mid, alpha=16) */
    Dt a = (Dt)2 * (f1 - f0);
    a = DFMA((Dt)-8, y1 + y0, a);
    a = DFMA((Dt)16, ymid, a);
    /* :153-159 */
    Dt b = (Dt)5 * f0;
    b = DFMA((Dt)-3, f1, b);
    b = DFMA((Dt)18, y0, b);
    b = DFMA((Dt)14, y1, b);
    b = DFMA((Dt)-32, ymid, b);
    /* :160-165 */
    Dt c = DFMA((Dt)-4, f0, f1);
    c = DFMA((Dt)-11, y0, c);
    c = DFMA((Dt)-5, y1, c);
    c = DFMA((Dt)16, ymid, c);
    co[0] = a;
    co[1] = b;
    co[2] = c;
    co[3] = f0;
    co[4] = y0;
  } else {
    /* tsit5.py:132-135: B = einsum("b,cs,sbf->cbf", dt, b_other, k); c,b,a = B[0..2] */
    co[2] = FN(weighted_sum)(dtD, tab->w[0], S, k, e);
    co[1] = FN(weighted_sum)(dtD, tab->w[1], S, k, e);
    co[0] = FN(weighted_sum)(dtD, tab->w[2], S, k, e);
    co[3] = dtD * k[0][e];
    co[4] = y0;
  }
}

/* interpolation.py:25-40 poly4eval: x = (t - t0) / (t1 - t0), t1 = t0 + dt */
static inline Dt FN(interp_x)(Tt t, Tt t0, Tt dt) {
  const Tt t1 = t0 + dt;
  Tt h = t1 - t0;
  if (!(TABS(h) > (Tt)0)) h = (Tt)1; /* where(dt.abs() > 0, dt, 1) */
  return (Dt)((t - t0) / h);
}
static inline Dt FN(horner4)(const Dt* co, Dt x) {
  Dt y = co[0];
  y = DFMA(y, x, co[1]);
  y = DFMA(y, x, co[2]);
  y = DFMA(y, x, co[3]);
  y = DFMA(y, x, co[4]);
  return y;
}

/* ------------------------------------------------------------------ */
/* runge_kutta.py:259-263                                              */
/* ------------------------------------------------------------------ */
static int FN(orc_erk_stage)(const tode_tableau* tab, int stage, const tode_state* st,
                             const void* const* kv, void* y_out_v) {
  if (stage < 1 || stage >= tab->n_stages) return TODE_EINVAL;
  if (st->ctl && st->ctl[TODE_CTL_STOP]) return 0;
  const int64_t B = st->B, F = st->F;
  const Dt* y = (const Dt*)st->y;
  const Tt* dt = (const Tt*)st->dt;
  const Dt* const* k = (const Dt* const*)kv;
  Dt* out = (Dt*)y_out_v;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    if (st->running && !st->running[b]) continue; /* finished rows: untouched */
    const Dt dtD = (Dt)dt[b];                        /* :247 */
    for (int64_t f = 0; f < F; ++f) {
      const int64_t e = b * F + f;
      /* :261 einsum("j,jbf->bf", a[i,:i], k[:i]) -- FMA chain, ascending j */
      Dt acc = (Dt)tab->a[stage][0] * k[0][e];
      for (int j = 1; j < stage; ++j) acc = DFMA((Dt)tab->a[stage][j], k[j][e], acc);
      out[e] = DFMA(dtD, acc, y[e]); /* :262 addcmul(y0, dt, acc) */
    }
  }
  return 0;
}

/* runge_kutta.py:268-269: y1 = y0 + einsum(dt, b, k) / error = einsum(dt, b_err, k) */
static int FN(orc_erk_weighted_sum)(const tode_tableau* tab, int which, int64_t B, int64_t F,
                                    const void* dtv, const void* const* kv, const void* basev,
                                    void* outv) {
  const Tt* dt = (const Tt*)dtv;
  const Dt* const* k = (const Dt* const*)kv;
  const Dt* base = (const Dt*)basev;
  const double* w = which == TODE_W_B ? tab->b : tab->b_err;
  Dt* out = (Dt*)outv;
  for (int64_t b = 0; b < B; ++b)
    for (int64_t f = 0; f < F; ++f) {
      const Dt acc = FN(weighted_sum)((Dt)dt[b], w, tab->n_stages, k, b * F + f);
      out[b * F + f] = base ? base[b * F + f] + acc : acc;
    }
  return 0;
}

/* step_size_controllers.py:393-429 / 738-774 as a stand-alone op */
static int FN(orc_adapt_step_size)(const tode_controller* c, int64_t B, int64_t F, const void* dtv,
                                   const void* y0v, const void* y1v, const void* errv,
                                   const void* r1v, const void* r2v, uint8_t* accept,
                                   void* dt_nextv, void* ratiov, void* r1ov, void* r2ov,
                                   int64_t* status) {
  const Tt* dt = (const Tt*)dtv;
  const Dt *y0 = (const Dt*)y0v, *y1 = (const Dt*)y1v, *err = (const Dt*)errv;
  const Dt *r1 = (const Dt*)r1v, *r2 = (const Dt*)r2v;
  Dt* q = (Dt*)malloc(sizeof(Dt) * (size_t)F);
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      const int64_t e = b * F + f;
      const Dt bounds = DFMA((Dt)c->rtol, FN(dmax)(DABS(y0[e]), DABS(y1[e])), (Dt)c->atol);
      q[f] = DABS(err[e]) / bounds;
    }
    const Dt nrm = FN(row_norm)(q, F, c->norm);
    FN(ctrl_out) o = FN(controller)(c, nrm, dt[b], r1 ? r1[b] : (Dt)1, r2 ? r2[b] : (Dt)1);
    accept[b] = (uint8_t)o.accept;
    ((Tt*)dt_nextv)[b] = o.dt_next;
    if (ratiov) ((Dt*)ratiov)[b] = o.ratio;
    if (r1ov) ((Dt*)r1ov)[b] = o.r1;
    if (r2ov) ((Dt*)r2ov)[b] = o.r2;
    status[b] = o.status;
  }
  free(q);
  return 0;
}

/* build_interpolation + evaluate for arbitrary (t, idx) pairs */
static int FN(orc_interp_eval)(const tode_tableau* tab, int64_t B, int64_t F, int64_t N,
                               const void* t0v, const void* dtv, const void* y0v, const void* y1v,
                               const void* const* kv, const void* tv, const int64_t* idx,
                               void* outv) {
  (void)B;
  const Tt *t0 = (const Tt*)t0v, *dt = (const Tt*)dtv, *t = (const Tt*)tv;
  const Dt *y0 = (const Dt*)y0v, *y1 = (const Dt*)y1v;
  const Dt* const* k = (const Dt* const*)kv;
  Dt* out = (Dt*)outv;
  for (int64_t n = 0; n < N; ++n) {
    const int64_t b = idx[n];
    const Dt x = FN(interp_x)(t[n], t0[b], dt[b]);
    for (int64_t f = 0; f < F; ++f) {
      Dt co[5];
      FN(interp_coeffs)(tab, (Dt)dt[b], y0[b * F + f], y1[b * F + f], k, b * F + f, co);
      out[n * F + f] = FN(horner4)(co, x);
    }
  }
  return 0;
}

/* t_nodes[i][b] = addcmul(t0, c[:,None], dt)  runge_kutta.py:259 */
static inline void FN(write_t_nodes)(const tode_tableau* tab, const tode_state* st, int64_t b, Tt t, Tt dt) {
  Tt* tn = (Tt*)st->t_nodes;
  if (!tn) return;
  for (int i = 0; i < tab->n_stages; ++i) tn[(int64_t)i * st->B + b] = TFMA((Tt)tab->c[i], dt, t);
}

/* Dense output for one sample after its commit (adjoints.py:215-234, 298-301).
 * (t0, dt, y0 row, y1 row, k) describe THIS step; t_new is t after the commit. */
static inline void FN(dense_output)(const tode_tableau* tab, const tode_state* st, int64_t b, Tt dir,
                                    Tt t0, Tt dt, Tt t_new, const Dt* y1, const Dt* const* k,
                                    int running_old, int running_new, int status) {
  const int64_t F = st->F, Tn = st->T;
  const Dt* y0 = (const Dt*)st->y; /* NOT yet overwritten by the commit (caller's order) */
  Dt* ye = (Dt*)st->y_eval;
  const Dt dtD = (Dt)dt;
  if (Tn == 0) {
    /* adjoints.py:298-301: the LAST iteration's interpolant at t_end.  A sample's
     * last effective step is the one in which it finishes or reports a failure. */
    if (!running_old || (running_new && status == TODE_SUCCESS)) return;
    const Tt t_end = ((const Tt*)st->t_end)[b];
    const Dt x = FN(interp_x)(t_end, t0, dt);
    for (int64_t f = 0; f < F; ++f) {
      Dt co[5];
      FN(interp_coeffs)(tab, dtD, y0[b * F + f], y1[b * F + f], k, b * F + f, co);
      ye[b * F + f] = FN(horner4)(co, x);
    }
    return;
  }
  const Tt* te = (const Tt*)st->t_eval + b * st->t_eval_stride_b;
  if (st->not_yet == NULL) {
    /* cursor mode: rows monotone in the direction of time, evaluated set is a prefix */
    int32_t cur = st->cursor[b];
    while (cur < Tn) {
      /* :216-223  addcmul(-dir*t_eval, dir, t) >= 0 */
      const Tt tej = te[cur];
      if (!(TFMA(dir, t_new, -dir * tej) >= (Tt)0)) break;
      const Dt x = FN(interp_x)(tej, t0, dt);
      for (int64_t f = 0; f < F; ++f) {
        Dt co[5];
        FN(interp_coeffs)(tab, dtD, y0[b * F + f], y1[b * F + f], k, b * F + f, co);
        ye[(b * Tn + cur) * F + f] = FN(horner4)(co, x);
      }
      ++cur;
    }
    st->cursor[b] = cur;
  } else {
    uint8_t* ny = st->not_yet + b * Tn;
    for (int64_t j = 0; j < Tn; ++j) {
      if (!ny[j]) continue;
      const Tt tej = te[j];
      if (!(TFMA(dir, t_new, -dir * tej) >= (Tt)0)) continue;
      const Dt x = FN(interp_x)(tej, t0, dt);
      for (int64_t f = 0; f < F; ++f) {
        Dt co[5];
        FN(interp_coeffs)(tab, dtD, y0[b * F + f], y1[b * F + f], k, b * F + f, co);
        ye[(b * Tn + j) * F + f] = FN(horner4)(co, x);
      }
      ny[j] = 0; /* :232 logical_xor */
    }
  }
}

/* ------------------------------------------------------------------ */
/* One loop iteration after the last stage: adjoints.py:140-255         */
/* ------------------------------------------------------------------ */
/* `exact_end` = 1 reproduces the reference literally when there is no t_eval: the
 * interpolant of the batch's LAST iteration is evaluated at t_end for every sample
 * (adjoints.py:298-301), which needs the batch-wide stop decision before the dense
 * output.  With 0 the product's path-A rule is used: a sample writes its end value in
 * the iteration in which it finishes or fails itself (identical unless ANOTHER sample's
 * failure aborts the batch while this one is still running). */
static int FN(orc_erk_finish_ex)(const tode_tableau* tab, const tode_controller* c,
                                 const tode_state* st, const void* const* kv, const void* y1v,
                                 int exact_end) {
  if (st->ctl[TODE_CTL_STOP]) return 0;
  const int64_t B = st->B, F = st->F;
  const int S = tab->n_stages;
  const Dt* const* k = (const Dt* const*)kv;
  const Dt* y1 = (const Dt*)y1v;
  Dt* y = (Dt*)st->y;
  Dt* f0 = (Dt*)st->f0;
  Tt* t = (Tt*)st->t;
  Tt* dtp = (Tt*)st->dt;
  const Tt* t_start = (const Tt*)st->t_start;
  const Tt* t_end = (const Tt*)st->t_end;
  Dt* r1p = (Dt*)st->r1;
  Dt* r2p = (Dt*)st->r2;
  FN(ctrl_out)* outs = (FN(ctrl_out)*)malloc(sizeof(FN(ctrl_out)) * (size_t)B);
  Tt* t_news = (Tt*)malloc(sizeof(Tt) * (size_t)B);
  uint8_t* run_new = (uint8_t*)malloc((size_t)B);
  int32_t* stat = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
  int64_t n_running = 0;
  int failed = 0;
  /* ---- pass A: controller decisions for every running sample ---- */
#pragma omp parallel for schedule(static) reduction(+ : n_running) reduction(| : failed)
  for (int64_t b = 0; b < B; ++b) {
    run_new[b] = 0;
    if (!st->running[b]) continue; /* finished: state frozen, status stays SUCCESS */
    const Tt t0 = t[b], dt = dtp[b];
    const Dt dtD = (Dt)dt; /* runge_kutta.py:247 */
    const Tt dir = FN(dir_of)(t_start[b], t_end[b]);
    /* error ratio: step_size_controllers.py:394-400 */
    Dt qs[16];
    Dt* q = F <= 16 ? qs : (Dt*)malloc(sizeof(Dt) * (size_t)F);
    for (int64_t f = 0; f < F; ++f) {
      const int64_t e = b * F + f;
      const Dt err = FN(weighted_sum)(dtD, tab->b_err, S, k, e); /* runge_kutta.py:269 */
      const Dt bounds = DFMA((Dt)c->rtol, FN(dmax)(DABS(y[e]), DABS(y1[e])), (Dt)c->atol);
      q[f] = DABS(err) / bounds;
    }
    const Dt nrm = FN(row_norm)(q, F, c->norm);
    if (q != qs) free(q);
    const Dt r1 = c->pid ? r1p[b] : (Dt)1, r2 = c->pid ? r2p[b] : (Dt)1;
    outs[b] = FN(controller)(c, nrm, dt, r1, r2);
    /* adjoints.py:150-162 */
    const int upd = outs[b].accept;
    t_news[b] = upd ? t0 + dt : t0;
    const int32_t ns = st->n_steps[b] + 1;
    st->n_steps[b] = ns;
    st->n_accepted[b] += upd;
    /* :169  running = addcmul(-dir*t_end, dir, t) < 0 */
    run_new[b] = (uint8_t)(TFMA(dir, t_news[b], -dir * t_end[b]) < (Tt)0);
    /* :171-181 */
    int status = outs[b].status;
    if (c->max_steps >= 0 && (int64_t)ns >= c->max_steps) status = TODE_REACHED_MAX_STEPS;
    stat[b] = status;
    st->status[b] = status;
    n_running += run_new[b];
    failed |= (status != TODE_SUCCESS);
  }
  /* adjoints.py:186-190 */
  const int stop_now = (n_running == 0 || failed);
  /* ---- pass B: dense output, commit, next dt ---- */
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    if (!st->running[b]) continue;
    const Tt t0 = t[b], dt = dtp[b];
    const Tt dir = FN(dir_of)(t_start[b], t_end[b]);
    const Tt t_min = t_start[b] < t_end[b] ? t_start[b] : t_end[b]; /* adjoints.py:66-67 */
    const Tt t_max = t_start[b] < t_end[b] ? t_end[b] : t_start[b];
    const Tt t_new = t_news[b];
    const int running_new = run_new[b];
    /* :215-234 / :298-301 dense output with the data of this step, BEFORE y is overwritten */
    FN(dense_output)(tab, st, b, dir, t0, dt, t_new, y1, k, 1, running_new && !(exact_end && stop_now),
                     stat[b]);
    /* commit: adjoints.py:151-155, runge_kutta.py:216-224 (FSAL slot) */
    if (outs[b].accept) {
      for (int64_t f = 0; f < F; ++f) {
        y[b * F + f] = y1[b * F + f];
        f0[b * F + f] = k[S - 1][b * F + f];
      }
    }
    t[b] = t_new;
    /* :247-251 */
    Tt dt_new = running_new ? outs[b].dt_next : dt;
    dt_new = FN(tclamp)(dt_new, t_min - t_new, t_max - t_new);
    dtp[b] = dt_new;
    /* :253-255 PIDController.merge_states :639-647 */
    if (c->pid && running_new) {
      r1p[b] = outs[b].r1;
      r2p[b] = outs[b].r2;
    }
    FN(write_t_nodes)(tab, st, b, t_new, dt_new);
  }
  for (int64_t b = 0; b < B; ++b) st->running[b] = run_new[b];
  st->ctl[TODE_CTL_ITERS] += 1;
  if (stop_now) st->ctl[TODE_CTL_STOP] = 1;
  free(outs); free(t_news); free(run_new); free(stat);
  return 0;
}

static int FN(orc_erk_finish)(const tode_tableau* tab, const tode_controller* c,
                              const tode_state* st, const void* const* kv, const void* y1v) {
  return FN(orc_erk_finish_ex)(tab, c, st, kv, y1v, 0);
}

/* Common state initialisation (adjoints.py:59-126) once dt is known. */
static void FN(init_state)(const tode_tableau* tab, const tode_controller* c, const tode_state* st) {
  const int64_t B = st->B, F = st->F, Tn = st->T;
  Tt* t = (Tt*)st->t;
  Tt* dtp = (Tt*)st->dt;
  const Tt* t_start = (const Tt*)st->t_start;
  const Tt* t_end = (const Tt*)st->t_end;
  int nonmono = 0;
  for (int64_t b = 0; b < B; ++b) {
    const Tt t_min = t_start[b] < t_end[b] ? t_start[b] : t_end[b];
    const Tt t_max = t_start[b] < t_end[b] ? t_end[b] : t_start[b];
    const Tt dir = FN(dir_of)(t_start[b], t_end[b]);
    t[b] = t_start[b];
    dtp[b] = FN(tclamp)(dtp[b], t_min - t[b], t_max - t[b]); /* :109 */
    st->running[b] = 1;
    st->n_steps[b] = 0;
    st->n_accepted[b] = 0;
    st->status[b] = 0;
    if (c->pid) { /* PIDState.default :542 */
      ((Dt*)st->r1)[b] = (Dt)1;
      ((Dt*)st->r2)[b] = (Dt)1;
    }
    if (st->cursor) st->cursor[b] = 0;
    if (Tn == 0) /* never hand out uninitialised memory (deviations) */
      for (int64_t f = 0; f < F; ++f) ((Dt*)st->y_eval)[b * F + f] = ((const Dt*)st->y)[b * F + f];
    if (Tn > 0) {
      const Tt* te = (const Tt*)st->t_eval + b * st->t_eval_stride_b;
      /* :123-126 */
      if (te[0] == t_start[b]) {
        for (int64_t f = 0; f < F; ++f) ((Dt*)st->y_eval)[(b * Tn) * F + f] = ((const Dt*)st->y)[b * F + f];
        if (st->cursor) st->cursor[b] = 1;
        if (st->not_yet) st->not_yet[b * Tn] = 0;
      }
      for (int64_t j = 1; j < Tn; ++j)
        if (dir * te[j] < dir * te[j - 1]) nonmono = 1;
    }
    FN(write_t_nodes)(tab, st, b, t[b], dtp[b]);
  }
  for (int i = 0; i < TODE_CTL_WORDS; ++i) st->ctl[i] = 0;
  st->ctl[TODE_CTL_NONMONO] = nonmono;
}

/* step_size_controllers.py:453-473 (Integral) / :798-818 (PID) */
static int FN(orc_init_step_a)(const tode_tableau* tab, const tode_controller* c, const tode_state* st,
                               void* y1v, void* t1v) {
  (void)tab;
  const int64_t B = st->B, F = st->F;
  if (st->scratch_elems < 2 * B) return TODE_EINVAL;
  const Dt* y0 = (const Dt*)st->y;
  const Dt* f0 = (const Dt*)st->f0;
  const Tt* t_start = (const Tt*)st->t_start;
  const Tt* t_end = (const Tt*)st->t_end;
  Dt* scr = (Dt*)st->scratch; /* [0,B): dt0, [B,2B): d1 */
  Dt* y1 = (Dt*)y1v;
  Tt* t1 = (Tt*)t1v;
  Dt* q0 = (Dt*)malloc(sizeof(Dt) * (size_t)F);
  Dt* q1 = (Dt*)malloc(sizeof(Dt) * (size_t)F);
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      const int64_t e = b * F + f;
      /* :461-462  reciprocal(add(atol, |y0|, alpha=rtol)) */
      const Dt inv = (Dt)1 / DFMA((Dt)c->rtol, DABS(y0[e]), (Dt)c->atol);
      q0[f] = y0[e] * inv;
      q1[f] = f0[e] * inv;
    }
    const Dt d0 = FN(row_norm)(q0, F, c->norm); /* :464-465 */
    const Dt d1 = FN(row_norm)(q1, F, c->norm);
    /* :467-468 */
    Dt dt0 = (d0 < (Dt)1e-5 || d1 < (Dt)1e-5) ? (Dt)1e-6 : ((Dt)0.01 * d0) / d1;
    /* :471  minimum(dt0, |t_end - t_start|.to(Dt)) */
    dt0 = FN(dmin)(dt0, (Dt)TABS(t_end[b] - t_start[b]));
    const Tt dir = FN(dir_of)(t_start[b], t_end[b]);
    const Dt sdt = (Dt)dir * dt0; /* (direction * dt0) */
    for (int64_t f = 0; f < F; ++f) y1[b * F + f] = DFMA(sdt, f0[b * F + f], y0[b * F + f]); /* :473 */
    t1[b] = TFMA(dir, (Tt)dt0, t_start[b]); /* :475 */
    scr[b] = dt0;
    scr[B + b] = d1;
  }
  free(q0);
  free(q1);
  return 0;
}

/* step_size_controllers.py:481-490 / :826-835, then adjoints.py:59-126 */
static int FN(orc_init_step_b)(const tode_tableau* tab, const tode_controller* c, const tode_state* st,
                               const void* f1v) {
  const int64_t B = st->B, F = st->F;
  const Dt* y0 = (const Dt*)st->y;
  const Dt* f0 = (const Dt*)st->f0;
  const Dt* f1 = (const Dt*)f1v;
  const Tt* t_start = (const Tt*)st->t_start;
  const Tt* t_end = (const Tt*)st->t_end;
  const Dt* scr = (const Dt*)st->scratch;
  Dt* q = (Dt*)malloc(sizeof(Dt) * (size_t)F);
  for (int64_t b = 0; b < B; ++b) {
    const Dt dt0 = scr[b], d1 = scr[B + b];
    for (int64_t f = 0; f < F; ++f) {
      const int64_t e = b * F + f;
      const Dt inv = (Dt)1 / DFMA((Dt)c->rtol, DABS(y0[e]), (Dt)c->atol);
      q[f] = (f1[e] - f0[e]) * inv;
    }
    Dt d2 = FN(row_norm)(q, F, c->norm) / dt0;
    if (!c->pid && dt0 == (Dt)0) d2 = (Dt)INFINITY; /* only the Integral copy guards (:481 vs :826) */
    const Dt m = FN(dmax)(d1, d2);
    /* :484-488; `0.01 / m` is Tensor.__rtruediv__ = m.reciprocal() * 0.01 */
    Dt dt1;
    if (m <= (Dt)1e-15)
      dt1 = FN(dmax)((Dt)1e-6, dt0 * (Dt)1e-3);
    else
      dt1 = DPOW(((Dt)1 / m) * (Dt)0.01, 1.0 / (double)tab->order);
    const Tt dir = FN(dir_of)(t_start[b], t_end[b]);
    /* :490 */
    ((Tt*)st->dt)[b] = (Tt)((Dt)dir * FN(dmin)((Dt)100 * dt0, dt1));
  }
  free(q);
  FN(init_state)(tab, c, st);
  return 0;
}

static int FN(orc_init_with_dt0)(const tode_tableau* tab, const tode_controller* c, const tode_state* st,
                                 const void* dt0v) {
  for (int64_t b = 0; b < st->B; ++b) ((Tt*)st->dt)[b] = ((const Tt*)dt0v)[b];
  FN(init_state)(tab, c, st);
  return 0;
}

/* ------------------------------------------------------------------ */
/* Built-in analytic vector fields (op order = the torch forward of     */
/* torchode_b200.fields.*: one rounding per torch op)                   */
/* ------------------------------------------------------------------ */
static void FN(eval_field)(int field, const double* p, int64_t B, int64_t F, const Dt* y, Dt* out) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < B; ++b) {
    const Dt* yb = y + b * F;
    Dt* ob = out + b * F;
    switch (field) {
      case TODE_FIELD_LINEAR:
        for (int64_t f = 0; f < F; ++f) ob[f] = (Dt)p[0] * yb[f];
        break;
      case TODE_FIELD_VAN_DER_POL: {
        const Dt x = yb[0], v = yb[1];
        const Dt xx = x * x;
        const Dt one_m = (Dt)1 - xx;
        const Dt m1 = (Dt)p[0] * one_m;
        const Dt m2 = m1 * v;
        ob[0] = v;
        ob[1] = m2 - x;
        break;
      }
      case TODE_FIELD_LOTKA_VOLTERRA: {
        const Dt x = yb[0], z = yb[1];
        const Dt xz = x * z;
        ob[0] = (Dt)p[0] * x - (Dt)p[1] * xz;
        ob[1] = (Dt)p[2] * xz - (Dt)p[3] * z;
        break;
      }
      default:
        break;
    }
  }
}

/* The whole of AutoDiffAdjoint.solve (adjoints.py:43-311) for a built-in field,
 * literally: lock-step over the batch, one iteration at a time. */
static int FN(orc_solve_builtin)(int field, const double* fp, const tode_tableau* tab,
                                 const tode_controller* c, const tode_problem* prob,
                                 const tode_solution* sol, int64_t iter_cap) {
  const int64_t B = prob->B, F = prob->F, Tn = prob->T;
  const int S = tab->n_stages;
  if ((field == TODE_FIELD_VAN_DER_POL || field == TODE_FIELD_LOTKA_VOLTERRA) && F != 2) return TODE_EINVAL;
  tode_state st;
  memset(&st, 0, sizeof(st));
  st.B = B;
  st.F = F;
  st.T = Tn;
  st.data_dtype = prob->data_dtype;
  st.time_dtype = prob->time_dtype;
  st.t_start = prob->t_start;
  st.t_end = prob->t_end;
  st.t_eval = prob->t_eval;
  st.t_eval_stride_b = prob->t_eval_stride_b;
  const size_t nBF = (size_t)(B * F);
  Dt* ybuf = (Dt*)malloc(sizeof(Dt) * nBF);
  memcpy(ybuf, prob->y0, sizeof(Dt) * nBF);
  Dt* kbuf = (Dt*)malloc(sizeof(Dt) * nBF * (size_t)S);
  Dt* ystage = (Dt*)malloc(sizeof(Dt) * nBF);
  Dt* f0 = (Dt*)malloc(sizeof(Dt) * nBF);
  st.y = ybuf;
  st.f0 = f0;
  st.t = malloc(sizeof(Tt) * (size_t)B);
  st.dt = malloc(sizeof(Tt) * (size_t)B);
  st.r1 = malloc(sizeof(Dt) * (size_t)B);
  st.r2 = malloc(sizeof(Dt) * (size_t)B);
  st.running = (uint8_t*)malloc((size_t)B);
  st.n_steps = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
  st.n_accepted = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
  st.status = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
  st.cursor = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
  st.not_yet = NULL;
  st.y_eval = sol->ys;
  st.t_nodes = NULL;
  int32_t ctl[TODE_CTL_WORDS];
  st.ctl = ctl;
  st.scratch = malloc(sizeof(Dt) * (size_t)(2 * B));
  st.scratch_elems = 2 * B;
  memcpy(st.t, prob->t_start, sizeof(Tt) * (size_t)B);

  /* ExplicitRungeKutta.init / controller.init: f0 = f(t_start, y0) */
  FN(eval_field)(field, fp, B, F, ybuf, f0);
  if (prob->dt0 == NULL) {
    Tt* t1 = (Tt*)malloc(sizeof(Tt) * (size_t)B);
    FN(orc_init_step_a)(tab, c, &st, ystage, t1);
    FN(eval_field)(field, fp, B, F, ystage, kbuf); /* f1 */
    FN(orc_init_step_b)(tab, c, &st, kbuf);
    free(t1);
  } else {
    FN(orc_init_with_dt0)(tab, c, &st, prob->dt0);
  }
  /* general (non-monotone) t_eval rows: switch to the mask */
  if (ctl[TODE_CTL_NONMONO] && Tn > 0) {
    st.not_yet = (uint8_t*)malloc((size_t)(B * Tn));
    memset(st.not_yet, 1, (size_t)(B * Tn));
    for (int64_t b = 0; b < B; ++b)
      if (st.cursor[b] == 1) st.not_yet[b * Tn] = 0;
  }
  const Dt* kp[TODE_MAX_STAGES];
  kp[0] = f0;
  for (int s = 1; s < S; ++s) kp[s] = kbuf + (size_t)s * nBF;
  int32_t first_fail = INT32_MAX;
  while (!ctl[TODE_CTL_STOP]) {
    for (int s = 1; s < S; ++s) {
      FN(orc_erk_stage)(tab, s, &st, (const void* const*)kp, ystage);
      /* rows of finished samples keep stale values; the field is evaluated on
       * every row like the reference does, results of finished rows are unused */
      FN(eval_field)(field, fp, B, F, ystage, kbuf + (size_t)s * nBF);
    }
    FN(orc_erk_finish_ex)(tab, c, &st, (const void* const*)kp, ystage,
                          /*exact_end=*/1);
    if (ctl[TODE_CTL_STOP]) {
      for (int64_t b = 0; b < B; ++b)
        if (st.status[b] != 0) first_fail = ctl[TODE_CTL_ITERS];
    }
    if (iter_cap > 0 && ctl[TODE_CTL_ITERS] >= iter_cap) break;
  }
  for (int64_t b = 0; b < B; ++b) {
    sol->n_steps[b] = st.n_steps[b];
    sol->n_accepted[b] = st.n_accepted[b];
    sol->status[b] = st.status[b];
    if (Tn == 0) {
      sol->n_initialized[b] = 1;
    } else if (st.not_yet == NULL) {
      sol->n_initialized[b] = st.cursor[b];
    } else {
      /* adjoints.py:289-292: searchsorted(not_yet.int(), 1) (left) on the row */
      int64_t lo = 0, hi = Tn;
      while (lo < hi) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (st.not_yet[b * Tn + mid] < 1) lo = mid + 1; else hi = mid;
      }
      sol->n_initialized[b] = lo;
    }
    if (sol->t_final) ((Tt*)sol->t_final)[b] = ((Tt*)st.t)[b];
    if (sol->dt_final) ((Tt*)sol->dt_final)[b] = ((Tt*)st.dt)[b];
  }
  if (sol->summary) {
    sol->summary[0] = ctl[TODE_CTL_ITERS];
    sol->summary[1] = first_fail;
    sol->summary[2] = ctl[TODE_CTL_NONMONO];
  }
  free(ybuf); free(kbuf); free(ystage); free(f0);
  free(st.t); free(st.dt); free(st.r1); free(st.r2); free(st.running);
  free(st.n_steps); free(st.n_accepted); free(st.status); free(st.cursor);
  free(st.not_yet); free(st.scratch);
  return 0;
}

#undef FN
#undef DFMA
#undef TFMA
#undef DSQRT
#undef DABS
#undef TABS
#undef DPOW
