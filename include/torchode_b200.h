/*
 * torchode_b200 -- C-ABI of the B200-native batch-parallel adaptive explicit
 * Runge-Kutta solve loop (drop-in for torchode's AutoDiffAdjoint.solve path).
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / PODs, is
 * asynchronous on the CUDA stream it is given (a `cudaStream_t` passed as
 * `void*`), never allocates, never synchronises and never throws: it returns
 * 0 on success, a positive `cudaError_t` value for a CUDA failure or a
 * negative TODE_E* code for an argument error.  All tensor arguments are
 * DEVICE pointers owned by the caller.
 *
 * The reference has no native boundary at all (it is pure Python / PyTorch);
 * each entry point therefore cites the reference *Python* function whose body
 * it replaces (paths relative to the torchode v1.0.1 tree):
 *
 *   tode_erk_stage          torchode/single_step_methods/runge_kutta.py:259-263
 *   tode_erk_finish         runge_kutta.py:265-279 (y1, error estimate, FSAL),
 *                           step_size_controllers.py:371-429 / 716-774 (error
 *                           ratio, accept, dt factor, status), :598-671 (PID),
 *                           adjoints.py:150-181 (commit, stats, running,
 *                           status), :215-234 (dense output), :247-255 (next
 *                           dt, controller state merge), :186-201 (termination)
 *   tode_init_step_a/_b     step_size_controllers.py:431-490 / 776-835 and
 *                           adjoints.py:59-126 (state initialisation)
 *   tode_init_with_dt0      adjoints.py:103-126 with a user-supplied dt0
 *   tode_solve_fused        the whole of adjoints.py:43-311 for a built-in
 *                           analytic vector field
 *   tode_adapt_step_size    step_size_controllers.py:371-429 / 716-774 alone
 *   tode_erk_weighted_sum   runge_kutta.py:268-269 (y1 for non-SSAL tableaux, error estimate)
 *   tode_time_nodes         runge_kutta.py:259
 *   tode_interp_eval        dopri5.py:54-60 / tsit5.py:124-139 +
 *                           interpolation.py:25-40,139-175
 */
#ifndef TORCHODE_B200_H
#define TORCHODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TODE_ABI_VERSION 3
#define TODE_SUMMARY_WORDS 8
#define TODE_MAX_STAGES 7
#define TODE_MAX_FIELD_PARAMS 8
#define TODE_MAX_PEERS 8

/* argument-error codes (negative; CUDA errors are returned as positive values) */
#define TODE_EINVAL (-1)   /* bad argument (NULL pointer, bad size, bad enum)    */
#define TODE_ENOSUP (-2)   /* combination not supported by this build            */
#define TODE_EALIGN (-3)   /* pointer not aligned as the layout contract demands */

enum tode_dtype { TODE_F32 = 0, TODE_F64 = 1 };
enum tode_norm { TODE_NORM_RMS = 0, TODE_NORM_MAX = 1 };

/* Dense-output recipe (which quartic the step method builds from its stages). */
enum tode_interp {
  TODE_INTERP_DOPRI5 = 0, /* w[0] = b_mid, Hermite-type quartic, interpolation.py:139-170 */
  TODE_INTERP_TSIT5 = 1   /* w[0..2] = weights of x^2,x^3,x^4, tsit5.py:124-139           */
};

/* torchode/status_codes.py:9-13 */
enum tode_status {
  TODE_SUCCESS = 0,
  TODE_GENERAL_ERROR = 1,
  TODE_REACHED_DT_MIN = 2,
  TODE_REACHED_MAX_STEPS = 3,
  TODE_INFINITE_NORM = 4
};

/* Built-in analytic vector fields of the fused whole-solve kernel. */
enum tode_field {
  TODE_FIELD_LINEAR = 0,         /* y' = p0 * y                           (any F<=4) */
  TODE_FIELD_VAN_DER_POL = 1,    /* x' = v ; v' = p0*(1-x*x)*v - x        (F = 2)    */
  TODE_FIELD_LOTKA_VOLTERRA = 2  /* x' = p0*x - p1*(x*z) ; z' = p2*(x*z) - p3*z (F=2) */
};

/* Butcher tableau in float64, exactly as ButcherTableau.from_lists builds it
 * (runge_kutta.py:60-105); kernels round to the data dtype (a, b, b_err, w)
 * and the time dtype (c) like ButcherTableau.to (runge_kutta.py:107-121).
 * Only FSAL + SSAL explicit tableaux are supported by the fused kernels
 * (both Dopri5 and Tsit5 are). */
typedef struct tode_tableau {
  int32_t n_stages; /* <= TODE_MAX_STAGES */
  int32_t interp;   /* enum tode_interp */
  int32_t order;    /* convergence_order() */
  int32_t reserved;
  double c[TODE_MAX_STAGES];
  double a[TODE_MAX_STAGES][TODE_MAX_STAGES];
  double b[TODE_MAX_STAGES];
  double b_err[TODE_MAX_STAGES];
  double w[3][TODE_MAX_STAGES];
} tode_tableau;

/* Step-size controller parameters (IntegralController / PIDController
 * constructor arguments, step_size_controllers.py:263-287 / 566-596).
 * atol/rtol are the float32-ROUNDED values (the reference keeps them in fp32
 * buffers, :278-279).  The three exponents are the python-float expressions
 * of dt_factor (:289-294 / :612-618) evaluated on the host:
 *   integral: exp_ratio = -(1/order), pid = 0
 *   PID:      exp_ratio = -(kI+kP+kD), exp_prev = kP+2kD, exp_prev2 = -kD */
typedef struct tode_controller {
  int32_t norm; /* enum tode_norm */
  int32_t pid;  /* 0 = integral, 1 = PID (uses r1/r2 history) */
  int32_t has_dt_min;
  int32_t has_dt_max;
  double atol, rtol;
  double safety, factor_min, factor_max;
  double exp_ratio, exp_prev, exp_prev2;
  double dt_min, dt_max;
  double almost_zero; /* 1e-38 (1e-5 for fp16), step_size_controllers.py:251-256 */
  int64_t max_steps;  /* < 0: unlimited (adjoints.py:176-181) */
  /* stage-wise path, replay after a failure (ABI 3): > 0 = the batch is known to stop after this
   * many loop iterations; without t_eval a sample still running then writes its end value from that
   * iteration's interpolant, as the reference does for a batch another sample aborted
   * (adjoints.py:298-301).  0: not a replay. */
  int64_t iter_cap;
} tode_controller;

/* Device control block shared by all launches of one solve (int32 words).
 * The finish kernel maintains it so that the loop never has to sync:
 *   [STOP]    1 once `any(running) & all(status==0)` became false
 *             (adjoints.py:186-190); all later launches are no-ops
 *   [ITERS]   number of effective loop iterations executed so far
 *   [RUNNING] scratch: running-sample count of the iteration in flight
 *   [FAILED]  scratch: any status != SUCCESS in the iteration in flight
 *   [TICKET]  scratch: CTA arrival counter of the finish kernel
 *   [NONMONO] set by tode_init_* if a t_eval row is not monotone in its
 *             direction of time (cursor fast path invalid) */
enum tode_ctl_word {
  TODE_CTL_STOP = 0,
  TODE_CTL_ITERS = 1,
  TODE_CTL_RUNNING = 2,
  TODE_CTL_FAILED = 3,
  TODE_CTL_TICKET = 4,
  TODE_CTL_NONMONO = 5,
  TODE_CTL_WORDS = 8
};

/* Per-solve state of the stage-wise path ("path A": opaque user f).
 * Layouts: (B) vectors contiguous; y, f0 are (B,F) row-major contiguous;
 * t_eval is (B,T) with element strides (t_eval_stride_b may be 0 for a
 * broadcast row, t_eval_stride_t must be 1 or T==0/1); y_eval is (B,T,F)
 * contiguous, or (B,1,F) when T == 0 (solution at t_end only).
 * Counters are int32 on device (widened to int64 by the host at the end). */
typedef struct tode_state {
  int64_t B, F, T;
  int32_t data_dtype; /* enum tode_dtype of y */
  int32_t time_dtype; /* enum tode_dtype of t */
  /* problem (read-only) */
  const void* t_start; /* (B) time */
  const void* t_end;   /* (B) time */
  const void* t_eval;  /* (B,T) time or NULL */
  int64_t t_eval_stride_b;
  /* solver state (read-write) */
  void* t;          /* (B) time */
  void* dt;         /* (B) time */
  void* y;          /* (B,F) data */
  void* f0;         /* (B,F) data: FSAL slot = f(t, y) */
  void* r1;         /* (B) data: PID prev_error_ratio (NULL if !pid) */
  void* r2;         /* (B) data: PID prev_prev_error_ratio */
  uint8_t* running; /* (B) */
  int32_t* n_steps;
  int32_t* n_accepted;
  int32_t* status;
  int32_t* cursor; /* (B): evaluated prefix length of the t_eval row = n_initialized */
  uint8_t* not_yet; /* (B,T) or NULL. NULL: cursor mode (rows monotone in time direction);
                       non-NULL: general mode, scan every not-yet-evaluated point
                       (adjoints.py:216-223) */
  void* y_eval;    /* (B,max(T,1),F) data */
  void* t_nodes;   /* (n_stages,B) time: t + c_i*dt of the NEXT step */
  int32_t* ctl;    /* TODE_CTL_WORDS control words */
  /* scratch for the multi-CTA ("large F") reduction and the initial step */
  void* scratch;         /* data dtype, at least tode_scratch_elems(B,F) elements */
  int64_t scratch_elems;
} tode_state;

/* Problem/solution descriptors of the fused whole-solve path ("path B"). */
typedef struct tode_problem {
  int64_t B, F, T;
  int32_t data_dtype, time_dtype;
  const void* y0;      /* (B,F) */
  const void* t_start; /* (B) */
  const void* t_end;   /* (B) */
  const void* t_eval;  /* (B,T) or NULL */
  int64_t t_eval_stride_b;
  const void* dt0;     /* (B) time or NULL -> Hairer initial-step heuristic */
} tode_problem;

typedef struct tode_solution {
  void* ys;               /* (B,max(T,1),F) data */
  int64_t* n_steps;       /* (B) */
  int64_t* n_accepted;    /* (B) */
  int64_t* n_initialized; /* (B) */
  int64_t* status;        /* (B) */
  void* t_final;          /* (B) time, may be NULL */
  void* dt_final;         /* (B) time, may be NULL */
  /* device int32[TODE_SUMMARY_WORDS] (8-byte aligned): [0] max n_steps over the batch (= loop
   * iterations of the reference), [1] first iteration (1-based) at which any sample reported a
   * status != SUCCESS or INT32_MAX, [2] non-monotone t_eval flag, [3] number of kernels the call launched, [4..5] scratch: the
   * 64-bit work queue of the persistent dense-output kernel, [6..7] reserved (ABI 3) */
  int32_t* summary;
  /* Multi-GPU, "write the all-gather while solving" (replaces the all_gather of ys / statistics
   * that follows a sharded solve): with n_peers > 0 every result of sample b --
   * its ys rows as they are produced and its statistics when it finishes -- is ALSO stored at
   * row peer_row0 + b of the gathered buffers of n_peers replicas.  The pointers must be valid
   * on THIS GPU (peer memory mapped over NVLink, e.g. one symmetric-memory allocation per rank;
   * a replica may be this GPU's own gathered buffer).  peer_ys[p] may be NULL: the ys rows are
   * then not replicated to p by the kernel (dense-output workloads write many small rows; the
   * caller pushes its finished ys block to the peers in bulk instead).  The four statistics
   * pointers of a replica may be NULL together (e.g. the caller's own replica when ys / n_steps /
   * ... above already point into it).  peer_global[p] = device int32[4] of
   * replica p, zeroed on every rank before the launches of a step: after the solve kernel a
   * one-thread epilogue kernel atomicMax-es (system scope) this launch's iteration count into [0]
   * and, if a sample failed, INT32_MAX - (first failing iteration) into [2] of every replica --
   * after a cross-GPU barrier every rank knows the batch-wide loop length and the batch-wide
   * first failure, and replays its shard with iter_cap = that iteration if its samples ran past
   * it ("any failure stops the whole batch", adjoints.py:186-190, across GPUs).  A replay launch
   * (iter_cap > 0) publishes nothing.  No collective, no host synchronisation; the caller
   * brackets the launches with two cross-GPU barriers. */
  int32_t n_peers;
  int32_t reserved;
  int64_t peer_row0;
  void* peer_ys[TODE_MAX_PEERS];
  int64_t* peer_n_steps[TODE_MAX_PEERS];
  int64_t* peer_n_accepted[TODE_MAX_PEERS];
  int64_t* peer_n_initialized[TODE_MAX_PEERS];
  int64_t* peer_status[TODE_MAX_PEERS];
  int32_t* peer_global[TODE_MAX_PEERS];
} tode_solution;

int tode_abi_version(void);
const char* tode_error_string(int code);
int64_t tode_scratch_elems(int64_t B, int64_t F);

/* y_out[b,:] = y[b,:] + dt[b] * sum_{j<stage} a[stage][j] * k[j][b,:]   (1 <= stage < n_stages)
 * k[j] are (B,F) device pointers (k[0] may be st->f0).  Rows with running==0
 * are neither read nor written. */
int tode_erk_stage(const tode_tableau* tab, int stage, const tode_state* st,
                   const void* const* k, void* y_out, void* stream);

/* Everything after the last stage of one loop iteration; y1 is the last
 * stage's y_out (SSAL), k[n_stages-1] = f(t+dt, y1). */
int tode_erk_finish(const tode_tableau* tab, const tode_controller* ctrl,
                    const tode_state* st, const void* const* k, const void* y1,
                    void* stream);

/* Initial step size, part a: needs st->y = y0, st->f0 = f(t_start, y0), st->t = t_start.
 * Writes y1_out = y0 + dir*dt0*f0 (B,F), t1_out = t_start + dir*dt0 (B) and
 * keeps d1, dt0 in st->scratch for part b. */
int tode_init_step_a(const tode_tableau* tab, const tode_controller* ctrl,
                     const tode_state* st, void* y1_out, void* t1_out, void* stream);

/* Part b: f1 = f(t1, y1).  Writes st->dt (clamped to the time domain), and
 * initialises running, counters, cursor (+ y_eval[:,0] where
 * t_eval[:,0]==t_start), PID history, t_nodes and ctl. */
int tode_init_step_b(const tode_tableau* tab, const tode_controller* ctrl,
                     const tode_state* st, const void* f1, void* stream);

/* Same initialisation with a user-supplied dt0 (B, time dtype) copied to st->dt. */
int tode_init_with_dt0(const tode_tableau* tab, const tode_controller* ctrl,
                       const tode_state* st, const void* dt0, void* stream);

/* Whole solve for a built-in field, one launch.  iter_cap > 0 limits every
 * sample to that many loop iterations (used to reproduce "any failure stops
 * the whole batch"). */
int tode_solve_fused(int field, const double* field_params, const tode_tableau* tab,
                     const tode_controller* ctrl, const tode_problem* prob,
                     const tode_solution* sol, int64_t iter_cap, void* stream);

/* Stand-alone protocol pieces (SingleStepMethod / StepSizeController API). */
/* out = (base ? base : 0) + sum_s (dt * w_s) * k[s] with w = tab->b (which = TODE_W_B) or
 * tab->b_err (TODE_W_BERR): einsum("b,s,sbf->bf", dt, w, k) of runge_kutta.py:268-269. */
enum tode_weights { TODE_W_B = 0, TODE_W_BERR = 1 };
int tode_erk_weighted_sum(const tode_tableau* tab, int which, int32_t data_dtype,
                          int32_t time_dtype, int64_t B, int64_t F, const void* dt,
                          const void* const* k, const void* base, void* out, void* stream);

/* out[i,b] = t0[b] + c[i] * dt[b] for every stage i: (n_stages,B) time tensor */
int tode_time_nodes(const tode_tableau* tab, int32_t time_dtype, int64_t B, const void* t0,
                    const void* dt, void* out, void* stream);

int tode_adapt_step_size(const tode_controller* ctrl, int32_t data_dtype, int32_t time_dtype,
                         int64_t B, int64_t F, const void* dt, const void* y0, const void* y1,
                         const void* err, const void* r1, const void* r2, uint8_t* accept_out,
                         void* dt_next_out, void* ratio_out, void* r1_out, void* r2_out,
                         int64_t* status_out, void* stream);

/* out[n,:] = quartic of sample idx[n] built from (t0, dt, y0, y1, k) evaluated at t[n]. */
int tode_interp_eval(const tode_tableau* tab, int32_t data_dtype, int32_t time_dtype,
                     int64_t B, int64_t F, int64_t N, const void* t0, const void* dt,
                     const void* y0, const void* y1, const void* const* k, const void* t,
                     const int64_t* idx, void* out, void* stream);

/* Neural-ODE vector field on the tensor cores (BASELINE.json configs[3]): out = MLP(y) for a
 * width-256 tanh MLP with n_layers linear layers (tanh between them, none after the last):
 * bf16 operands, fp32 accumulation (tcgen05.mma into TMEM), fp32 bias, fp32 in/out.
 * y, out: (B,256) float32 row-major; weights_bf16: (n_layers,256,256) bfloat16, [layer][out][in]
 * (the layout of torch.nn.Linear.weight); biases_f32: (n_layers,256) float32.
 * This is a user-level f (terms.py:60-63 calls it), not part of the solver arithmetic. */
int tode_mlp_tanh256_forward(const void* y, const void* weights_bf16, const void* biases_f32,
                             void* out, int64_t B, int32_t n_layers, void* stream);

/* Stage-fused evaluation: out = MLP(y_i) with y_i = st->y + st->dt * sum_{j<stage} a[stage][j] k[j]
 * (runge_kutta.py:259-263) formed while the kernel loads its activation tile -- one launch instead of
 * tode_erk_stage + tode_mlp_tanh256_forward, same bits (y_i is never written unless y_out != NULL:
 * pass the last stage's y_out, which tode_erk_finish needs as y1).  st->F == 256, data dtype f32;
 * k[j], y_out, out 16-byte aligned; a no-op once st->ctl (if non-NULL) carries the stop flag. */
int tode_mlp_tanh256_stage_forward(const tode_tableau* tab, int stage, const tode_state* st,
                                   const void* const* k, void* y_out, const void* weights_bf16,
                                   const void* biases_f32, void* out, int32_t n_layers, void* stream);

/* Step-fused evaluation (ABI 3): ALL stages of one explicit RK step around the MLP field in ONE launch --
 * for i = 1 .. n_stages-1:  y_i = y + dt * sum_{j<i} a[i][j] k[j],  k[i] = f(y_i)   (runge_kutta.py:259-263;
 * f autonomous).  k[0] = f(t, y) on entry (FSAL slot), k[1 .. n_stages-1] are outputs ((B,256) fp32 each);
 * y1_out (may be NULL) receives y_{n_stages-1} (the step's y1 for an SSAL tableau).  f acts row by row, so
 * a CTA runs the stages of its rows back to back with no grid-wide synchronisation.  Same bits as
 * n_stages-1 calls of tode_mlp_tanh256_stage_forward. */
int tode_mlp_tanh256_step_forward(const tode_tableau* tab, const tode_state* st, void* const* k,
                                  void* y1_out, const void* weights_bf16, const void* biases_f32,
                                  int32_t n_layers, void* stream);

/* Multi-GPU (ABI 3): ship `bytes` (multiple of 16) from `src` (this GPU's memory) to n_dst peer buffers
 * (pointers valid on THIS GPU: peer memory mapped over NVLink) by SM stores -- every vector is read once and
 * stored to all peers, all links busy at once.  For the dense-output block of a sharded solve (the all-gather
 * of solutions); the caller's cross-GPU barrier afterwards covers arrival. */
int tode_peer_push(const void* src, void* const* dst, int32_t n_dst, int64_t bytes, void* stream);

/* Method-of-lines vector field of the 1-D heat equation with Dirichlet ends (configs[4]), one
 * HBM pass: out[b,i] = kappa * ((y[b,i+1] - 2 y[b,i]) + y[b,i-1]) for 0 < i < N-1, 0 at the ends;
 * y, out (B,N) row-major, 16-byte aligned, N divisible by 4 (f32) / 2 (f64).  A user-level f like
 * the MLP field. */
int tode_heat1d_forward(const void* y, void* out, int64_t B, int64_t N, double kappa, int32_t dtype,
                        void* stream);

/* One whole loop iteration (runge_kutta.py:227-279, step_size_controllers.py:371-429 / 716-774,
 * adjoints.py:150-234) of a problem whose f is the heat field above (t_eval rows, if any, monotone in the
 * direction of time: cursor mode, st->not_yet == NULL): the six stage
 * combinations, the six stencil evaluations, the error estimate and the per-chunk error norms in ONE
 * pass over y and the FSAL slot (stage values never leave the SM), then the per-sample controller --
 * two launches, 4 rows of HBM traffic instead of 56.  Replaces 6 x (tode_erk_stage,
 * tode_heat1d_forward) + tode_erk_finish; same bits.  The state of sample b lives in (st->y, st->f0)
 * while sel[b] == 0 and in (y_alt, f_alt) while sel[b] == 1: a step reads one pair, writes y1 and
 * k[S-1] into the other, and accepting it toggles sel[b] (no commit copy).  sel: (B) bytes, zeroed by
 * the caller before the first iteration; y_alt, f_alt: (B,F), 16-byte aligned; F divisible by 4 (f32)
 * / 2 (f64); st->scratch as for tode_erk_finish.  st->y_eval receives the dense output of the points a
 * step covers if it is accepted (T == 0: the value at t_end of every step that reaches t_end), written
 * before the accept decision is known and rewritten by the step that does cover them; a step that ends
 * with status != SUCCESS has no end-point value unless it reaches t_end: if any sample fails,
 * st->y_eval is not meaningful and the caller must re-solve through the stage-wise entry points. */
int tode_heat_step(const tode_tableau* tab, const tode_controller* ctrl, const tode_state* st,
                   double kappa, void* y_alt, void* f_alt, uint8_t* sel, void* stream);

/* Measurement aid for bench.py (not on the solve path): every thread of a machine-filling
 * grid runs `iters` rounds of 8 independent double-precision FMA chains; writes one double
 * per thread to `sink` (at least tode_bench_fp64_fma_threads() doubles).  Returns the number
 * of FMAs executed through *n_fma_out (host pointer). */
int64_t tode_bench_fp64_fma_threads(void);
int tode_bench_fp64_fma(int64_t iters, void* sink, int64_t* n_fma_out, void* stream);

/* Test aid (not on the solve path): checks on `n` pseudo-random and adversarial operands that
 * the branch-free scalar functions of the fused kernel (division, log2, exp2, controller) return
 * the bits of the checked functions wherever their range flag stays set.  counts8: 8 x uint64
 * on the device, zeroed by the caller; [0..3] receive the number of mismatches per function,
 * [4..7] how often the fast flag stayed set. */
int tode_selftest_fast_math(int64_t n, uint64_t seed, const tode_controller* ctrl, void* counts8, void* stream);
/* Test aid: the fp32 fast-path division / square root of the packed kernels (erk_fused_f2.cuh: div_fast,
 * div_fast2, sqrt_fast) against div.rn.f32 / sqrt.rn.f32 on n pseudo-random and adversarial operands;
 * counts6 = device uint64[6] (zeroed by the caller): [0..2] mismatches (division, division by sqrt(2),
 * square root) where the range flag is set, [3..5] how often it was set. */
int tode_selftest_fast_math_f32(int64_t n, uint64_t seed, void* counts6, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TORCHODE_B200_H */
