/*
 * popdyn_b200.h -- C-ABI of libpopdyn_b200.so, the B200 engine behind popdynamics' BaseModel
 * integrators.
 *
 * The reference (popdynamics/popdynamics) is pure Python and has no FFI layer: the drop-in
 * boundary is the Python class API of basepop.BaseModel (reference basepop.py:184-558,
 * integrate() at :854-903).  This C-ABI therefore sits UNDER that class: it is exactly what a
 * ctypes binding inside the reference's `integrate` would call instead of its three Python loops.
 * Each entry point cites the reference code it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every data pointer may be HOST memory (numpy) or
 * DEVICE memory (torch / cudaMalloc) -- the library looks at the pointer's attributes, stages
 * host buffers through the device inside the call and leaves device buffers in place.  All
 * functions return 0 on success or a negative pd_status; pd_last_error() gives the message of
 * the calling thread's last failure.  Handles are thread-compatible, not thread-safe.  There is
 * no CPU fallback: without a CUDA device the run functions fail with PD_E_NO_DEVICE.
 *
 * Staging of HOST per-member buffers (y0 [C][M], member_params, soln / final_state, mask): when
 * they add up to more than about 192 MB the member range is cut into chunks that flow through
 * two streams -- H2D(c+1) and D2H(c-1) overlap the kernel of chunk c -- with one
 * cudaMemcpy2DAsync per buffer per chunk, so a 10^8-member sweep never needs its inputs resident
 * at once and the copies hide behind the integration.  pd_run_result.n_launch is the number of
 * chunks; kernel_ms is then the span of the pipelined region.  Events, the error block and the
 * staging scratch are kept by the handle between runs (nothing is created per call).
 */
#ifndef POPDYN_B200_H
#define POPDYN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PD_OK = 0,
  PD_E_INVALID = -1,   /* bad argument                                                        */
  PD_E_COMPILE = -2,   /* NVRTC rejected the generated model (log in pd_last_error)           */
  PD_E_CUDA = -3,      /* CUDA runtime / driver error                                         */
  PD_E_NO_DEVICE = -4, /* no usable GPU                                                       */
  PD_E_MODEL = -5      /* device-side error flag raised (see pd_run_result.err_*)             */
} pd_status;

/* integrator selector: which reference loop the compiled kernel replaces */
enum { PD_EXPLICIT = 0,   /* integrate_explicit,                      basepop.py:658-688      */
       PD_TAU_LEAP = 1,   /* integrate_discrete_time_stochastic,      basepop.py:797-852      */
       PD_GILLESPIE = 2,  /* integrate_continuous_time_stochastic,    basepop.py:737-795      */
       PD_DIAGNOSTICS = 3,/* calculate_diagnostics over stored times,  basepop.py:911-960     */
       PD_ADAPTIVE = 4 }; /* integrate(method='scipy'): odeint on the derivative closure,
                             basepop.py:882-886 -- adaptive Dormand-Prince 5(4) per member     */
/* output selector */
enum { PD_FULL = 0,       /* soln [T][C][M] fp64 (the reference's soln_array per member)      */
       PD_FINAL = 1,      /* final state [C][M] fp64                                          */
       PD_STATS = 2,      /* streaming ensemble moments + histograms (stochastic only)        */
       PD_FINAL_VARS = 3 };/* ODE integrators: the diagnostics pass (basepop.py:927-948) of the
                             FINAL state fused into the kernel -- final_state receives
                             [n_final_var][M] vars (e.g. model.vars['proportion'],
                             examples/rmit_tb_model.py:143) instead of the compartments       */
/* uniform source for the stochastic integrators */
enum { PD_PHILOX = 0,     /* native counter-based Philox4x32-10                               */
       PD_INJECT = 1 };   /* caller-supplied uniform stream replacing random.random (:174,
                             :776) and the uniforms inside numpy.random.poisson (:830)        */

typedef struct pd_model pd_model; /* one NVRTC-specialised kernel: (model graph, integrator,
                                     output, rng, count width, block size)                    */

typedef struct {
  int32_t method, out_mode, rng_mode;
  int32_t count_bits;     /* 64 (default) or 32: width of the integer compartment counts      */
  int32_t block;          /* threads per block, multiple of 32 (0 -> 256)                     */
  int32_t device;         /* CUDA ordinal                                                     */
  int32_t n_comp, n_param, n_member_param, n_scaleup, n_flow; /* must match the model header  */
  int32_t min_blocks;     /* __launch_bounds__ second argument: resident blocks per SM the
                             register allocation must allow (0 -> 1, i.e. no cap)             */
  int32_t flags;          /* PD_FLAG_* bits                                                   */
  int32_t reserved;
} pd_model_desc;

#define PD_FLAG_MEMBER_MASK 1   /* the kernel honours pd_*_run.mask (a separate specialisation:
                                   testing for a mask in the general kernel costs 6 % of the
                                   tau-leap step)                                              */
#define PD_FLAG_HIST_MOMENTS 2  /* PD_STATS kernels without moment accumulators: every run of
                                   the model sets pd_stoch_run.moments_from_hist and derives the
                                   moments with pd_hist_moments (a separate specialisation: the
                                   in-kernel moments are warp-collective, which keeps idle lanes
                                   and a per-step activity test in the loop)                    */

typedef struct {
  double kernel_ms;       /* device time of the integrator kernel (CUDA events)               */
  double h2d_ms, d2h_ms;  /* staging time when host buffers were passed                       */
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t n_steps;       /* derivative evaluations / tau-leap steps / Gillespie events, total */
  int32_t err_code;       /* 0, or the PD_ERR_* code of the first failing member              */
  int32_t err_where;      /* target index where it was raised (explicit) or 0                 */
  int64_t err_member;     /* global member id                                                 */
  int32_t grid, block, regs_per_thread, smem_bytes, blocks_per_sm, n_launch;
} pd_run_result;

/* Inputs common to the stochastic runs.  y0: [C] (shared, y0_per_member = 0) or [C][M].
 * member_params: [n_member_param][M].  target_times: the host grid produced by make_times
 * (basepop.py:355-376).                                                                      */
typedef struct {
  int64_t n_member, member0;
  const double *y0; int32_t y0_per_member, n_time;
  const double *uniform_params;   /* [n_param] HOST: values of all parameter slots            */
  const double *member_params;
  const double *target_times;     /* [n_time] HOST                                            */
  uint64_t seed;
  const double *uniforms; int64_t n_uniform;           /* PD_INJECT: [n_uniform][M]           */
  double *soln;                   /* PD_FULL  [T][C][M]                                       */
  double *final_state;            /* PD_FINAL [C][M]                                          */
  uint64_t *moments;              /* PD_STATS [T][C][2] = (sum x, sum x^2), exact while they fit
                                     64 bits; ACCUMULATED into (caller zeroes)                */
  uint32_t *hist;                 /* PD_STATS [n_hist_row][C][hist_bins] or NULL, accumulated */
  const int32_t *hist_row;        /* [T] HOST: row per target index or -1                     */
  const int32_t *hist_shift;      /* [C] HOST: bin = min(count >> shift, bins - 1)            */
  int32_t hist_bins, n_hist_row;
  void *stream;                   /* cudaStream_t or NULL                                     */
  const uint8_t *mask;            /* [M] or NULL: only members with mask[m] == mask_value are
                                     integrated, the outputs of the others are left untouched
                                     (PD_FULL / PD_FINAL only)                                */
  int32_t mask_value;
  int32_t moments_from_hist;      /* 1: skip the in-kernel moment reduction; the caller runs
                                     pd_hist_moments on the (lossless) histograms afterwards:
                                     needs hist at every target time with hist_shift == 0; a
                                     count in the last bin raises the device error flag         */
  const double *const *member_param_rows; /* [n_member_param] HOST array of row pointers, each to
                                     [M] doubles (host or device): the swept parameter columns as
                                     the caller holds them, instead of one [P][M] block; used when
                                     member_params is NULL.  A NULL row is a column the compiled
                                     kernel never reads: nothing is staged for it               */
} pd_stoch_run;

typedef struct {
  int64_t n_member;
  const double *y0; int32_t y0_per_member, n_time;
  const double *uniform_params;   /* [n_param] HOST                                           */
  const double *member_params;    /* [n_member_param][M]                                      */
  const int32_t *n_sub;           /* [T-1] HOST   } from pd_explicit_schedule                 */
  const double *dt;               /* [n_steps] HOST                                           */
  const double *t_eval;           /* [n_steps] HOST (may be NULL when the model ignores time) */
  const double *scaleup_table;    /* [n_steps][n_scaleup] HOST or NULL                        */
  int64_t n_steps;
  double *soln;                   /* PD_FULL  [T][C][M]                                       */
  double *final_state;            /* PD_FINAL [C][M]                                          */
  void *stream;
  const uint8_t *mask;            /* as in pd_stoch_run                                       */
  int32_t mask_value;
  int32_t n_final_var;            /* PD_FINAL_VARS: rows of final_state (PD_NFVAR of the header) */
  const double *const *member_param_rows; /* as in pd_stoch_run                                */
  double t_final;                 /* PD_FINAL_VARS: self.time of the diagnostics pass, i.e. the
                                     last target time                                         */
} pd_explicit_run;

/* The diagnostics pass over stored trajectories (model built with method PD_DIAGNOSTICS, the
 * header generated from compile_model(..., diagnostics=True)).                                */
typedef struct {
  int64_t n_member;
  int32_t n_time, n_var;          /* n_var must match the model header (PD_NDIAG)              */
  const double *soln;             /* [T][C][M]: PD_FULL output, or [1][C][M] = a PD_FINAL one  */
  const double *uniform_params;   /* [n_param] HOST                                            */
  const double *member_params;    /* [n_member_param][M]                                       */
  const double *times;            /* [T] HOST                                                  */
  const double *scaleup_row;      /* [n_scaleup] HOST or NULL (explicit: last derivative call) */
  double *var_soln;               /* [T][n_var][M]                                             */
  double *flow_soln;              /* [T][C][M] or NULL                                         */
  void *stream;
  const double *const *member_param_rows; /* as in pd_stoch_run                                */
} pd_diag_run;

/* The adaptive integrator (model built with method PD_ADAPTIVE; scale-up functions are traced on
 * time, so n_scaleup is 0).  Replaces `odeint(derivative, y, self.target_times)`,
 * basepop.py:886: the state of every member at every target time, local error controlled by
 * rtol / atol (odeint's defaults are 1.49012e-8 for both).                                     */
typedef struct {
  int64_t n_member, member0;
  const double *y0; int32_t y0_per_member, n_time;
  const double *uniform_params;   /* [n_param] HOST                                           */
  const double *member_params;    /* [n_member_param][M]                                      */
  const double *const *member_param_rows; /* as in pd_stoch_run                               */
  const double *target_times;     /* [n_time] HOST, increasing                                */
  double rtol, atol;
  double first_step;              /* 0: estimated per member                                  */
  int64_t max_steps;              /* attempted-step budget per member (PD_ERR_MAX_STEPS)      */
  double *soln;                   /* PD_FULL  [T][C][M]                                       */
  double *final_state;            /* PD_FINAL [C][M], PD_FINAL_VARS [n_final_var][M]          */
  void *stream;
  int32_t n_final_var, reserved;
} pd_adaptive_run;

/* ---- library ---------------------------------------------------------------------------- */
int pd_version(void);
const char *pd_last_error(void);
int pd_device_count(void);
/* sizeof() of the ABI structs as compiled: 0 pd_model_desc, 1 pd_run_result, 2 pd_stoch_run,
 * 3 pd_explicit_run, 4 pd_diag_run, 5 pd_adaptive_run -- lets a foreign-language binding verify
 * its struct layout.                                                                           */
int pd_struct_size(int which);

/* ---- flow compiler back end -------------------------------------------------------------- */
/* Compile the generated model header (popdynamics_b200/codegen.py) together with the
 * hand-written kernel for desc->method.  Needs no GPU (NVRTC only).  Replaces nothing in the
 * reference: it is the lowering of set_flows()/calculate_vars() (basepop.py:553, :628).       */
int pd_model_create(const char *model_header, const pd_model_desc *desc, pd_model **out);
void pd_model_destroy(pd_model *m);
/* size of the sm_100a cubin and, once loaded on a device, registers per thread               */
int pd_model_cubin_size(const pd_model *m, size_t *bytes);
/* copy the cubin (for cuobjdump / offline inspection)                                        */
int pd_model_cubin_copy(const pd_model *m, void *dst, size_t capacity);

/* ---- host helper: sub-step schedule of integrate_explicit, basepop.py:672-682 ------------- */
/* Fills up to `cap` entries of t_eval / dt (any may be NULL) and n_sub[n_time-1]; returns the
 * total number of derivative evaluations in *n_steps.                                        */
int pd_explicit_schedule(int32_t n_time, const double *target_times, double min_dt,
                         double *t_eval, double *dt, int32_t *n_sub, int64_t cap,
                         int64_t *n_steps);

/* ---- the three integrators ---------------------------------------------------------------- */
int pd_run_explicit(pd_model *m, const pd_explicit_run *run, pd_run_result *res);
int pd_run_tau_leap(pd_model *m, const pd_stoch_run *run, pd_run_result *res);
int pd_run_gillespie(pd_model *m, const pd_stoch_run *run, pd_run_result *res);
/* replaces the odeint call of integrate(method='scipy') (basepop.py:882-886) for a batch of
 * members; pd_run_result.n_steps = attempted steps, total */
int pd_run_adaptive(pd_model *m, const pd_adaptive_run *run, pd_run_result *res);
/* replaces BaseModel.calculate_diagnostics (basepop.py:911-960) for a batch of members */
int pd_run_diagnostics(pd_model *m, const pd_diag_run *run, pd_run_result *res);

/* ---- ensemble statistics ------------------------------------------------------------------ */
/* Quantiles from (all-reduced) histograms: for every (row, c) and every q the smallest bin b with
 * cumulative count >= q * total (inverted CDF); value = b << shift[c].  hist / out may be host
 * or device.  out: [n_row][C][n_q] int64.                                                    */
/* moments [n_row][C][2] = (sum x, sum x^2) derived from width-1 histograms (overwritten) */
int pd_hist_moments(const uint32_t *hist, int32_t n_row, int32_t n_comp, int32_t n_bins,
                    uint64_t *moments, int32_t device, void *stream);
int pd_hist_quantiles(const uint32_t *hist, int32_t n_row, int32_t n_comp, int32_t n_bins,
                      const int32_t *hist_shift, const double *q, int32_t n_q, int64_t *out,
                      int32_t device, void *stream);

/* ---- device utilities --------------------------------------------------------------------- */
/* Fill out[n][M] with the Philox uniforms member (member0 + j) would draw: the same stream the
 * integrators consume (tests, and device-side generation of injected streams).               */
int pd_philox_fill(double *out, int64_t n, int64_t n_member, int64_t member0, uint64_t seed,
                   int32_t device, void *stream);
/* Pipe-peak microbenchmarks for the roofline denominators MEASURED_PEAKS.json lacks:
 * kind 0 = fp64 DADD+DMUL (non-fused, what --fmad=false code issues), 1 = fp64 DFMA,
 * 2 = int32 IMAD/LOP3 mix as in Philox (3 operations per element, issued as IMAD.WIDE.U32 + LOP3).
 * Kinds 3, 4, 5 count ELEMENTS per second of a dependent chain of IMAD.HI.U32 + LOP3, IMAD (low
 * half) + LOP3, and SHF + LOP3: they show what one Philox multiply costs on the FMA-heavy pipe.
 * Kinds 6 .. 10 are co-issue probes, elements per second again: 6 = DADD alone, 7 / 8 = DADD with
 * one / two LOP3 on chains of their own, 9 / 10 = IMAD.WIDE.U32 + LOP3 with one / two more LOP3:
 * they show what hides behind a two-cycle fp64 or a four-cycle wide-multiply issue and what adds up.
 * Returns giga-operations (giga-elements) per second.                                          */
int pd_microbench(int32_t kind, int32_t device, double *gops);
/* Compares the kernels' own forms of exp, log and division (csrc/kernels/pd_math.cuh) with the CUDA
 * math library's on `n` generated arguments each: n_bad[0] exp, [1] log on the 53-bit uniform
 * grid, [2] log on raw bit patterns, [3] division; first_bad[k] is one offending argument.  All
 * four counts must be zero: the forms are restatements, not approximations.                  */
int pd_math_selfcheck(int32_t device, int64_t n, uint64_t seed, uint64_t *n_bad, double *first_bad);

#ifdef __cplusplus
}
#endif
#endif
