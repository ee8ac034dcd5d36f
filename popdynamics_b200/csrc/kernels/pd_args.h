/*
 * pd_args.h -- launch descriptors shared by the host C-ABI (csrc/popdyn_capi.cu) and the device
 * kernels (compiled by NVRTC).  Plain C, fixed layout: only 8-byte members first, then 4-byte
 * members in pairs, so host and device agree without packing pragmas.  The model-specific uniform
 * parameter block (PdUniform, generated) is appended AFTER this fixed part in the kernel argument.
 */
#ifndef PD_ARGS_H
#define PD_ARGS_H

#define PD_OUT_FULL 0   /* soln  [T][C][M] fp64, member fastest (coalesced stores)              */
#define PD_OUT_FINAL 1  /* final [C][M] fp64                                                     */
#define PD_OUT_STATS 2  /* streaming moments u64 [T][C][2] + histograms u32 [rows][C][bins]      */
#define PD_OUT_FINAL_VARS 3 /* ODE kernels: diagnostics vars of the final state [PD_NFVAR][M]    */

#define PD_RNG_PHILOX 0 /* Philox4x32-10, key = seed, counter = (draw block, global member id)   */
#define PD_RNG_INJECT 1 /* uniforms[n_uniform][M] fp64 consumed in order (parity tests)          */

#define PD_METHOD_EXPLICIT 0
#define PD_METHOD_TAU_LEAP 1
#define PD_METHOD_GILLESPIE 2
#define PD_METHOD_DIAGNOSTICS 3
#define PD_METHOD_ADAPTIVE 4

#define PD_MAX_HIST_COMP 64

/* error codes (also oracle/popdyn_oracle.h) */
#define PD_ERR_CHECKS 1        /* checks() assertion failed -> AssertionError on the host        */
#define PD_ERR_POISSON_MEAN 2  /* negative / NaN Poisson mean (numpy raises ValueError)          */
#define PD_ERR_ZERO_RATE 3     /* total event rate 0 with events present (ZeroDivisionError)     */
#define PD_ERR_STREAM 4        /* injected uniform stream exhausted                              */
#define PD_ERR_LOG0 5          /* log(0) in the waiting time (math domain error)                 */
#define PD_ERR_COUNT_RANGE 6   /* count too large for the statistics accumulators                */
#define PD_ERR_MAX_STEPS 7     /* adaptive solver: step budget of a member exhausted             */
#define PD_ERR_STEP_SIZE 8     /* adaptive solver: step size underflow (t + h == t)              */
#define PD_ERR_SQUEEZE 9       /* PD_PTRS_CHECK builds: squeeze and exact PTRS test disagree     */

typedef struct {
  long long n_member;          /* members in this launch                                         */
  long long member0;           /* global id of member 0 (Philox counter; shard invariance)       */
  const double *y0;            /* initial compartments                                           */
  long long y0_stride_c;       /* element stride between compartments                            */
  long long y0_stride_m;       /* element stride between members (0 = shared)                    */
  const double *pm;            /* per-member parameters [PD_NPM][M]                              */
  const double *target;        /* target times [T]                                               */
  unsigned long long seed;
  const double *uniforms;      /* injected stream [n_uniform][M] or NULL                         */
  long long n_uniform;
  double *soln;                /* FULL                                                           */
  double *final_state;         /* FINAL                                                          */
  unsigned long long *moments; /* STATS: [T][C][2] (sum x, sum x^2), exact integers              */
  unsigned int *hist;          /* STATS: [n_hist_row][C][hist_bins] or NULL                      */
  const int *hist_row;         /* [T]: histogram row recorded at each target index, -1 = none    */
  unsigned long long *n_steps; /* total steps / events executed (atomic)                         */
  int *err;                    /* [4]: code, member lo, member hi, target index                  */
  int n_time;
  int hist_bins;               /* even; 0 = no histograms                                        */
  int hist_shift[PD_MAX_HIST_COMP]; /* bin = min(count >> shift[c], bins - 1)                    */
  unsigned int philox_rk[20];  /* Philox round keys (k0, k1) of rounds 0..9, filled by the host
                                  from `seed`: operands straight from the constant bank          */
  const double *code_table;    /* tau-leap: [PD_NTAB][code_table_n] Poisson codes by count, built
                                  by pd_tau_table_kernel when every step has the same dt; NULL
                                  otherwise                                                      */
  double code_table_dt;        /* the dt the table was built for                                 */
  const unsigned char *mask;   /* [M] or NULL: only members with mask[m] == mask_value run, the
                                  others are left untouched (hybrid switching, one launch per
                                  integrator over the same state)                                */
  int code_table_n;
  int mask_value;
  int moments_from_hist;       /* 1: the caller derives the moments from the (lossless: shift 0,
                                  every time recorded) histograms afterwards; the kernel skips
                                  their reduction and raises PD_ERR_COUNT_RANGE for a count that
                                  lands in the last (overflow) bin                               */
  int pad2;
  long long pm_ld;             /* row stride (elements) of pm: M, or the chunk size when the host
                                  staged this launch's member range on its own                   */
  long long out_ld;            /* row stride of soln / final_state                               */
  const double *code_table2;   /* tau-leap: [PD_NTAB2][code_table2_nb][code_table2_na] Poisson
                                  codes of the count x count x uniform-parameter rows (a force
                                  of infection: susceptibles x infectious x beta), for the same
                                  dt as code_table; NULL when not built                          */
  int code_table2_na;          /* counts of the row's own compartment covered                    */
  int code_table2_nb;          /* counts of the compartment inside the rate                      */
} PdStochArgs;

/* The tau-leap kernel keeps per-warp partial moment sums in shared memory and moves them to
 * global memory once per EPOCH of PD_EPOCH recorded target times (pd_common.cuh).              */
#ifndef PD_EPOCH
#define PD_EPOCH 16
#endif

/* Dynamic shared memory of the stochastic kernels (each kernel exports its total as the device
 * symbol pd_smem_bytes, which the host reads after loading the module):
 *   tau-leap : the draw list     double   [n_row][block]   (n_row = flow-table rows that can fire)
 *   Gillespie, FULL / STATS output: the recording window  count_t [slots][n_comp][block]       */
#define PD_ALIGN8(x) (((x) + 7) & ~(size_t)7)
#define PD_CODE_SMEM_BYTES(n_row, block) ((size_t)((n_row) > 0 ? (n_row) : 1) * (size_t)(block) * 8)
typedef struct {
  long long n_member;
  const double *y0;
  long long y0_stride_c;
  long long y0_stride_m;
  const double *pm;            /* per-member parameters [PD_NPM][M]                              */
  const int *n_sub;            /* [T-1] sub-steps per target interval (host replay of :675-682)  */
  const double *dt;            /* [n_steps] step size of every sub-step                          */
  const double *t_eval;        /* [n_steps] evaluation times (only read when PD_USES_TIME)       */
  const double *su_table;      /* [n_steps][PD_NSU] scale-up values (host-evaluated closures)    */
  double *soln;                /* FULL  [T][C][M]                                                */
  double *final_state;         /* FINAL [C][M]                                                   */
  int *err;
  const unsigned char *mask;   /* [M] or NULL: only members with mask[m] == mask_value run       */
  int n_time;
  int mask_value;
  long long member0;           /* global id of member 0 of this launch (error reports)           */
  long long pm_ld;             /* row stride (elements) of pm                                    */
  long long out_ld;            /* row stride of soln / final_state                               */
  double t_final;              /* last target time (PD_OUT_FINAL_VARS: time of the diagnostics)  */
  long long n_steps;           /* sub-steps in total (PD_OUT_FINAL_VARS: last scale-up row)      */
} PdExplicitArgs;

typedef struct {
  long long n_member;
  long long member0;           /* global id of member 0 of this launch (error reports)           */
  const double *y0;
  long long y0_stride_c;
  long long y0_stride_m;
  const double *pm;            /* per-member parameters [PD_NPM][.]                              */
  long long pm_ld;
  const double *target;        /* target times [T]                                               */
  double *soln;                /* FULL  [T][C][.]                                                */
  double *final_state;         /* FINAL [C][.]                                                   */
  long long out_ld;
  unsigned long long *n_steps; /* attempted steps, total (atomic)                                */
  int *err;
  double rtol, atol;           /* error tolerances (odeint's defaults: 1.49012e-8 both)          */
  double first_step;           /* 0: estimated per member                                        */
  long long max_steps;         /* attempted-step budget per member                               */
  int n_time;
  int pad;
} PdAdaptiveArgs;

typedef struct {
  long long n_member;
  const double *soln;          /* [T][C][M] stored compartments                                  */
  const double *pm;            /* per-member parameters [PD_NPM][M]                              */
  const double *times;         /* [T] stored times                                               */
  const double *su;            /* [PD_NSU] scale-up values of the last derivative call, or NULL  */
  double *var_soln;            /* [T][PD_NDIAG][M]                                               */
  double *flow_soln;           /* [T][C][M] or NULL                                              */
  int n_time;
  int pad;
  long long soln_ld;           /* row strides (elements) of soln, pm, and var_soln / flow_soln   */
  long long pm_ld;
  long long out_ld;
} PdDiagArgs;

#endif
