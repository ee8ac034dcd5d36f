/*
 * popdyn_oracle.h -- CPU restatement of popdynamics' BaseModel integrator loops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under popdynamics_b200/ may include, link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and there only as the checker / the CPU arm.  Every function cites the reference lines
 * (relative to /root/reference) it restates.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks this restatement against golden
 * vectors generated from the unmodified reference (oracle/make_golden.py, fixtures in
 * tests/golden/), including seeded stochastic runs reproduced draw for draw through the MT19937
 * restatements below, and tests/test_oracle_vs_reference.py re-runs the comparison live when
 * /root/reference is present.
 */
#ifndef POPDYN_ORACLE_H
#define POPDYN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- expression program for the traced calculate_vars hook ------------------------------ */
/* Value slots (all fp64):  [0,C) compartments | [C,C+NP) params | C+NP time |
 * [C+NP+1, C+NP+1+NSU) scale-up values | then one slot per instruction, in order.          */
enum {
  PDO_OP_CONST = 0, /* a = index into consts[]                                             */
  PDO_OP_ADD, PDO_OP_SUB, PDO_OP_MUL, PDO_OP_DIV,
  PDO_OP_NEG, PDO_OP_FABS, PDO_OP_FLOOR,
  PDO_OP_LT, PDO_OP_LE, PDO_OP_GT, PDO_OP_GE, PDO_OP_EQ, PDO_OP_NE, /* -> 0.0 / 1.0        */
  PDO_OP_AND, PDO_OP_OR, PDO_OP_NOT,
  PDO_OP_ISFINITE,
  PDO_OP_SELECT, /* a ? b : c                                                              */
  PDO_OP_EXP     /* libm exp: scale-up curves, basepop.py:122                              */
};

/* flow kinds in the canonical evaluation order, basepop.py:566-598 / :703-735 */
enum { PDO_ENTRY = 0, PDO_VAR = 1, PDO_FIXED = 2, PDO_BGDEATH = 3, PDO_INFDEATH = 4 };

typedef struct {
  int32_t n_comp, n_param, n_su, n_ins, n_flow, n_check, n_post_check, pad1;
  const int32_t *ins;    /* n_ins x 4: op, a, b, c (slot indices)                          */
  const double *consts;
  const int32_t *kind;   /* per flow                                                       */
  const int32_t *from;   /* compartment index, -1 = entry                                  */
  const int32_t *to;     /* compartment index, -1 = death                                  */
  const int32_t *slot;   /* VALUE slot holding the rate (var value, or param value)        */
  const int32_t *is_int; /* per flow: 1 when the Python value of the rate is an int (only
                            possible for entry flows); steers Python's sum() path, :775    */
  const int32_t *check;  /* value slots that must be non-zero (traced checks() asserts)    */
  const int32_t *post_check; /* n_post_check value slots of a user-defined checks() traced for
                            the stochastic loops (basepop.py:789, :846): evaluated AFTER the
                            update, on the updated compartments (a second set of n_comp leaf
                            slots behind the scale-ups) and the vars of the step's start       */
} pdo_model;

/* ---- uniform sources -------------------------------------------------------------------- */
enum { PDO_RNG_MT = 0, PDO_RNG_INJECT = 1, PDO_RNG_PHILOX = 2 };

typedef struct {
  int32_t kind, mti;
  uint32_t mt[624];
  /* injected stream */
  const double *u; int64_t n_u, stride, cursor;
  /* philox4x32-10 */
  uint32_t key[2], ctr[4]; double spare; int32_t has_spare, exhausted;
} pdo_rng;

void pdo_rng_seed_numpy(pdo_rng *r, uint32_t seed);       /* numpy.random.seed(int)        */
void pdo_rng_seed_python(pdo_rng *r, uint64_t seed);      /* random.seed(int)              */
void pdo_rng_inject(pdo_rng *r, const double *u, int64_t n, int64_t stride);
void pdo_rng_philox(pdo_rng *r, uint64_t seed, uint64_t member);
double pdo_rng_double(pdo_rng *r);
void pdo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int64_t pdo_poisson(pdo_rng *r, double lam);               /* numpy legacy poisson          */
int32_t pdo_pick_event(const double *intervals, int32_t n, double u); /* basepop.py:161-178 */

/* ---- integrators ------------------------------------------------------------------------ */
/* error codes written to *err: 0 ok, 1 checks() failed, 2 bad Poisson mean, 3 zero total
 * rate, 4 uniform stream exhausted, 5 log(0)                                               */
int64_t pdo_explicit_schedule(int32_t n_time, const double *target, double min_dt,
                              double *t_eval, double *dt, int32_t *interval, int64_t cap);
int32_t pdo_explicit_first_failure(const pdo_model *m, const double *y0, const double *params,
                                   int32_t n_time, const double *target, double min_dt,
                                   const double *su_table);
int64_t pdo_explicit(const pdo_model *m, const double *y0, const double *params,
                     int32_t n_time, const double *target, double min_dt,
                     const double *su_table, double *soln, int32_t *err);
int64_t pdo_tau_leap(const pdo_model *m, const double *y0, const double *params,
                     int32_t n_time, const double *target, pdo_rng *rng,
                     double *soln, int32_t *err);
int64_t pdo_gillespie(const pdo_model *m, const double *y0, const double *params,
                      int32_t n_time, const double *target, pdo_rng *rng,
                      double *soln, int32_t *err);
void pdo_eval_vars2(const pdo_model *m, const double *y, const double *y2, const double *params,
                    double time, const double *su, double *v);
void pdo_eval_vars(const pdo_model *m, const double *y, const double *params, double time,
                   const double *su, double *values);

/* batch drivers (OpenMP over members) used by the CPU baseline; y0/params have member stride
 * 0 (shared) or n_comp / n_param; soln may be NULL (then only moments are accumulated).     */
int64_t pdo_tau_leap_batch(const pdo_model *m, int64_t n_member, int64_t member0,
                           const double *y0, int64_t y0_stride,
                           const double *params, int64_t p_stride,
                           int32_t n_time, const double *target, uint64_t seed,
                           double *soln /* [M][T][C] or NULL */,
                           int64_t *sum /* [T][C] */, int64_t *sumsq /* [T][C] */,
                           int32_t n_thread, int32_t *err);
int64_t pdo_gillespie_batch(const pdo_model *m, int64_t n_member, int64_t member0,
                            const double *y0, int64_t y0_stride,
                            const double *params, int64_t p_stride,
                            int32_t n_time, const double *target, uint64_t seed,
                            double *soln, int64_t *sum, int64_t *sumsq,
                            int32_t n_thread, int32_t *err);
int64_t pdo_explicit_batch(const pdo_model *m, int64_t n_member,
                           const double *y0, int64_t y0_stride,
                           const double *params, int64_t p_stride,
                           int32_t n_time, const double *target, double min_dt,
                           const double *su_table, double *soln /* [M][T][C] */,
                           double *final_state /* [M][C] or NULL */,
                           int32_t n_thread, int32_t *err);
int32_t pdo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
