/*
 * popdyn_oracle.c -- CPU restatement of popdynamics' BaseModel integrator loops.
 *
 * TEST INFRASTRUCTURE ONLY (see popdyn_oracle.h).  Plain C99 + libm (+ OpenMP for the batch
 * drivers).  Citations are file:line relative to /root/reference.
 *
 * Third-party arithmetic the reference reaches outside its own tree, restated here from the
 * published algorithms (none of it is vendored in /root/reference):
 *   - numpy legacy RandomState.poisson / random_sample  (requirements.txt:1 "numpy>=1.10.1",
 *     installed 2.3.5): multiplication method for lam < 10, Hoermann PTRS for lam >= 10.
 *   - CPython `random.random`, `random.seed(int)`       (MT19937, 53-bit doubles).
 *   - CPython >= 3.12 builtin `sum` over floats          (Neumaier compensated summation).
 *   - Philox4x32-10 (Salmon et al., SC'11) is NOT used by the reference; it is the native
 *     generator of the CUDA kernels and restated here so that kernel and oracle can be compared
 *     draw for draw.
 */
#include "popdyn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura 1998), as used by numpy's legacy RandomState and CPython.
 * ---------------------------------------------------------------------------------------- */
static void mt_init_genrand(pdo_rng *r, uint32_t s) {
  r->mt[0] = s;
  for (int i = 1; i < 624; i++)
    r->mt[i] = 1812433253u * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = 624;
}

static void mt_init_by_array(pdo_rng *r, const uint32_t *key, int klen) {
  mt_init_genrand(r, 19650218u);
  int i = 1, j = 0;
  int k = 624 > klen ? 624 : klen;
  for (; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
    i++; j++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
    if (j >= klen) j = 0;
  }
  for (k = 623; k; k--) {
    r->mt[i] = (r->mt[i] ^ ((r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
    i++;
    if (i >= 624) { r->mt[0] = r->mt[623]; i = 1; }
  }
  r->mt[0] = 0x80000000u;
  r->mti = 624;
}

static uint32_t mt_next32(pdo_rng *r) {
  if (r->mti >= 624) {
    uint32_t *mt = r->mt, y;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    r->mti = 0;
  }
  uint32_t y = r->mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* numpy.random.seed(int): legacy seeding = Knuth's init_genrand on the 32-bit seed.          */
void pdo_rng_seed_numpy(pdo_rng *r, uint32_t seed) {
  memset(r, 0, sizeof(*r));
  r->kind = PDO_RNG_MT;
  mt_init_genrand(r, seed);
}

/* random.seed(int): init_by_array over the little-endian 32-bit words of abs(seed).          */
void pdo_rng_seed_python(pdo_rng *r, uint64_t seed) {
  memset(r, 0, sizeof(*r));
  r->kind = PDO_RNG_MT;
  uint32_t key[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};
  mt_init_by_array(r, key, key[1] ? 2 : 1);
}

void pdo_rng_inject(pdo_rng *r, const double *u, int64_t n, int64_t stride) {
  memset(r, 0, sizeof(*r));
  r->kind = PDO_RNG_INJECT;
  r->u = u; r->n_u = n; r->stride = stride; r->cursor = 0;
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10.  Counter layout shared with the CUDA kernels (csrc/kernels/pd_philox.cuh):
 *   key = (seed_lo, seed_hi);  ctr = (block_lo, block_hi, member_lo, member_hi)
 * Each block yields two 53-bit doubles: (out0,out1) then (out2,out3), numpy's
 * ((a >> 5) * 2^26 + (b >> 6)) / 2^53 packing.
 * ---------------------------------------------------------------------------------------- */
void pdo_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int round = 0; round < 10; round++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void pdo_rng_philox(pdo_rng *r, uint64_t seed, uint64_t member) {
  memset(r, 0, sizeof(*r));
  r->kind = PDO_RNG_PHILOX;
  r->key[0] = (uint32_t)seed; r->key[1] = (uint32_t)(seed >> 32);
  r->ctr[0] = 0; r->ctr[1] = 0;
  r->ctr[2] = (uint32_t)member; r->ctr[3] = (uint32_t)(member >> 32);
}

static double pack53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
}

double pdo_rng_double(pdo_rng *r) {
  switch (r->kind) {
  case PDO_RNG_MT: {
    /* random.random / RandomState.random_sample: two 32-bit outputs -> 53 bits */
    uint32_t a = mt_next32(r), b = mt_next32(r);
    return pack53(a, b);
  }
  case PDO_RNG_INJECT:
    if (r->cursor >= r->n_u) { r->exhausted = 1; return 0.5; }
    return r->u[(r->cursor++) * r->stride];
  default: {
    if (r->has_spare) { r->has_spare = 0; return r->spare; }
    uint32_t o[4];
    pdo_philox4x32_10(r->ctr, r->key, o);
    if (++r->ctr[0] == 0) ++r->ctr[1];
    r->spare = pack53(o[2], o[3]);
    r->has_spare = 1;
    return pack53(o[0], o[1]);
  }
  }
}

/* ------------------------------------------------------------------------------------------
 * numpy legacy Poisson (what basepop.py:830 `numpy.random.poisson(mean, 1)[0]` runs).
 * ---------------------------------------------------------------------------------------- */
static double np_loggam(double x) {
  static const double a[10] = {8.333333333333333e-02, -2.777777777777778e-03,
                               7.936507936507937e-04, -5.952380952380952e-04,
                               8.417508417508418e-04, -1.917526917526918e-03,
                               6.410256410256410e-03, -2.955065359477124e-02,
                               1.796443723688307e-01, -1.39243221690590e+00};
  double x0, x2, gl, gl0;
  int64_t k, n;
  if (x == 1.0 || x == 2.0) return 0.0;
  n = (x < 7.0) ? (int64_t)(7 - x) : 0;
  x0 = x + (double)n;
  x2 = (1.0 / x0) * (1.0 / x0);
  gl0 = a[9];
  for (k = 8; k >= 0; k--) { gl0 *= x2; gl0 += a[k]; }
  gl = gl0 / x0 + 0.5 * 1.8378770664093453e+00 + (x0 - 0.5) * log(x0) - x0;
  if (x < 7.0) {
    for (k = 1; k <= n; k++) { gl -= log(x0 - 1.0); x0 -= 1.0; }
  }
  return gl;
}

static int64_t np_poisson_mult(pdo_rng *r, double lam) {
  int64_t X = 0;
  double enlam = exp(-lam), prod = 1.0;
  for (;;) {
    prod *= pdo_rng_double(r);
    if (prod > enlam) X += 1; else return X;
    if (r->exhausted) return X;
  }
}

static int64_t np_poisson_ptrs(pdo_rng *r, double lam) {
  int64_t k;
  double U, V, us;
  double slam = sqrt(lam), loglam = log(lam);
  double b = 0.931 + 2.53 * slam;
  double a = -0.059 + 0.02483 * b;
  double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  double vr = 0.9277 - 3.6224 / (b - 2);
  for (;;) {
    U = pdo_rng_double(r) - 0.5;
    V = pdo_rng_double(r);
    us = 0.5 - fabs(U);
    k = (int64_t)floor((2 * a / us + b) * U + lam + 0.43);
    if (r->exhausted) return k < 0 ? 0 : k;
    if (us >= 0.07 && V <= vr) return k;
    if (k < 0 || (us < 0.013 && V > us)) continue;
    if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <=
        (-lam + (double)k * loglam - np_loggam((double)(k + 1))))
      return k;
  }
}

int64_t pdo_poisson(pdo_rng *r, double lam) {
  if (lam >= 10) return np_poisson_ptrs(r, lam);
  if (lam == 0) return 0;
  return np_poisson_mult(r, lam);
}

/* basepop.py:161-178 pick_event: sequential cumulative sum, strict `>` linear search.        */
int32_t pdo_pick_event(const double *intervals, int32_t n, double u) {
  double cumul[256];
  double c = 0.0;
  int32_t i_event = 0;
  for (int32_t k = 0; k < n; k++) { c += intervals[k]; cumul[k] = c; }
  double i = u * c;
  while (i > cumul[i_event] && i_event < n - 1) i_event++;
  return i_event;
}

/* CPython >= 3.12 builtin sum(list) with start=0 (Python/bltinmodule.c builtin_sum_impl):
 * leading ints accumulate exactly; the first float is added plainly; further floats go through
 * Neumaier compensation; ints met on the float path are added plainly.  Used for
 * `total_rate = sum(event_rates)` (basepop.py:775).                                          */
static double py_sum(const double *x, const int32_t *is_int, int32_t n) {
  double f = 0.0, c = 0.0;
  int32_t k = 0, on_float_path = 0;
  for (; k < n && is_int[k]; k++) f += x[k]; /* exact: small integers */
  if (k < n) { f = f + x[k]; k++; on_float_path = 1; }
  for (; k < n; k++) {
    double v = x[k];
    if (is_int[k]) { f += v; continue; }
    double t = f + v;
    if (fabs(f) >= fabs(v)) c += (f - t) + v; else c += (v - t) + f;
    f = t;
  }
  if (on_float_path && c != 0.0 && isfinite(c)) f += c;
  return f;
}

/* ------------------------------------------------------------------------------------------
 * Expression program (the traced calculate_vars hook, basepop.py:553 and its overrides).
 * ---------------------------------------------------------------------------------------- */
/* value slots: [compartments | parameters | time | scale-ups | updated compartments (only with
 * post checks) | one per instruction] */
static int n_leaves(const pdo_model *m) {
  return m->n_comp + m->n_param + 1 + m->n_su + (m->n_post_check > 0 ? m->n_comp : 0);
}
static int n_values(const pdo_model *m) { return n_leaves(m) + m->n_ins; }

void pdo_eval_vars(const pdo_model *m, const double *y, const double *params, double time,
                   const double *su, double *v) {
  pdo_eval_vars2(m, y, y, params, time, su, v);
}

/* y: the compartments the hooks saw (start of the step); y2: the updated compartments a traced
 * stochastic checks() looks at */
void pdo_eval_vars2(const pdo_model *m, const double *y, const double *y2, const double *params,
                    double time, const double *su, double *v) {
  int base = 0;
  for (int i = 0; i < m->n_comp; i++) v[base + i] = y[i];
  base += m->n_comp;
  for (int i = 0; i < m->n_param; i++) v[base + i] = params[i];
  base += m->n_param;
  v[base++] = time;
  for (int i = 0; i < m->n_su; i++) v[base + i] = su ? su[i] : 0.0;
  base += m->n_su;
  if (m->n_post_check > 0) {
    for (int i = 0; i < m->n_comp; i++) v[base + i] = y2[i];
    base += m->n_comp;
  }
  for (int k = 0; k < m->n_ins; k++) {
    const int32_t *in = m->ins + 4 * k;
    double a = 0, b = 0, c = 0, r = 0;
    if (in[0] != PDO_OP_CONST) {
      a = v[in[1]];
      b = in[2] >= 0 ? v[in[2]] : 0.0;
      c = in[3] >= 0 ? v[in[3]] : 0.0;
    }
    switch (in[0]) {
    case PDO_OP_CONST: r = m->consts[in[1]]; break;
    case PDO_OP_ADD: r = a + b; break;
    case PDO_OP_SUB: r = a - b; break;
    case PDO_OP_MUL: r = a * b; break;
    case PDO_OP_DIV: r = a / b; break;
    case PDO_OP_NEG: r = -a; break;
    case PDO_OP_FABS: r = fabs(a); break;
    case PDO_OP_FLOOR: r = floor(a); break;
    case PDO_OP_LT: r = a < b; break;
    case PDO_OP_LE: r = a <= b; break;
    case PDO_OP_GT: r = a > b; break;
    case PDO_OP_GE: r = a >= b; break;
    case PDO_OP_EQ: r = a == b; break;
    case PDO_OP_NE: r = a != b; break;
    case PDO_OP_AND: r = (a != 0.0) && (b != 0.0); break;
    case PDO_OP_OR: r = (a != 0.0) || (b != 0.0); break;
    case PDO_OP_NOT: r = !(a != 0.0); break;
    case PDO_OP_ISFINITE: r = isfinite(a) ? 1.0 : 0.0; break;
    case PDO_OP_SELECT: r = (a != 0.0) ? b : c; break;
    case PDO_OP_EXP: r = exp(a); break;
    default: r = NAN;
    }
    v[base + k] = r;
  }
}

/* a user-defined checks() in the stochastic loops: 1 when an assertion fails */
static int run_post_checks(const pdo_model *m, const int64_t *y_new, const double *y_old,
                           const double *params, double time, double *v, double *scratch) {
  if (m->n_post_check <= 0) return 0;
  for (int i = 0; i < m->n_comp; i++) scratch[i] = (double)y_new[i];
  pdo_eval_vars2(m, y_old, scratch, params, time, 0, v);
  for (int k = 0; k < m->n_post_check; k++)
    if (!(v[m->post_check[k]] != 0.0)) return 1;
  return 0;
}

static int run_checks(const pdo_model *m, const double *v) {
  for (int k = 0; k < m->n_check; k++)
    if (!(v[m->check[k]] != 0.0)) return 1;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Explicit Euler: basepop.py:658-688, derivative basepop.py:613-626, flows basepop.py:560-598.
 * ---------------------------------------------------------------------------------------- */

/* The sub-step schedule of integrate_explicit is member independent: fp64 time accumulation,
 * clamp to the target on overshoot (basepop.py:675-682).  Returns the number of derivative
 * evaluations; fills t_eval/dt/interval up to `cap` entries (any may be NULL).               */
int64_t pdo_explicit_schedule(int32_t n_time, const double *target, double min_dt,
                              double *t_eval, double *dt_out, int32_t *interval, int64_t cap) {
  int64_t n = 0;
  double time = target[0];
  for (int32_t i = 1; i < n_time; i++) {
    double new_time = target[i];
    while (time < new_time) {
      double old_time = time;
      double dt = min_dt;
      time = time + min_dt;
      if (time > new_time) { dt = new_time - old_time; time = new_time; }
      if (n < cap) {
        if (t_eval) t_eval[n] = old_time;
        if (dt_out) dt_out[n] = dt;
        if (interval) interval[n] = i;
      }
      n++;
    }
  }
  return n;
}

static void calc_flows(const pdo_model *m, const double *y, const double *v, double *flows) {
  /* basepop.py:566-598, literal order: zero, entry, var, fixed, background death, infection death */
  for (int i = 0; i < m->n_comp; i++) flows[i] = 0.;
  for (int e = 0; e < m->n_flow; e++) {
    double rate = v[m->slot[e]];
    if (m->kind[e] == PDO_ENTRY) {
      flows[m->to[e]] += rate;
    } else {
      double val = y[m->from[e]] * rate;
      flows[m->from[e]] -= val;
      if (m->to[e] >= 0) flows[m->to[e]] += val;
    }
  }
}

/* Target index of the interval in which checks() first fails (basepop.py:623 inside the loop of
 * :672-688), 0 when it never does: the index the device error flag has to report.             */
int32_t pdo_explicit_first_failure(const pdo_model *m, const double *y0, const double *params,
                                   int32_t n_time, const double *target, double min_dt,
                                   const double *su_table) {
  const int C = m->n_comp;
  double *v = (double *)malloc(sizeof(double) * (size_t)n_values(m));
  double *y = (double *)malloc(sizeof(double) * (size_t)C);
  double *f = (double *)malloc(sizeof(double) * (size_t)C);
  int64_t n_step = 0;
  int32_t where = 0;
  for (int i = 0; i < C; i++) y[i] = y0[i];
  double time = target[0];
  for (int32_t it = 1; it < n_time && !where; it++) {
    double new_time = target[it];
    while (time < new_time) {
      pdo_eval_vars(m, y, params, time, su_table ? su_table + n_step * m->n_su : 0, v);
      calc_flows(m, y, v, f);
      if (run_checks(m, v)) { where = it; break; }
      double old_time = time;
      time = time + min_dt;
      double dt = min_dt;
      if (time > new_time) { dt = new_time - old_time; time = new_time; }
      for (int i = 0; i < C; i++) {
        y[i] = y[i] + dt * f[i];
        if (y[i] < 0.) y[i] = 0.;
      }
      n_step++;
    }
  }
  free(v); free(y); free(f);
  return where;
}

int64_t pdo_explicit(const pdo_model *m, const double *y0, const double *params,
                     int32_t n_time, const double *target, double min_dt,
                     const double *su_table, double *soln, int32_t *err) {
  const int C = m->n_comp;
  double *v = (double *)malloc(sizeof(double) * (size_t)n_values(m));
  double *y = (double *)malloc(sizeof(double) * (size_t)C);
  double *f = (double *)malloc(sizeof(double) * (size_t)C);
  int64_t n_step = 0;
  int32_t e = 0;
  for (int i = 0; i < C; i++) { y[i] = y0[i]; soln[i] = y0[i]; }   /* :670-671 */
  double time = target[0];
  for (int32_t it = 1; it < n_time; it++) {                          /* :672 */
    double new_time = target[it];
    while (time < new_time) {                                        /* :675 */
      /* derivative_fn(y, time): :618-624 */
      pdo_eval_vars(m, y, params, time, su_table ? su_table + n_step * m->n_su : 0, v);
      calc_flows(m, y, v, f);
      if (!e && run_checks(m, v)) e = 1;                             /* :623 */
      double old_time = time;
      time = time + min_dt;                                          /* :678 */
      double dt = min_dt;
      if (time > new_time) { dt = new_time - old_time; time = new_time; } /* :680-682 */
      for (int i = 0; i < C; i++) {
        y[i] = y[i] + dt * f[i];                                     /* :684 */
        if (y[i] < 0.) y[i] = 0.;                                    /* :686-687 */
      }
      n_step++;
    }
    for (int i = 0; i < C; i++) soln[(size_t)it * C + i] = y[i];     /* :688 */
  }
  free(v); free(y); free(f);
  if (err) *err = e;
  return n_step;
}

/* ------------------------------------------------------------------------------------------
 * Event table: basepop.py:690-735.  Entry events are always present; all others only when
 * val > 0.  Returns the number of live events; ev_* arrays are indexed by live position.
 * ---------------------------------------------------------------------------------------- */
static int calc_events(const pdo_model *m, const int64_t *y, const double *v,
                       int32_t *ev_flow, double *ev_rate, int32_t *ev_is_int) {
  int n = 0;
  for (int e = 0; e < m->n_flow; e++) {
    double rate = v[m->slot[e]];
    if (m->kind[e] == PDO_ENTRY) {
      ev_flow[n] = e; ev_rate[n] = rate; ev_is_int[n] = m->is_int ? m->is_int[e] : 0; n++;
    } else {
      double val = (double)y[m->from[e]] * rate;
      if (val > 0) { ev_flow[n] = e; ev_rate[n] = val; ev_is_int[n] = 0; n++; }
    }
  }
  return n;
}

/* Discrete-time stochastic (clamped Poisson tau-leap): basepop.py:797-852.                   */
int64_t pdo_tau_leap(const pdo_model *m, const double *y0, const double *params,
                     int32_t n_time, const double *target, pdo_rng *rng,
                     double *soln, int32_t *err) {
  const int C = m->n_comp, S = m->n_flow;
  double *v = (double *)malloc(sizeof(double) * (size_t)n_values(m));
  double *yd = (double *)malloc(sizeof(double) * (size_t)C);
  double *y2 = (double *)malloc(sizeof(double) * (size_t)C);
  int64_t *y = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
  int32_t *ev_flow = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
  int32_t *ev_int = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
  double *ev_rate = (double *)malloc(sizeof(double) * (size_t)(S + 1));
  int32_t e = 0;
  int64_t n_step = 0;
  for (int i = 0; i < C; i++) { y[i] = (int64_t)y0[i]; if (soln) soln[i] = y0[i]; } /* :805-806, :813 */
  double time = target[0];
  for (int32_t it = 1; it < n_time; it++) {
    double dt = target[it] - target[it - 1];                        /* :820 */
    for (int i = 0; i < C; i++) yd[i] = (double)y[i];
    pdo_eval_vars(m, yd, params, time, 0, v);                       /* :822-823 */
    int n_ev = calc_events(m, y, v, ev_flow, ev_rate, ev_int);      /* :824 */
    for (int k = 0; k < n_ev; k++) {                                /* :826 */
      int fl = ev_flow[k];
      double mean = ev_rate[k] * dt;                                /* :829 */
      if (!(mean >= 0.0)) { if (!e) e = 2; continue; }              /* numpy: ValueError */
      int64_t delta = pdo_poisson(rng, mean);                       /* :830 */
      int from = m->from[fl], to = m->to[fl];
      if (from >= 0) {                                              /* :832-841 */
        if (delta > y[from]) delta = y[from];
        y[from] -= delta;
        if (to >= 0) y[to] += delta;
      } else {
        y[to] += delta;                                             /* :842-844 */
      }
    }
    for (int i = 0; i < C; i++) if (!(y[i] >= 0) && !e) e = 1;      /* :846 */
    if (!e && run_post_checks(m, y, yd, params, time, v, y2)) e = 1;  /* :846, user-defined */
    time += dt;                                                     /* :848 */
    if (soln) for (int i = 0; i < C; i++) soln[(size_t)it * C + i] = (double)y[i]; /* :850-852 */
    n_step++;
  }
  if (rng->exhausted && !e) e = 4;
  if (!soln) for (int i = 0; i < C; i++) yd[i] = (double)y[i];
  free(v); free(yd); free(y2); free(y); free(ev_flow); free(ev_int); free(ev_rate);
  if (err) *err = e;
  return n_step;
}

/* Continuous-time stochastic (Gillespie direct method): basepop.py:737-795.                  */
int64_t pdo_gillespie(const pdo_model *m, const double *y0, const double *params,
                      int32_t n_time, const double *target, pdo_rng *rng,
                      double *soln, int32_t *err) {
  const int C = m->n_comp, S = m->n_flow;
  double *v = (double *)malloc(sizeof(double) * (size_t)n_values(m));
  double *yd = (double *)malloc(sizeof(double) * (size_t)C);
  double *y2 = (double *)malloc(sizeof(double) * (size_t)C);
  int64_t *y = (int64_t *)malloc(sizeof(int64_t) * (size_t)C);
  int32_t *ev_flow = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
  int32_t *ev_int = (int32_t *)malloc(sizeof(int32_t) * (size_t)(S + 1));
  double *ev_rate = (double *)malloc(sizeof(double) * (size_t)(S + 1));
  int32_t e = 0;
  int64_t n_event = 0;
  for (int i = 0; i < C; i++) { y[i] = (int64_t)y0[i]; soln[i] = y0[i]; } /* :747-748, :755 */
  double time = target[0];
  for (int32_t it = 1; it < n_time && !e; it++) {
    double new_time = target[it];
    while (time < new_time) {                                       /* :762 */
      double dt;
      for (int i = 0; i < C; i++) yd[i] = (double)y[i];
      pdo_eval_vars(m, yd, params, time, 0, v);                     /* :763-764 */
      int n_ev = calc_events(m, y, v, ev_flow, ev_rate, ev_int);    /* :765 */
      if (n_ev == 0) {
        dt = new_time - time;                                       /* :767-770 */
      } else {
        int k = pdo_pick_event(ev_rate, n_ev, pdo_rng_double(rng)); /* :773 */
        double total = py_sum(ev_rate, ev_int, n_ev);               /* :775 */
        double u2 = pdo_rng_double(rng);
        if (rng->exhausted) { e = 4; break; }
        if (total == 0.0) { e = 3; break; }                         /* ZeroDivisionError */
        if (u2 == 0.0) { e = 5; break; }                            /* math domain error */
        dt = -log(u2) / total;                                      /* :776 */
        int fl = ev_flow[k];
        int from = m->from[fl], to = m->to[fl];
        if (from >= 0) y[from] -= 1;                                /* :778-787 */
        if (to >= 0) y[to] += 1;
        for (int i = 0; i < C; i++) if (!(y[i] >= 0) && !e) e = 1;  /* :789 */
        if (!e && run_post_checks(m, y, yd, params, time, v, y2)) e = 1;
        n_event++;
        if (e) break;
      }
      time += dt;                                                   /* :790 */
    }
    for (int i = 0; i < C; i++) soln[(size_t)it * C + i] = (double)y[i]; /* :793-795 */
  }
  free(v); free(yd); free(y2); free(y); free(ev_flow); free(ev_int); free(ev_rate);
  if (err) *err = e;
  return n_event;
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers: the reference's ensemble is `for i in range(n_replica): Model().integrate()`
 * (examples/stochastic_strains_model.py:210-218); here the loop runs over OpenMP threads with
 * Philox keyed by the global member id, exactly like the CUDA kernels.
 * ---------------------------------------------------------------------------------------- */
int32_t pdo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

typedef int64_t (*stoch_fn)(const pdo_model *, const double *, const double *, int32_t,
                            const double *, pdo_rng *, double *, int32_t *);

static int64_t stoch_batch(stoch_fn fn, const pdo_model *m, int64_t n_member, int64_t member0,
                           const double *y0, int64_t y0_stride, const double *params,
                           int64_t p_stride, int32_t n_time, const double *target,
                           uint64_t seed, double *soln, int64_t *sum, int64_t *sumsq,
                           int32_t n_thread, int32_t *err) {
  const size_t TC = (size_t)n_time * (size_t)m->n_comp;
  int64_t total = 0;
  int32_t first_err = 0;
  if (n_thread <= 0) n_thread = pdo_max_threads();
  if (sum) memset(sum, 0, sizeof(int64_t) * TC);
  if (sumsq) memset(sumsq, 0, sizeof(int64_t) * TC);
#pragma omp parallel num_threads(n_thread) reduction(+ : total)
  {
    double *buf = (double *)malloc(sizeof(double) * TC);
    int64_t *ls = sum ? (int64_t *)calloc(TC, sizeof(int64_t)) : 0;
    int64_t *lq = sumsq ? (int64_t *)calloc(TC, sizeof(int64_t)) : 0;
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_member; i++) {
      pdo_rng rng;
      int32_t e = 0;
      pdo_rng_philox(&rng, seed, (uint64_t)(member0 + i));
      double *out = soln ? soln + (size_t)i * TC : buf;
      total += fn(m, y0 + i * y0_stride, params + i * p_stride, n_time, target, &rng, out, &e);
      if (e) {
#pragma omp critical
        if (!first_err) first_err = e;
      }
      if (ls) for (size_t k = 0; k < TC; k++) { int64_t x = (int64_t)out[k]; ls[k] += x; if (lq) lq[k] += x * x; }
    }
#pragma omp critical
    {
      if (ls) for (size_t k = 0; k < TC; k++) sum[k] += ls[k];
      if (lq) for (size_t k = 0; k < TC; k++) sumsq[k] += lq[k];
    }
    free(buf); free(ls); free(lq);
  }
  if (err) *err = first_err;
  return total;
}

int64_t pdo_tau_leap_batch(const pdo_model *m, int64_t n_member, int64_t member0,
                           const double *y0, int64_t y0_stride, const double *params,
                           int64_t p_stride, int32_t n_time, const double *target,
                           uint64_t seed, double *soln, int64_t *sum, int64_t *sumsq,
                           int32_t n_thread, int32_t *err) {
  return stoch_batch(pdo_tau_leap, m, n_member, member0, y0, y0_stride, params, p_stride,
                     n_time, target, seed, soln, sum, sumsq, n_thread, err);
}

int64_t pdo_gillespie_batch(const pdo_model *m, int64_t n_member, int64_t member0,
                            const double *y0, int64_t y0_stride, const double *params,
                            int64_t p_stride, int32_t n_time, const double *target,
                            uint64_t seed, double *soln, int64_t *sum, int64_t *sumsq,
                            int32_t n_thread, int32_t *err) {
  return stoch_batch(pdo_gillespie, m, n_member, member0, y0, y0_stride, params, p_stride,
                     n_time, target, seed, soln, sum, sumsq, n_thread, err);
}

int64_t pdo_explicit_batch(const pdo_model *m, int64_t n_member, const double *y0,
                           int64_t y0_stride, const double *params, int64_t p_stride,
                           int32_t n_time, const double *target, double min_dt,
                           const double *su_table, double *soln, double *final_state,
                           int32_t n_thread, int32_t *err) {
  const size_t TC = (size_t)n_time * (size_t)m->n_comp;
  int64_t total = 0;
  int32_t first_err = 0;
  if (n_thread <= 0) n_thread = pdo_max_threads();
#pragma omp parallel num_threads(n_thread) reduction(+ : total)
  {
    double *buf = (double *)malloc(sizeof(double) * TC);
#pragma omp for schedule(dynamic, 4)
    for (int64_t i = 0; i < n_member; i++) {
      int32_t e = 0;
      double *out = soln ? soln + (size_t)i * TC : buf;
      total += pdo_explicit(m, y0 + i * y0_stride, params + i * p_stride, n_time, target,
                            min_dt, su_table, out, &e);
      if (final_state)
        for (int c = 0; c < m->n_comp; c++)
          final_state[(size_t)i * m->n_comp + c] = out[TC - m->n_comp + c];
      if (e) {
#pragma omp critical
        if (!first_err) first_err = e;
      }
    }
    free(buf);
  }
  if (err) *err = first_err;
  return total;
}
