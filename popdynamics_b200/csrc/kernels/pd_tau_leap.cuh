/*
 * pd_tau_leap.cuh -- discrete-time stochastic integrator (clamped Poisson tau-leap), one thread
 * per ensemble member, integer compartment counts in registers.
 *
 * Restates /root/reference/basepop.py:797-852 for a batch of independent members:
 *   per target interval: dt = target[i] - target[i-1]            (:820, ONE step per interval)
 *   rates from the start-of-step counts (calculate_vars + calculate_events, :822-824)
 *   per event, in flow-table order: delta = Poisson(rate * dt), clamped to the CURRENT source
 *   count for transfers and deaths, unclamped for births          (:826-844)
 *   record the counts                                             (:850-852)
 * Non-entry events exist only when rate > 0 (:712, :718, :725, :733); entry events always.
 * Poisson(mean) is numpy's legacy transform (pd_common.cuh): a zero mean draws nothing, a mean
 * below 10 multiplies uniforms until the product drops to exp(-mean), PTRS above.
 *
 * The rates of a step are frozen at its start and a Poisson draw does not look at the counts, so
 * only the clamp-and-move has to run in event order.  One step is therefore three passes, none
 * of which changes which uniform goes to which event:
 *   1. RATE PASS, all lanes in lockstep: every live event's code -- exp(-mean) for the
 *      multiplication method, -mean for PTRS -- is appended to the lane's list in shared memory
 *      [k][thread]; a bit mask remembers which flow-table rows are in the list.  Codes that
 *      depend on one or two integer counts only (count * parameter; count * count' * parameter,
 *      a force of infection) come from tables built once per launch; one range test per step
 *      decides whether every such row is a plain lookup.  A cell holding 0.0 closes the list.
 *   2. DRAW LOOP, flattened: the lane walks its OWN list while the warp walks the uniform stream
 *      together.  One trip = one Philox block = two uniforms for every lane; each uniform either
 *      continues the lane's current product or ends its event, whose draw X then overwrites the
 *      code in the list; the lane is done when the code it loads next is the closing 0.0.  The
 *      trip count of a warp is the largest TOTAL number of uniforms any lane needs in the step,
 *      instead of the sum over events of the per-event maxima that the nested event-major loop
 *      pays.
 *   3. APPLY PASS, lockstep again: the flow-table rows in order, compartment indices literal,
 *      clamp to the current source count and move.
 */
/* the traced rate expressions use the guarded division and the constant-bank exp of pd_math.cuh
 * (same values; see pd_gillespie.cuh): counts of zero are as common here */
#include "pd_math.cuh"
#ifndef PD_TAU_PLAIN_MATH
#define PD_DIV(a, b) pd_div_zero_guard(a, b)
#define PD_EXP(x) pd_exp_any(x)
#endif
#include "pd_model.h"
#include "pd_common.cuh"

struct PdTauLaunch {
  PdStochArgs a;
  PdUniform U;
};

#if PD_LIVE_NROW <= 32
typedef pd_u32 pd_mask;
#elif PD_LIVE_NROW <= 64
typedef pd_u64 pd_mask;
#else
#error "the tau-leap kernel handles at most 64 flow-table rows"
#endif

/* the kernel's dynamic shared memory (the draw list), read by the host after loading the module
 * (csrc/popdyn_capi.cu) */
/* PD_TAU_SENTINEL: the list ends with a cell holding 0.0 -- no code is ever 0.0, a row that
 * draws nothing is not in the list -- so the draw loop learns that a lane is done from the code it
 * has just loaded: a finished event is a store, an add, a load and three moves, with no compare
 * of the cursor against the end of the list and nothing predicated on it. */
#ifndef PD_TAU_SENTINEL
#define PD_TAU_SENTINEL 3
#endif
extern "C" __device__ unsigned int pd_smem_bytes =
    (unsigned int)PD_CODE_SMEM_BYTES(PD_LIVE_NROW + (PD_TAU_SENTINEL ? 1 : 0), PD_BLOCK);

/* Poisson code of a mean: exp(-mean) for the multiplication method, -mean for PTRS */
__device__ __forceinline__ double pd_tau_code(double mean) {
  return (mean >= 10.0) ? -mean : pd_exp_nonpos(-mean);
}

/* Code tables (see codegen.py): a row whose rate is count * U.p[k] has a code that is a function of
 * the integer count alone while dt is constant.  This kernel evaluates exactly the expression of
 * the rate pass -- rate = (double)n * p; mean = rate * dt; pd_tau_code(mean) -- for n = 0 ..
 * n_entry - 1, so a looked-up code is bit-identical to a computed one.  0.0 marks "the row does
 * not exist / draws nothing" (:712-733, a zero mean), NaN a negative or NaN mean (error): the host
 * builds a table for a positive dt only, so NaN can only appear among the single codes of the
 * entry rows. */
extern "C" __device__ unsigned int pd_tab_count = PD_NTAB;       /* tables of n_entry codes  */
extern "C" __device__ unsigned int pd_uni_count = PD_NUNI;       /* single codes behind them */
__device__ const int pd_tab_param[PD_NTAB_DIM] = PD_TAB_PARAMS;

extern "C" __global__ void pd_tau_table_kernel(const __grid_constant__ PdTauLaunch L, double *table,
                                               int n_entry, double dt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {
    /* entry rows with an ensemble-wide rate: one code each (basepop.py:706-707, always present;
     * a zero mean draws nothing) */
    const PdUniform &U = L.U;
    const double uni_rate[PD_NUNI_DIM] = PD_UNI_RATES;
    for (int j = 0; j < PD_NUNI; ++j) {
      const double mean = uni_rate[j] * dt;
      double code = 0.0;
      if (mean > 0.0) code = pd_tau_code(mean);
      else if (!(mean == 0.0)) code = __longlong_as_double(0x7ff8000000000000LL);
      table[(size_t)PD_NTAB * n_entry + j] = code;
    }
  }
  if (n >= n_entry) return;
  for (int t = 0; t < PD_NTAB; ++t) {
    const double rate = (double)n * L.U.p[pd_tab_param[t]];
    double code = 0.0;
    if (rate > 0) {
      const double mean = rate * dt;          /* dt > 0: positive, zero (underflow) or +inf */
      if (mean > 0.0) code = pd_tau_code(mean);
    }
    table[(size_t)t * n_entry + n] = code;
  }
}

/* Pair tables: a row whose rate is count_a * (count_b * U.p[k]) -- the model's `rate_force = beta *
 * infectious` var times the susceptibles -- has a code that depends on the two integer counts
 * alone while dt is constant.  Exactly the expressions of pd_event_mults (count_b * parameter,
 * either order: the product is the same double) and of the rate pass (count_a * that, the mean,
 * pd_tau_code), so a looked-up code is bit-identical to a computed one; 0.0 = the row does not
 * fire.  Layout [table][n_b][n_a], count_a fastest. */
extern "C" __device__ unsigned int pd_tab2_count = PD_NTAB2;
__device__ const int pd_tab2_param[PD_NTAB2_DIM] = PD_TAB2_PARAMS;

extern "C" __global__ void pd_tau_table2_kernel(const __grid_constant__ PdTauLaunch L, double *table,
                                                int n_a, int n_b, double dt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_a * n_b) return;
  const int na = (int)(i % n_a), nb = (int)(i / n_a);
  for (int t = 0; t < PD_NTAB2; ++t) {
    const double mult = (double)nb * L.U.p[pd_tab2_param[t]];
    const double rate = (double)na * mult;
    double code = 0.0;
    if (rate > 0) {
      const double mean = rate * dt;          /* dt > 0: positive, zero (underflow) or +inf */
      if (mean > 0.0) code = pd_tau_code(mean);
    }
    table[(size_t)t * n_a * n_b + (size_t)i] = code;
  }
}

/* a finished draw is parked in the list cell of its code; with 32-bit counts it saturates at
 * 2^31 - 1, which changes nothing: a clamp brings it down to a count anyway and an unclamped
 * gain of that size raises PD_ERR_COUNT_RANGE either way */
__device__ __forceinline__ void pd_park_draw(double *cell, pd_i64 k) {
  if (sizeof(pd_count) == 4) *(int *)cell = (k > 0x7fffffffLL) ? 0x7fffffff : (int)k;
  else *(pd_i64 *)cell = k;
}
__device__ __forceinline__ void pd_park_draw(double *cell, int x) {
  if (sizeof(pd_count) == 4) *(int *)cell = x;
  else *(pd_i64 *)cell = (pd_i64)x;
}

/* PD_TAU_ABS_CURSOR: the draw list is walked with 32-bit shared-memory ADDRESSES instead of
 * element offsets from the lane's column: an access is then [register] with no multiply-add in
 * front of it (one IMAD per appended code, per finished event and per row of the apply pass). */
#ifndef PD_TAU_ABS_CURSOR
#define PD_TAU_ABS_CURSOR 1
#endif

#ifndef PD_TAU_ONE_TABLE_TEST
#define PD_TAU_ONE_TABLE_TEST 1   /* hoist the rows' "count inside the code table" tests */
#endif
#if PD_TAU_ABS_CURSOR
typedef unsigned pd_cursor;
#define PD_CURSOR_STEP ((unsigned)(PD_BLOCK * sizeof(double)))
__device__ __forceinline__ double pd_list_code(pd_cursor at) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(at) : "memory");
  return v;
}
__device__ __forceinline__ void pd_list_put_code(pd_cursor at, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(at), "d"(v) : "memory");
}
__device__ __forceinline__ void pd_list_park(pd_cursor at, pd_i64 k) {
  if (sizeof(pd_count) == 4) {
    const int d = (k > 0x7fffffffLL) ? 0x7fffffff : (int)k;
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(at), "r"(d) : "memory");
  } else asm volatile("st.shared.s64 [%0], %1;" :: "r"(at), "l"(k) : "memory");
}
__device__ __forceinline__ void pd_list_park(pd_cursor at, int x) {
  if (sizeof(pd_count) == 4) asm volatile("st.shared.s32 [%0], %1;" :: "r"(at), "r"(x) : "memory");
  else asm volatile("st.shared.s64 [%0], %1;" :: "r"(at), "l"((pd_i64)x) : "memory");
}
__device__ __forceinline__ pd_count pd_list_draw(pd_cursor at) {
  if (sizeof(pd_count) == 4) {
    int d;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(d) : "r"(at) : "memory");
    return (pd_count)d;
  }
  pd_i64 d;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(d) : "r"(at) : "memory");
  return (pd_count)d;
}
#else
typedef int pd_cursor;
#define PD_CURSOR_STEP PD_BLOCK
#endif

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_tau_leap_kernel(const __grid_constant__ PdTauLaunch L) {
  const PdStochArgs &a = L.a;
  const PdUniform &U = L.U;
  const int tid = threadIdx.x;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + tid;
#if PD_USE_MASK
  const bool has_member = m < a.n_member && a.mask[m < a.n_member ? m : 0] == (unsigned char)a.mask_value;
#else
  const bool has_member = m < a.n_member;
#endif
#if PD_OUT_MODE == PD_OUT_STATS && !PD_HIST_MOMENTS
  /* the in-kernel moments are warp- and block-collective: lanes without a member stay */
  const bool active = has_member;
#else
  /* nothing below is collective: lanes without a member leave, and the step loop carries no
   * per-lane activity test */
  if (!has_member) return;
  const bool active = true;
#endif
  const pd_i64 M = a.n_member;

  extern __shared__ double s_dyn[];
  double *const s_list = s_dyn + tid;                       /* [k][PD_BLOCK], this lane's column */
#if PD_TAU_ABS_CURSOR
  const pd_cursor list0 = (pd_cursor)__cvta_generic_to_shared(s_list);
#define PD_LIST_CODE(at) pd_list_code(at)
#define PD_LIST_PUT_CODE(at, v) pd_list_put_code(at, v)
#define PD_LIST_PARK(at, d) pd_list_park(at, d)
#define PD_LIST_DRAW(at) pd_list_draw(at)
#else
  const pd_cursor list0 = 0;
#define PD_LIST_CODE(at) s_list[at]
#define PD_LIST_PUT_CODE(at, v) s_list[at] = (v)
#define PD_LIST_PARK(at, d) pd_park_draw(&s_list[at], d)
#define PD_LIST_DRAW(at)                                                             \
  ((sizeof(pd_count) == 4) ? (pd_count)*(const int *)&s_list[at] : (pd_count)*(const pd_i64 *)&s_list[at])
#endif
#if PD_OUT_MODE == PD_OUT_STATS
  __shared__ PdStatsPart s_acc;
#endif

  pd_count y[PD_C];
  double pm[PD_NPM_DIM];
  PdRng rng;
  pm[0] = 0.0;
  if (active) {
#pragma unroll
    for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * a.pm_ld + m];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) {
      const double v = pd_load_y0(a.y0, a.y0_stride_c, a.y0_stride_m, c, m);
      y[c] = (pd_count)v;                                   /* int() truncation, :805-806 */
#if PD_OUT_MODE == PD_OUT_FULL
      a.soln[(size_t)c * a.out_ld + m] = v;                        /* row 0 = untruncated y, :813 */
#endif
    }
    rng.init(a, m);
  } else {
#pragma unroll
    for (int c = 0; c < PD_C; ++c) y[c] = 0;
  }
#if PD_OUT_MODE == PD_OUT_STATS
  pd_stats_step(a, s_acc, y, active, 0, 0, a.member0 + m);
  if (PD_EPOCH == 1 || a.n_time == 1) pd_stats_part_flush(a, s_acc, 0, 1);
#endif

  int bad = 0, bad_at = 0;
  double time = a.target[0];
  for (int it = 1; it < a.n_time; ++it) {
    if (active) {
      const double dt = a.target[it] - a.target[it - 1];    /* :820 */
      const bool use_table = (PD_NTAB + PD_NUNI) > 0 && a.code_table != 0 &&
                             dt == a.code_table_dt;
      const bool use_table2 = PD_NTAB2 > 0 && a.code_table2 != 0 && dt == a.code_table_dt;

      /* ---- 1. rate pass (:822-824): start-of-step counts as doubles; every rate of this step
       * is formed from them, the live counts only enter through the clamps (:833-841) */
      double yd[PD_C], mult[PD_NMULT_DIM];
#pragma unroll
      for (int c = 0; c < PD_C; ++c) yd[c] = (double)y[c];
      pd_event_mults(yd, U, pm, time, (const double *)0, mult);
      pd_mask live = 0;
      pd_cursor end = list0;                                /* append cursor */
#define PD_TAU_APPEND(e, code)                                                       \
  { PD_LIST_PUT_CODE(end, code); end += PD_CURSOR_STEP; live |= (pd_mask)1 << e; }
#define PD_TAU_CODE(e, mean)                                                         \
  if (mean > 0.0) PD_TAU_APPEND(e, pd_tau_code(mean))                                \
  else if (!(mean == 0.0)) { if (!bad) bad = PD_ERR_POISSON_MEAN; }
#define PD_TAU_COMPUTED(e, rate)                                                     \
  { const double r_ = rate; if (r_ > 0) { const double mean = r_ * dt; PD_TAU_CODE(e, mean) } }
/* a row of the form count * uniform parameter: looked up by count when the launch has a table */
/* 32-bit index arithmetic: the host builds pair tables of fewer than 2^31 cells in all */
#define PD_TAU_PAIR_INDEX(e, from)                                                   \
  (((unsigned)(PD_ROW_TAB2_##e < 0 ? 0 : PD_ROW_TAB2_##e) * (unsigned)a.code_table2_nb + \
    (unsigned)y[PD_ROW_TAB2B_##e < 0 ? 0 : PD_ROW_TAB2B_##e]) * (unsigned)a.code_table2_na + \
   (unsigned)y[from])
#define PD_TAU_FIXED(e, from, rate)                                                  \
  if (PD_ROW_TAB_##e >= 0 && use_table && y[from] < (pd_count)a.code_table_n) {      \
    const double code = a.code_table[(PD_ROW_TAB_##e < 0 ? 0 : PD_ROW_TAB_##e) *     \
                                     a.code_table_n + (int)y[from]];                 \
    if (code != 0.0) PD_TAU_APPEND(e, code)       /* never NaN: the table's dt is > 0 */ \
  } else if (PD_ROW_TAB2_##e >= 0 && use_table2 &&                                   \
             y[from] < (pd_count)a.code_table2_na &&                                 \
             y[PD_ROW_TAB2B_##e < 0 ? 0 : PD_ROW_TAB2B_##e] < (pd_count)a.code_table2_nb) { \
    const double code = a.code_table2[PD_TAU_PAIR_INDEX(e, from)];                   \
    if (code != 0.0) PD_TAU_APPEND(e, code)                                          \
  } else PD_TAU_COMPUTED(e, rate)
#define PD_TAU_ENTRY(e, to, is_int, rate)                                            \
  if (PD_ROW_UNI_##e >= 0 && use_table) {                                            \
    const double code = a.code_table[PD_NTAB * a.code_table_n +                      \
                                     (PD_ROW_UNI_##e < 0 ? 0 : PD_ROW_UNI_##e)];     \
    if (code != 0.0) {                                                               \
      if (code != code) { if (!bad) bad = PD_ERR_POISSON_MEAN; }                     \
      else PD_TAU_APPEND(e, code)                                                    \
    }                                                                                \
  } else { const double mean = rate * dt; PD_TAU_CODE(e, mean) }
#define PD_TAU_TRANSFER(e, from, to, rate) PD_TAU_FIXED(e, from, rate)
#define PD_TAU_DEATH(e, from, rate) PD_TAU_FIXED(e, from, rate)
#if PD_TAU_ONE_TABLE_TEST && (PD_NTAB > 0 || PD_NTAB2 > 0)
      /* One test per step instead of one per row: when the launch has a table for this dt and
       * the largest count a table row looks up is inside it, every such row is a plain lookup
       * (PD_TAU_LOOKUP*); any other step takes the general rows, which test and fall back to
       * computing the code row by row. */
      bool all_in_table = false;
      if ((PD_NTAB == 0 || use_table) && (PD_NTAB2 == 0 || use_table2)) {
        pd_count top = 0, top_a = 0, top_b = 0;
#define PD_TAU_TOP_ENTRY(e, to, is_int, rate)
#define PD_TAU_TOP_FIXED(e, from)                                                    \
  if (PD_ROW_TAB_##e >= 0) top = top > y[from] ? top : y[from];                      \
  else if (PD_ROW_TAB2_##e >= 0) {                                                   \
    top_a = top_a > y[from] ? top_a : y[from];                                       \
    top_b = top_b > y[PD_ROW_TAB2B_##e < 0 ? 0 : PD_ROW_TAB2B_##e]                   \
                ? top_b : y[PD_ROW_TAB2B_##e < 0 ? 0 : PD_ROW_TAB2B_##e];            \
  }
#define PD_TAU_TOP_TRANSFER(e, from, to, rate) PD_TAU_TOP_FIXED(e, from)
#define PD_TAU_TOP_DEATH(e, from, rate) PD_TAU_TOP_FIXED(e, from)
        PD_FOR_EACH_LIVE_ROW(PD_TAU_TOP_ENTRY, PD_TAU_TOP_TRANSFER, PD_TAU_TOP_DEATH)
#undef PD_TAU_TOP_ENTRY
#undef PD_TAU_TOP_FIXED
#undef PD_TAU_TOP_TRANSFER
#undef PD_TAU_TOP_DEATH
        /* counts are never negative */
        all_in_table = (PD_NTAB == 0 || top < (pd_count)a.code_table_n) &&
                       (PD_NTAB2 == 0 || (top_a < (pd_count)a.code_table2_na &&
                                          top_b < (pd_count)a.code_table2_nb));
      }
      if (all_in_table) {
#define PD_TAU_LOOKUP_FIXED(e, from, rate)                                           \
  if (PD_ROW_TAB_##e >= 0) {                                                         \
    const double code = a.code_table[(PD_ROW_TAB_##e < 0 ? 0 : PD_ROW_TAB_##e) *     \
                                     a.code_table_n + (int)y[from]];                 \
    if (code != 0.0) PD_TAU_APPEND(e, code)                                          \
  } else if (PD_ROW_TAB2_##e >= 0) {                                                 \
    const double code = a.code_table2[PD_TAU_PAIR_INDEX(e, from)];                   \
    if (code != 0.0) PD_TAU_APPEND(e, code)                                          \
  } else PD_TAU_COMPUTED(e, rate)
#define PD_TAU_LOOKUP_ENTRY(e, to, is_int, rate)                                     \
  if (PD_ROW_UNI_##e >= 0 && use_table) {                                                         \
    const double code = a.code_table[PD_NTAB * a.code_table_n +                      \
                                     (PD_ROW_UNI_##e < 0 ? 0 : PD_ROW_UNI_##e)];     \
    if (code != 0.0) {                                                               \
      if (code != code) { if (!bad) bad = PD_ERR_POISSON_MEAN; }                     \
      else PD_TAU_APPEND(e, code)                                                    \
    }                                                                                \
  } else { const double mean = rate * dt; PD_TAU_CODE(e, mean) }
#define PD_TAU_LOOKUP_TRANSFER(e, from, to, rate) PD_TAU_LOOKUP_FIXED(e, from, rate)
#define PD_TAU_LOOKUP_DEATH(e, from, rate) PD_TAU_LOOKUP_FIXED(e, from, rate)
        PD_FOR_EACH_LIVE_ROW(PD_TAU_LOOKUP_ENTRY, PD_TAU_LOOKUP_TRANSFER, PD_TAU_LOOKUP_DEATH)
#undef PD_TAU_LOOKUP_ENTRY
#undef PD_TAU_LOOKUP_FIXED
#undef PD_TAU_LOOKUP_TRANSFER
#undef PD_TAU_LOOKUP_DEATH
      } else {
        PD_FOR_EACH_LIVE_ROW(PD_TAU_ENTRY, PD_TAU_TRANSFER, PD_TAU_DEATH)
      }
#else
      PD_FOR_EACH_LIVE_ROW(PD_TAU_ENTRY, PD_TAU_TRANSFER, PD_TAU_DEATH)
#endif
#undef PD_TAU_ENTRY
#undef PD_TAU_TRANSFER
#undef PD_TAU_DEATH
#undef PD_TAU_FIXED
#undef PD_TAU_PAIR_INDEX
#undef PD_TAU_COMPUTED
#undef PD_TAU_CODE
#undef PD_TAU_APPEND

      /* ---- 2. draw loop.  Lane state: list cursor, the current event's code, its running
       * product and count */
#if PD_TAU_SENTINEL
      PD_LIST_PUT_CODE(end, 0.0);
#endif
      pd_cursor cur = list0;
      double code = PD_LIST_CODE(list0);
      double prod = 1.0;
      int X = 0;
/* the current event ends with draw d: park it, step to the lane's next event */
#if PD_TAU_SENTINEL
#define PD_TAU_MORE (__double2hiint(code) != 0)     /* codes are normal numbers or below -10 */
#define PD_TAU_MORE_MULT (code > 0.0)               /* ... and not a PTRS event either */
/* the running product restarts at 1.0 as ONE fp64 instruction, code * 0 + 1 (codes are finite
 * wherever the product is used: a multiplication event's code is exp(-mean) in (0, 1], the
 * sentinel is 0): the FP64 pipe is idle here, the two moves of a 64-bit constant are not free */
#if PD_TAU_SENTINEL & 2
#define PD_TAU_ONE(code_) __fma_rn((code_), 0.0, 1.0)
#else
#define PD_TAU_ONE(code_) 1.0
#endif
#define PD_TAU_NEXT(d)                                                               \
  {                                                                                  \
    PD_LIST_PARK(cur, d);                                                            \
    cur += PD_CURSOR_STEP;                                                           \
    code = PD_LIST_CODE(cur); prod = PD_TAU_ONE(code); X = 0;                        \
  }
#else
#define PD_TAU_MORE (cur != end)
#define PD_TAU_MORE_MULT (cur != end && code >= 0.0)
#define PD_TAU_NEXT(d)                                                               \
  {                                                                                  \
    PD_LIST_PARK(cur, d);                                                            \
    cur += PD_CURSOR_STEP;                                                           \
    if (cur != end) { code = PD_LIST_CODE(cur); prod = 1.0; X = 0; }                 \
  }
#endif
/* one uniform of the multiplication method */
#define PD_TAU_MULT(u)                                                               \
  {                                                                                  \
    prod *= (u);                                                                     \
    if (prod > code) X += 1; else PD_TAU_NEXT(X)                                     \
  }
      while (PD_TAU_MORE) {
        if (code < 0.0) {
          /* PTRS event (mean >= 10): one attempt = the next two uniforms of the same stream,
           * judged out of line; the lane rejoins the block-wise walk at the next trip */
          double u, v;
#if PD_RNG_MODE == PD_RNG_PHILOX
          {
            double x0, x1;
            rng.block(a, x0, x1);
            if (rng.has_spare) { u = rng.spare; v = x0; rng.spare = x1; }
            else { u = x0; v = x1; }
          }
#else
          u = rng.next(a);
          v = rng.next(a);
          if (rng.exhausted) break;
#endif
          const pd_i64 k = pd_poisson_ptrs_attempt(-code, u, v);
#if PD_PTRS_CHECK
          if (k == -2 && !bad) bad = PD_ERR_SQUEEZE;
#endif
          if (k >= 0) PD_TAU_NEXT(k)
          continue;
        }
#if PD_RNG_MODE == PD_RNG_PHILOX
        if (!rng.has_spare) {               /* every lane without a pending spare: one block */
          double u0;
          rng.block(a, u0, rng.spare);
          rng.has_spare = 1;
          PD_TAU_MULT(u0)
        }
        if (PD_TAU_MORE_MULT) {             /* the block's second uniform, or the pending spare */
          rng.has_spare = 0;
          PD_TAU_MULT(rng.spare)
        }
#else
        const double u = rng.next(a);
        if (rng.exhausted) break;
        PD_TAU_MULT(u)
#endif
      }
#undef PD_TAU_MULT
#undef PD_TAU_NEXT
#undef PD_TAU_MORE
#undef PD_TAU_MORE_MULT

      /* ---- 3. apply pass, flow-table order: clamp to the CURRENT source count (:833-834,
       * :839-840), move; births unclamped (:842-844).  Branch-free: a row that is not in the
       * list takes d = 0.  With 32-bit counts a gain that passes 2^31 - 1 sets the sign bit at
       * that moment (both terms are non-negative, the sum cannot wrap past it); the OR of every
       * updated count is checked once per step and raises the error flag (numpy / Python ints
       * never wrap). */
      if (!rng.exhausted) {
        pd_cursor get = list0;
        pd_count signs = 0;
#define PD_TAU_TAKE(e, d)                                                            \
  const bool on_##e = (live >> e) & 1;                                               \
  pd_count d = 0;                                                                    \
  if (on_##e) d = PD_LIST_DRAW(get);                                                 \
  get += on_##e ? PD_CURSOR_STEP : 0;
#define PD_TAU_ENTRY(e, to, is_int, rate)                                            \
  { PD_TAU_TAKE(e, d) y[to] += d; signs |= y[to]; }
#define PD_TAU_TRANSFER(e, from, to, rate)                                           \
  { PD_TAU_TAKE(e, d) d = d < y[from] ? d : y[from]; y[from] -= d; y[to] += d; signs |= y[to]; }
#define PD_TAU_DEATH(e, from, rate)                                                  \
  { PD_TAU_TAKE(e, d) d = d < y[from] ? d : y[from]; y[from] -= d; }
        PD_FOR_EACH_LIVE_ROW(PD_TAU_ENTRY, PD_TAU_TRANSFER, PD_TAU_DEATH)
#undef PD_TAU_ENTRY
#undef PD_TAU_TRANSFER
#undef PD_TAU_DEATH
#undef PD_TAU_TAKE
        if (sizeof(pd_count) == 4 && signs < 0 && !bad) bad = PD_ERR_COUNT_RANGE;
#if PD_HAS_STEP_CHECKS
        {
          /* a user-defined checks() (:846): the updated counts, the vars of the step's start */
          double y2[PD_C];
#pragma unroll
          for (int c = 0; c < PD_C; ++c) y2[c] = (double)y[c];
          if (!pd_step_checks(yd, y2, U, pm, time, (const double *)0) && !bad) {
            bad = PD_ERR_CHECKS;
            bad_at = it;
          }
        }
#endif
      }
      time += dt;                                           /* :848 */
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        a.soln[((size_t)it * PD_C + c) * a.out_ld + m] = (double)y[c];   /* :850-852 */
#endif
    }
#if PD_OUT_MODE == PD_OUT_STATS
    {
      /* the warp records this target time; at the end of an epoch the block moves its
       * accumulators to global memory (two barriers inside pd_stats_part_flush) */
      const int slot = it % PD_EPOCH;
      pd_stats_step(a, s_acc, y, active, it, slot, a.member0 + m);
      if (slot == PD_EPOCH - 1 || it == a.n_time - 1)       /* block-uniform */
        pd_stats_part_flush(a, s_acc, it - slot, slot + 1);
    }
#endif
  }

  if (active) {
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * a.out_ld + m] = (double)y[c];
#endif
    if (!bad && rng.exhausted) bad = PD_ERR_STREAM;
    if (bad) pd_raise(a.err, bad, a.member0 + m, bad_at);
  }
  if (a.n_steps && active) {
    /* whichever lanes arrive together elect one of them; the groups' counts add up */
    const pd_u32 group = __activemask();
    if ((tid & 31) == __ffs(group) - 1)
      atomicAdd(a.n_steps, (pd_u64)__popc(group) * (pd_u64)(a.n_time - 1));
  }
}
