/*
 * pd_gillespie.cuh -- continuous-time stochastic integrator (Gillespie direct method), one
 * thread per ensemble member, integer compartment counts in registers.
 *
 * Restates /root/reference/basepop.py:737-795 and pick_event (:161-178):
 *   while time < target[i]:
 *     rates from the current counts (calculate_vars + calculate_events, :763-765)
 *     no live event  -> dt = target[i] - time                               (:767-770)
 *     else uniform #1 picks the event: sequential cumulative sum over the LIVE events in
 *          flow-table order, i = u1 * total, first k with not (i > cumul[k]), else the last
 *          live event (:168-177); uniform #2 gives dt = -log(u2) / sum(rates)  (:775-776);
 *          the event is applied BEFORE time advances                        (:778-790)
 *   the counts recorded at target[i] therefore contain the overshooting event; time is never
 *   snapped back.
 * `sum(rates)` is Python's builtin: Neumaier-compensated on CPython >= 3.12 (PD_NAIVE_SUM 0) or
 * the plain left fold (PD_NAIVE_SUM 1), in which case it equals the last cumulative sum.
 *
 * Divergence: every lane walks its own target index (flattened loop), so a warp only idles at the
 * very end of its longest member.  Recording (FULL / STATS output) must not happen inside that
 * divergent walk -- a handful of lanes cross a target time in any given trip, and the record code
 * run for them alone cost 75 % of the kernel.  Instead a lane PARKS its counts in its own column
 * of a shared-memory window of PD_GIL_SLOTS target times (five cheap stores) and keeps going; once
 * per trip one REDUX.MIN finds the oldest target time every lane of the WARP has passed, and the
 * warp records those slots together, all 32 lanes on: coalesced rows of the trajectory, one RED
 * per count into the histograms, one REDUX per moment.  A lane that gets PD_GIL_SLOTS target times
 * ahead of the slowest lane of its warp waits for it; there is no block-level barrier.
 */
/* Per-event arithmetic shaped for this kernel (pd_math.cuh), one bit each so that a build can be
 * compared with the plain forms (POPDYN_DEFINE=-DPD_GIL_OPT=0): 1 the traced divisions skip the
 * divider's slow path for a zero numerator, 2 dead rows are zeroed through the integer halves, 4 the
 * traced exp and 8 the waiting-time log keep their coefficients in the constant bank, 16 the picked
 * row is applied through compile-time row masks instead of a jump table, 32 the picked row is found
 * by counting instead of through a liveness mask, 64 the first pass takes every rate as it is
 * when one integer maximum shows that none has a sign bit, is infinite or NaN (PD_GIL_FAST below).
 * Values are the same in every combination. */
#ifndef PD_GIL_OPT
#define PD_GIL_OPT 127
#endif
#include "pd_math.cuh"
#if PD_GIL_OPT & 1
#define PD_DIV(a, b) pd_div_zero_guard(a, b)
#endif
#if PD_GIL_OPT & 4
#define PD_EXP(x) pd_exp_any(x)
#endif
#include "pd_model.h"
#include "pd_common.cuh"

struct PdGilLaunch {
  PdStochArgs a;
  PdUniform U;
};

/* the kernel's dynamic shared memory (the recording window), read by the host after loading the
 * module (csrc/popdyn_capi.cu) */
#define PD_GIL_WINDOW (PD_OUT_MODE != PD_OUT_FINAL)
/* window slots: the largest power of two in [2, 16] that keeps the window near 40 KB per block */
#define PD_GIL_SLOT_BYTES ((int)(PD_C * PD_BLOCK * sizeof(pd_count)))
#ifndef PD_GIL_SLOTS
#define PD_GIL_SLOTS                                                                   \
  (16 * PD_GIL_SLOT_BYTES <= 40960 ? 16 : 8 * PD_GIL_SLOT_BYTES <= 40960 ? 8           \
   : 4 * PD_GIL_SLOT_BYTES <= 40960 ? 4 : 2)
#endif
extern "C" __device__ unsigned int pd_smem_bytes =
    PD_GIL_WINDOW ? (unsigned int)(PD_GIL_SLOTS * PD_GIL_SLOT_BYTES) : 0u;

/* the two uniforms of one reaction: exactly one Philox block when no spare is pending */
__device__ __forceinline__ void pd_two_uniforms(const PdStochArgs &a, PdPhilox &rng, double &u1,
                                                double &u2) {
  if (!rng.has_spare) rng.block(a, u1, u2);
  else { u1 = rng.next(a); u2 = rng.next(a); }
}
__device__ __forceinline__ void pd_two_uniforms(const PdStochArgs &a, PdInject &rng, double &u1,
                                                double &u2) {
  u1 = rng.next(a);
  u2 = rng.next(a);
}

#define PD_GIL_WORDS ((PD_NFLOW_DIM + 31) / 32)   /* words of a flow-table row mask */

/* Compile-time facts about the flow table: does it hold an entry row (always live, :706-707), is
 * row 0 one, does an entry row carry a Python int (sum() then starts on its exact-int path). */
#define PD_GF_NONE3(e, from, to, rate)
#define PD_GF_NONE2(e, from, rate)
#define PD_GF_ANY_ENTRY(e, to, is_int, rate) | 1
#define PD_GF_ROW0_ENTRY(e, to, is_int, rate) | ((e) == 0 ? 1 : 0)
#define PD_GF_INT_ENTRY(e, to, is_int, rate) | ((is_int) ? 1 : 0)
#define PD_GIL_HAS_ENTRY (0 PD_FOR_EACH_LIVE_ROW(PD_GF_ANY_ENTRY, PD_GF_NONE3, PD_GF_NONE2))
#define PD_GIL_ROW0_ENTRY (0 PD_FOR_EACH_LIVE_ROW(PD_GF_ROW0_ENTRY, PD_GF_NONE3, PD_GF_NONE2))
#define PD_GIL_HAS_INT_ENTRY (0 PD_FOR_EACH_LIVE_ROW(PD_GF_INT_ENTRY, PD_GF_NONE3, PD_GF_NONE2))
/* The fast first pass.  A rate that is +0.0 or positive (+inf aside) joins the running sums
 * exactly as the reference's filtered event list would have it: a zero adds nothing, a positive
 * one is live.  A rate is a count (+0.0 or positive) times a multiplier, so when the unsigned
 * maximum of the multipliers' high words is below 0x7ff00000 -- no sign bit, no infinity, no NaN
 * anywhere -- the `r > 0` test of a row, the two selects that zero a dead row and the liveness
 * bookkeeping (5 of a row's 21 issue slots on the TB model) do nothing and are left out;
 * whatever pass 2 needs to know about liveness it reads off cumul[].  Any other event takes the
 * exact pass.  Needs the counting search (bit 32) and a sum() without an
 * exact-int prefix (with f == 0 in front of the first live rate its correction term is 0 by
 * itself, so `on_float_path` has nothing to suppress). */
#define PD_GIL_FAST (((PD_GIL_OPT) & 64) && ((PD_GIL_OPT) & 32) && !PD_GIL_HAS_INT_ENTRY && PD_LIVE_NROW > 0)

/* one reaction: returns the error code (0 = ok); advances time, mutates y */
__device__ __forceinline__ int pd_gillespie_event(const PdStochArgs &a, const PdUniform &U,
                                                  const double (&pm)[PD_NPM_DIM],
                                                  pd_count (&y)[PD_C], double &time,
                                                  const double new_time, PdRng &rng,
                                                  pd_u64 &n_event) {
  double yd[PD_C], mult[PD_NMULT_DIM], r[PD_NFLOW_DIM], cumul[PD_NFLOW_DIM];
#pragma unroll
  for (int c = 0; c < PD_C; ++c) yd[c] = (double)y[c];
  pd_event_mults(yd, U, pm, time, (const double *)0, mult);

  /* pass 1: cumulative intervals over the live events + Python's sum().  Which rows are live is
   * kept as a bit mask (row indices are literals, so word and bit resolve at compile time). */
  double cum = 0.0;            /* cumul_interval, :170-173                      */
  double f = 0.0, comp = 0.0;  /* sum(): running value and Neumaier compensation */
  bool on_float_path = false;  /* sum() has met its first float                  */
  pd_u32 live[PD_GIL_WORDS];
#pragma unroll
  for (int w = 0; w < PD_GIL_WORDS; ++w) live[w] = 0;
  int last_live = -1;          /* PD_GIL_OPT & 32: all pass 2 needs to know about liveness */
  bool row0_live = false;
#define PD_SETBIT(words, e, p) words[(e) >> 5] |= (pd_u32)(p) << ((e) & 31)
#if PD_GIL_OPT & 32
#define PD_MARK_LIVE(e, p) { if (p) last_live = (e); if ((e) == 0) row0_live = (p); }
#else
#define PD_MARK_LIVE(e, p) PD_SETBIT(live, e, p)
#endif
/* One row joins the running sums.  Branch-free: a dead row (rate not > 0, :712-733) adds +0.0,
 * which changes neither the cumulative sum nor the compensated one (t == f, the correction term
 * is (f - f) + 0), so every lane executes the same straight-line code and nothing has to be
 * merged after a branch; its cumul[] entry repeats the previous one and is never looked at
 * (pass 2 masks with the live bits).  Python's sum() adds its first float with a plain `+` and
 * compensates from the second one on (on_float_path). */
#define PD_ACC_INT(e)                                                                \
  { cum += r[e]; cumul[e] = cum; PD_MARK_LIVE(e, true); if (!PD_NAIVE_SUM) f += r[e]; }
#define PD_ACC_FLOAT(e, live_)                                                       \
  {                                                                                  \
    const bool lv = (live_);                                                         \
    const double rr = (PD_GIL_OPT & 2) ? pd_keep_if(lv, r[e]) : (lv ? r[e] : 0.0);   \
    cum += rr; cumul[e] = cum; PD_MARK_LIVE(e, lv);                               \
    if (!PD_NAIVE_SUM) {                                                             \
      /* CPython's correction term (hi - t) + lo, hi the larger magnitude, IS the rounding  \
       * error of f + rr, which is exactly representable; Knuth's branch-free TwoSum yields  \
       * the same number without the compare and four selects */                           \
      const double t = f + rr;                                                       \
      const double bb = t - f;                                                       \
      const double ci = (f - (t - bb)) + (rr - bb);                                  \
      comp += on_float_path ? ci : 0.0;                                              \
      f = t; on_float_path |= lv;                                                    \
    }                                                                                \
  }
#define PD_G1_ENTRY(e, to, is_int, rate)                                             \
  r[e] = rate; if (is_int) PD_ACC_INT(e) else PD_ACC_FLOAT(e, true)
#define PD_G1_TRANSFER(e, from, to, rate) r[e] = rate; PD_ACC_FLOAT(e, r[e] > 0)
#define PD_G1_DEATH(e, from, rate) r[e] = rate; PD_ACC_FLOAT(e, r[e] > 0)
/* the exact pass: every row tested, dead rows zeroed, liveness recorded */
#define PD_GIL_PASS1_EXACT                                                           \
  {                                                                                  \
    cum = 0.0; f = 0.0; comp = 0.0; on_float_path = false;                           \
    last_live = -1; row0_live = false;                                               \
    PD_FOR_EACH_LIVE_ROW(PD_G1_ENTRY, PD_G1_TRANSFER, PD_G1_DEATH)                   \
  }
#if PD_GIL_FAST
  /* the guard, on the rows' MULTIPLIERS (the rate expressions with every count set to 1.0, which
   * folds to the multiplier itself): counts are +0.0 or positive, so a multiplier without a sign
   * bit, finite and not NaN gives a rate that is +0.0 or positive.  Parameters sit in the
   * constant bank or are fixed per member: their share of the maximum leaves the event loop. */
  unsigned top = 0u;
  {
    double ones[PD_C];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) ones[c] = 1.0;
#define PD_GM_ROW(rate)                                                              \
  { const double (&yd)[PD_C] = ones; top = max(top, (unsigned)__double2hiint(rate)); }
#define PD_GM_ENTRY(e, to, is_int, rate) PD_GM_ROW(rate)
#define PD_GM_TRANSFER(e, from, to, rate) PD_GM_ROW(rate)
#define PD_GM_DEATH(e, from, rate) PD_GM_ROW(rate)
    PD_FOR_EACH_LIVE_ROW(PD_GM_ENTRY, PD_GM_TRANSFER, PD_GM_DEATH)
#undef PD_GM_ENTRY
#undef PD_GM_TRANSFER
#undef PD_GM_DEATH
#undef PD_GM_ROW
  }
  const bool fast = top < 0x7ff00000u;
  if (fast) {
#define PD_GF_ROW(e, rate)                                                           \
  {                                                                                  \
    const double r_ = (rate);                                                        \
    const double t = cum + r_;                                                       \
    if (!PD_NAIVE_SUM) {                                                             \
      const double bb = t - cum;                                                     \
      comp += (cum - (t - bb)) + (r_ - bb);                                          \
    }                                                                                \
    cum = t; cumul[e] = t;                                                           \
  }
#define PD_GF_ENTRY(e, to, is_int, rate) PD_GF_ROW(e, rate)
#define PD_GF_TRANSFER(e, from, to, rate) PD_GF_ROW(e, rate)
#define PD_GF_DEATH(e, from, rate) PD_GF_ROW(e, rate)
    PD_FOR_EACH_LIVE_ROW(PD_GF_ENTRY, PD_GF_TRANSFER, PD_GF_DEATH)
#undef PD_GF_ENTRY
#undef PD_GF_TRANSFER
#undef PD_GF_DEATH
#undef PD_GF_ROW
    f = cum;                                    /* sum()'s running value IS the cumulative one */
    /* a live row: an entry row (always), or a rate above zero -- and rates that are all +0.0
     * or positive add up to more than zero exactly when one of them is */
    const bool any_fast = PD_GIL_HAS_ENTRY || cum > 0.0;
    on_float_path = any_fast;
    last_live = any_fast ? 0 : -1;
  } else PD_GIL_PASS1_EXACT
#else
  const bool fast = false;
  PD_GIL_PASS1_EXACT
#endif

  pd_u32 any = (PD_GIL_OPT & 32) ? (pd_u32)(last_live >= 0) : 0u;
#pragma unroll
  for (int w = 0; w < PD_GIL_WORDS; ++w) any |= live[w];
  if (!any) {                                   /* :767-770 */
    const double dt = new_time - time;
    time += dt;
    return 0;
  }
  double u1, u2;                                /* :773 -> :174, then :776 */
  pd_two_uniforms(a, rng, u1, u2);
  const double point = u1 * cum;
  double total;
  if (PD_NAIVE_SUM) total = cum;
  else {
    total = f;
    if (on_float_path && comp != 0.0 && pd_isfinite(comp)) total = f + comp;
  }
  if (rng.exhausted) return PD_ERR_STREAM;
  if (total == 0.0) return PD_ERR_ZERO_RATE;
  if (u2 == 0.0) return PD_ERR_LOG0;
  const double dt = -((PD_GIL_OPT & 8) ? pd_log_any(u2) : log(u2)) / total;

  /* pass 2: first live event with not (point > cumul), else the last live one (:176-177) */
  int sel = -1;
#if PD_GIL_OPT & 32
  {
    /* cumul[] never decreases and a dead row repeats its predecessor's entry, so the first row
     * with not (point > cumul) -- the NUMBER of rows with point > cumul -- raised the cumulative
     * sum, i.e. is live, unless it sits before the first live row (cumul still 0, which takes
     * point == 0).  No liveness mask then: a count, the index of the last live row for a point
     * beyond every interval, and a cold search for the first live row. */
    int n_above = 0;
#define PD_G2_ENTRY(e, to, is_int, rate) if (point > cumul[e]) ++n_above;
#define PD_G2_TRANSFER(e, from, to, rate) if (point > cumul[e]) ++n_above;
#define PD_G2_DEATH(e, from, rate) if (point > cumul[e]) ++n_above;
    PD_FOR_EACH_LIVE_ROW(PD_G2_ENTRY, PD_G2_TRANSFER, PD_G2_DEATH)
#undef PD_G2_ENTRY
#undef PD_G2_TRANSFER
#undef PD_G2_DEATH
    bool resolved = false;
#if PD_GIL_FAST
    if (fast) {
      /* after the fast pass: a row in the middle of the table raised the sum past the point, so
       * it is live; row 0 is live when it is an entry row or its rate is above zero.  Else the
       * point is 0 in front of dead rows, and the first live row is the first that raised the
       * sum (the total is above zero here, PD_ERR_ZERO_RATE has been dealt with).  A point
       * beyond every interval does not happen with finite rates: u1 < 1, so u1 * cum <= cum. */
      resolved = true;
      if (n_above > 0 || PD_GIL_ROW0_ENTRY || cumul[0] > 0.0)
        sel = n_above < PD_LIVE_NROW ? n_above : PD_LIVE_NROW - 1;
      else {
        sel = PD_LIVE_NROW - 1;
#pragma unroll
        for (int e = PD_LIVE_NROW - 1; e >= 0; --e)
          if (cumul[e] > 0.0) sel = e;
      }
    }
#endif
    if (!resolved) {
      if (n_above >= PD_LIVE_NROW) sel = last_live;
      else if (n_above > 0 || row0_live) sel = n_above;
      else {
        sel = last_live;
#pragma unroll
        for (int e = PD_LIVE_NROW - 1; e >= 0; --e)
          if (cumul[e] > 0.0) sel = e;            /* the first row that raised the sum */
      }
    }
  }
#else
  pd_u32 above[PD_GIL_WORDS];                   /* rows with point > cumul[e] */
#pragma unroll
  for (int w = 0; w < PD_GIL_WORDS; ++w) above[w] = 0;
#define PD_G2_ENTRY(e, to, is_int, rate) PD_SETBIT(above, e, point > cumul[e]);
#define PD_G2_TRANSFER(e, from, to, rate) PD_SETBIT(above, e, point > cumul[e]);
#define PD_G2_DEATH(e, from, rate) PD_SETBIT(above, e, point > cumul[e]);
  PD_FOR_EACH_LIVE_ROW(PD_G2_ENTRY, PD_G2_TRANSFER, PD_G2_DEATH)
#undef PD_G2_ENTRY
#undef PD_G2_TRANSFER
#undef PD_G2_DEATH
#pragma unroll
  for (int w = 0; w < PD_GIL_WORDS; ++w) {
    const pd_u32 hit = live[w] & ~above[w];
    if (sel < 0 && hit) sel = w * 32 + __ffs(hit) - 1;
  }
  if (sel < 0) {
#pragma unroll
    for (int w = PD_GIL_WORDS - 1; w >= 0; --w)
      if (sel < 0 && live[w]) sel = w * 32 + 31 - __clz(live[w]);
  }
#endif
#undef PD_SETBIT
#undef PD_GIL_PASS1_EXACT
#undef PD_G1_ENTRY
#undef PD_G1_TRANSFER
#undef PD_G1_DEATH
#undef PD_ACC_INT
#undef PD_ACC_FLOAT
#undef PD_MARK_LIVE

  /* apply +-1 (:778-787) */
#if (PD_GIL_OPT & 16) && PD_NFLOW_DIM <= 64
  {
    /* Branch-free: which rows add to / take from compartment c is a compile-time bit mask over
     * the rows (indices are literals, the loops fold), so the update is two mask tests per
     * compartment with every lane on.  The `if (sel == e)` chain below compiles to a jump table --
     * an indexed constant load and an indirect branch that the warp takes once per distinct row
     * its lanes picked (9 percent of the kernel's stall samples on the TB model). */
#if PD_NFLOW_DIM <= 32
    typedef pd_u32 row_mask;
#else
    typedef pd_u64 row_mask;
#endif
    row_mask gain[PD_C], loss[PD_C];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) gain[c] = loss[c] = 0;
#define PD_G3_ENTRY(e, to, is_int, rate) gain[to] |= (row_mask)1 << (e);
#define PD_G3_TRANSFER(e, from, to, rate) loss[from] |= (row_mask)1 << (e); gain[to] |= (row_mask)1 << (e);
#define PD_G3_DEATH(e, from, rate) loss[from] |= (row_mask)1 << (e);
    PD_FOR_EACH_LIVE_ROW(PD_G3_ENTRY, PD_G3_TRANSFER, PD_G3_DEATH)
#undef PD_G3_ENTRY
#undef PD_G3_TRANSFER
#undef PD_G3_DEATH
    const row_mask picked = (row_mask)1 << sel;
#pragma unroll
    for (int c = 0; c < PD_C; ++c)
      y[c] += (pd_count)((picked & gain[c]) != 0) - (pd_count)((picked & loss[c]) != 0);
  }
#else
#define PD_G3_ENTRY(e, to, is_int, rate) if (sel == e) y[to] += 1;
#define PD_G3_TRANSFER(e, from, to, rate) if (sel == e) { y[from] -= 1; y[to] += 1; }
#define PD_G3_DEATH(e, from, rate) if (sel == e) y[from] -= 1;
  PD_FOR_EACH_LIVE_ROW(PD_G3_ENTRY, PD_G3_TRANSFER, PD_G3_DEATH)
#undef PD_G3_ENTRY
#undef PD_G3_TRANSFER
#undef PD_G3_DEATH
#endif
#if PD_HAS_STEP_CHECKS
  {
    /* a user-defined checks() (:789): the updated counts, the vars computed before the event */
    double y2[PD_C];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) y2[c] = (double)y[c];
    if (!pd_step_checks(yd, y2, U, pm, time, (const double *)0)) return PD_ERR_CHECKS;
  }
#endif
  time += dt;                                   /* :790 */
  ++n_event;
  return 0;
}

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_gillespie_kernel(const __grid_constant__ PdGilLaunch L) {
  const PdStochArgs &a = L.a;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + threadIdx.x;
#if PD_USE_MASK
  const bool active = m < a.n_member && a.mask[m < a.n_member ? m : 0] == (unsigned char)a.mask_value;
#else
  const bool active = m < a.n_member;
#endif
  const pd_i64 M = a.n_member;

#if PD_GIL_WINDOW
  extern __shared__ double s_dyn[];
  pd_count *const s_col = (pd_count *)s_dyn + threadIdx.x;  /* [slot][c][PD_BLOCK], own column */
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
      y[c] = (pd_count)v;                                   /* :747-748 */
#if PD_OUT_MODE == PD_OUT_FULL
      a.soln[(size_t)c * a.out_ld + m] = v;                        /* :755 */
#endif
    }
    rng.init(a, m);
  } else {
#pragma unroll
    for (int c = 0; c < PD_C; ++c) y[c] = 0;
  }

  int bad = 0;
  pd_u64 n_event = 0;
  double time = a.target[0];

#if PD_GIL_WINDOW
  /* FULL: row 0 is the untruncated y written above; STATS: target 0 goes through the window */
  const int first = (PD_OUT_MODE == PD_OUT_FULL) ? 1 : 0;
  int it = active ? first : a.n_time;          /* next target time this lane has to park     */
  int flushed = first;                         /* warp-uniform: targets below it are recorded */
  for (;;) {
    /* park every target time already reached (:762 false -> :793-795), window permitting; a
     * failed member parks its last counts for the rest so that the warp can finish */
    while (it < a.n_time && it - flushed < PD_GIL_SLOTS &&
           (it == 0 || bad || !(time < a.target[it]))) {
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        s_col[((it & (PD_GIL_SLOTS - 1)) * PD_C + c) * PD_BLOCK] = y[c];
      ++it;
    }
    /* record, all lanes together, what every lane of the warp has parked */
    const int ready = __reduce_min_sync(PD_FULL_MASK, it);
    while (flushed < ready) {
      pd_count ys[PD_C];
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        ys[c] = s_col[((flushed & (PD_GIL_SLOTS - 1)) * PD_C + c) * PD_BLOCK];
#if PD_OUT_MODE == PD_OUT_FULL
      if (active) {
#pragma unroll
        for (int c = 0; c < PD_C; ++c)
          a.soln[((size_t)flushed * PD_C + c) * a.out_ld + m] = (double)ys[c];
      }
#elif PD_HIST_MOMENTS
      if (active) pd_hist_record(a, ys, flushed, a.member0 + m);
#else
      pd_stats_warp_record(a, ys, active, flushed, a.member0 + m);
#endif
      ++flushed;
    }
    if (flushed >= a.n_time) break;
    /* one reaction for every lane that is still inside its current interval (:762) */
    if (it < a.n_time && !bad && time < a.target[it])
      bad = pd_gillespie_event(a, L.U, pm, y, time, a.target[it], rng, n_event);
  }
#else
  if (active) {
    /* Every trip of the loop is one reaction for EVERY unfinished lane: a lane first steps over
     * the target times it has already reached (a short predicated loop, usually no trip). */
    int it = 1;
    for (;;) {
      while (it < a.n_time && !(time < a.target[it])) ++it;  /* :762 false */
      if (it >= a.n_time) break;
      bad = pd_gillespie_event(a, L.U, pm, y, time, a.target[it], rng, n_event);
      if (bad) break;
    }
  }
#endif

  if (active) {
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * a.out_ld + m] = (double)y[c];
#endif
    if (bad) pd_raise(a.err, bad, a.member0 + m, 0);
  }
  if (a.n_steps) {
    const pd_u64 warp_events = pd_warp_sum_u64(n_event);
    if ((threadIdx.x & 31) == 0 && warp_events) atomicAdd(a.n_steps, warp_events);
  }
}
