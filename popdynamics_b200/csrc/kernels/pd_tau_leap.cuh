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
 */
#include "pd_model.h"
#include "pd_common.cuh"

struct PdTauLaunch {
  PdStochArgs a;
  PdUniform U;
};

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_tau_leap_kernel(const __grid_constant__ PdTauLaunch L) {
  const PdStochArgs &a = L.a;
  const PdUniform &U = L.U;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + threadIdx.x;
  const bool active = m < a.n_member;
  const pd_i64 M = a.n_member;

#if PD_OUT_MODE == PD_OUT_STATS
  __shared__ PdStatsSmem s_stats;
  extern __shared__ pd_u32 s_hist[];
  for (int w = threadIdx.x; w < PD_C * (a.hist_bins >> 1); w += PD_BLOCK) s_hist[w] = 0;
  __syncthreads();
#endif

  pd_count y[PD_C];
  double pm[PD_NPM_DIM];
  PdRng rng;
  pm[0] = 0.0;
  if (active) {
#pragma unroll
    for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * M + m];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) {
      const double v = pd_load_y0(a.y0, a.y0_stride_c, a.y0_stride_m, c, m);
      y[c] = (pd_count)v;                                   /* int() truncation, :805-806 */
#if PD_OUT_MODE == PD_OUT_FULL
      a.soln[(size_t)c * M + m] = v;                        /* row 0 = untruncated y, :813 */
#endif
    }
    rng.init(a, m);
  } else {
#pragma unroll
    for (int c = 0; c < PD_C; ++c) y[c] = 0;
  }
#if PD_OUT_MODE == PD_OUT_STATS
  pd_stats_record(a, s_stats, s_hist, y, active, 0, m);
#endif

  int bad = 0;
  double time = a.target[0];
  for (int it = 1; it < a.n_time; ++it) {
    if (active) {
      const double dt = a.target[it] - a.target[it - 1];    /* :820 */
      /* start-of-step counts as doubles: every rate of this step is formed from them, the
       * live counts y[] only enter through the clamps (:824 vs :833-841) */
      double yd[PD_C], mult[PD_NMULT_DIM];
#pragma unroll
      for (int c = 0; c < PD_C; ++c) yd[c] = (double)y[c];
      pd_event_mults(yd, U, pm, time, (const double *)0, mult);   /* :822-824 */

/* draws are made in 64 bits; with 32-bit counts a gain that would overflow raises the
 * error flag instead of wrapping (numpy / Python ints never wrap) */
#define PD_TAU_GAIN(to, d)                                                           \
  {                                                                                  \
    if (sizeof(pd_count) == 4 && (pd_i64)y[to] + (d) > 0x7fffffffLL) {               \
      if (!bad) bad = PD_ERR_COUNT_RANGE;                                            \
    } else y[to] += (pd_count)(d);                                                   \
  }
#define PD_TAU_ENTRY(e, to, is_int, rate)                                            \
  {                                                                                  \
    const double mean = rate * dt;                                                   \
    if (!(mean >= 0.0)) { if (!bad) bad = PD_ERR_POISSON_MEAN; }                     \
    else { const pd_i64 d = pd_poisson(rng, mean); PD_TAU_GAIN(to, d) } /* :842-844 */\
  }
#define PD_TAU_TRANSFER(e, from, to, rate)                                           \
  if (rate > 0) {                                                                    \
    const double mean = rate * dt;                                                   \
    if (!(mean >= 0.0)) { if (!bad) bad = PD_ERR_POISSON_MEAN; }                     \
    else {                                                                           \
      pd_i64 d = pd_poisson(rng, mean);                                              \
      if (d > (pd_i64)y[from]) d = (pd_i64)y[from];              /* :833-834 */      \
      y[from] -= (pd_count)d; PD_TAU_GAIN(to, d)                                     \
    }                                                                                \
  }
#define PD_TAU_DEATH(e, from, rate)                                                  \
  if (rate > 0) {                                                                    \
    const double mean = rate * dt;                                                   \
    if (!(mean >= 0.0)) { if (!bad) bad = PD_ERR_POISSON_MEAN; }                     \
    else {                                                                           \
      pd_i64 d = pd_poisson(rng, mean);                                              \
      if (d > (pd_i64)y[from]) d = (pd_i64)y[from];              /* :839-840 */      \
      y[from] -= (pd_count)d;                                                        \
    }                                                                                \
  }
      PD_FOR_EACH_FLOW(PD_TAU_ENTRY, PD_TAU_TRANSFER, PD_TAU_DEATH)
#undef PD_TAU_ENTRY
#undef PD_TAU_TRANSFER
#undef PD_TAU_DEATH
#undef PD_TAU_GAIN
      time += dt;                                           /* :848 */
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        a.soln[((size_t)it * PD_C + c) * M + m] = (double)y[c];   /* :850-852 */
#endif
    }
#if PD_OUT_MODE == PD_OUT_STATS
    pd_stats_record(a, s_stats, s_hist, y, active, it, m);
#endif
  }

  if (active) {
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * M + m] = (double)y[c];
#endif
    if (!bad && rng.exhausted) bad = PD_ERR_STREAM;
    if (bad) pd_raise(a.err, bad, a.member0 + m, 0);
  }
  if (threadIdx.x == 0 && a.n_steps) {
    const pd_i64 first = (pd_i64)blockIdx.x * PD_BLOCK;
    const pd_i64 here = (M - first < PD_BLOCK) ? (M - first) : PD_BLOCK;
    atomicAdd(a.n_steps, (pd_u64)here * (pd_u64)(a.n_time - 1));
  }
}
