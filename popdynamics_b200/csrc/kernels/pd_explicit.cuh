/*
 * pd_explicit.cuh -- explicit Euler ODE integrator, one thread per ensemble member, fp64
 * compartment state and flows in registers.
 *
 * Restates /root/reference/basepop.py:658-688 (integrate_explicit) with the derivative of
 * :613-626 / :560-598 generated into pd_derivative():
 *   per target interval: sub-steps of min_dt with fp64 time accumulation and a snap to the target
 *   on overshoot (:675-682).  That schedule does not depend on the member, so the host replays
 *   it once (pd_explicit_schedule in the C-ABI) and ships dt[] / n_sub[] -- and, because the
 *   scale-up closures (:544-551) depend on time only, their values at every sub-step;
 *   all flows are computed from the OLD state before any component is updated (:676, :683-684);
 *   y[i] = y[i] + dt * f[i], then negatives are clamped to zero (:684-687);
 *   checks() runs on the state handed to the derivative (:623).
 * Uniform loads (dt, scale-ups) are the same address for every lane: one broadcast transaction.
 *
 * The kernel is bound by the FP64 pipe, and a fifth of its FP64 instructions were comparisons
 * (DSETP): per compartment the clamp `y < 0` and the default checks() `y >= 0` (:1016-1025), per
 * traced sum() the `c != 0 and isfinite(c)` of CPython's compensated fold.  PD_EXPL_OPT takes them
 * out of the common path without changing a single result (moving a comparison to the integer
 * ALU alone bought little, leaving instructions out is what paid: DESIGN.md 3.3):
 *   bit 0  the two tests of the sum are exact bit tests (pd_math.cuh);
 *   bit 1  the default checks() is fused with the clamp.  After the update z = y + dt f the next
 *          derivative call sees clamp(z), and `clamp(z) >= 0` fails exactly when z is NaN.  One
 *          unsigned maximum over the high words of the z decides the sub-step: below 0x7ff00000
 *          every z is a finite number without a sign bit -- nothing to clamp, the next checks()
 *          passes, no comparison at all.  Otherwise some z has a sign bit or an all-ones
 *          exponent: negative finite values are clamped on the sign bit (`z < 0` is the sign
 *          bit for anything but NaN and -0.0), and an infinity or a NaN sends the sub-step to the
 *          fp64 comparisons themselves.  -0.0 can only come from -0.0 + -0.0 (round to nearest),
 *          i.e. from a compartment that was -0.0 in the initial state: a member that starts
 *          with one takes the exact path throughout.
 * POPDYN_DEFINE=-DPD_EXPL_OPT=0 rebuilds the plain forms (tests/test_gpu_explicit_forms.py compares them).
 */
#ifndef PD_EXPL_OPT
#define PD_EXPL_OPT 3
#endif
#ifndef PD_EXPL_UNROLL
#define PD_EXPL_UNROLL 4          /* sub-steps per trip of the inner loop */
#endif
#include "pd_math.cuh"
#if PD_EXPL_OPT & 1
#define PD_ISFINITE(x) pd_isfinite_bits(x)
#define PD_NE0(x) pd_ne_zero_bits(x)
#endif
#if PD_EXPL_OPT & 2
#define PD_SKIP_DEFAULT_CHECKS 1
#endif
#include "pd_model.h"
#include "pd_common.cuh"
#define PD_EXPL_FUSED (((PD_EXPL_OPT) & 2) && PD_DEFAULT_CHECKS)

struct PdExplicitLaunch {
  PdExplicitArgs a;
  PdUniform U;
};

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_explicit_kernel(const __grid_constant__ PdExplicitLaunch L) {
  const PdExplicitArgs &a = L.a;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + threadIdx.x;
  const pd_i64 M = a.n_member;
  if (m >= M) return;
#if PD_USE_MASK
  if (a.mask[m] != (unsigned char)a.mask_value) return;
#endif

  double y[PD_C], f[PD_C], pm[PD_NPM_DIM];
  pm[0] = 0.0;
#pragma unroll
  for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * a.pm_ld + m];
#pragma unroll
  for (int c = 0; c < PD_C; ++c) {
    y[c] = pd_load_y0(a.y0, a.y0_stride_c, a.y0_stride_m, c, m);
#if PD_OUT_MODE == PD_OUT_FULL
    a.soln[(size_t)c * a.out_ld + m] = y[c];                       /* :671 */
#endif
  }

  bool ok = true;
  int bad_at = 0;
  pd_i64 step = 0;
#if PD_EXPL_FUSED
  /* the first sub-step whose derivative call is handed a state that fails checks() (:623), and
   * whether this member can hold a -0.0 */
  pd_i64 fail_step = -1;
  bool exact_only = false;
#pragma unroll
  for (int c = 0; c < PD_C; ++c) {
    if (!(y[c] >= 0.)) fail_step = 0;
    exact_only = exact_only || (__double2hiint(y[c]) == (int)0x80000000);
  }
#endif
  for (int it = 1; it < a.n_time; ++it) {                   /* :672 */
    const int n_sub = a.n_sub[it - 1];
#pragma unroll PD_EXPL_UNROLL
    for (int s = 0; s < n_sub; ++s, ++step) {               /* :675 */
      const double dt = a.dt[step];
#if PD_USES_TIME
      const double t = a.t_eval[step];
#else
      const double t = 0.0;
#endif
#if PD_NSU > 0
      const double *su = a.su_table + step * PD_NSU;
#else
      const double *su = (const double *)0;
#endif
#if PD_EXPL_FUSED
      pd_derivative(y, L.U, pm, t, su, f);                  /* :676, default checks left out */
      unsigned hi_top = 0;
#pragma unroll
      for (int c = 0; c < PD_C; ++c) {
        y[c] = y[c] + dt * f[c];                            /* :684 */
        hi_top = max(hi_top, (unsigned)__double2hiint(y[c]));
      }
      /* largest high word below 0x7ff00000: every new value is a finite number without a sign
       * bit -- nothing to clamp (:686-687), and the next checks() passes */
      if (exact_only || hi_top >= 0x7ff00000u) {
        int hi_signed = 0;
        unsigned hi_unsigned = 0;
#pragma unroll
        for (int c = 0; c < PD_C; ++c) {
          const int hi = __double2hiint(y[c]);
          hi_signed = max(hi_signed, hi);
          hi_unsigned = max(hi_unsigned, (unsigned)hi);
        }
        if (exact_only || hi_signed >= 0x7ff00000 || hi_unsigned >= 0xfff00000u) {
          /* an infinity, a NaN or a possible -0.0 among them: the comparisons themselves */
          bool next_ok = true;
#pragma unroll
          for (int c = 0; c < PD_C; ++c) {
            if (y[c] < 0.) y[c] = 0.;
            next_ok = next_ok && (y[c] >= 0.);
          }
          if (!next_ok && fail_step < 0) fail_step = step + 1;
        } else {
          /* negative finite values (and nothing else with a sign bit): clamp on the sign bit */
#pragma unroll
          for (int c = 0; c < PD_C; ++c)
            if (__double2hiint(y[c]) < 0) y[c] = 0.;
        }
      }
#else
      const bool step_ok = pd_derivative(y, L.U, pm, t, su, f);   /* :676 */
      if (!step_ok && ok) { ok = false; bad_at = it; }            /* :623 */
#pragma unroll
      for (int c = 0; c < PD_C; ++c) {
        y[c] = y[c] + dt * f[c];                            /* :684 */
        if (y[c] < 0.) y[c] = 0.;                           /* :686-687 */
      }
#endif
    }
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.soln[((size_t)it * PD_C + c) * a.out_ld + m] = y[c];  /* :688 */
#endif
  }
#if PD_EXPL_FUSED
  if (fail_step >= 0 && fail_step < step) {
    /* a derivative call did see that state (the one after the last sub-step never happens):
     * the target interval of that call, as the plain form reports it */
    pd_i64 seen = 0;
    int it = 1;
    for (; it < a.n_time - 1; ++it) {
      seen += a.n_sub[it - 1];
      if (fail_step < seen) break;
    }
    ok = false;
    bad_at = it;
  }
#endif
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
  for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * a.out_ld + m] = y[c];
#endif
#if PD_OUT_MODE == PD_OUT_FINAL_VARS
  {
    /* calculate_diagnostics of the final state (:927-948), fused: scale-up vars keep the values
     * of the last derivative call (the reference never re-evaluates them in that pass) */
#if PD_NSU > 0
    const double *su_last = a.su_table + (a.n_steps > 0 ? a.n_steps - 1 : 0) * PD_NSU;
#else
    const double *su_last = (const double *)0;
#endif
    double fv[PD_NFVAR_DIM];
    pd_final_vars(y, L.U, pm, a.t_final, su_last, fv);
#pragma unroll
    for (int v = 0; v < PD_NFVAR; ++v) a.final_state[(size_t)v * a.out_ld + m] = fv[v];
  }
#endif
  if (!ok) pd_raise(a.err, PD_ERR_CHECKS, a.member0 + m, bad_at);
}
