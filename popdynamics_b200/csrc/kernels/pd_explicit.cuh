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
 */
#include "pd_model.h"
#include "pd_common.cuh"

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
  for (int it = 1; it < a.n_time; ++it) {                   /* :672 */
    const int n_sub = a.n_sub[it - 1];
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
      const bool step_ok = pd_derivative(y, L.U, pm, t, su, f);   /* :676 */
      if (!step_ok && ok) { ok = false; bad_at = it; }            /* :623 */
#pragma unroll
      for (int c = 0; c < PD_C; ++c) {
        y[c] = y[c] + dt * f[c];                            /* :684 */
        if (y[c] < 0.) y[c] = 0.;                           /* :686-687 */
      }
    }
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.soln[((size_t)it * PD_C + c) * a.out_ld + m] = y[c];  /* :688 */
#endif
  }
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
