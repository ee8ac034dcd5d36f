/*
 * pd_adaptive.cuh -- adaptive ODE integrator, one thread per ensemble member: the device
 * counterpart of integrate(method='scipy') (/root/reference/basepop.py:882-886), where the
 * reference hands the derivative closure of :613-626 to scipy.integrate.odeint (LSODA, default
 * rtol = atol = 1.49012e-8) and stores the state at every target time.
 *
 * LSODA's Fortran step / order / method control cannot be reproduced bit for bit, so parity for
 * this method is agreement within a stated tolerance, not equality.  What runs here is
 * the explicit Dormand-Prince 5(4) pair with first-same-as-last and a per-member step
 * controller -- the algorithm of scipy.integrate.RK45, restated:
 *   error norm   rms( err_i / (atol + rtol * max(|y_i|, |ynew_i|)) )
 *   accept       norm < 1, next h *= min(10, 0.9 * norm^-1/5)  (at most 1 after a rejection)
 *   reject       h *= max(0.2, 0.9 * norm^-1/5)
 *   first step   Hairer / Norsett / Wanner's estimate (RK45.select_initial_step)
 * A step never crosses a target time: it is shortened to land on it exactly and the state is
 * recorded there, so no dense output is involved.  The derivative is the same generated
 * pd_derivative() the explicit Euler kernel uses (traced calculate_vars + calculate_flows);
 * scale-up closures are traced on the symbolic time, because every member is at its own time.
 * checks() (:623) is evaluated on accepted states; like the Euler kernel, a member that fails it
 * is reported (first failure) and integrated to the end.  Bound: FP64 pipe.
 */
#include "pd_model.h"
#include "pd_common.cuh"

#ifndef PD_ADAPT_TAB
#define PD_ADAPT_TAB 1
#endif
#if PD_ADAPT_TAB
__constant__ double pd_dp_tab[30] = {
    1.0 / 5.0,             /* A21 */
    3.0 / 40.0,            /* A31 */
    9.0 / 40.0,            /* A32 */
    44.0 / 45.0,           /* A41 */
    -56.0 / 15.0,          /* A42 */
    32.0 / 9.0,            /* A43 */
    19372.0 / 6561.0,      /* A51 */
    -25360.0 / 2187.0,     /* A52 */
    64448.0 / 6561.0,      /* A53 */
    -212.0 / 729.0,        /* A54 */
    9017.0 / 3168.0,       /* A61 */
    -355.0 / 33.0,         /* A62 */
    46732.0 / 5247.0,      /* A63 */
    49.0 / 176.0,          /* A64 */
    -5103.0 / 18656.0,     /* A65 */
    35.0 / 384.0,          /* B1 */
    500.0 / 1113.0,        /* B3 */
    125.0 / 192.0,         /* B4 */
    -2187.0 / 6784.0,      /* B5 */
    11.0 / 84.0,           /* B6 */
    -71.0 / 57600.0,       /* E1 */
    71.0 / 16695.0,        /* E3 */
    -71.0 / 1920.0,        /* E4 */
    17253.0 / 339200.0,    /* E5 */
    -22.0 / 525.0,         /* E6 */
    1.0 / 40.0,            /* E7 */
    1.0 / 5.0,             /* C2 */
    3.0 / 10.0,            /* C3 */
    4.0 / 5.0,             /* C4 */
    8.0 / 9.0,             /* C5 */
};
#endif

struct PdAdaptiveLaunch {
  PdAdaptiveArgs a;
  PdUniform U;
};

/* err^(-1/5) of the step controller as exp(-0.2 log(err)) with the constant-bank forms of
 * pd_math.cuh: within 2 ulp of pow(err, -0.2), which is itself only within an ulp of the libm value
 * the CPU oracle uses (parity of this method is a tolerance, see the file header), at half the
 * instructions of the library's general pow, once per attempted step (out of line, like pow: its
 * temporaries must not add to the registers of the stage loop).  err == 0 gives +inf, as pow. */
__device__ __noinline__ double pd_inv_fifth_root(double err) {
  return pd_exp_any(-0.2 * pd_log_any(err));
}

__device__ __forceinline__ double pd_rms_scaled(const double (&v)[PD_C], const double (&scale)[PD_C]) {
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < PD_C; ++c) { const double r = v[c] / scale[c]; s += r * r; }
  return sqrt(s / (double)PD_C);
}

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_adaptive_kernel(const __grid_constant__ PdAdaptiveLaunch L) {
  const PdAdaptiveArgs &a = L.a;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + threadIdx.x;
  if (m >= a.n_member) return;

  /* Dormand-Prince 5(4) tableau: in the constant bank (PD_ADAPT_TAB, the default), so that a
   * coefficient is a c[][] operand of its DMUL; as literals every use costs two UMOVs (92 of the
   * 746 instructions of an attempted step) */
#if PD_ADAPT_TAB
#define A21 pd_dp_tab[0]
#define A31 pd_dp_tab[1]
#define A32 pd_dp_tab[2]
#define A41 pd_dp_tab[3]
#define A42 pd_dp_tab[4]
#define A43 pd_dp_tab[5]
#define A51 pd_dp_tab[6]
#define A52 pd_dp_tab[7]
#define A53 pd_dp_tab[8]
#define A54 pd_dp_tab[9]
#define A61 pd_dp_tab[10]
#define A62 pd_dp_tab[11]
#define A63 pd_dp_tab[12]
#define A64 pd_dp_tab[13]
#define A65 pd_dp_tab[14]
#define B1 pd_dp_tab[15]
#define B3 pd_dp_tab[16]
#define B4 pd_dp_tab[17]
#define B5 pd_dp_tab[18]
#define B6 pd_dp_tab[19]
#define E1 pd_dp_tab[20]
#define E3 pd_dp_tab[21]
#define E4 pd_dp_tab[22]
#define E5 pd_dp_tab[23]
#define E6 pd_dp_tab[24]
#define E7 pd_dp_tab[25]
#define C2 pd_dp_tab[26]
#define C3 pd_dp_tab[27]
#define C4 pd_dp_tab[28]
#define C5 pd_dp_tab[29]
#else
  const double A21 = 1.0 / 5.0;
  const double A31 = 3.0 / 40.0;
  const double A32 = 9.0 / 40.0;
  const double A41 = 44.0 / 45.0;
  const double A42 = -56.0 / 15.0;
  const double A43 = 32.0 / 9.0;
  const double A51 = 19372.0 / 6561.0;
  const double A52 = -25360.0 / 2187.0;
  const double A53 = 64448.0 / 6561.0;
  const double A54 = -212.0 / 729.0;
  const double A61 = 9017.0 / 3168.0;
  const double A62 = -355.0 / 33.0;
  const double A63 = 46732.0 / 5247.0;
  const double A64 = 49.0 / 176.0;
  const double A65 = -5103.0 / 18656.0;
  const double B1 = 35.0 / 384.0;
  const double B3 = 500.0 / 1113.0;
  const double B4 = 125.0 / 192.0;
  const double B5 = -2187.0 / 6784.0;
  const double B6 = 11.0 / 84.0;
  const double E1 = -71.0 / 57600.0;
  const double E3 = 71.0 / 16695.0;
  const double E4 = -71.0 / 1920.0;
  const double E5 = 17253.0 / 339200.0;
  const double E6 = -22.0 / 525.0;
  const double E7 = 1.0 / 40.0;
  const double C2 = 1.0 / 5.0;
  const double C3 = 3.0 / 10.0;
  const double C4 = 4.0 / 5.0;
  const double C5 = 8.0 / 9.0;
#endif

  double y[PD_C], pm[PD_NPM_DIM];
  pm[0] = 0.0;
#pragma unroll
  for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * a.pm_ld + m];
#pragma unroll
  for (int c = 0; c < PD_C; ++c) {
    y[c] = pd_load_y0(a.y0, a.y0_stride_c, a.y0_stride_m, c, m);
#if PD_OUT_MODE == PD_OUT_FULL
    a.soln[(size_t)c * a.out_ld + m] = y[c];
#endif
  }
  const double *su = (const double *)0;
  const double rtol = a.rtol, atol = a.atol;

  double k1[PD_C], k2[PD_C], k3[PD_C], k4[PD_C], k5[PD_C], k6[PD_C], k7[PD_C], yn[PD_C], sc[PD_C];
  double t = a.target[0];
  bool ok = pd_derivative(y, L.U, pm, t, su, k1);
  int bad = ok ? 0 : PD_ERR_CHECKS, bad_at = 0;

  /* first step (RK45.select_initial_step, order 4) */
  double h = a.first_step;
  if (!(h > 0.0)) {
#pragma unroll
    for (int c = 0; c < PD_C; ++c) sc[c] = atol + fabs(y[c]) * rtol;
    const double d0 = pd_rms_scaled(y, sc), d1 = pd_rms_scaled(k1, sc);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
#pragma unroll
    for (int c = 0; c < PD_C; ++c) yn[c] = y[c] + h0 * k1[c];
    pd_derivative(yn, L.U, pm, t + h0, su, k2);
#pragma unroll
    for (int c = 0; c < PD_C; ++c) k3[c] = k2[c] - k1[c];
    const double d2 = pd_rms_scaled(k3, sc) / h0;
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3)
                                                   : pow(0.01 / fmax(d1, d2), 0.2);
    h = fmin(100.0 * h0, h1);
  }

  pd_u64 n_step = 0;
  bool rejected = false;
  for (int it = 1; it < a.n_time; ++it) {
    const double t_end = a.target[it];
    while (t < t_end) {
      if (n_step >= (pd_u64)a.max_steps) { bad = PD_ERR_MAX_STEPS; bad_at = it; break; }
      ++n_step;
      /* never step over the target: land on it exactly */
      double hs = h;
      bool last = false;
      if (!(t + hs < t_end)) { hs = t_end - t; last = true; }
      if (!(hs > 0.0) || !(t + hs > t)) { bad = PD_ERR_STEP_SIZE; bad_at = it; break; }
#pragma unroll
      for (int c = 0; c < PD_C; ++c) yn[c] = y[c] + hs * (A21 * k1[c]);
      pd_derivative(yn, L.U, pm, t + C2 * hs, su, k2);
#pragma unroll
      for (int c = 0; c < PD_C; ++c) yn[c] = y[c] + hs * (A31 * k1[c] + A32 * k2[c]);
      pd_derivative(yn, L.U, pm, t + C3 * hs, su, k3);
#pragma unroll
      for (int c = 0; c < PD_C; ++c) yn[c] = y[c] + hs * (A41 * k1[c] + A42 * k2[c] + A43 * k3[c]);
      pd_derivative(yn, L.U, pm, t + C4 * hs, su, k4);
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        yn[c] = y[c] + hs * (A51 * k1[c] + A52 * k2[c] + A53 * k3[c] + A54 * k4[c]);
      pd_derivative(yn, L.U, pm, t + C5 * hs, su, k5);
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        yn[c] = y[c] + hs * (A61 * k1[c] + A62 * k2[c] + A63 * k3[c] + A64 * k4[c] + A65 * k5[c]);
      pd_derivative(yn, L.U, pm, t + hs, su, k6);
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        yn[c] = y[c] + hs * (B1 * k1[c] + B3 * k3[c] + B4 * k4[c] + B5 * k5[c] + B6 * k6[c]);
      const double t_new = last ? t_end : t + hs;
      const bool ok_new = pd_derivative(yn, L.U, pm, t_new, su, k7);
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < PD_C; ++c) {
        const double e = hs * (E1 * k1[c] + E3 * k3[c] + E4 * k4[c] + E5 * k5[c] + E6 * k6[c] +
                               E7 * k7[c]);
        /* e is exactly zero for a compartment at rest (a fixed point of the model, an empty
         * class): the guarded division keeps such lanes off the divider's slow path, which a
         * warp otherwise runs 1.6 times per step (pd_math.cuh) */
        const double r = pd_div_zero_guard(e, atol + pd_py_max(fabs(y[c]), fabs(yn[c])) * rtol);
        s += r * r;
      }
      const double err = sqrt(s / (double)PD_C);
      const double growth = 0.9 * pd_inv_fifth_root(err);
      if (err < 1.0) {
        double factor = (err == 0.0) ? 10.0 : fmin(10.0, growth);
        if (rejected) factor = fmin(1.0, factor);
        rejected = false;
        /* a step shortened to reach the target says nothing about the step the error allows:
         * keep the longer of the two proposals */
        h = last ? fmax(h, hs * factor) : hs * factor;
        t = t_new;
#pragma unroll
        for (int c = 0; c < PD_C; ++c) { y[c] = yn[c]; k1[c] = k7[c]; }
        if (!ok_new && !bad) { bad = PD_ERR_CHECKS; bad_at = it; }      /* :623 */
      } else {
        /* NaN lands here as well: shrink */
        const double factor = (err == err) ? fmax(0.2, growth) : 0.2;
        h = hs * factor;
        rejected = true;
      }
    }
    if (bad == PD_ERR_MAX_STEPS || bad == PD_ERR_STEP_SIZE) break;
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.soln[((size_t)it * PD_C + c) * a.out_ld + m] = y[c];
#endif
  }
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
  for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * a.out_ld + m] = y[c];
#endif
#if PD_OUT_MODE == PD_OUT_FINAL_VARS
  {
    double fv[PD_NFVAR_DIM];
    pd_final_vars(y, L.U, pm, a.target[a.n_time - 1], su, fv);
#pragma unroll
    for (int v = 0; v < PD_NFVAR; ++v) a.final_state[(size_t)v * a.out_ld + m] = fv[v];
  }
#endif
  if (bad) pd_raise(a.err, bad, a.member0 + m, bad_at);
  if (a.n_steps) {
    /* a full warp that arrives together adds up its counts through lane 0 */
    const pd_u32 group = __activemask();
    if (group == PD_FULL_MASK) {
      const pd_u64 total = pd_warp_sum_u64(n_step);
      if ((threadIdx.x & 31) == 0) atomicAdd(a.n_steps, total);
    } else atomicAdd(a.n_steps, n_step);
  }
}
