/*
 * pd_diag.cuh -- the diagnostics pass for an ensemble: one thread per (member, stored time).
 *
 * Restates /root/reference/basepop.py:911-960 (calculate_diagnostics) for a batch: the
 * compartments of a stored time are loaded as the doubles of soln_array (:932-933), the traced
 * calculate_vars / calculate_flows / calculate_diagnostic_vars (generated pd_diag) run on them,
 * every var goes to var_soln [T][V][M] (var_array, :947-948) and the net flow of every compartment
 * to flow_soln [T][C][M] (flow_array, :949).  Scale-up vars of the explicit integrator keep the
 * value of the LAST derivative call, as in the reference, which never re-evaluates them here
 * (SURVEY.md 3.4): the host passes that one row.  Elementwise and HBM-bound: 8 C bytes read,
 * 8 (V + C) bytes written per (member, time), all accesses coalesced over members.
 */
#include "pd_model.h"
#include "pd_common.cuh"

struct PdDiagLaunch {
  PdDiagArgs a;
  PdUniform U;
};

extern "C" __global__ void __launch_bounds__(PD_BLOCK)
pd_diag_kernel(const __grid_constant__ PdDiagLaunch L) {
  const PdDiagArgs &a = L.a;
  const pd_i64 M = a.n_member;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + threadIdx.x;
  if (m >= M) return;
  double pm[PD_NPM_DIM];
  pm[0] = 0.0;
#pragma unroll
  for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * a.pm_ld + m];
  for (int it = blockIdx.y; it < a.n_time; it += gridDim.y) {
    double y[PD_C], out[PD_NDIAG_DIM], flow[PD_C];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) y[c] = a.soln[((size_t)it * PD_C + c) * a.soln_ld + m];
    pd_diag(y, L.U, pm, a.times[it], a.su, out, flow);
#pragma unroll
    for (int v = 0; v < PD_NDIAG; ++v) a.var_soln[((size_t)it * PD_NDIAG + v) * a.out_ld + m] = out[v];
    if (a.flow_soln) {
#pragma unroll
      for (int c = 0; c < PD_C; ++c) a.flow_soln[((size_t)it * PD_C + c) * a.out_ld + m] = flow[c];
    }
  }
}
