/*
 * pd_tau_leap.cuh -- discrete-time stochastic integrator (clamped Poisson tau-leap), one thread
 * per ensemble member.
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
 * Shape of one step (this is where the kernel differs from the reference's nested loops, without
 * changing which uniform goes to which event):
 *   1. RATE PASS, all lanes in lockstep: every event's mean and its code -- exp(-mean) for the
 *      multiplication method, -mean for PTRS -- go to shared memory [event][thread]; a bit mask
 *      remembers which events draw at all.
 *   2. DRAW LOOP, flattened: the lane walks its OWN live events while the warp walks the uniform
 *      stream together.  One trip = one Philox block = two uniforms for every lane; each uniform
 *      either continues the lane's current product or finishes its event (clamp, move, advance
 *      to the lane's next live event).  The trip count of a warp is the largest TOTAL number of
 *      uniforms any lane needs in the step, instead of the sum over events of the per-event
 *      maxima that the nested event-major loop pays.
 *   3. the live counts sit in shared memory [compartment][thread] (conflict-free whatever event a
 *      lane is at), so an event's source / destination may be a run-time index.
 */
#include "pd_model.h"
#include "pd_common.cuh"

struct PdTauLaunch {
  PdStochArgs a;
  PdUniform U;
};

#if PD_TAU_NROW <= 32
typedef pd_u32 pd_mask;
#define PD_MASK_FIRST(x) (__ffs((int)(x)) - 1)
#elif PD_TAU_NROW <= 64
typedef pd_u64 pd_mask;
#define PD_MASK_FIRST(x) (__ffsll((long long)(x)) - 1)
#else
#error "the tau-leap kernel handles at most 64 flow-table rows"
#endif

/* event table word: source | destination << 16, 0xffff = none (entry has no source, death no
 * destination); generated per model as PD_TAU_EVENT_TABLE over the rows that can fire */
__device__ const pd_u32 pd_event_table[PD_TAU_NROW_DIM] = PD_TAU_EVENT_TABLE;

/* ring slots: the live counts of target index `it` sit in slot it % PD_NSLOT */
#if PD_OUT_MODE == PD_OUT_STATS
#define PD_NSLOT PD_EPOCH
#else
#define PD_NSLOT 1
#endif
/* the kernel's dynamic shared memory, read by the host after loading the module
 * (csrc/popdyn_capi.cu) */
extern "C" __device__ unsigned int pd_smem_bytes =
    (unsigned int)(PD_CODE_SMEM_BYTES(PD_TAU_NROW, PD_BLOCK) +
                   PD_RING_SMEM_BYTES(PD_NSLOT, PD_C, PD_BLOCK, sizeof(pd_count)) +
                   PD_TAB_SMEM_BYTES(PD_TAU_NROW));

extern "C" __global__ void __launch_bounds__(PD_BLOCK, PD_MIN_BLOCKS)
pd_tau_leap_kernel(const __grid_constant__ PdTauLaunch L) {
  const PdStochArgs &a = L.a;
  const PdUniform &U = L.U;
  const int tid = threadIdx.x;
  const pd_i64 m = (pd_i64)blockIdx.x * PD_BLOCK + tid;
  const bool active = m < a.n_member;
  const pd_i64 M = a.n_member;

  extern __shared__ double s_dyn[];
  double *s_code = s_dyn;                                          /* [PD_TAU_NROW][PD_BLOCK]   */
  pd_count *s_ring = (pd_count *)((char *)s_dyn + PD_CODE_SMEM_BYTES(PD_TAU_NROW, PD_BLOCK));
  pd_u32 *s_tab = (pd_u32 *)((char *)s_ring +
                             PD_RING_SMEM_BYTES(PD_NSLOT, PD_C, PD_BLOCK, sizeof(pd_count)));
#if PD_OUT_MODE == PD_OUT_STATS
  const pd_i64 block_member0 = (pd_i64)blockIdx.x * PD_BLOCK;
  const int n_active = (M - block_member0 < PD_BLOCK) ? (int)(M - block_member0) : PD_BLOCK;
#endif
  for (int e = tid; e < PD_TAU_NROW_DIM; e += PD_BLOCK) s_tab[e] = pd_event_table[e];
  __syncthreads();

  double pm[PD_NPM_DIM];
  PdRng rng;
  pm[0] = 0.0;
  pd_count *s_y = s_ring;                                   /* slot of target index 0 */
  if (active) {
#pragma unroll
    for (int j = 0; j < PD_NPM; ++j) pm[j] = a.pm[(size_t)j * M + m];
#pragma unroll
    for (int c = 0; c < PD_C; ++c) {
      const double v = pd_load_y0(a.y0, a.y0_stride_c, a.y0_stride_m, c, m);
      s_y[c * PD_BLOCK + tid] = (pd_count)v;                /* int() truncation, :805-806 */
#if PD_OUT_MODE == PD_OUT_FULL
      a.soln[(size_t)c * M + m] = v;                        /* row 0 = untruncated y, :813 */
#endif
    }
    rng.init(a, m);
  }
#if PD_OUT_MODE == PD_OUT_STATS
  if (a.n_time == 1) pd_stats_flush(a, s_ring, 0, 1, n_active, a.member0 + block_member0);
#endif

  int bad = 0;
  double time = a.target[0];
  for (int it = 1; it < a.n_time; ++it) {
#if PD_NSLOT > 1
    const int slot = it % PD_NSLOT;
    const pd_count *s_prev = s_ring + ((slot ? slot : PD_NSLOT) - 1) * (PD_C * PD_BLOCK);
    s_y = s_ring + slot * (PD_C * PD_BLOCK);
#else
    const int slot = 0;
    const pd_count *s_prev = s_ring;
#endif
    if (active) {
      const double dt = a.target[it] - a.target[it - 1];    /* :820 */

      /* ---- 1. rate pass (:822-824): start-of-step counts as doubles; every rate of this step
       * is formed from them, the live counts only enter through the clamps (:833-841) */
      double yd[PD_C], mult[PD_NMULT_DIM];
#pragma unroll
      for (int c = 0; c < PD_C; ++c) {
        const pd_count have = s_prev[c * PD_BLOCK + tid];
#if PD_NSLOT > 1
        s_y[c * PD_BLOCK + tid] = have;                     /* this step's slot starts as a copy */
#endif
        yd[c] = (double)have;
      }
      pd_event_mults(yd, U, pm, time, (const double *)0, mult);
      pd_mask live = 0;
#define PD_TAU_CODE(e, mean)                                                         \
  if (mean > 0.0) {                                                                  \
    double code;                                                                     \
    if (mean >= 10.0) code = -mean; else code = pd_exp_nonpos(-mean);                \
    s_code[e * PD_BLOCK + tid] = code;                                               \
    live |= (pd_mask)1 << e;                                                         \
  } else if (!(mean == 0.0)) { if (!bad) bad = PD_ERR_POISSON_MEAN; }
#define PD_TAU_ENTRY(e, to, is_int, rate)                                            \
  { const double mean = rate * dt; PD_TAU_CODE(e, mean) }
#define PD_TAU_TRANSFER(e, from, to, rate)                                           \
  { const double r_ = rate; if (r_ > 0) { const double mean = r_ * dt; PD_TAU_CODE(e, mean) } }
#define PD_TAU_DEATH(e, from, rate)                                                  \
  { const double r_ = rate; if (r_ > 0) { const double mean = r_ * dt; PD_TAU_CODE(e, mean) } }
      PD_FOR_EACH_TAU_ROW(PD_TAU_ENTRY, PD_TAU_TRANSFER, PD_TAU_DEATH)
#undef PD_TAU_ENTRY
#undef PD_TAU_TRANSFER
#undef PD_TAU_DEATH
#undef PD_TAU_CODE

      /* ---- 2. draw loop.  Lane state: remaining live events (lowest bit = current event e),
       * the current event's code, running product and count */
      int e = live ? PD_MASK_FIRST(live) : 0;
      double code = s_code[e * PD_BLOCK + tid];
      double prod = 1.0;
      int X = 0;

/* step to the lane's next live event */
#define PD_TAU_ADVANCE                                                               \
  live &= live - 1;                                                                  \
  if (live) {                                                                        \
    e = PD_MASK_FIRST(live);                                                         \
    code = s_code[e * PD_BLOCK + tid];                                               \
    prod = 1.0; X = 0;                                                               \
  }
/* finish the current event with the draw X of the multiplication method: clamp to the CURRENT
 * source count (:833-834, :839-840), move, births unclamped (:842-844).  With 32-bit counts a
 * gain that would pass 2^31 - 1 raises the error flag instead of wrapping (numpy / Python ints
 * never wrap): both terms are non-negative, so the unsigned sum cannot wrap itself */
#define PD_TAU_FINISH_SMALL                                                          \
  {                                                                                  \
    const pd_u32 ft = s_tab[e];                                                      \
    const int from = (int)(ft & 0xffffu), to = (int)(ft >> 16);                      \
    pd_count d = (pd_count)X;                                                        \
    if (from != 0xffff) {                                                            \
      const pd_count have = s_y[from * PD_BLOCK + tid];                              \
      d = d < have ? d : have;                                                       \
      s_y[from * PD_BLOCK + tid] = have - d;                                         \
    }                                                                                \
    if (to != 0xffff) {                                                              \
      const pd_count have = s_y[to * PD_BLOCK + tid];                                \
      if (sizeof(pd_count) == 4) {                                                   \
        const pd_u32 sum = (pd_u32)have + (pd_u32)d;                                 \
        if (sum > 0x7fffffffu) { if (!bad) bad = PD_ERR_COUNT_RANGE; }               \
        else s_y[to * PD_BLOCK + tid] = (pd_count)sum;                               \
      } else s_y[to * PD_BLOCK + tid] = have + d;                                    \
    }                                                                                \
    PD_TAU_ADVANCE                                                                   \
  }
/* the same after PTRS, whose draw may exceed 32 bits */
#define PD_TAU_FINISH_BIG(k_in)                                                      \
  {                                                                                  \
    pd_i64 d = (k_in);                                                               \
    const pd_u32 ft = s_tab[e];                                                      \
    const int from = (int)(ft & 0xffffu), to = (int)(ft >> 16);                      \
    if (from != 0xffff) {                                                            \
      const pd_count have = s_y[from * PD_BLOCK + tid];                              \
      if (d > (pd_i64)have) d = (pd_i64)have;                                        \
      s_y[from * PD_BLOCK + tid] = have - (pd_count)d;                               \
    }                                                                                \
    if (to != 0xffff) {                                                              \
      const pd_count have = s_y[to * PD_BLOCK + tid];                                \
      if (sizeof(pd_count) == 4 && (pd_i64)have + d > 0x7fffffffLL) {                \
        if (!bad) bad = PD_ERR_COUNT_RANGE;                                          \
      } else s_y[to * PD_BLOCK + tid] = have + (pd_count)d;                          \
    }                                                                                \
    PD_TAU_ADVANCE                                                                   \
  }
/* one uniform of the multiplication method */
#define PD_TAU_MULT(u)                                                               \
  {                                                                                  \
    prod *= (u);                                                                     \
    if (prod > code) X += 1; else PD_TAU_FINISH_SMALL                                \
  }

      while (live) {
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
          if (k >= 0) PD_TAU_FINISH_BIG(k)
          continue;
        }
#if PD_RNG_MODE == PD_RNG_PHILOX
        if (!rng.has_spare) {               /* every lane without a pending spare: one block */
          double u0;
          rng.block(a, u0, rng.spare);
          rng.has_spare = 1;
          PD_TAU_MULT(u0)
        }
        if (live && code >= 0.0) {          /* the block's second uniform, or the pending spare */
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
#undef PD_TAU_FINISH_SMALL
#undef PD_TAU_FINISH_BIG
#undef PD_TAU_ADVANCE
      time += dt;                                           /* :848 */
#if PD_OUT_MODE == PD_OUT_FULL
#pragma unroll
      for (int c = 0; c < PD_C; ++c)
        a.soln[((size_t)it * PD_C + c) * M + m] = (double)s_y[c * PD_BLOCK + tid];   /* :850-852 */
#endif
    }
#if PD_OUT_MODE == PD_OUT_STATS
    /* end of an epoch: the block reduces the ring (two barriers inside) */
    if (slot == PD_NSLOT - 1 || it == a.n_time - 1)   /* block-uniform */
      pd_stats_flush(a, s_ring, it - slot, slot + 1, n_active, a.member0 + block_member0);
#endif
  }

  if (active) {
#if PD_OUT_MODE == PD_OUT_FINAL
#pragma unroll
    for (int c = 0; c < PD_C; ++c) a.final_state[(size_t)c * M + m] = (double)s_y[c * PD_BLOCK + tid];
#endif
    if (!bad && rng.exhausted) bad = PD_ERR_STREAM;
    if (bad) pd_raise(a.err, bad, a.member0 + m, 0);
  }
  if (tid == 0 && a.n_steps) {
    const pd_i64 first = (pd_i64)blockIdx.x * PD_BLOCK;
    const pd_i64 here = (M - first < PD_BLOCK) ? (M - first) : PD_BLOCK;
    atomicAdd(a.n_steps, (pd_u64)here * (pd_u64)(a.n_time - 1));
  }
}
