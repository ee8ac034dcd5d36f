# -*- coding: utf-8 -*-
"""
The BASELINE.json workloads as synthetic inputs, shared by ``bench.py``, the
GPU tests and the tools.  Every builder takes a GLOBAL member range ``[start, stop)`` so that a
rank of a multi-GPU run creates its own shard only; swept values are a pure function of the global
member id (numpy Philox streams keyed by ``(seed, column, block of 2**20 members)``), so the
ensemble does not depend on how it is sharded.

A few members at the start of each sweep are pinned to parameter sets the unmodified reference
was run on (``oracle/make_golden.py`` -> ``tests/golden/reference_vectors.npz``): they are the
parity spot-checks embedded in the full-size runs.
"""

import numpy as np

from examples import models

BLOCK = 1 << 20


def sweep_uniform(seed, column, low, high, start, stop):
    """U[low, high) for the global member ids [start, stop): member i takes draw ``i % 2**20`` of
    the numpy Philox stream keyed by (seed, column, i // 2**20)."""
    out = np.empty(max(stop - start, 0), dtype=np.float64)
    pos = start
    while pos < stop:
        block = pos // BLOCK
        first = block * BLOCK
        last = min(stop, first + BLOCK)
        rng = np.random.Generator(np.random.Philox(key=[seed, (column << 32) | block]))
        draws = rng.uniform(low, high, last - first)       # draws 0 .. last - first - 1
        out[pos - start:last - start] = draws[pos - first:]
        pos = last
    return out


def _pin(values, start, pinned):
    """Overwrite the global members 0 .. len(pinned) - 1 that fall into this shard."""
    for i, v in enumerate(pinned):
        if start <= i < start + len(values):
            values[i - start] = v
    return values


# ---- config 2: SEIR with demography, parameter sweep, explicit, 50 years ----------------------
# /root/reference/examples/seir_model.py:147-228, run :458-464.  Member 0 is the measles base case.
SEIR_TIMES = (0, 365 * 50, 1)
SEIR_GOLDEN_KEY = 'seird_explicit_18250_final'


def seir_sweep(start, stop, seed=0):
    d = dict(models.SEIR_PARAMS['measles'])
    base = models.SEIR_PARAMS['measles']
    d['r0'] = _pin(sweep_uniform(seed, 0, 2., 18., start, stop), start, [base['r0']])
    d['duration_infectious'] = _pin(sweep_uniform(seed, 1, 2., 10., start, stop), start,
                                    [base['duration_infectious']])
    d['duration_preinfectious'] = _pin(sweep_uniform(seed, 2, 2., 10., start, stop), start,
                                       [base['duration_preinfectious']])
    model = models.SeirDemographyModel(d)
    model.make_times(*SEIR_TIMES)
    return model


# ---- config 5: RMIT TB calibration sweep, explicit, final-state proportion -------------------
# /root/reference/examples/rmit_tb_model.py:23-93, sweep loop :133-143.  Members 0..5 are six beta
# points of the reference's figure 2 (sigma = 0, rho = 0.1).  sigma is swept over [0, .25], the
# range of the reference's own figures (:150-178): with sigma up to .5 and beta up to 500 explicit
# Euler at dt = .05 diverges for ~7 % of the members and checks() raises, in the reference as here.
RMIT_TIMES = (0, 500., 1.)
RMIT_GOLDEN_BETAS = [1., 7., 20., 99., 250., 500.]


def rmit_sweep(start, stop, seed=0):
    p = dict(models.RMIT_FIXED_PARAMS)
    n_pin = len(RMIT_GOLDEN_BETAS)
    p['beta'] = _pin(sweep_uniform(seed, 10, 1., 500., start, stop), start, RMIT_GOLDEN_BETAS)
    p['rho'] = _pin(sweep_uniform(seed, 11, 0., 10., start, stop), start, [0.1] * n_pin)
    for k in (1, 2, 3):
        p['sigma%d' % k] = _pin(sweep_uniform(seed, 11 + k, 0., .25, start, stop), start,
                                [0.] * n_pin)
    model = models.RmitTbModel(p)
    model.make_times(*RMIT_TIMES)
    return model


# ---- config 4: TB model, Gillespie, 50 simulated years ------------------------------------------
# /root/reference/examples/tb_model.py:41-156 with its scale-ups evaluated inside the stochastic
# loop (the declared extension ScaleupTbModel).  No sweep: members differ by stream.
TB_TIMES = (1950, 2000, 1.)


def tb_gillespie(start=0, stop=0):
    model = models.ScaleupTbModel([])
    model.make_times(*TB_TIMES)
    return model


# ---- config 5b: hybrid stochastic / deterministic switching of mix_model.py --------------------
# /root/reference/examples/mix_model.py:107-136: one-day segments of step 0.2 up to day 50, first
# one Gillespie, then explicit whenever the smallest compartment exceeds 3.
MIX_CUTOFF = 3
MIX_SEGMENTS = [(day, day + 1, 0.2) for day in range(50)]


def mix_model(start=0, stop=0):
    return models.MixSirModel({'start_population': 100})


# ---- config 3 (headline): two-strain SIR, tau-leap ------------------------------------------------
STRAINS_TIMES = (0, 1000, 1)


def strains(start=0, stop=0):
    model = models.StrainsModel()
    model.make_times(*STRAINS_TIMES)
    return model
