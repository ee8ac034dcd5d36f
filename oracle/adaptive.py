# -*- coding: utf-8 -*-
"""
CPU restatement of the adaptive integrator (popdynamics_b200/csrc/kernels/pd_adaptive.cuh) in plain
Python floats: Dormand-Prince 5(4) with the step controller of scipy.integrate.RK45, steps
shortened to land on every target time.  TEST INFRASTRUCTURE ONLY.

The reference's ``integrate('scipy')`` (/root/reference/basepop.py:882-886) hands its derivative
closure to ``scipy.integrate.odeint`` (LSODA), whose step control cannot be restated; this oracle
therefore plays two roles: the GPU kernel is compared with it operation for operation (same
arithmetic order; ``pow`` differs by an ulp between libm and the device, so agreement is to
about 1e-12, not bitwise), and it is itself pinned against the reference's odeint goldens within
the tolerance DESIGN.md states for the method.
"""

import math

A = [[],
     [1 / 5],
     [3 / 40, 9 / 40],
     [44 / 45, -56 / 15, 32 / 9],
     [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
     [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656]]
B = [35 / 384, 0., 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84]
E = [-71 / 57600, 0., 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40]
C = [0., 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.]


def _rms(v, scale):
    return math.sqrt(sum((a / s) ** 2 for a, s in zip(v, scale)) / len(v))


def dopri5(f, y0, times, rtol=1.49012e-8, atol=1.49012e-8, first_step=0.0, max_steps=10 ** 7):
    """Integrate ``dy/dt = f(y, t)`` (``f`` returns a list) through ``times``; returns
    (rows at every target time, attempted steps)."""
    n = len(y0)
    y = [float(v) for v in y0]
    t = float(times[0])
    rows = [list(y)]
    k1 = [float(v) for v in f(list(y), t)]
    h = first_step
    if not h > 0.0:
        scale = [atol + abs(v) * rtol for v in y]
        d0, d1 = _rms(y, scale), _rms(k1, scale)
        h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        y1 = [a + h0 * b for a, b in zip(y, k1)]
        f1 = f(y1, t + h0)
        d2 = _rms([a - b for a, b in zip(f1, k1)], scale) / h0
        if d1 <= 1e-15 and d2 <= 1e-15:
            h1 = max(1e-6, h0 * 1e-3)
        else:
            h1 = (0.01 / max(d1, d2)) ** 0.2
        h = min(100.0 * h0, h1)
    steps, rejected = 0, False
    for t_end in times[1:]:
        t_end = float(t_end)
        while t < t_end:
            if steps >= max_steps:
                raise RuntimeError('step budget exhausted')
            steps += 1
            hs, last = h, False
            if not (t + hs < t_end):
                hs, last = t_end - t, True
            if not hs > 0.0 or not t + hs > t:
                raise RuntimeError('step size underflow')
            ks = [k1]
            for stage in range(1, 6):
                yn = [y[c] + hs * sum(A[stage][j] * ks[j][c] for j in range(stage)
                                      if A[stage][j] != 0. or True)
                      for c in range(n)]
                ks.append([float(v) for v in f(yn, t + C[stage] * hs)])
            yn = [y[c] + hs * (B[0] * ks[0][c] + B[2] * ks[2][c] + B[3] * ks[3][c] +
                               B[4] * ks[4][c] + B[5] * ks[5][c]) for c in range(n)]
            t_new = t_end if last else t + hs
            k7 = [float(v) for v in f(yn, t_new)]
            s = 0.0
            for c in range(n):
                e = hs * (E[0] * ks[0][c] + E[2] * ks[2][c] + E[3] * ks[3][c] + E[4] * ks[4][c] +
                          E[5] * ks[5][c] + E[6] * k7[c])
                r = e / (atol + max(abs(y[c]), abs(yn[c])) * rtol)
                s += r * r
            err = math.sqrt(s / n)
            if err < 1.0:
                factor = 10.0 if err == 0.0 else min(10.0, 0.9 * err ** -0.2)
                if rejected:
                    factor = min(1.0, factor)
                rejected = False
                h = max(h, hs * factor) if last else hs * factor
                t, y, k1 = t_new, yn, k7
            else:
                factor = max(0.2, 0.9 * err ** -0.2) if err == err else 0.2
                h = hs * factor
                rejected = True
        rows.append(list(y))
    return rows, steps
