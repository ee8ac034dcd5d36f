# -*- coding: utf-8 -*-
"""
Runs the UNMODIFIED reference from its compiled form ``oracle/_ref/`` (see oracle/build_ref.py):
the CPU baseline "the reference's own loop on the box's host cores" of bench.py.

TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

Timing protocol: one worker PROCESS per host core; each worker seeds ``random`` and
``numpy.random``, then runs sequential ``StrainsModel(); make_times(0, 1000, 1);
integrate('discrete_time_stochastic')`` for at least the requested number of seconds and times
itself (pool start-up excluded).  Two figures: integrator only (``calculate_diagnostics``
stubbed -- what the GPU path computes) and the whole ``integrate()``.
"""

import importlib.machinery
import importlib.util
import json
import marshal
import math
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')
_modules = {}


def available():
    return os.path.isfile(os.path.join(REF, 'basepop.pycode'))


def basepop():
    """The reference's ``basepop`` module, loaded from bytecode."""
    if 'basepop' not in _modules:
        shim = os.path.join(HERE, 'ref_shim')
        if shim not in sys.path:
            sys.path.insert(0, shim)
        loader = importlib.machinery.SourcelessFileLoader('reference_basepop_compiled',
                                                          os.path.join(REF, 'basepop.pycode'))
        spec = importlib.util.spec_from_loader(loader.name, loader)
        module = importlib.util.module_from_spec(spec)
        loader.exec_module(module)
        _modules['basepop'] = module
    return _modules['basepop']


def example_classes(script):
    """Model classes of ``examples/<script>.py`` bound to the reference's own basepop."""
    import numpy
    module = basepop()
    from past.utils import old_div
    with open(os.path.join(REF, 'classes_' + script + '.pycode'), 'rb') as handle:
        code = marshal.loads(handle.read()[16:])
    namespace = {'basepop': module, 'old_div': old_div, 'math': math, 'numpy': numpy,
                 'range': range, 'map': map, '__name__': 'reference_example'}
    exec(code, namespace)
    return {k: v for k, v in namespace.items() if isinstance(v, type)}


def _worker(seconds, seed):
    import random
    import numpy
    Strains = example_classes('stochastic_strains_model')['StrainsModel']
    random.seed(seed)
    numpy.random.seed(seed)

    def run(stub):
        members, t0 = 0, time.perf_counter()
        while True:
            model = Strains()
            if stub:
                model.calculate_diagnostics = lambda: None
            model.make_times(0, 1000, 1)
            model.integrate('discrete_time_stochastic')
            members += 1
            elapsed = time.perf_counter() - t0
            if elapsed >= (seconds if stub else max(0.5, seconds / 4.0)):
                return members, elapsed
    m1, e1 = run(True)
    m2, e2 = run(False)
    return {'members': m1, 'seconds': e1, 'whole_members': m2, 'whole_seconds': e2}


def _hybrid_worker(seconds, seed):
    """The day-by-day switching loop of examples/mix_model.py:107-136 (a script-level function
    there, restated here around the reference's own SirModel and integrate())."""
    import random
    import numpy
    Sir = example_classes('mix_model')['SirModel']
    random.seed(seed)
    numpy.random.seed(seed)
    members, t0 = 0, time.perf_counter()
    while True:
        model = Sir({'start_population': 100})
        model.calculate_diagnostics = lambda: None
        day, final, cutoff, dt = 1, 50, 3, 0.2
        model.make_times(0, day, dt)
        model.integrate(method='continuous_time_stochastic')
        while day < final:
            start, day = day, min(final, day + 1)
            method = 'explicit' if min(model.compartments.values()) > cutoff \
                else 'continuous_time_stochastic'
            model.make_times(start, day, dt)
            model.integrate(method=method, is_continue=True)
        members += 1
        elapsed = time.perf_counter() - t0
        if elapsed >= seconds:
            return {'members': members, 'seconds': elapsed}


def _run_workers(flag, seconds, n_proc):
    env = dict(os.environ, OMP_NUM_THREADS='1', OPENBLAS_NUM_THREADS='1', MKL_NUM_THREADS='1')
    env['PYTHONPATH'] = os.path.dirname(HERE) + os.pathsep + env.get('PYTHONPATH', '')
    procs = [subprocess.Popen([sys.executable, '-m', 'oracle.ref_runner', flag, str(seconds),
                               str(1000 + k)], stdout=subprocess.PIPE, env=env,
                              cwd=os.path.dirname(HERE), text=True) for k in range(n_proc)]
    results = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError('reference worker failed')
        results.append(json.loads(out.strip().splitlines()[-1]))
    return results


def time_mix_hybrid(seconds, n_proc):
    """Members per second of the reference's mix_model hybrid loop on ``n_proc`` host cores."""
    results = _run_workers('--hybrid-worker', seconds, n_proc)
    rate = sum(r['members'] / r['seconds'] for r in results)
    members = sum(r['members'] for r in results)
    return {'value': rate, 'unit': 'members/s', 'cores': n_proc, 'kind': 'reference',
            'sample': '%d members x 50 one-day segments in %d worker processes (>= %.1f s each); '
                      'unmodified reference integrate() from oracle/_ref driven by the loop of '
                      'examples/mix_model.py:107-136, diagnostics stubbed'
                      % (members, n_proc, seconds)}


def time_strains_dts(seconds, n_proc):
    """Aggregate comp-steps/s of ``n_proc`` worker processes, >= ``seconds`` each."""
    rate = whole = 0.0
    members = 0
    for r in _run_workers('--worker', seconds, n_proc):
        rate += r['members'] * 1000 * 5 / r['seconds']
        whole += r['whole_members'] * 1000 * 5 / r['whole_seconds']
        members += r['members']
    return {'value': rate, 'unit': 'comp-steps/s', 'cores': n_proc, 'kind': 'reference',
            'members': members, 'comp_steps': members * 1000 * 5,
            'whole_integrate_value': whole,
            'sample': '%d members x 1000 steps x 5 compartments in %d worker processes '
                      '(>= %.1f s each, timed inside the workers); unmodified basepop.py:797-852 '
                      'compiled into oracle/_ref; `value` is integrator only, '
                      '`whole_integrate_value` includes calculate_diagnostics'
                      % (members, n_proc, seconds)}


if __name__ == '__main__':
    if len(sys.argv) >= 4 and sys.argv[1] == '--worker':
        print(json.dumps(_worker(float(sys.argv[2]), int(sys.argv[3]))))
    elif len(sys.argv) >= 4 and sys.argv[1] == '--hybrid-worker':
        print(json.dumps(_hybrid_worker(float(sys.argv[2]), int(sys.argv[3]))))
    else:
        print(json.dumps(time_strains_dts(2.0, len(os.sched_getaffinity(0))), indent=1))
