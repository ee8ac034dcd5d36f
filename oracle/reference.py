# -*- coding: utf-8 -*-
"""
Loader for the UNMODIFIED reference (``/root/reference``), used only to pin the oracle.

TEST INFRASTRUCTURE ONLY.  The reference tree exists in the authoring container but not on the
GPU box, so everything here is optional: ``available()`` says whether it can be used, the golden
fixtures under tests/golden/ (written by oracle/make_golden.py) carry its outputs to the box.

* ``import basepop`` needs ``past.utils.old_div`` from the un-vendored package ``future``; the
  restatement in oracle/ref_shim/past is put on ``sys.path`` first.
* every example script runs its main routine and imports pylab at import time, so model classes
  are taken by extracting the ``ClassDef`` nodes from the script's AST and exec-ing only those.
"""

import ast
import importlib
import math
import os
import random
import sys

import numpy

REFERENCE_DIR = os.environ.get('POPDYN_REFERENCE_DIR', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))
_ref_module = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, 'basepop.py'))


def basepop():
    """The reference's own ``basepop`` module (imported once)."""
    global _ref_module
    if _ref_module is None:
        shim = os.path.join(_HERE, 'ref_shim')
        if shim not in sys.path:
            sys.path.insert(0, shim)
        spec = importlib.util.spec_from_file_location(
            'reference_basepop', os.path.join(REFERENCE_DIR, 'basepop.py'))
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        _ref_module = module
    return _ref_module


def example_classes(script, against=None):
    """Model classes defined in ``examples/<script>``; ``against`` is the module bound to the
    name ``basepop`` inside them (default: the reference's own; pass popdynamics_b200.basepop to
    exercise the drop-in with the reference's unmodified model code)."""
    module = basepop() if against is None else against
    old_div = getattr(module, 'old_div', None)
    if old_div is None:
        from past.utils import old_div
    namespace = {'basepop': module, 'old_div': old_div, 'math': math, 'numpy': numpy,
                 'range': range, 'map': map, '__name__': 'reference_example'}
    with open(os.path.join(REFERENCE_DIR, 'examples', script)) as handle:
        tree = ast.parse(handle.read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            code = compile(ast.Module(body=[node], type_ignores=[]), script, 'exec')
            exec(code, namespace)
    return {k: v for k, v in namespace.items() if isinstance(v, type)}


class injected_uniforms(object):
    """Context manager replacing the reference's three RNG call sites (basepop.py:174, :776,
    :830) by draws from a supplied uniform array: ``random.random`` reads the next value,
    ``numpy.random.poisson`` runs numpy's legacy transform (oracle.Rng.poisson) on the same
    stream.  This is the injected-stream protocol of the parity tests."""

    def __init__(self, uniforms):
        from . import oracle
        self.rng = oracle.Rng('inject', uniforms=numpy.ascontiguousarray(uniforms, dtype=float))

    def __enter__(self):
        ref = basepop()
        self._saved = (ref.random.random, ref.numpy.random.poisson)
        rng = self.rng

        def poisson(lam, size=None):
            return numpy.array([rng.poisson(lam)], dtype=numpy.int64)

        ref.random.random = rng.random
        ref.numpy.random.poisson = poisson
        return self

    def __exit__(self, *exc):
        ref = basepop()
        ref.random.random, ref.numpy.random.poisson = self._saved
        return False


def seed_all(seed):
    random.seed(seed)
    numpy.random.seed(seed)
