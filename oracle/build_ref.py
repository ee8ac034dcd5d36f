# -*- coding: utf-8 -*-
"""
Builds ``oracle/_ref/``: the UNMODIFIED reference compiled to CPython bytecode, so that its own
Python loops can be timed on the GPU box, where ``/root/reference`` does not exist.

    make -C oracle _ref            (or: python -m oracle.build_ref)

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing is copied: ``basepop.py`` is byte-compiled where it
lies (``py_compile``), and of every example script only the model ``class`` definitions are
compiled (the scripts run their main routine and import pylab at import time, so they cannot be
imported as modules -- same AST extraction as oracle/reference.py).  Outputs are ``.pycode`` (CPython bytecode) files
under ``oracle/_ref/`` only; the directory is git-ignored and travels to the box with the built
``.so`` files.  The one third-party name the reference needs, ``past.utils.old_div`` (package
``future``, not installed), is restated in oracle/ref_shim/.
"""

import ast
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REFERENCE_DIR = os.environ.get('POPDYN_REFERENCE_DIR', '/root/reference')
EXAMPLES = ('sir_model.py', 'seir_model.py', 'stochastic_strains_model.py', 'tb_model.py',
            'rmit_tb_model.py', 'mix_model.py')


def build():
    source = os.path.join(REFERENCE_DIR, 'basepop.py')
    if not os.path.isfile(source):
        return False
    os.makedirs(OUT, exist_ok=True)
    py_compile.compile(source, cfile=os.path.join(OUT, 'basepop.pycode'), dfile='basepop.py',
                       doraise=True)
    for script in EXAMPLES:
        path = os.path.join(REFERENCE_DIR, 'examples', script)
        with open(path) as handle:
            tree = ast.parse(handle.read())
        classes = [node for node in tree.body if isinstance(node, ast.ClassDef)]
        code = compile(ast.Module(body=classes, type_ignores=[]), 'examples/' + script, 'exec')
        data = importlib._bootstrap_external._code_to_timestamp_pyc(code)
        with open(os.path.join(OUT, 'classes_' + script[:-3] + '.pycode'), 'wb') as handle:
            handle.write(data)
    with open(os.path.join(OUT, 'BUILT_FROM'), 'w') as handle:
        handle.write('%s (python %s)\n' % (REFERENCE_DIR, sys.version.split()[0]))
    return True


if __name__ == '__main__':
    ok = build()
    print('oracle/_ref built' if ok else 'reference tree not found: nothing built')
