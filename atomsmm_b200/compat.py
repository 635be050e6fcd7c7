"""
Drop-in entry point for code written against OpenMM's Python API -- in particular the UNMODIFIED
reference package (atoms-ufrj/atomsmm), which does ``from simtk import openmm, unit`` and
``from simtk.openmm import app``.

    import atomsmm_b200.compat as compat
    compat.install()                     # registers simtk / simtk.openmm / simtk.openmm.app / simtk.unit
    import atomsmm                       # the reference package, from its own sources
    ...                                  # Platform.getPlatformByName('CUDA') is the B200 engine

``install`` puts this repository's description layer (atomsmm_b200.mm / .app / .unit) under the module
names OpenMM 7.x used, so every System, Force and CustomIntegrator the caller builds is described to
-- and every Context executed by -- the B200 engine through the C ABI (include/atomsmm_b200.h).
Nothing is emulated on the CPU.  ``install(reference_source=...)`` additionally puts a checkout of the
reference on sys.path and applies two import-time compatibility fixes that do not change what the
reference emits: ``np.int`` (removed from numpy >= 1.24; reference computers.py:33) and sympy's
parsing of single-letter names such as ``Q`` (reference integrators.py:103).
"""

import re
import sys
import types


def install(reference_source=None, also_openmm=False):
    """Register the description layer as ``simtk``; returns the reference package if
    ``reference_source`` (the directory holding ``atomsmm/``) is given."""
    from . import app, mm, unit
    simtk = types.ModuleType('simtk')
    simtk.__doc__ = 'simtk namespace provided by atomsmm_b200.compat (B200 engine)'
    simtk.openmm = mm
    simtk.unit = unit
    mm.app = app
    # classes the reference subclasses at import time but which lie outside the hot path (out of scope)
    for module, name in ((app, 'StateDataReporter'), (mm, 'CustomCVForce')):
        if not hasattr(module, name):
            setattr(module, name, type(name, (object,), {}))
    sys.modules['simtk'] = simtk
    sys.modules['simtk.openmm'] = mm
    sys.modules['simtk.openmm.app'] = app
    sys.modules['simtk.unit'] = unit
    if also_openmm:
        sys.modules.setdefault('openmm', mm)
        sys.modules.setdefault('openmm.app', app)
        sys.modules.setdefault('openmm.unit', unit)
    if reference_source is None:
        return None
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int
    if reference_source not in sys.path:
        sys.path.insert(0, reference_source)
    import sympy
    from sympy.parsing import sympy_parser
    original = getattr(sympy_parser.parse_expr, '_b200_original', sympy_parser.parse_expr)

    def safe_parse(text, *args, **kwargs):
        names = set(re.findall(r'[A-Za-z_][A-Za-z_0-9]*', text))
        local = {n: sympy.Symbol(n) for n in names
                 if n not in ('sqrt', 'exp', 'log', 'sin', 'cos', 'erf', 'erfc', 'step', 'select', 'deriv')}
        kwargs.setdefault('local_dict', local)
        return original(text, *args, **kwargs)
    safe_parse._b200_original = original
    sympy_parser.parse_expr = safe_parse
    import atomsmm
    atomsmm.integrators.parse_expr = safe_parse
    if hasattr(atomsmm.systems, 'parse_expr'):
        atomsmm.systems.parse_expr = safe_parse
    return atomsmm
