"""
Capture what the REFERENCE's own Python layer emits (energy strings, force settings, step
programs) and store it as tests/golden/ref_programs.json.

Runs only in the build container: it imports the unmodified reference sources from
/root/reference/src on top of a stub ``simtk`` package whose ``openmm`` / ``unit`` modules are
this repository's description layer (atomsmm_b200.mm / .unit).  OpenMM itself is not needed
because atomsmm only *describes* forces and integrators.  Two compatibility shims are applied
to the reference at import time (neither changes what it emits): ``np.int`` (removed from
numpy >= 1.24, used at computers.py:33) and sympy's parsing of the name ``Q``
(integrators.py:103).

    python tests/golden/capture_reference.py
"""

import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REFERENCE = '/root/reference'


def install_stub():
    from atomsmm_b200 import app, mm, unit
    simtk = types.ModuleType('simtk')
    simtk.openmm = mm
    simtk.unit = unit
    mm.app = app
    # classes the reference subclasses at import time but which are outside the hot path
    for module, name in ((app, 'StateDataReporter'), (mm, 'CustomCVForce')):
        if not hasattr(module, name):
            setattr(module, name, type(name, (object,), {}))
    sys.modules['simtk'] = simtk
    sys.modules['simtk.openmm'] = mm
    sys.modules['simtk.openmm.app'] = app
    sys.modules['simtk.unit'] = unit
    if not hasattr(np, 'int'):
        np.int = int
    sys.path.insert(0, os.path.join(REFERENCE, 'src'))
    import sympy
    from sympy.parsing import sympy_parser
    original = sympy_parser.parse_expr

    def safe_parse(text, *args, **kwargs):
        import re
        names = set(re.findall(r'[A-Za-z_][A-Za-z_0-9]*', text))
        local = {n: sympy.Symbol(n) for n in names
                 if n not in ('sqrt', 'exp', 'log', 'sin', 'cos', 'erf', 'erfc', 'step', 'select', 'deriv')}
        kwargs.setdefault('local_dict', local)
        return original(text, *args, **kwargs)
    sympy_parser.parse_expr = safe_parse
    import atomsmm
    atomsmm.integrators.parse_expr = safe_parse
    if hasattr(atomsmm.systems, 'parse_expr'):
        atomsmm.systems.parse_expr = safe_parse
    return atomsmm


if __name__ == '__main__':
    reference = install_stub()
    from atomsmm_b200 import app, unit
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cases import build_cases
    data = os.path.join(REFERENCE, 'tests', 'data')

    def loader(case):
        return app.PDBFile(os.path.join(data, case + '.pdb')), app.ForceField(os.path.join(data, case + '.xml'))
    result = build_cases(reference, unit, app, loader)
    target = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ref_programs.json')
    with open(target, 'w') as handle:
        json.dump(result, handle, indent=1, sort_keys=True)
    print('wrote', target, {k: len(v) for k, v in result.items()})
