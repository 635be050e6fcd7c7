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
    """The product's own entry point for code written against simtk.openmm (atomsmm_b200/compat.py)."""
    from atomsmm_b200 import compat
    return compat.install(os.path.join(REFERENCE, 'src'))


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
