"""The unmodified reference package runs on this repository's description layer through
atomsmm_b200.compat.install() (the `simtk` entry point).  Needs the reference sources, which exist only
in the build container (/root/reference): skipped elsewhere.  No GPU: everything up to the C ABI."""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference/src'

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + '/tests')
from atomsmm_b200 import compat
reference = compat.install(%(ref)r)
from simtk import openmm, unit
from simtk.openmm import app
import systems
assert reference.__file__.startswith(%(ref)r)
system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
respa = reference.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
assert sorted(f.getForceGroup() for f in respa.getForces())[-1] == 31
dof = reference.countDegreesOfFreedom(respa)
nh = reference.propagators.NoseHooverPropagator(300*unit.kelvin, dof, 100*unit.femtoseconds)
integrator = reference.propagators.TrotterSuzukiPropagator(
    reference.propagators.RespaPropagator([4, 2, 1]),
    reference.propagators.SuzukiYoshidaPropagator(nh, 3)).integrator(4*unit.femtoseconds)
# the engine's front end accepts what the reference built: pair-force families and the lowered step program
from atomsmm_b200 import lowering
families = []
for force in respa.getForces():
    if isinstance(force, openmm.CustomNonbondedForce):
        families.append(lowering.classify_pair_force(force, {})[0])
assert families == [lowering.PAIR_NEAR, lowering.PAIR_NEAR]
program = lowering.lower_program(integrator, 0b111 | (1 << 31), {}, True, constrained=False)
kinds = [op[0] for op in program.ops]
assert lowering.OP_KICK in kinds and lowering.OP_EVAL in kinds
# and it is byte-identical to what this repository's own classes emit for the same construction
import atomsmm_b200 as ours
mine = ours.TrotterSuzukiPropagator(ours.RespaPropagator([4, 2, 1]), ours.SuzukiYoshidaPropagator(
    ours.NoseHooverPropagator(300*unit.kelvin, dof, 100*unit.femtoseconds), 3)).integrator(4*unit.femtoseconds)
steps = lambda it: [tuple(it.getComputationStep(k)) for k in range(it.getNumComputations())]
assert steps(integrator) == steps(mine)
print('compat ok', len(kinds))
"""


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='reference sources are only in the build container')
def test_unmodified_reference_package_runs_on_the_description_layer():
    result = subprocess.run([sys.executable, '-c', SCRIPT % dict(root=ROOT, ref=REFERENCE)], capture_output=True,
                            text=True, timeout=600)
    assert result.returncode == 0, result.stderr[-3000:]
    assert 'compat ok' in result.stdout
