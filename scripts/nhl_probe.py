"""Massive NHL (NHL_R_Integrator) on the engine vs the oracle interpreter, step by step."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
import atomsmm_b200 as atomsmm  # noqa: E402
from atomsmm_b200 import mm, unit  # noqa: E402
import systems  # noqa: E402
from test_gpu_integrators import thermal_velocities  # noqa: E402
from oracle import interp  # noqa: E402
fs, ps, K = unit.femtoseconds, unit.picoseconds, unit.kelvin
respa, pdb = systems.respa_water()
pos = systems.positions_of(pdb)
vel = thermal_velocities(respa, 300.0, 1234)
for fast in ('true', 'false'):
    factory = lambda: atomsmm.NHL_R_Integrator(2*fs, [2, 1, 1], 300*K, 50*fs, 1e-8/ps)
    integrator = factory()
    context = mm.Context(respa, integrator, mm.Platform.getPlatformByName('B200'), {'FastPaths': fast})
    context.setPositions(pos)
    context.setVelocities(vel)
    reference = interp.Interpreter(respa, factory(), pos, vel)
    print('FastPaths', fast, 'ops', [op[0] for op in context._program.ops][:40], 'perdof', context._program.perdof_names)
    for step in range(3):
        integrator.step(1)
        reference.step(1)
        s = context.getState(getPositions=True, getVelocities=True)
        line = 'step %d: dx %.3e dv %.3e' % (step + 1, np.max(np.abs(s._positions - reference.x)), np.max(np.abs(s._velocities - reference.v)))
        for name in context._program.perdof_names:
            ours = np.array(integrator.getPerDofVariableByName(name))
            theirs = reference.perdof.get(name)
            if theirs is not None:
                line += ' %s %.3e (scale %.3e)' % (name, np.max(np.abs(ours - theirs)), np.max(np.abs(theirs)))
        print(line)
