"""
Parity at the sizes the benchmark numbers are quoted on ("no parity at benchmark
scale"): BASELINE config 2 (98 304-atom water, RESPASystem) and one frame of config 3 (179 200-atom
ionic liquid, DampedSmoothedForce + NonbondedExceptionsForce) against the float64 C restatement
(oracle/cport.py, itself checked against the generic oracle in tests/test_oracle_c.py): forces of every
force group within 1e-5 relative RMS, energies within 1e-6 relative, and the interacting pair sets
of both neighbour lists EQUAL (count and checksum) -- "neighbour lists bit-exact" at a size where the
cell grid is 9+ cells wide and most pairs do not cross a periodic image.  Config 5 (4.2 M atoms) runs the
same gate inside bench.py (`parity` key of the JSON line).
"""

import os

import numpy as np
import pytest

from atomsmm_b200 import mm, unit

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench', os.path.join(ROOT, 'bench.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def _compare(context, system, pos, groups, pair_forces):
    from oracle import cport
    port = cport.CPort(system, threads=os.cpu_count(), verify=True)
    for g in groups:
        state = context.getState(getForces=True, getEnergy=True, groups={g})
        f_ref, e_ref, _ = port.evaluate(pos, groups={g})
        rms = float(np.sqrt(np.sum((state._forces - f_ref)**2)/np.sum(f_ref**2)))
        assert rms < 1e-5, (g, rms)
        assert abs(state._potential - e_ref) <= 1e-6*abs(e_ref), (g, state._potential, e_ref)
    for force in pair_forces:
        assert context.pair_set(force) == port.pair_set(pos, force), force.getForceGroup()
    return port


def test_config2_first_frame_against_c_oracle(cuda_platform):
    bench = _bench()
    system, pos, vel = bench.build_workload(4)
    assert system.getNumParticles() == 98304
    context = mm.Context(system, mm.CustomIntegrator(0.001), cuda_platform)
    context.setPositions(pos)
    pair_forces = [f for f in system.getForces()
                   if isinstance(f, (mm.CustomNonbondedForce, mm.NonbondedForce)) and f.getForceGroup() != 31]
    assert sorted(f.getForceGroup() for f in pair_forces) == [1, 2]
    port = _compare(context, system, pos, [0, 1, 2], pair_forces)
    # the far list of this box: ~204 pairs per atom within 1.0 nm, the near one ~69 within 0.7 nm
    far = next(f for f in pair_forces if f.getForceGroup() == 2)
    near = next(f for f in pair_forces if f.getForceGroup() == 1)
    assert 190 < port.pair_set(pos, far)[0]/98304 < 215
    assert 60 < port.pair_set(pos, near)[0]/98304 < 75


def test_config2_after_dynamics_against_c_oracle(cuda_platform):
    """The same gate on a frame the engine itself produced: 40 RESPA + Nose-Hoover steps move every atom, the
    lists have been rebuilt on the device several times, and the pair sets must still equal the oracle's."""
    bench = _bench()
    system, pos, vel = bench.build_workload(4)
    integrator, dof = bench.make_integrator(system)
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    integrator.step(40)
    assert context.list_stats()['rebuilds'] >= 3
    now = context.getState(getPositions=True)._positions
    pair_forces = [f for f in system.getForces()
                   if isinstance(f, (mm.CustomNonbondedForce, mm.NonbondedForce)) and f.getForceGroup() != 31]
    _compare(context, system, now, [0, 1, 2], pair_forces)


def test_config3_first_frame_against_c_oracle(cuda_platform):
    bench = _bench()
    system, pos, vel, integrator = bench.build_c3(4)
    assert system.getNumParticles() == 179200
    context = mm.Context(system, mm.CustomIntegrator(0.001), cuda_platform)
    context.setPositions(pos)
    pair_forces = [f for f in system.getForces() if isinstance(f, mm.CustomNonbondedForce)]
    assert len(pair_forces) == 1
    _compare(context, system, pos, [0, 1], pair_forces)
