"""
NPT: openmm.MonteCarloBarostat acting through the UpdateContextState hook that every
atomsmm step program starts with (reference: integrators.py:115-122).  The engine's volume moves
(csrc/barostat.cu) are replayed by the float64 oracle interpreter with the same SplitMix64 stream: the
accept / reject sequence and the volumes must agree; a longer run checks the adaptive step size and
that density and energy stay sane.  Also Context.setPeriodicBoxVectors on a live context.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import systems
from systems import positions_of
from test_gpu_integrators import thermal_velocities

pytestmark = pytest.mark.gpu
fs, K = unit.femtoseconds, unit.kelvin
kJ = unit.kilojoules_per_mole


def _npt_system(frequency, seed):
    system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
    barostat = mm.MonteCarloBarostat(1.0*unit.bar, 300*K, frequency)
    barostat.setRandomNumberSeed(seed)
    system.addForce(barostat)
    return system, pdb


def test_volume_moves_match_the_oracle_interpreter(cuda_platform):
    from oracle import interp
    system, pdb = _npt_system(2, 77)
    pos = positions_of(pdb)
    vel = thermal_velocities(system, 300.0, 3)
    factory = lambda: atomsmm.propagators.UnconstrainedVelocityVerletPropagator().integrator(1*fs)
    integrator = factory()
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    reference = interp.Interpreter(system, factory(), pos, vel)
    volumes = []
    for _ in range(8):
        integrator.step(2)
        volumes.append(context.getState().getPeriodicBoxVolume().value_in_unit(unit.nanometer**3))
    reference.step(16)
    log = reference.barostats[0]['log']
    assert len(log) == 8
    assert [v for _, v, _ in log] == pytest.approx(volumes, rel=1e-12)
    stats = context.barostat_statistics()
    assert stats['attempts'] == 8 and stats['accepted'] == sum(1 for a, _, _ in log if a)
    state = context.getState(getPositions=True, getEnergy=True)
    assert np.max(np.abs(state._positions - reference.x)) < 2e-5
    assert state._potential == pytest.approx(reference.potential_energy(), rel=1e-6)
    assert state._box == pytest.approx(reference.box, rel=1e-12)


def test_npt_run_adapts_the_step_and_keeps_the_density(cuda_platform):
    respa, pdb = systems.respa_water()
    barostat = mm.MonteCarloBarostat(1.0*unit.bar, 300*K, 10)
    barostat.setRandomNumberSeed(5)
    respa.addForce(barostat)
    dof = atomsmm.countDegreesOfFreedom(respa)
    nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                 atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(2*fs)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(positions_of(pdb))
    context.setVelocities(thermal_velocities(respa, 300.0, 9))
    v0 = 2.5**3
    integrator.step(2000)                       # 200 attempts
    stats = context.barostat_statistics()
    assert stats['attempts'] == 200
    assert 0.15 < stats['accepted']/stats['attempts'] < 0.85
    assert 0 < stats['volume_scale'] <= 0.3*v0   # adapted only when acceptance leaves the 25-75 % window
    volume = context.getState().getPeriodicBoxVolume().value_in_unit(unit.nanometer**3)
    assert 0.93*v0 < volume < 1.07*v0           # liquid water stays at liquid density
    state = context.getState(getEnergy=True)
    temperature = 2*state.getKineticEnergy().value_in_unit(kJ)/(dof*8.314472471220217e-3)
    assert 280 < temperature < 320
    assert context.counters()['graph_launches'] > 1000      # the step graph is re-captured after accepted moves


def test_set_periodic_box_vectors_on_a_live_context(cuda_platform):
    """Same configuration, 1 % larger box: energies and forces equal those of a fresh context in that box."""
    from oracle import refmath
    system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
    pos = positions_of(pdb)
    context = mm.Context(system, mm.VerletIntegrator(0.0), cuda_platform)
    context.setPositions(pos)
    before = context.getState(getEnergy=True)._potential
    box = np.array([2.5, 2.5, 2.5])*1.01
    context.setPeriodicBoxVectors(mm.Vec3(box[0], 0, 0), mm.Vec3(0, box[1], 0), mm.Vec3(0, 0, box[2]))
    state = context.getState(getEnergy=True, getForces=True)
    ref = refmath.evaluate_system(system, pos, box)
    assert state._potential != pytest.approx(before, rel=1e-6)
    assert state._potential == pytest.approx(ref.energy, rel=1e-6)
    assert np.sqrt(np.sum((state._forces - ref.forces)**2)/np.sum(ref.forces**2)) < 1e-5
    assert state.getPeriodicBoxVolume().value_in_unit(unit.nanometer**3) == pytest.approx(np.prod(box))
