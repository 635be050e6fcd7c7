"""
Constrained dynamics on the CUDA engine: SHAKE / RATTLE behind
CustomIntegrator.addConstrainPositions / addConstrainVelocities, exercised through the reference's own
constrained integrators -- VelocityVerletPropagator and RespaPropagator([4,1]) with constrained boost and
move (tests/test_propagators.py:37-49) on the H-bond-constrained ionic liquid, and rigid water.

The reference's goldens for these runs depend on OpenMM's velocity RNG and cannot be reproduced
(parity unpinned); the comparison is the float64 oracle interpreter executing the same
step program with a Gauss-Seidel SHAKE/RATTLE converged to 1e-13.
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


def constraint_errors(system, x, v=None):
    worst_x, worst_v = 0.0, 0.0
    for k in range(system.getNumConstraints()):
        i, j, d = system.getConstraintParameters(k)
        d = d.value_in_md_units()
        r = x[i] - x[j]
        worst_x = max(worst_x, abs(np.sqrt(r.dot(r)) - d)/d)
        if v is not None:
            worst_v = max(worst_v, abs(r.dot(v[i] - v[j]))/(d*d))
    return worst_x, worst_v


def constrained_start(system, pos, seed):
    """Velocities with no component along the constraints (as OpenMM's setVelocitiesToTemperature gives)."""
    from oracle import interp
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(system.getNumParticles())])
    constraints = [(c[0], c[1], c[2]) for c in system._constraints]
    return interp.rattle(constraints, mass, pos, thermal_velocities(system, 300.0, seed))


def run_both(system, pos, vel, factory, steps, platform):
    from oracle import interp
    integrator = factory()
    integrator.setConstraintTolerance(1e-9)
    context = mm.Context(system, integrator, platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    reference = interp.Interpreter(system, factory(), pos, vel)
    integrator.step(steps)
    reference.step(steps)
    state = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
    x = state.getPositions(asNumpy=True).value_in_unit(unit.nanometer)
    v = state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)
    assert np.max(np.abs(x - reference.x)) < 2e-5
    assert np.sqrt(np.sum((v - reference.v)**2)/np.sum(reference.v**2)) < 1e-4
    ex, ev = constraint_errors(system, x, v)
    assert ex < 1e-7 and ev < 1e-6
    return context, state, reference


def ionic_liquid():
    pdb, ff = systems.fixtures.load('emim_BCN4_Jiung2014')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.PME, constraints=app.HBonds, removeCMMotion=False)
    nb = system.getForce(atomsmm.findNonbondedForce(system))
    nb.setReciprocalSpaceForceGroup(1)
    assert system.getNumConstraints() > 0
    return system, positions_of(pdb)


def test_constrained_velocity_verlet_matches_interpreter(cuda_platform):
    """tests/test_propagators.py:37-40"""
    system, pos = ionic_liquid()
    vel = constrained_start(system, pos, 1)
    factory = lambda: atomsmm.GlobalThermostatIntegrator(1*fs, atomsmm.VelocityVerletPropagator())
    run_both(system, pos, vel, factory, 3, cuda_platform)


def test_constrained_respa_matches_interpreter(cuda_platform):
    """tests/test_propagators.py:43-49: RESPA [4,1], reciprocal space in group 1, constrained boost / move"""
    system, pos = ionic_liquid()
    vel = constrained_start(system, pos, 2)

    def factory():
        boost = atomsmm.propagators.VelocityBoostPropagator(constrained=True)
        move = atomsmm.propagators.TranslationPropagator(constrained=True)
        return atomsmm.GlobalThermostatIntegrator(1*fs, atomsmm.RespaPropagator([4, 1], boost=boost, move=move))
    run_both(system, pos, vel, factory, 2, cuda_platform)


def test_rigid_water_nve(cuda_platform):
    """Default createSystem (rigidWater=True): 1 ps of velocity Verlet at 2 fs keeps every O-H and H-H
    distance at the tolerance and conserves energy."""
    pdb, ff = systems.fixtures.load('q-SPC-FW')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, removeCMMotion=False)
    assert system.getNumConstraints() == system.getNumParticles()
    pos = positions_of(pdb)
    # the PDB geometry is flexible water: put it on the constraint manifold first
    from oracle import interp
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(system.getNumParticles())])
    constraints = [(c[0], c[1], c[2]) for c in system._constraints]
    pos = interp.shake(constraints, mass, pos, pos)
    vel = constrained_start(system, pos, 3)
    integrator = atomsmm.GlobalThermostatIntegrator(2*fs, atomsmm.VelocityVerletPropagator())
    integrator.setConstraintTolerance(1e-8)
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    dof = atomsmm.countDegreesOfFreedom(system)
    energies = []
    for _ in range(10):
        integrator.step(50)
        state = context.getState(getEnergy=True, getPositions=True, getVelocities=True)
        energies.append((state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(unit.kilojoules_per_mole))
    x = state.getPositions(asNumpy=True).value_in_unit(unit.nanometer)
    v = state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)
    ex, ev = constraint_errors(system, x, v)
    assert ex < 1e-6 and ev < 1e-5
    drift = np.polyfit(0.1*np.arange(10), energies, 1)[0]/dof       # kJ/mol/ps per degree of freedom
    assert abs(drift) < 0.02
    assert np.std(energies) < 0.05*abs(np.mean(energies))/100
