"""
AFED / extended-variable dynamics (BASELINE config 4) on the CUDA engine:
``AlchemicalSystem`` (soft-core solute-solvent coupling through an interaction group, systems.py:318-410)
driven by ``AdiabaticDynamicsIntegrator`` with ``lambda_vdw`` as an ``ExtendedSystemVariable``
(integrators.py:642-860).  The construction is the one of the reference's (disabled) test_afed.py:21-45,
without constraints.  The engine keeps lambda on the device: the pair kernels read it from the integrator's
globals, `deriv(energy, lambda_vdw)` is evaluated by the fp64 pair-energy kernel, and the whole step is one
CUDA graph.  Parity: the float64 oracle interpreter executing the same step program (the reference has no
runnable AFED golden: parity unpinned).
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import systems
from test_gpu_integrators import thermal_velocities

pytestmark = pytest.mark.gpu
fs, K = unit.femtoseconds, unit.kelvin


def build():
    pdb, ff = systems.fixtures.load('methane-in-water')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False,
                             removeCMMotion=False)
    residues = [atom.residue.name for atom in pdb.topology.atoms()]
    solute = set(i for i, name in enumerate(residues) if name == 'C1')
    assert len(solute) == 1
    return atomsmm.AlchemicalSystem(system, solute), pdb


def afed_integrator(system, seed=1234):
    nvt = atomsmm.TrotterSuzukiPropagator(
        atomsmm.VelocityVerletPropagator(),
        atomsmm.NoseHooverPropagator(300*K, atomsmm.countDegreesOfFreedom(system), 10*fs)).integrator(1*fs)
    variable = atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)
    integrator = atomsmm.AdiabaticDynamicsIntegrator(nvt, 2, [variable])
    integrator.setRandomNumberSeed(seed)
    return integrator


def test_afed_matches_interpreter(cuda_platform):
    from oracle import interp
    system, pdb = build()
    pos = systems.positions_of(pdb)
    vel = thermal_velocities(system, 300.0, 7)
    integrator = afed_integrator(system)
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    context.setParameter('lambda_vdw', 0.998)
    integrator.step(0)                                   # runs initialize(): random v_lambda, v_eta
    integrator.setGlobalVariableByName('_v_lambda_vdw', 0.9)    # heads for the upper wall: reflection within 2 steps
    reference = interp.Interpreter(system, afed_integrator(system), pos, vel, parameters={'lambda_vdw': 0.998})
    for k in range(integrator.getNumGlobalVariables()):
        reference.globals[integrator.getGlobalVariableName(k)] = integrator.getGlobalVariable(k)
    # the interpreter needs the per-DOF variable the base class fills in at the first step
    reference.perdof['ndof'] = np.full((system.getNumParticles(), 3), 3.0*system.getNumParticles())
    steps = 3
    integrator.step(steps)
    reference.step(steps)
    state = context.getState(getPositions=True, getVelocities=True, getParameters=True)
    x = state.getPositions(asNumpy=True).value_in_unit(unit.nanometer)
    v = state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)
    assert np.max(np.abs(x - reference.x)) < 2e-5
    assert np.sqrt(np.sum((v - reference.v)**2)/np.sum(reference.v**2)) < 1e-4
    lam = state.getParameters()['lambda_vdw']
    assert lam == pytest.approx(reference.parameters['lambda_vdw'], abs=1e-7)
    assert 0.0 <= lam <= 1.0
    assert integrator.getGlobalVariableByName('_v_lambda_vdw') == pytest.approx(reference.globals['_v_lambda_vdw'], rel=1e-5)
    assert reference.globals['_v_lambda_vdw'] < 0          # it did bounce off the wall
    assert context.getParameter('lambda_vdw') == lam


def test_afed_long_run_stays_inside_the_walls(cuda_platform):
    system, pdb = build()
    integrator = afed_integrator(system, seed=3)
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(pdb.positions)
    context.setVelocitiesToTemperature(300*K, 11)
    seen = []
    for _ in range(10):
        integrator.step(50)
        seen.append(context.getParameter('lambda_vdw'))
    assert all(0.0 <= value <= 1.0 for value in seen)
    assert max(seen) - min(seen) > 1e-3                      # lambda does move
    state = context.getState(getEnergy=True, getParameterDerivatives=True)
    total = state.getPotentialEnergy() + state.getKineticEnergy()
    assert np.isfinite(total.value_in_unit(unit.kilojoules_per_mole))
    assert 'lambda_vdw' in state.getEnergyParameterDerivatives()
    assert context.counters()['graph_launches'] > 400        # the step program runs as a CUDA graph
