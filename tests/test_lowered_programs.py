"""
CPU check of the integrator lowering: the lowered op list + bytecode (what the CUDA engine executes),
run by tests/lowered_executor.py with the engine's op semantics, must reproduce oracle/interp.py
running the ORIGINAL CustomIntegrator program of the reference's classes -- for RESPA with a
Suzuki-Yoshida Nose-Hoover chain (BASELINE config 2's integrator: kick fusion, one-launch thermostat
blocks, chained reductions), Bussi and Langevin baths (scalar programs with loops, random numbers),
constrained velocity Verlet (constraint ops) and AFED (derivative and invalidation ops).
Small clusters cut from the reference's data sets keep the float64 oracle fast.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, lowering, unit

import systems
from lowered_executor import Executor

fs, ps, K = unit.femtoseconds, unit.picoseconds, unit.kelvin
KB = 8.314472471220217e-3


def velocities(system, seed, temperature=300.0):
    n = system.getNumParticles()
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(n)])
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3))*np.sqrt(KB*temperature/mass)[:, None]


def water_cluster(nmol=48, **kwargs):
    pdb, ff = systems.subset('q-SPC-FW', nmol)
    options = dict(nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False, removeCMMotion=False)
    options.update(kwargs)
    return ff.createSystem(pdb.topology, **options), systems.positions_of(pdb)


def compare(system, factory, pos, vel, steps, parameters=None, derivative_slots=None, group_mask=0xffffffff,
            tol=1e-9):
    from oracle import interp
    executor = Executor(system, factory(), pos, vel, parameters=parameters, derivative_slots=derivative_slots,
                        group_mask=group_mask, seed=11)
    executor.program_derivative_slots = derivative_slots
    reference = interp.Interpreter(system, factory(), pos, vel, seed=11, parameters=parameters)
    executor.step(steps)
    reference.step(steps)
    scale_x = max(1e-3, float(np.max(np.abs(reference.x - pos))))
    assert np.max(np.abs(executor.x - reference.x)) <= tol*max(1.0, scale_x/1e-3)
    assert np.max(np.abs(executor.v - reference.v)) <= tol*max(1.0, float(np.max(np.abs(reference.v))))
    names = executor.program.global_names
    for name, value in reference.globals.items():
        if name in names and not name.startswith('_coef') and 'RESPA' not in name:
            assert executor.globals[names.index(name)] == pytest.approx(value, rel=1e-8, abs=1e-10), name
    return executor, reference


def test_respa_nose_hoover_chain_program():
    system, pos = water_cluster()
    respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
    dof = atomsmm.countDegreesOfFreedom(respa)

    def factory():
        nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs)
        return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                               atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    executor, _ = compare(respa, factory, pos, velocities(respa, 1), 3, group_mask=0b111 | (1 << 31))
    kinds = [op[0] for op in executor.program.ops]
    assert lowering.OP_SUM not in kinds and lowering.OP_SCALE not in kinds       # everything fused


def test_bussi_and_langevin_programs():
    system, pos = water_cluster(36)
    respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
    dof = atomsmm.countDegreesOfFreedom(respa)

    def bussi():
        thermostat = atomsmm.VelocityRescalingPropagator(300*K, dof, 0.05*ps)
        integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 1, 1]), thermostat).integrator(2*fs)
        return integrator
    compare(respa, bussi, pos, velocities(respa, 2), 3, group_mask=0b111 | (1 << 31))

    def langevin():
        return atomsmm.Langevin_R_Integrator(2*fs, [2, 1, 1], 300*K, 10/ps)
    compare(respa, langevin, pos, velocities(respa, 3), 2, group_mask=0b111 | (1 << 31))


def test_constrained_velocity_verlet_program():
    from oracle import interp
    system, pos = water_cluster(30, rigidWater=True)
    assert system.getNumConstraints() == 90
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(system.getNumParticles())])
    constraints = [(c[0], c[1], c[2]) for c in system._constraints]
    pos = interp.shake(constraints, mass, pos, pos)
    vel = interp.rattle(constraints, mass, pos, velocities(system, 4))
    factory = lambda: atomsmm.GlobalThermostatIntegrator(2*fs, atomsmm.VelocityVerletPropagator())
    executor, reference = compare(system, factory, pos, vel, 3)
    kinds = [op[0] for op in executor.program.ops]
    assert lowering.OP_CONSTRAIN_X in kinds and lowering.OP_CONSTRAIN_V in kinds


def test_afed_program():
    pdb, ff = systems.subset('methane-in-water', 40)
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False,
                             removeCMMotion=False)
    alchemical = atomsmm.AlchemicalSystem(system, {0})
    pos = systems.positions_of(pdb)

    def factory():
        nvt = atomsmm.TrotterSuzukiPropagator(
            atomsmm.VelocityVerletPropagator(),
            atomsmm.NoseHooverPropagator(300*K, atomsmm.countDegreesOfFreedom(alchemical), 10*fs)).integrator(1*fs)
        variable = atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)
        integrator = atomsmm.AdiabaticDynamicsIntegrator(nvt, 2, [variable])
        integrator.setGlobalVariableByName('_v_lambda_vdw', 1.5)     # 0.9995 + 0.5*dt*v > 1: hits the wall at once
        integrator.setPerDofVariableByName('ndof', np.full((alchemical.getNumParticles(), 3),
                                                           3.0*alchemical.getNumParticles()))
        return integrator
    slots = {'lambda_vdw': lowering.ENERGY_SLOT_DLAMBDA_VDW}
    executor, reference = compare(alchemical, factory, pos, velocities(alchemical, 5), 3,
                                  parameters={'lambda_vdw': 0.9995}, derivative_slots=slots, tol=1e-8)
    assert executor.parameters['lambda_vdw'] == pytest.approx(reference.parameters['lambda_vdw'], abs=1e-9)
    assert 0.0 <= executor.parameters['lambda_vdw'] <= 1.0


def _nh(dof, nloops=1):
    return atomsmm.NoseHooverPropagator(300*K, dof, 100*fs, nloops)


CASES = {
    # propagator algebra + thermostats of SURVEY rows a11-a19, each through the whole lowering pipeline
    'nhc-trotter': lambda dof: atomsmm.TrotterSuzukiPropagator(
        atomsmm.RespaPropagator([2, 2, 1]), atomsmm.NoseHooverChainPropagator(300*K, dof, 100*fs)).integrator(2*fs),
    'nh-nloops4-sy7': lambda dof: atomsmm.TrotterSuzukiPropagator(
        atomsmm.RespaPropagator([2, 1, 1]), atomsmm.SuzukiYoshidaPropagator(_nh(dof, 4), 7)).integrator(2*fs),
    'nhl-stochastic': lambda dof: atomsmm.TrotterSuzukiPropagator(
        atomsmm.RespaPropagator([2, 1, 1]),
        atomsmm.NoseHooverLangevinPropagator(300*K, dof, 100*fs, 10/ps)).integrator(2*fs),
    'chained-split': lambda dof: atomsmm.ChainedPropagator(
        [atomsmm.RespaPropagator([2, 1, 1]), atomsmm.SplitPropagator(_nh(dof), 2)]).integrator(2*fs),
    'mts-xo-respa': lambda dof: atomsmm.MultipleTimeScaleIntegrator(2*fs, [2, 1, 1], bath=_nh(dof), scheme='xo-respa'),
    'mts-xi-respa': lambda dof: atomsmm.MultipleTimeScaleIntegrator(2*fs, [2, 1, 1], bath=_nh(dof), scheme='xi-respa'),
    'mts-side-nsy3': lambda dof: atomsmm.MultipleTimeScaleIntegrator(2*fs, [2, 1, 1], bath=_nh(dof), scheme='side',
                                                                   location=1, nsy=3),
    'mts-middle-nres2': lambda dof: atomsmm.MultipleTimeScaleIntegrator(2*fs, [2, 1, 1], bath=_nh(dof), nres=2),
    'respa-memory': lambda dof: atomsmm.RespaPropagator([2, 2, 1], has_memory=True).integrator(2*fs),
    'nhl-r-massive': lambda dof: atomsmm.NHL_R_Integrator(2*fs, [2, 1, 1], 300*K, 50*fs, 10/ps),
}


@pytest.mark.parametrize('case', sorted(CASES))
def test_reference_integrators_lower_faithfully(case):
    system, pos = water_cluster(24)
    respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
    dof = atomsmm.countDegreesOfFreedom(respa)
    compare(respa, lambda: CASES[case](dof), pos, velocities(respa, 6), 2, group_mask=0b111 | (1 << 31))
