"""
GPU parity tests for the integrator path: the lowered step program running on the CUDA engine
(fused kernels + CUDA graph) against the float64 oracle interpreter executing the SAME step
program on the same inputs.  The reference itself publishes no reproducible trajectory goldens
(tests/test_propagators.py depends on OpenMM's RNG stream and constraint solver), so
deterministic schemes are compared step for step and stochastic ones statistically.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import systems
from systems import positions_of

pytestmark = pytest.mark.gpu

fs, ps, K = unit.femtoseconds, unit.picoseconds, unit.kelvin
KB = 8.314472471220217e-3


def thermal_velocities(system, temperature, seed):
    n = system.getNumParticles()
    mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(n)])
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))*np.sqrt(KB*temperature/mass)[:, None]
    v -= (mass[:, None]*v).sum(0)/mass.sum()
    return v


def run_both(system, pdb, integrator_factory, steps, platform, properties=None):
    from oracle import interp
    pos = positions_of(pdb)
    vel = thermal_velocities(system, 300.0, 1234)
    integrator = integrator_factory()
    context = mm.Context(system, integrator, platform, properties or {})
    context.setPositions(pos)
    context.setVelocities(vel)
    reference = interp.Interpreter(system, integrator_factory(), pos, vel)
    integrator.step(steps)
    reference.step(steps)
    state = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
    return context, integrator, state, reference


def compare(state, reference, x_tol=2e-5, v_rel=1e-4):
    x = state.getPositions(asNumpy=True).value_in_unit(unit.nanometer)
    v = state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)
    assert np.max(np.abs(x - reference.x)) < x_tol
    assert np.sqrt(np.sum((v - reference.v)**2)/np.sum(reference.v**2)) < v_rel


def test_respa_nve_matches_interpreter(cuda_platform):
    """RespaPropagator([4,2,1]) at 4 fs on the RESPASystem water box (BASELINE config 1)."""
    respa, pdb = systems.respa_water()
    factory = lambda: atomsmm.RespaPropagator([4, 2, 1]).integrator(4*fs)
    context, integrator, state, reference = run_both(respa, pdb, factory, 5, cuda_platform)
    compare(state, reference)
    total = (state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(unit.kilojoules_per_mole)
    expected = reference.potential_energy() + reference.kinetic_energy()
    assert total == pytest.approx(expected, rel=1e-6)
    counters = context.counters()
    assert counters['graph_launches'] >= 3          # steps 3.. replay the captured CUDA graph


def test_generic_vm_equals_fast_paths(cuda_platform):
    """The dedicated kick/drift/scale kernels and the generic per-DOF VM give the same trajectory."""
    respa, pdb = systems.respa_water()
    factory = lambda: atomsmm.RespaPropagator([2, 2, 1]).integrator(2*fs)
    out = []
    for fast in ('true', 'false'):
        integrator = factory()
        context = mm.Context(respa, integrator, cuda_platform, {'FastPaths': fast})
        context.setPositions(positions_of(pdb))
        context.setVelocities(thermal_velocities(respa, 300.0, 7))
        integrator.step(6)
        state = context.getState(getPositions=True, getVelocities=True)
        out.append((state.getPositions(asNumpy=True).value_in_unit(unit.nanometer),
                    state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)))
    assert np.max(np.abs(out[0][0] - out[1][0])) < 1e-7
    assert np.max(np.abs(out[0][1] - out[1][1])) < 2e-5


def test_nose_hoover_respa_matches_interpreter(cuda_platform):
    """TrotterSuzuki(Respa, SuzukiYoshida(NoseHoover, 3)): BASELINE config 2's integrator."""
    respa, pdb = systems.respa_water()
    dof = atomsmm.countDegreesOfFreedom(respa)

    def factory():
        nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs, 2)
        return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                               atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    context, integrator, state, reference = run_both(respa, pdb, factory, 4, cuda_platform)
    compare(state, reference)
    assert integrator.getGlobalVariableByName('p_eta') == pytest.approx(reference.globals['p_eta'], rel=1e-5)
    # the engine carries sum(m v.v) across the final rescaling of a thermostat chain (it is s^2 times the last
    # summed value), so its `mvv` belongs to the CURRENT velocities; the literal program leaves the last summed one
    assert integrator.getGlobalVariableByName('mvv') == pytest.approx(
        float(np.sum(reference.mass*reference.v**2)), rel=1e-5)


def test_nose_hoover_chain_and_loops(cuda_platform):
    """NoseHooverChain and NoseHoover(nloops=4: a data-independent while block over globals that
    runs inside the single-thread VM kernel) on the ionic liquid with explicit 1-4 exceptions."""
    system, pdb = systems.flexible('emim_BCN4_Jiung2014', app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A)
    dof = atomsmm.countDegreesOfFreedom(respa)
    for thermostat in (lambda: atomsmm.NoseHooverChainPropagator(300*K, dof, 100*fs),
                       lambda: atomsmm.NoseHooverPropagator(300*K, dof, 100*fs, 4)):
        factory = lambda: atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 1, 1]), thermostat()).integrator(1*fs)
        context, integrator, state, reference = run_both(respa, pdb, factory, 3, cuda_platform)
        compare(state, reference)


def test_nve_energy_conservation(cuda_platform):
    """Velocity Verlet at 0.5 fs on flexible water: total-energy drift over 2 000 steps is small
    and list rebuilds happen on the device (no overflow, several rebuilds)."""
    system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
    nb = system.getForce(atomsmm.findNonbondedForce(system))
    nb.setUseSwitchingFunction(True)
    nb.setSwitchingDistance(9*systems.A)
    integrator = atomsmm.propagators.UnconstrainedVelocityVerletPropagator().integrator(0.5*fs)
    context = mm.Context(system, integrator, cuda_platform)
    context.setPositions(positions_of(pdb))
    context.setVelocities(thermal_velocities(system, 300.0, 3))
    energies = []
    for _ in range(20):
        integrator.step(100)
        state = context.getState(getEnergy=True)
        energies.append((state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(unit.kilojoules_per_mole))
    energies = np.array(energies)
    dof = 3*system.getNumParticles()
    slope = np.polyfit(np.arange(len(energies))*0.05, energies, 1)[0]      # kJ/mol/ps
    assert abs(slope)/dof < 0.02*KB*300                                    # < 2% kT per ps per DOF
    assert np.std(energies) < 5.0
    assert context.counters()['rebuilds'] >= 2


def test_langevin_and_bussi_statistics(cuda_platform):
    """Stochastic baths: Langevin_R (OU core in RESPA) and Bussi velocity rescaling drive the
    kinetic temperature to the bath value; different seeds give different trajectories, the same
    seed reproduces bit-identical velocities."""
    respa, pdb = systems.respa_water()
    dof = atomsmm.countDegreesOfFreedom(respa)

    def temperature(context):
        ke = context.getState(getEnergy=True).getKineticEnergy().value_in_unit(unit.kilojoules_per_mole)
        return 2*ke/(dof*KB)

    def langevin():
        return atomsmm.Langevin_R_Integrator(2*fs, [4, 1, 1], 300*K, 20/ps)

    def bussi():
        thermostat = atomsmm.VelocityRescalingPropagator(300*K, dof, 0.05*ps)
        return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 1, 1]), thermostat).integrator(2*fs)
    for factory in (langevin, bussi):
        finals = []
        for seed in (11, 11, 12):
            integrator = factory()
            integrator.setRandomNumberSeed(seed)
            context = mm.Context(respa, integrator, cuda_platform)
            context.setPositions(positions_of(pdb))
            context.setVelocities(thermal_velocities(respa, 100.0, 5))     # start cold
            integrator.step(400)
            temps = []
            for _ in range(10):
                integrator.step(20)
                temps.append(temperature(context))
            assert 270 < np.mean(temps) < 330
            finals.append(context.getState(getVelocities=True).getVelocities(asNumpy=True).value_in_unit(
                unit.nanometer/unit.picosecond))
        assert np.array_equal(finals[0], finals[1])
        assert not np.allclose(finals[0], finals[2])


def test_long_run_refreshes_the_spatial_order(cuda_platform):
    """Molecules diffuse away from the order the groups were built for; the engine re-sorts between
    steps of a long run (state carried across, lists stay exact) instead of letting the lists grow
    until they overflow."""
    import os
    respa, pdb = systems.respa_water()
    pos = positions_of(pdb)
    vel = thermal_velocities(respa, 600.0, 5)            # hot: fast diffusion
    integrator = atomsmm.RespaPropagator([4, 2, 1]).integrator(4*fs)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    e0 = context.getState(getEnergy=True)
    total0 = (e0.getPotentialEnergy() + e0.getKineticEnergy()).value_in_unit(unit.kilojoules_per_mole)
    integrator.step(3000)                                # 12 ps
    state = context.getState(getEnergy=True, getPositions=True)
    total = (state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(unit.kilojoules_per_mole)
    dof = atomsmm.countDegreesOfFreedom(respa)
    assert abs(total - total0)/dof < 0.05                # energy conserved across the re-orderings
    out = (__import__('ctypes').c_longlong*8)()
    context._call('b2_get_counters', out)
    assert out[7] >= 1000000                             # at least one re-ordering happened
    stats = context.list_stats()
    assert stats['largest_list'] < 2300
