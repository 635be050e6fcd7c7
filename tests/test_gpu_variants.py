"""
GPU tests for paths the reference emits but round 1 exercised only on the CPU (VERDICT round 1,
"missing" item 5 and "weak" items 3-5):

* RespaPropagator keyword variants has_memory / use_respa_switch / blitz (propagators.py:922-956) and
  RESPASystem(fastExceptions=False) with its NearExceptionForce pair of CustomBondForces
  (forces.py:673-680, systems.py:67,78) against the float64 oracle;
* NoseHooverLangevinPropagator (propagators.py:1452-1534): deterministic limit against the oracle
  interpreter, and the canonical statistics of its thermostat momentum;
* the NVE gate of SURVEY 8d: config 1, RESPA [4,2,1], 1 000 steps, drift per degree of freedom on the
  GPU beside the drift of the float64 oracle executing the same step program;
* stochastic thermostats: variance of the kinetic energy and a chi-square test of the gamma deviates
  of Bussi's Marsaglia-Tsang rejection loop (propagators.py:1198-1215), Gaussian velocities for Langevin;
* bit-reproducibility: two identical runs give identical positions and velocities.
"""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, mm, unit

import systems
from systems import positions_of
from test_gpu_integrators import compare, run_both, thermal_velocities

pytestmark = pytest.mark.gpu

fs, ps, K = unit.femtoseconds, unit.picoseconds, unit.kelvin
KB = 8.314472471220217e-3
kJ = unit.kilojoules_per_mole


@pytest.mark.parametrize('variant', ['has_memory', 'use_respa_switch', 'blitz'])
def test_respa_keyword_variants_match_interpreter(cuda_platform, variant):
    respa, pdb = systems.respa_water()
    factory = lambda: atomsmm.RespaPropagator([2, 2, 1], **{variant: True}).integrator(2*fs)
    context, integrator, state, reference = run_both(respa, pdb, factory, 4, cuda_platform)
    compare(state, reference)
    total = (state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(kJ)
    assert total == pytest.approx(reference.potential_energy() + reference.kinetic_energy(), rel=1e-6)


def test_respa_system_without_fast_exceptions(cuda_platform):
    """fastExceptions=False: the 1-4 pairs of the ionic liquid stay at the near/far time scales as two
    CustomBondForces carrying the near expression and its negative (NearExceptionForce's construction)."""
    from oracle import refmath
    system, pdb = systems.flexible('emim_BCN4_Jiung2014', app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*systems.A, 5*systems.A, fastExceptions=False)
    bond_groups = sorted(f.getForceGroup() for f in respa.getForces() if isinstance(f, mm.CustomBondForce))
    assert bond_groups == [1, 31]
    pos = positions_of(pdb)
    context = mm.Context(respa, mm.VerletIntegrator(0.0), cuda_platform)
    context.setPositions(pos)
    for groups in ({0}, {1}, {2}, {31}):
        state = context.getState(getEnergy=True, getForces=True, groups=groups)
        ref = refmath.evaluate_system(respa, pos, groups=groups)
        assert state.getPotentialEnergy().value_in_unit(kJ) == pytest.approx(ref.energy, rel=1e-6, abs=1e-6)
        forces = state.getForces(asNumpy=True).value_in_unit(kJ/unit.nanometer)
        assert np.sqrt(np.sum((forces - ref.forces)**2)/np.sum(ref.forces**2)) < 1e-5
    # and the time stepping with the exceptions at the middle time scale
    factory = lambda: atomsmm.RespaPropagator([2, 2, 1]).integrator(1*fs)
    context, integrator, state, reference = run_both(respa, pdb, factory, 3, cuda_platform)
    compare(state, reference)


def test_near_exception_force(cuda_platform):
    """NearExceptionForce (forces.py:673-680) on the ionic liquid's exceptions, single point."""
    from oracle import refmath
    pdb, ff = systems.fixtures.load('emim_BCN4_Jiung2014')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic)
    nb = atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))
    force = atomsmm.forces.NearExceptionForce(7*systems.A, 5*systems.A, 'force-switch')
    force.importFrom(nb).addTo(system)
    pos = positions_of(pdb)
    context = mm.Context(system, mm.VerletIntegrator(0.0), cuda_platform)
    context.setPositions(pos)
    group = force.getForceGroup()
    state = context.getState(getEnergy=True, getForces=True, groups={group})
    ref = refmath.evaluate_system(system, pos, groups={group})
    assert abs(ref.energy) > 1.0
    assert state.getPotentialEnergy().value_in_unit(kJ) == pytest.approx(ref.energy, rel=1e-6)
    forces = state.getForces(asNumpy=True).value_in_unit(kJ/unit.nanometer)
    assert np.sqrt(np.sum((forces - ref.forces)**2)/np.sum(ref.forces**2)) < 1e-5


def test_nose_hoover_langevin(cuda_platform):
    respa, pdb = systems.respa_water()
    dof = atomsmm.countDegreesOfFreedom(respa)

    def factory(friction):
        nhl = atomsmm.NoseHooverLangevinPropagator(300*K, dof, 100*fs, friction)
        return lambda: atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 2, 1]), nhl).integrator(2*fs)
    # (a) vanishing friction: the noise amplitude sqrt(kT Q (1 - exp(-2 gamma h))) is ~1e-7 of the thermostat
    # momentum, so the GPU (Philox) and the oracle interpreter (numpy) must agree although their streams differ
    context, integrator, state, reference = run_both(respa, pdb, factory(1e-12/ps), 4, cuda_platform)
    compare(state, reference)
    assert integrator.getGlobalVariableByName('p_NHL') == pytest.approx(reference.globals['p_NHL'], rel=1e-4)
    # (b) real friction: the temperature is held and the thermostat variable fluctuates around zero.  (The
    # reference's program mixes a momentum-like drive with a velocity-like noise amplitude kT/Q,
    # propagators.py:1521-1526, so its stationary variance is not the canonical Q kT; the emitted program is
    # reproduced as it is -- tests/test_program_parity.py -- and only model-independent facts are asserted.)
    integrator = factory(50/ps)()
    integrator.setRandomNumberSeed(3)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(positions_of(pdb))
    context.setVelocities(thermal_velocities(respa, 300.0, 5))
    integrator.step(500)
    p, temps = [], []
    for _ in range(1500):
        integrator.step(4)
        p.append(integrator.getGlobalVariableByName('p_NHL'))
        temps.append(2*context.getState(getEnergy=True).getKineticEnergy().value_in_unit(kJ)/(dof*KB))
    assert np.mean(temps) == pytest.approx(300.0, rel=0.02)
    assert np.std(p) > 0 and abs(np.mean(p)) < 3*np.std(p)


def test_massive_nose_hoover_langevin_through_the_generic_per_dof_path(cuda_platform):
    """NHL_R_Integrator (integrators.py:349-352 family): a MASSIVE thermostat -- one Nose-Hoover-Langevin
    variable per degree of freedom, i.e. per-DOF expressions that no dedicated kernel matches and that run
    through the generic per-DOF path (kernels compiled from the bytecode at run time, csrc/jit.cu; the device-side
    virtual machine without NVRTC) -- against the float64 oracle interpreter, in the limit of
    vanishing friction where the two random streams cannot matter."""
    from oracle import interp
    respa, pdb = systems.respa_water()
    pos = positions_of(pdb)
    vel = thermal_velocities(respa, 300.0, 1234)
    # friction 1e-8/ps: noise ~5e-6 of the thermostat velocity, while 1 - exp(-gamma h) = 1e-11 is still resolved
    factory = lambda: atomsmm.NHL_R_Integrator(2*fs, [2, 1, 1], 300*K, 50*fs, 1e-8/ps)
    integrator = factory()
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    integrator.step(0)        # the integrator's first step() draws the thermostat velocities (integrators.py initialize hook)
    reference = interp.Interpreter(respa, factory(), pos, vel)
    names = [integrator.getPerDofVariableName(k) for k in range(integrator.getNumPerDofVariables())]
    assert 'v2' in names
    for name in names:        # same initial thermostat state on both sides
        reference.perdof[name] = np.array(integrator.getPerDofVariableByName(name), dtype=np.float64)
    reference.globals['NDOF'] = integrator.getGlobalVariableByName('NDOF')
    assert np.std(reference.perdof['v2']) > 1.0
    integrator.step(3)
    reference.step(3)
    state = context.getState(getPositions=True, getVelocities=True)
    compare(state, reference)
    ours = np.array(integrator.getPerDofVariableByName('v2'))
    theirs = reference.perdof['v2']
    assert np.max(np.abs(ours - theirs)) < 2e-4*float(np.max(np.abs(theirs)))
    # the generic per-DOF steps ran as kernels compiled from their bytecode at run time (csrc/jit.cu), not through
    # the interpreter: NVRTC and the driver library are part of the CUDA installation on a GPU box
    stats = context.jit_stats()
    assert stats['compiled_steps'] > 0 and stats['launches'] > 0, stats


def test_compiled_per_dof_steps_give_the_bits_of_the_interpreter(cuda_platform):
    """The kernels csrc/jit.cu compiles from per-DOF bytecode (NVRTC, --fmad=false, the interpreter's Philox stream)
    against the device-side interpreter they replace: a massive Nose-Hoover-Langevin thermostat WITH friction --
    per-DOF expressions with exponentials, square roots and Gaussian draws -- must give bit-identical positions,
    velocities and thermostat variables with the compiler on and off."""
    respa, pdb = systems.respa_water()
    pos = positions_of(pdb)
    vel = thermal_velocities(respa, 300.0, 4321)
    results = []
    for compiler in ('true', 'false'):
        integrator = atomsmm.NHL_R_Integrator(2*fs, [2, 1, 1], 300*K, 50*fs, 10/ps)
        integrator.setRandomNumberSeed(99)
        context = mm.Context(respa, integrator, cuda_platform, {'PerDofCompiler': compiler})
        context.setPositions(pos)
        context.setVelocities(vel)
        integrator.step(6)
        state = context.getState(getPositions=True, getVelocities=True)
        stats = context.jit_stats()
        assert (stats['compiled_steps'] > 0) == (compiler == 'true'), stats
        results.append((state._positions, state._velocities, np.array(integrator.getPerDofVariableByName('v2'))))
    for a, b in zip(*results):
        assert np.array_equal(a, b)


def test_nve_drift_no_worse_than_the_float64_oracle(cuda_platform):
    """config 1 (1 536-atom water), RespaPropagator([4,2,1]) at 4 fs, 1 000 steps; slope of the
    total energy per degree of freedom on the GPU and on the float64 oracle executing the same step program
    (oracle/c/oracle.c's RESPA driver, checked against oracle/interp.py in tests/test_oracle_c.py)."""
    from oracle import cport
    respa, pdb = systems.respa_water()
    pos = positions_of(pdb)
    vel = thermal_velocities(respa, 300.0, 1234)
    dof = atomsmm.countDegreesOfFreedom(respa)
    mass = np.array(respa._masses)
    integrator = atomsmm.RespaPropagator([4, 2, 1]).integrator(4*fs)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(pos)
    context.setVelocities(vel)
    port = cport.CPort(respa)
    groups = {0, 2}          # the conserved energy: RESPA kicks with f0 + f1 + (f2 - f1); group 1 is the near PART of 2
    x, v = pos.copy(), vel.copy()
    gpu, cpu = [], []
    for _ in range(50):
        integrator.step(20)
        state = context.getState(getEnergy=True, groups=groups)
        gpu.append((state.getPotentialEnergy() + state.getKineticEnergy()).value_in_unit(kJ))
        x, v, _ = port.respa(x, v, 20, 0.004, 4, 2)
        cpu.append(port.evaluate(x, groups)[1] + 0.5*float(np.sum(mass[:, None]*v*v)))
    t = np.arange(1, 51)*20*0.004                       # ps
    slope_gpu = np.polyfit(t, gpu, 1)[0]/dof            # kJ/mol/ps per degree of freedom
    slope_cpu = np.polyfit(t, cpu, 1)[0]/dof
    # "no worse than the reference's": within 2x of the float64 oracle, with a floor for the case where the
    # oracle's slope is accidentally ~0 (the fitted slope of a 4 ps window has a statistical spread of its own)
    floor = 0.002*KB*300
    assert abs(slope_gpu) <= max(2*abs(slope_cpu), floor), (slope_gpu, slope_cpu)
    # same fluctuation amplitude of the shadow-Hamiltonian error
    assert np.std(np.array(gpu) - np.polyval(np.polyfit(t, gpu, 1), t)) < \
        2*np.std(np.array(cpu) - np.polyval(np.polyfit(t, cpu, 1), t)) + 0.5
    # both start from the same energy
    assert gpu[0] == pytest.approx(cpu[0], rel=1e-5)


def test_bussi_kinetic_energy_distribution_and_gamma_deviates(cuda_platform):
    """Canonical kinetic-energy statistics under stochastic velocity rescaling (mean N_f kT/2, variance
    N_f (kT)^2/2) and the Marsaglia-Tsang gamma deviates drawn inside the step program on the device."""
    from scipy import stats
    respa, pdb = systems.respa_water()
    dof = atomsmm.countDegreesOfFreedom(respa)
    kT = KB*300
    thermostat = atomsmm.VelocityRescalingPropagator(300*K, dof, 0.02*ps)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 1, 1]), thermostat).integrator(2*fs)
    integrator.setRandomNumberSeed(2024)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(positions_of(pdb))
    context.setVelocities(thermal_velocities(respa, 300.0, 5))
    integrator.step(1000)
    shape = (dof - 2 + dof % 2)/2
    d = shape - 1/3
    kinetic, gamma = [], []
    for _ in range(3000):
        integrator.step(5)                              # tau = 10 steps: samples 0.5 tau apart
        kinetic.append(context.getState(getEnergy=True).getKineticEnergy().value_in_unit(kJ))
        gamma.append(d*integrator.getGlobalVariableByName('V'))
    kinetic, gamma = np.array(kinetic), np.array(gamma)
    assert np.mean(kinetic) == pytest.approx(0.5*dof*kT, rel=0.01)
    # 3 000 correlated samples (~1 500 independent): the variance estimate has ~4 % standard error
    assert np.var(kinetic) == pytest.approx(0.5*dof*kT*kT, rel=0.2)
    # gamma deviates: independent draws; 10 equiprobable bins of Gamma(shape) -> chi-square with 9 dof
    edges = stats.gamma.ppf(np.linspace(0, 1, 11), shape)
    observed = np.histogram(gamma, bins=edges)[0]
    assert observed.sum() == len(gamma)
    chi2 = float(np.sum((observed - len(gamma)/10)**2/(len(gamma)/10)))
    assert chi2 < 27.9, (chi2, observed)                # p = 0.001
    assert np.mean(gamma) == pytest.approx(shape, rel=0.003)


def test_langevin_velocities_are_gaussian_with_the_bath_variance(cuda_platform):
    """Ornstein-Uhlenbeck core in RESPA (Langevin_R_Integrator): per-DOF m v^2 averages to kT for both atom
    types, velocities are Gaussian (kurtosis 3), kinetic energy has the canonical variance."""
    respa, pdb = systems.respa_water()
    dof = 3*respa.getNumParticles()
    kT = KB*300
    integrator = atomsmm.Langevin_R_Integrator(2*fs, [4, 1, 1], 300*K, 20/ps)
    integrator.setRandomNumberSeed(7)
    context = mm.Context(respa, integrator, cuda_platform)
    context.setPositions(positions_of(pdb))
    context.setVelocities(thermal_velocities(respa, 300.0, 5))
    integrator.step(1000)
    mass = np.array(respa._masses)
    mv2, kurt, kinetic = [], [], []
    for _ in range(400):
        integrator.step(25)                             # 1/gamma = 25 steps
        v = context.getState(getVelocities=True)._velocities
        z = v*np.sqrt(mass/kT)[:, None]
        mv2.append([np.mean(z[mass > 2]**2), np.mean(z[mass < 2]**2)])
        kurt.append(np.mean(z**4)/np.mean(z**2)**2)
        kinetic.append(0.5*float(np.sum(mass[:, None]*v*v)))
    mv2 = np.mean(mv2, axis=0)
    # (hydrogens read ~1 % cold at 2 fs: the known configurational-vs-kinetic temperature gap of the splitting)
    assert mv2[0] == pytest.approx(1.0, abs=0.01) and mv2[1] == pytest.approx(1.0, abs=0.02)
    assert np.mean(kurt) == pytest.approx(3.0, abs=0.03)
    assert np.mean(kinetic) == pytest.approx(0.5*dof*kT, rel=0.015)
    assert np.var(kinetic) == pytest.approx(0.5*dof*kT*kT, rel=0.25)


@pytest.mark.parametrize('case', ['water-nh', 'ionic-liquid'])
def test_trajectories_are_bit_reproducible(cuda_platform, case):
    """Two identical runs (fresh contexts, same inputs) end in bit-identical positions and velocities: pair
    tiles are reduced in a fixed order, cutoff-band pairs are settled in fixed point, the fused inner loop
    accumulates bonded forces in fixed point, reductions use fixed trees.  The runs include list rebuilds."""
    if case == 'water-nh':
        system, pdb = systems.respa_water()
        big, pos = systems.replicate(system, positions_of(pdb), np.array([2.5, 2.5, 2.5]), 2)
        dof = atomsmm.countDegreesOfFreedom(big)

        def factory():
            nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs)
            return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                   atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
        steps = 60
    else:
        base, pdb = systems.flexible('emim_BCN4_Jiung2014', app.CutoffPeriodic)
        big = atomsmm.RESPASystem(base, 7*systems.A, 5*systems.A)
        pos = positions_of(pdb)
        factory = lambda: atomsmm.RespaPropagator([2, 2, 1]).integrator(2*fs)
        steps = 80
    vel = thermal_velocities(big, 300.0, 11)
    finals = []
    for _ in range(2):
        integrator = factory()
        context = mm.Context(big, integrator, cuda_platform)
        context.setPositions(pos)
        context.setVelocities(vel)
        integrator.step(steps)
        state = context.getState(getPositions=True, getVelocities=True)
        finals.append((state._positions, state._velocities, context.list_stats()['rebuilds']))
    assert finals[0][2] >= 2
    assert np.array_equal(finals[0][0], finals[1][0])
    assert np.array_equal(finals[0][1], finals[1][1])
