"""Host-side lowering: family recognition of energy strings and step-program lowering."""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, lowering, mm, unit

import systems

A, fs, K, ps = unit.angstroms, unit.femtoseconds, unit.kelvin, unit.picoseconds


@pytest.mark.parametrize('adjustment,variant', [(None, 0), ('shift', 1), ('force-switch', 2)])
def test_near_family_recognition(adjustment, variant):
    system, pdb, force = systems.water_near(adjustment)
    family, cutoff, params, info = lowering.classify_pair_force(force)
    assert family == lowering.PAIR_NEAR and int(params[0]) == variant and params[4] == 1.0
    assert cutoff == pytest.approx(1.0) and params[1] == pytest.approx(0.95) and params[2] == pytest.approx(1.0)


def test_discount_and_damped_and_virial_recognition():
    respa, pdb = systems.respa_water()
    kinds = {}
    for force in respa.getForces():
        if isinstance(force, mm.CustomNonbondedForce):
            family, cutoff, params, info = lowering.classify_pair_force(force)
            kinds[force.getForceGroup()] = (family, params[4])
    assert kinds[1] == (lowering.PAIR_NEAR, 1.0) and kinds[31] == (lowering.PAIR_NEAR, -1.0)
    for degree in (1, 2):
        system, pdb, force = systems.water_damped(degree)
        family, cutoff, params, info = lowering.classify_pair_force(force)
        assert family == lowering.PAIR_DAMPED and int(params[3]) == degree and params[0] == pytest.approx(2.9)
    system, pdb = systems.flexible('q-SPC-FW')
    computing = atomsmm.ComputingSystem(system)
    families = [lowering.classify_pair_force(f)[0] for f in computing.getForces() if isinstance(f, mm.CustomNonbondedForce)]
    assert families == [lowering.PAIR_LJ_VIRIAL]


def test_unknown_energy_is_rejected():
    system, pdb, force = systems.water_near(None)
    force.setEnergyFunction('sin(r)*chargeprod; chargeprod=charge1*charge2')
    with pytest.raises(lowering.UnsupportedDescription):
        lowering.classify_pair_force(force)


def test_respa_program_lowering():
    integrator = atomsmm.RespaPropagator([4, 2, 1]).integrator(4*fs)
    program = lowering.lower_program(integrator, 0b111)
    kinds = [op[0] for op in program.ops]
    assert kinds.count(lowering.OP_KICK) == 9            # 22 kicks + 8 drifts fused into 9 launches
    assert kinds.count(lowering.OP_DRIFT) == 0 and kinds.count(lowering.OP_PERDOF) == 0
    assert kinds.count(lowering.OP_GLOBAL) == 1          # only the per-step coefficient prologue


def test_nose_hoover_block_is_one_velocity_kernel():
    dof = 4605
    nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                 atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)
    program = lowering.lower_program(integrator, 0b111)
    kinds = [op[0] for op in program.ops]
    # sum(m v v) + scalar update + rescaling fold into the velocity kernels: no SUM / SCALE ops left,
    # and the only stand-alone scalar program is the coefficient prologue
    assert kinds.count(lowering.OP_SUM) == 0 and kinds.count(lowering.OP_SCALE) == 0
    assert kinds.count(lowering.OP_GLOBAL) == 1
    reductions = [op for op in program.ops if op[0] == lowering.OP_KICK and op[5] >= 0]
    # ... and a Suzuki-Yoshida chain of 3 blocks shares ONE reduction (m*v*v scales by s*s), with the
    # product of the three factors applied by the next velocity kernel
    assert len(reductions) == 2 and all(op[7] > 100 for op in reductions)
    assert kinds.count(lowering.OP_KICK) == 9 + 2      # 9 RESPA kicks + opening reduction + closing scale
    assert '_vscale_product' in program.global_names


def test_bussi_program_keeps_rejection_loop_on_device():
    thermostat = atomsmm.VelocityRescalingPropagator(300*K, 4605, 0.1*ps)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 1]), thermostat).integrator(1*fs)
    program = lowering.lower_program(integrator, 0b11)
    ops = [op[0] for op in program.ops]
    # the two velocity rescalings ride on velocity kernels (pre-scale of a kick / bare scale op)
    assert lowering.OP_PERDOF not in ops and ops.count(lowering.OP_SCALE) == 0
    assert sum(1 for op in program.ops if op[0] == lowering.OP_KICK and op[4] >= 0) == 2
    unfused = lowering.lower_program(atomsmm.TrotterSuzukiPropagator(
        atomsmm.RespaPropagator([2, 1]), atomsmm.VelocityRescalingPropagator(300*K, 4605, 0.1*ps)).integrator(1*fs),
        0b11, fast=False)
    assert lowering.OP_KICK not in [op[0] for op in unfused.ops]
    assert 33 in program.bc.code[::2]                    # VM_JMPZ: while/if compiled into the scalar VM


def test_constraint_steps_are_lowered():
    integrator = atomsmm.GlobalThermostatIntegrator(1*fs, atomsmm.VelocityVerletPropagator())
    program = lowering.lower_program(integrator, 1)
    kinds = [op[0] for op in program.ops]
    assert kinds.count(lowering.OP_CONSTRAIN_X) == 1 and kinds.count(lowering.OP_CONSTRAIN_V) == 1
    free = lowering.lower_program(integrator, 1, constrained=False)       # a System without constraints
    assert lowering.OP_CONSTRAIN_X not in [op[0] for op in free.ops]


def test_afed_program_lowering():
    """AdiabaticDynamicsIntegrator on an AlchemicalSystem: deriv(energy, lambda) becomes a derivative
    evaluation + VM_PUSHE, moving lambda invalidates the cached forces, constraint steps of an
    unconstrained System vanish (integrators.py:735-737,782-860; systems.py:318-410)."""
    from atomsmm_b200 import mm
    pdb, ff = systems.fixtures.load('methane-in-water')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False)
    alchemical = atomsmm.AlchemicalSystem(system, {0})
    softcore = [f for f in alchemical.getForces() if isinstance(f, mm.CustomNonbondedForce)][0]
    family, cutoff, params, info = lowering.classify_pair_force(softcore, {'lambda_vdw': 0.7})
    assert family == lowering.PAIR_SOFTCORE and params[1] == 0.7 and params[6] == 1.0 and info['partition'] == [0]
    nvt = atomsmm.TrotterSuzukiPropagator(
        atomsmm.VelocityVerletPropagator(),
        atomsmm.NoseHooverPropagator(300*K, atomsmm.countDegreesOfFreedom(alchemical), 10*fs)).integrator(1*fs)
    variable = atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)
    integrator = atomsmm.AdiabaticDynamicsIntegrator(nvt, 2, [variable])
    program = lowering.lower_program(integrator, 1, {'lambda_vdw': 1.0}, True, constrained=False,
                                     derivative_slots={'lambda_vdw': lowering.ENERGY_SLOT_DLAMBDA_VDW})
    kinds = [op[0] for op in program.ops]
    assert kinds.count(lowering.OP_ENERGY) == 8           # one per lambda kick: 2 per inner iteration x 4
    assert kinds.count(lowering.OP_INVALIDATE) == 4       # two half moves of lambda, each with a wall check
    assert 35 in program.bc.code[::2]                     # VM_PUSHE
    with pytest.raises(lowering.UnsupportedDescription):  # nobody provides d/d(lambda) of the energy
        lowering.lower_program(integrator, 1, {'lambda_vdw': 1.0}, True, constrained=False)


def test_alchemical_system_oracle_interaction_group():
    """The soft-core force only couples the solute to the solvent, and at lambda = 1 it is plain LJ."""
    from oracle import refmath
    from atomsmm_b200 import mm
    pdb, ff = systems.fixtures.load('methane-in-water')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False)
    alchemical = atomsmm.AlchemicalSystem(system, {0})
    index = [k for k, f in enumerate(alchemical.getForces()) if isinstance(f, mm.CustomNonbondedForce)][0]
    pos = systems.positions_of(pdb)
    result = refmath.eval_custom_nonbonded(alchemical.getForce(index), pos, refmath.system_box(alchemical),
                                           {'lambda_vdw': 1.0}, want_pairs=True)
    i, j, r = result.pairs
    assert len(i) > 50 and np.all((i == 0) ^ (j == 0))
    nb = system.getForce(atomsmm.findNonbondedForce(system))
    table = np.array([[v.value_in_md_units() for v in nb.getParticleParameters(k)] for k in range(len(pos))])
    sig = 0.5*(table[i, 1] + table[j, 1])
    eps = np.sqrt(table[i, 2]*table[j, 2])
    assert result.energy == pytest.approx(float(np.sum(4*eps*((sig/r)**12 - (sig/r)**6))), rel=1e-12)
