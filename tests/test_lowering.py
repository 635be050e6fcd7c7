"""Host-side lowering: family recognition of energy strings and step-program lowering."""

import numpy as np
import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import lowering, mm, unit

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


def test_constraints_are_refused():
    integrator = atomsmm.GlobalThermostatIntegrator(1*fs, atomsmm.VelocityVerletPropagator())
    with pytest.raises(lowering.UnsupportedDescription):
        lowering.lower_program(integrator, 1)
