"""
Constructions replayed both by the REFERENCE's Python layer (capture_reference.py, build
container only) and by atomsmm_b200 (tests/test_program_parity.py).  ``loader(case)`` returns an
object with ``.topology`` and ``.positions`` plus a ForceField.
"""

import os

def integrator_record(integrator):
    return dict(
        globals={integrator.getGlobalVariableName(k): integrator._global_values[k]
                 for k in range(integrator.getNumGlobalVariables())},
        perdof=[integrator.getPerDofVariableName(k) for k in range(integrator.getNumPerDofVariables())],
        steps=[list(integrator.getComputationStep(k)) for k in range(integrator.getNumComputations())],
        dt=integrator._dt)


def force_record(force):
    from atomsmm_b200 import mm
    rec = dict(cls=[c.__name__ for c in type(force).__mro__ if c.__module__ == mm.__name__][0],
               group=force.getForceGroup())
    if hasattr(force, 'getEnergyFunction'):
        rec['energy'] = force.getEnergyFunction()
        rec['globals'] = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
                          for k in range(force.getNumGlobalParameters())}
    if hasattr(force, 'getCutoffDistance'):
        rec['cutoff'] = force.getCutoffDistance().value_in_md_units()
        rec['use_switch'] = force.getUseSwitchingFunction()
        rec['switch'] = force.getSwitchingDistance().value_in_md_units()
        rec['method'] = force.getNonbondedMethod()
    if hasattr(force, 'getUseLongRangeCorrection'):
        rec['lrc'] = force.getUseLongRangeCorrection()
    if hasattr(force, 'getNumExclusions'):
        rec['n_exclusions'] = force.getNumExclusions()
    if hasattr(force, 'getNumBonds'):
        rec['n_bonds'] = force.getNumBonds()
    if hasattr(force, 'getNumExceptions'):
        rec['n_exceptions'] = force.getNumExceptions()
    if hasattr(force, 'getNumInteractionGroups') and force.getNumInteractionGroups():
        rec['interaction_groups'] = [[sorted(int(i) for i in s1)[:8], len(s1), len(s2)]
                                     for s1, s2 in (force.getInteractionGroupParameters(k)
                                                    for k in range(force.getNumInteractionGroups()))]
    if hasattr(force, 'getNumEnergyParameterDerivatives') and force.getNumEnergyParameterDerivatives():
        rec['derivatives'] = [force.getEnergyParameterDerivativeName(k)
                              for k in range(force.getNumEnergyParameterDerivatives())]
    return rec


def build_cases(atomsmm, unit, app, loader):
    """The same constructions are replayed with atomsmm_b200 by tests/test_program_parity.py."""
    out = {'integrators': {}, 'forces': {}, 'systems': {}}
    P = atomsmm.propagators
    K, fs, ps = unit.kelvin, unit.femtoseconds, unit.picoseconds
    dof = 8397

    def nh(nloops=1):
        return P.NoseHooverPropagator(300*K, dof, 100*fs, nloops)
    cases = {
        'respa_231': lambda: P.RespaPropagator([2, 3, 1]).integrator(4*fs),
        'respa_421': lambda: P.RespaPropagator([4, 2, 1]).integrator(4*fs),
        'respa_41_constrained': lambda: atomsmm.GlobalThermostatIntegrator(
            1*fs, P.RespaPropagator([4, 1], boost=P.VelocityBoostPropagator(constrained=True),
                                    move=P.TranslationPropagator(constrained=True))),
        'respa_memory': lambda: P.RespaPropagator([2, 2, 1], has_memory=True).integrator(4*fs),
        'respa_switch': lambda: P.RespaPropagator([2, 1], use_respa_switch=True).integrator(2*fs),
        'vv': lambda: atomsmm.GlobalThermostatIntegrator(1*fs, P.VelocityVerletPropagator()),
        'uvv_nh': lambda: P.TrotterSuzukiPropagator(P.UnconstrainedVelocityVerletPropagator(), nh()).integrator(1*fs),
        'respa_nh_sy3': lambda: P.TrotterSuzukiPropagator(
            P.RespaPropagator([4, 2, 1]), P.SuzukiYoshidaPropagator(nh(2), 3)).integrator(4*fs),
        'nh_loops4': lambda: P.TrotterSuzukiPropagator(P.UnconstrainedVelocityVerletPropagator(), nh(4)).integrator(1*fs),
        'nhc': lambda: P.TrotterSuzukiPropagator(P.UnconstrainedVelocityVerletPropagator(),
                                                 P.NoseHooverChainPropagator(300*K, dof, 100*fs)).integrator(1*fs),
        'nhl': lambda: P.TrotterSuzukiPropagator(P.UnconstrainedVelocityVerletPropagator(),
                                                 P.NoseHooverLangevinPropagator(300*K, dof, 100*fs, 10/ps)).integrator(1*fs),
        'bussi_reference_literal': lambda: P.TrotterSuzukiPropagator(
            P.UnconstrainedVelocityVerletPropagator(),
            P.VelocityRescalingPropagator(300*K, dof, 0.1*ps)).integrator(1*fs),
        'langevin_r': lambda: atomsmm.integrators.Langevin_R_Integrator(2*fs, [4, 1], 300*K, 10/ps),
        'mts_xo_sy3': lambda: atomsmm.integrators.MultipleTimeScaleIntegrator(
            4*fs, [4, 2, 1], None, None, nh(), scheme='xo-respa', nsy=3),
        'mts_xi_nres2': lambda: atomsmm.integrators.MultipleTimeScaleIntegrator(
            4*fs, [2, 2, 1], None, None, nh(), scheme='xi-respa', nres=2),
        'mts_side': lambda: atomsmm.integrators.MultipleTimeScaleIntegrator(
            4*fs, [2, 2, 1], None, None, nh(), scheme='side', location=1),
        'mts_blitz': lambda: atomsmm.integrators.MultipleTimeScaleIntegrator(
            4*fs, [2, 2, 1], None, None, P.OrnsteinUhlenbeckPropagator(300*K, 10/ps), scheme='blitz'),
        'chained_split': lambda: P.ChainedPropagator(
            [P.SplitPropagator(P.UnconstrainedVelocityVerletPropagator(), 3), nh()]).integrator(1*fs),
        'nhl_r': lambda: atomsmm.NHL_R_Integrator(2*fs, [4, 1], 300*K, 10*fs, 10/ps),
        'afed_nh': lambda: atomsmm.AdiabaticDynamicsIntegrator(
            P.TrotterSuzukiPropagator(P.VelocityVerletPropagator(), P.NoseHooverPropagator(300*K, 4491, 10*fs)).integrator(1*fs),
            2, [atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)]),
        'afed_langevin_periodic': lambda: atomsmm.AdiabaticDynamicsIntegrator(
            P.UnconstrainedVelocityVerletPropagator().integrator(1*fs), 1,
            [atomsmm.ExtendedSystemVariable('lam', 500*unit.dalton, 2.5*unit.kilojoules_per_mole, 40*fs, -1, 1, True,
                                            'Langevin', 0.05/fs)]),
        'extended_system_propagator': lambda: P.ExtendedSystemPropagator(
            'phi', 30, 6.283185307179586, P.TrotterSuzukiPropagator(
                P.UnconstrainedVelocityVerletPropagator(), P.GenericScalingPropagator('v', 'eta', perDof=True))).integrator(1*fs),
    }
    for name, make in cases.items():
        out['integrators'][name] = integrator_record(make())

    A, nm = unit.angstroms, unit.nanometers
    forces = {
        'near_none': lambda: atomsmm.NearNonbondedForce(10*A, 9.5*A, None),
        'near_shift': lambda: atomsmm.NearNonbondedForce(10*A, 9.5*A, 'shift'),
        'near_fs': lambda: atomsmm.NearNonbondedForce(10*A, 9.5*A, 'force-switch'),
        'near_fs_sub_actual': lambda: atomsmm.NearNonbondedForce(7*A, 5*A, 'force-switch', subtract=True, actual_cutoff=10*A),
        'damped1': lambda: atomsmm.DampedSmoothedForce(0.29/A, 10*A, 9.5*A, degree=1),
        'damped2': lambda: atomsmm.DampedSmoothedForce(0.29/A, 10*A, 9.5*A, degree=2),
        'exceptions': lambda: atomsmm.NonbondedExceptionsForce(),
        'near_exception': lambda: atomsmm.NearExceptionForce(7*A, 5*A, 'shift', subtract=True),
        'softcore': lambda: atomsmm.SoftcoreForce(1.0*nm, 0.9*nm),
        'softcore_lj': lambda: atomsmm.SoftcoreLennardJonesForce(1.0*nm, True, 0.9*nm, True, 'lambda_vdw'),
    }
    pdb, ff = loader('q-SPC-FW')
    for name, make in forces.items():
        system = ff.createSystem(pdb.topology, nonbondedMethod=app.PME)
        nb = atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))
        force = make().importFrom(nb)
        out['forces'][name] = force_record(force)
    for adj in (None, 'shift', 'force-switch'):
        system = ff.createSystem(pdb.topology, nonbondedMethod=app.PME)
        nb = atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))
        inner = atomsmm.NearNonbondedForce(7*A, 6.5*A, adj).importFrom(nb)
        outer = atomsmm.FarNonbondedForce(inner, 10*A, 9.5*A).setForceGroup(2).importFrom(nb)
        out['forces']['far_%s' % adj] = [force_record(f) for f in outer]
        out['forces']['near_lists_%s' % adj] = dict(near=atomsmm.forces.nearForceExpressions(7*A, 5*A, adj))
    for case in ('q-SPC-FW', 'emim_BCN4_Jiung2014'):
        pdb, ff = loader(case)
        system = ff.createSystem(pdb.topology, nonbondedMethod=app.PME, constraints=None, rigidWater=False,
                                 removeCMMotion=False)
        respa = atomsmm.RESPASystem(system, 7*A, 5*A)
        out['systems']['respa_' + case] = [force_record(f) for f in respa.getForces()]
        respa_slow = atomsmm.RESPASystem(system, 7*A, 5*A, adjustment='shift', fastExceptions=False)
        out['systems']['respa_slowexc_' + case] = [force_record(f) for f in respa_slow.getForces()]
        computing = atomsmm.ComputingSystem(system)
        out['systems']['computing_' + case] = [force_record(f) for f in computing.getForces()]
    # AlchemicalSystem of the (disabled) AFED test of the reference, tests/test_afed.py:21-35
    pdb, ff = loader('methane-in-water')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.PME, constraints=None, rigidWater=False,
                             removeCMMotion=False)
    solute = set(i for i, atom in enumerate(pdb.topology.atoms()) if atom.residue.name == 'C1')
    alchemical = atomsmm.systems.AlchemicalSystem(system, solute)
    records = [force_record(f) for f in alchemical.getForces()]
    nb = alchemical.getForce(atomsmm.findNonbondedForce(alchemical))
    records.append(dict(solute=sorted(solute), solute_parameters=[
        [float(getattr(v, 'value_in_md_units', lambda: v)()) for v in nb.getParticleParameters(i)] for i in sorted(solute)]))
    out['systems']['alchemical_methane'] = records
    return out


