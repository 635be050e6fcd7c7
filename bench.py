#!/usr/bin/env python
"""
Benchmark of the hot path (BASELINE.json metric: atom-steps/s and ns/day; pair-kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c5|c2|c3|c4] [--replicas]

Default workload (config.workload) = BASELINE config 5, the configuration the metric's multi-GPU half
is quoted on and the largest one: the q-SPC-FW water cell replicated 14x14x14 (4 214 784 atoms,
L = 35 nm, +-0.002 nm jitter, seed 1), RESPASystem (near force-switch 0.7/0.5 nm in group 1, full
LJ + reaction-field Coulomb rc 1.0 nm in group 2, bonded terms in group 0), integrator
TrotterSuzuki(Respa([4,2,1]), SuzukiYoshida(NoseHoover(300 K, dof, 100 fs), 3)) at 4 fs.  It fits one
B200 (N = 1) and is the one configuration that shards: with N > 1 ranks ONE system is integrated by
all ranks under spatial domain decomposition ("scaling": "strong"; halo positions are pulled over
NVLink peer memory before every pair-force evaluation, csrc/dist.cu).  `--workload c2` is BASELINE
config 2 (98 304 atoms, same forces and integrator), c3 the ionic liquid, c4 the AFED replica;
`--replicas` runs N independent replicas instead of one decomposed system ("scaling": "weak").

A bench "step" is one call integrator.step(md_steps_per_step) (default 100 outer MD steps; every MD
step is 8 inner iterations, 3 pair-force evaluations and 2 thermostat chains): one pass of the hot
path over one batch.  `value` times K such calls with state resident in HBM; `e2e` times the same
through the public API with HOST buffers: upload of positions+velocities from pinned host memory, the
MD steps, download of positions+velocities+energies, every step.

Before anything is timed the first frame is checked against the oracle (`parity` in the line): forces
of every force group within 1e-5 relative RMS, energies within 1e-6 relative, and the interacting
pair sets of both neighbour lists equal (count and checksum) to those of the float64 C restatement
(oracle/cport.py).  A failed gate makes the run exit non-zero after printing the line.
"""

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

MD_STEPS_PER_CALL = 100
DT_FS = 4.0
LOOPS = [4, 2, 1]


def build_workload(reps):
    import numpy as np
    import atomsmm_b200 as atomsmm
    from atomsmm_b200 import app, unit
    import systems
    system, pdb = systems.flexible('q-SPC-FW', app.CutoffPeriodic)
    respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
    base_pos = systems.positions_of(pdb)
    box = np.array([2.5, 2.5, 2.5])
    if reps > 1:
        big, pos = systems.replicate(respa, base_pos, box, reps)
    else:
        big, pos = respa, base_pos
    n = big.getNumParticles()
    mass = np.array(big._masses, dtype=np.float64)
    rng = np.random.Generator(np.random.Philox(1234))
    vel = rng.standard_normal((n, 3))*np.sqrt(8.314472471220217e-3*300.0/mass)[:, None]
    vel -= (mass[:, None]*vel).sum(0)/mass.sum()
    return big, pos, vel


def build_c3(reps):
    """BASELINE config 3: emim/B(CN)4 ionic liquid replicated reps^3 (4 -> 179 200 atoms), DampedSmoothedForce
    (alpha 2.9/nm, rc 1.0, rs 0.95 nm) in group 1, NonbondedExceptionsForce + bonded terms in group 0,
    RESPA [4,1] at 2 fs with Bussi velocity rescaling (tau 0.1 ps)."""
    import numpy as np
    import atomsmm_b200 as atomsmm
    from atomsmm_b200 import app, unit
    import systems
    pdb, ff = systems.fixtures.load('emim_BCN4_Jiung2014')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False,
                             removeCMMotion=False)
    nb = atomsmm.hijackForce(system, atomsmm.findNonbondedForce(system))
    exceptions = atomsmm.forces.NonbondedExceptionsForce()
    exceptions.importFrom(nb)
    exceptions.setForceGroup(0)
    system.addForce(exceptions)
    damped = atomsmm.DampedSmoothedForce(2.9/unit.nanometer, 1.0*unit.nanometer, 0.95*unit.nanometer)
    damped.importFrom(nb)
    damped.setForceGroup(1)
    system.addForce(damped)
    box = np.array([v.value_in_md_units()[k] for k, v in enumerate(system.getDefaultPeriodicBoxVectors())])
    pos = systems.positions_of(pdb)
    big, pos = systems.replicate(system, pos, box, reps) if reps > 1 else (system, pos)
    n = big.getNumParticles()
    mass = np.array(big._masses, dtype=np.float64)
    rng = np.random.Generator(np.random.Philox(1234))
    vel = rng.standard_normal((n, 3))*np.sqrt(8.314472471220217e-3*300.0/mass)[:, None]
    vel -= (mass[:, None]*vel).sum(0)/mass.sum()
    dof = atomsmm.countDegreesOfFreedom(big)
    bussi = atomsmm.VelocityRescalingPropagator(300*unit.kelvin, dof, 0.1*unit.picoseconds)
    integrator = atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 1]), bussi).integrator(2*unit.femtoseconds)
    return big, pos, vel, integrator


def build_c4():
    """BASELINE config 4: AFED on methane in water (1 498 atoms), soft-core solute-solvent coupling with
    lambda_vdw as extended variable (the construction of the reference's tests/test_afed.py:21-35)."""
    import numpy as np
    import atomsmm_b200 as atomsmm
    from atomsmm_b200 import app, unit
    import systems
    pdb, ff = systems.fixtures.load('methane-in-water')
    system = ff.createSystem(pdb.topology, nonbondedMethod=app.CutoffPeriodic, constraints=None, rigidWater=False,
                             removeCMMotion=False)
    solute = set(i for i, atom in enumerate(pdb.topology.atoms()) if atom.residue.name == 'C1')
    alchemical = atomsmm.AlchemicalSystem(system, solute)
    fs = unit.femtoseconds
    nvt = atomsmm.TrotterSuzukiPropagator(
        atomsmm.VelocityVerletPropagator(),
        atomsmm.NoseHooverPropagator(300*unit.kelvin, atomsmm.countDegreesOfFreedom(alchemical), 10*fs)).integrator(1*fs)
    variable = atomsmm.ExtendedSystemVariable('lambda_vdw', 1000, 5, 40*fs)
    integrator = atomsmm.AdiabaticDynamicsIntegrator(nvt, 2, [variable])
    n = alchemical.getNumParticles()
    mass = np.array(alchemical._masses, dtype=np.float64)
    rng = np.random.Generator(np.random.Philox(1234))
    vel = rng.standard_normal((n, 3))*np.sqrt(8.314472471220217e-3*300.0/mass)[:, None]
    return alchemical, systems.positions_of(pdb), vel, integrator


def make_integrator(system):
    import atomsmm_b200 as atomsmm
    from atomsmm_b200 import unit
    dof = atomsmm.countDegreesOfFreedom(system)
    nh = atomsmm.NoseHooverPropagator(300*unit.kelvin, dof, 100*unit.femtoseconds)
    return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator(LOOPS),
                                           atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(DT_FS*unit.femtoseconds), dof


class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []           # (host time, csv line)
        self.proc = None
        self.window = [None, None]

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.window[0] = time.perf_counter()

    def mark_end(self):
        self.window[1] = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        inside = [line for t, line in self.lines
                  if self.window[0] is None or (self.window[0] <= t <= (self.window[1] or t))]
        if not inside:            # a timed region shorter than the sampling period: nearest samples
            inside = [line for _, line in self.lines[-3:]]
        for line in inside:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[5:9]):
                if value.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm)//2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as handle:
            return float(json.load(handle)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def workload_text(args, n):
    if args.workload in ('c2', 'c5'):
        return ('%s: q-SPC-FW water x%d^3, %d atoms, RESPASystem near 0.7/0.5 nm force-switch + LJ/reaction-field '
                '1.0 nm, RESPA [4,2,1] + NoseHoover SY3, dt 4 fs' % (args.workload, args.reps, n))
    if args.workload == 'c3':
        return ('c3: emim/B(CN)4 ionic liquid x%d^3, %d atoms, DampedSmoothedForce + NonbondedExceptionsForce, '
                'RESPA [4,1] + Bussi, dt 2 fs' % (args.reps, n))
    return 'c4: AFED, methane in water, %d atoms, soft-core lambda_vdw extended variable, outer step 4 x 1 fs' % n


def workload_config(args, n, dt_fs=DT_FS):
    """Identical for both arms: names the workload, nothing about how it is executed."""
    return dict(workload=workload_text(args, n), atoms=n, dt_fs=dt_fs,
                loops=LOOPS if args.workload in ('c2', 'c5') else ([4, 1] if args.workload == 'c3' else None),
                l2_policy='no flush needed: a bench step is consecutive MD steps of ONE trajectory (positions, lists and '
                          'forces change every step), and at the default workload the per-step working set (state '
                          '0.4 GB + neighbour lists ~3 GB) is 25x the 126 MB L2')


def run_reference(args, rank, world):
    """CPU arm: the float64 C/OpenMP restatement of the reference algorithm (oracle/c/oracle.c; OpenMM
    itself is not installable offline) on all host cores.  Each step is a bounded
    sample of the same workload: `--cpu-md-steps` outer MD steps (default 1)."""
    if rank != 0:
        return
    import numpy as np
    from oracle import cport
    if args.workload not in ('c2', 'c5'):
        print(json.dumps(dict(impl='reference', unavailable='the CPU restatement drives RESPA [n0,n1,1] + Nose-Hoover '
                                                            'only (workloads c2, c5)')))
        return
    system, pos, vel = build_workload(args.reps)
    n = system.getNumParticles()
    integrator, dof = make_integrator(system)
    g = dict(zip([integrator.getGlobalVariableName(k) for k in range(integrator.getNumGlobalVariables())],
                 integrator._global_values))
    cores = os.cpu_count()
    port = cport.CPort(system, threads=cores, verify=False)
    sample = max(1, args.cpu_md_steps)
    x, v, p_eta = pos.copy(), vel.copy(), 0.0
    for _ in range(args.warmup):
        x, v, p_eta = port.respa(x, v, 1, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], p_eta))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        x, v, p_eta = port.respa(x, v, sample, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], p_eta))
    elapsed = time.perf_counter() - t0
    value = n*sample*args.steps/elapsed
    line = dict(metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3*elapsed/args.steps, higher_is_better=True,
                scaling='strong' if (args.gpus > 1 and not args.replicas) else 'weak',
                vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=workload_config(args, n), md_steps_per_step=sample,
                ns_per_day=sample*args.steps*DT_FS*1e-6*86400/elapsed,
                cpu_baseline=dict(value=value, unit='atom-steps/s', cores=cores, kind='port',
                                  sample='%d outer MD step(s) per step of the %d-atom workload (throughput is per '
                                         'atom-step, so the sample length does not bias it), C/OpenMP float64 '
                                         'restatement of the reference algorithm; NOT OpenMM-CPU, which is not '
                                         'installable offline' % (sample, n)),
                e2e=dict(value=value, unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def parity_gate(args, context, system, pos, rank):
    """First-frame parity against the oracle (SURVEY 8d "parity gates run with every benchmark").  Every
    rank takes part in the engine calls (collective under domain decomposition); rank 0 runs the oracle."""
    import numpy as np
    from atomsmm_b200 import mm
    groups = sorted(set(f.getForceGroup() for f in system.getForces() if not isinstance(f, mm.CMMotionRemover)))
    groups = [g for g in groups if g != 31]          # RESPASystem's group 31 is a report-only duplicate
    engine = {}
    for g in groups:
        state = context.getState(getForces=True, getEnergy=True, groups={g})
        engine[g] = (state._forces, state._potential)
    pair_forces = [f for f in system.getForces()
                   if isinstance(f, (mm.CustomNonbondedForce, mm.NonbondedForce)) and f.getForceGroup() != 31
                   and f.getNumParticles() > 0]
    sets = {}
    if context._nranks == 1:
        for f in pair_forces:
            sets[f.getForceGroup()] = context.pair_set(f)
    if rank != 0:
        return None
    from oracle import cport
    t0 = time.perf_counter()
    port = cport.CPort(system, threads=os.cpu_count(), verify=system.getNumParticles() <= 200000)
    result = dict(oracle='oracle/cport.py: float64 C/OpenMP restatement, pinned on the reference goldens '
                         '(tests/test_oracle_c.py, tests/test_oracle_goldens.py)',
                  atoms=system.getNumParticles(), force_rel_rms={}, energy_rel={}, pair_sets={},
                  tolerance=dict(force_rel_rms=1e-5, energy_rel=1e-6, pair_sets='equal count and checksum'))
    ok = True
    for g in groups:
        f_ref, e_ref, _ = port.evaluate(pos, groups={g})
        f, e = engine[g]
        rms = float(np.sqrt(np.sum((f - f_ref)**2)/max(np.sum(f_ref**2), 1e-300)))
        rel = abs(e - e_ref)/max(abs(e_ref), 1e-300)
        result['force_rel_rms'][str(g)] = rms
        result['energy_rel'][str(g)] = rel
        ok = ok and rms <= 1e-5 and rel <= 1e-6 and bool(np.isfinite(rms))
    for f in pair_forces:
        g = f.getForceGroup()
        if g in sets:
            expect = port.pair_set(pos, f)
            same = tuple(sets[g]) == tuple(expect)
            result['pair_sets'][str(g)] = dict(pairs=int(expect[0]), equal=bool(same))
            ok = ok and same
    if not sets:
        result['pair_sets'] = 'single-GPU diagnostic (checked at N=1 and in tests/test_gpu_scale.py)'
    result['ok'] = bool(ok)
    result['seconds'] = round(time.perf_counter() - t0, 1)
    return result


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200')
    parser.add_argument('--reps', type=int, default=None, help='replication of the 1 536-atom cell per axis')
    parser.add_argument('--md-steps', type=int, default=MD_STEPS_PER_CALL)
    parser.add_argument('--cpu-md-steps', type=int, default=1)
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-parity', action='store_true', help='skip the first-frame parity gate (profiling runs)')
    parser.add_argument('--workload', default='c5', choices=['c2', 'c3', 'c4', 'c5'])
    parser.add_argument('--replicas', action='store_true', help='N independent replicas instead of one decomposed system')
    parser.add_argument('--dd', action='store_true', help='(default for N > 1; kept for older command lines)')
    parser.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg')
    parser.add_argument('--no-profile', action='store_true', help='skip the per-launch pair-kernel timing pass')
    args = parser.parse_args()
    if args.reps is None:
        args.reps = {'c5': 14, 'c2': 4, 'c3': 4, 'c4': 1}[args.workload]
    if args.workload == 'c4':
        args.replicas = True            # AFED shards as independent replicas
    args.warmup = max(args.warmup, 3) if args.impl != 'reference' else max(args.warmup, 1)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from atomsmm_b200 import mm, unit
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU path')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    dt_fs = DT_FS
    if args.workload == 'c3':
        system, pos, vel, integrator = build_c3(args.reps)
        dt_fs = 2.0
    elif args.workload == 'c4':
        system, pos, vel, integrator = build_c4()
    else:
        system, pos, vel = build_workload(args.reps)
        integrator, dof = make_integrator(system)
    n = system.getNumParticles()
    dd = world > 1 and not args.replicas
    integrator.setRandomNumberSeed(1 if dd else 1 + rank)
    properties = {'DeviceIndex': local}
    if os.environ.get('B2_SKIN'):
        properties['Skin'] = float(os.environ['B2_SKIN'])       # neighbour-list skin sweep
    if dd:
        properties['DomainDecomposition'] = 'true'
    context = mm.Context(system, integrator, mm.Platform.getPlatformByName('B200'), properties)
    replicas = 1 if dd else world
    context.setPositions(pos)
    context.setVelocities(vel)
    md = args.md_steps
    stream = context._stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if not args.no_parity and args.workload in ('c2', 'c5'):
        parity = parity_gate(args, context, system, pos, rank)
        barrier()

    # ---- resident-state timing ---------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()               # nvidia-smi needs a moment to come up: start it before the warm-up
    for _ in range(args.warmup):
        integrator.step(md)
    context.synchronize()
    before = context.counters()
    barrier()
    sampler.mark_begin()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        start.record(stream)
        for _ in range(args.steps):
            integrator.step(md)
        stop.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    context.synchronize()
    elapsed = start.elapsed_time(stop)*1e-3
    after = context.counters()
    launches = (after['launches'] - before['launches']) + \
        (after['graph_launches'] - before['graph_launches'])*after['kernels_per_step']
    t = torch.tensor([elapsed], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed = float(t.item())
    value = replicas*n*md*args.steps/elapsed
    comm = context.comm_info()
    if dd:
        # where the exchanges spend their time, per rank: microseconds per exchange waiting for the slowest
        # peer's post, copying the halo out of peer memory, and waiting for the readers' acknowledgements
        clock = comm.get('exchange_clock') or {}
        k = max(1, clock.get('exchanges', 0))
        mine = dict(rank=rank, wait_for_posts_us=round(1e6*clock.get('wait_for_posts_s', 0.0)/k, 1),
                    halo_copy_us=round(1e6*clock.get('halo_copy_s', 0.0)/k, 1),
                    wait_for_acks_us=round(1e6*clock.get('wait_for_acks_s', 0.0)/k, 1), exchanges=clock.get('exchanges', 0))
        every = [None]*world
        dist.all_gather_object(every, mine)
        comm['exchange_clock_by_rank'] = every

    # ---- end to end through the public API with host buffers --------------------------------------
    host_x = torch.from_numpy(pos.copy()).pin_memory()
    host_v = torch.from_numpy(vel.copy()).pin_memory()
    state = context.getState(getPositions=True, getVelocities=True)
    host_x.copy_(torch.from_numpy(state._positions))
    host_v.copy_(torch.from_numpy(state._velocities))
    e2e_steps = max(3, args.steps//2)

    def e2e_step():
        context.setPositions(host_x)
        context.setVelocities(host_v)
        integrator.step(md)
        s = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
        host_x.copy_(torch.from_numpy(s._positions))
        host_v.copy_(torch.from_numpy(s._velocities))
        return s._potential + s._kinetic
    energy = None
    e2e_value = None
    if not args.no_e2e:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        laps = []
        for _ in range(e2e_steps):
            lap = time.perf_counter()
            energy = e2e_step()
            laps.append(round(1e3*(time.perf_counter() - lap), 1))
        barrier()
        if os.environ.get('B2_BENCH_VERBOSE'):
            sys.stderr.write('e2e laps (ms): %s\n' % laps)
        e2e_elapsed = time.perf_counter() - t0
        t = torch.tensor([e2e_elapsed], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = replicas*n*md*e2e_steps/float(t.item())
        if energy is None or not np.isfinite(energy):
            raise SystemExit('bench.py: the trajectory produced a non-finite energy')
    state_bytes = 2*n*3*8

    # ---- roofline of the dominant kernel: CUDA events around every pair launch, eager pass --------
    roofline = None
    if not args.no_profile:
        context.set_profiling(True)
        integrator.step(8)
        profile = context.pair_profile()
        phases = context.phase_profile()
        all_phases = [phases]
        if world > 1:
            all_phases = [None]*world
            dist.all_gather_object(all_phases, phases)
        context.set_profiling(False)
        peak, peak_kind = measured_peak()
        used = [p for p in profile if p['launches'] > 0]
        dominant = max(used, key=lambda p: p['total_ms'])
        avg_s = dominant['total_ms']*1e-3/dominant['launches']
        owned = comm['hi'] - comm['lo']
        bytes_per_launch = owned*(24 + 16 + 16) + 4*dominant['entries']
        achieved = bytes_per_launch/avg_s/1e9
        pair_ms_per_md_step = sum(p['total_ms'] for p in used)/8.0
        traffic, traffic_source = None, None
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as handle:
                entry = json.load(handle).get('%s/%s' % (args.workload, dominant['name']))
                if entry and world == 1:
                    traffic, traffic_source = entry['dram_bytes_per_launch'], entry['source']
        except (OSError, ValueError, KeyError):
            pass
        step_bytes = 1176.0 if args.workload in ('c2', 'c5') else None      # [4,2,1] + SY3-NH
        # throughput of every pair kernel in list slots (8 per entry) and, where the first-frame
        # gate counted the interacting pairs, in useful pair evaluations (every pair is evaluated in both tiles),
        # beside what one SM sub-partition issuing one warp instruction per cycle would allow per slot
        pair_sets = (parity or {}).get('pair_sets') if isinstance(parity, dict) else None
        pair_kernels = []
        for p in used:
            t = p['total_ms']*1e-3/p['launches']
            rec = dict(kernel=p['name'], group=p['group'], launches=p['launches'], avg_launch_us=round(t*1e6, 1),
                       list_entries=p['entries'], slots_per_s=8.0*p['entries']/t)
            rec['issue_cycles_per_warp_step'] = round(148*4*clocks['sm_mhz']*1e6*t/(8.0*p['entries']/32.0), 1) \
                if clocks and clocks.get('sm_mhz') else None
            if isinstance(pair_sets, dict) and world == 1:
                # the near list serves group 1, the far list group 2 (RESPASystem layout)
                counted = pair_sets.get(str(p['group']))
                if counted:
                    rec['useful_pair_evaluations_per_s'] = 2.0*counted['pairs']/t
                    rec['slots_inside_cutoff'] = round(2.0*counted['pairs']/(8.0*p['entries']), 3)
            pair_kernels.append(rec)
        roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved/peak, traffic=traffic,
                        traffic_source=traffic_source, peak_kind=peak_kind,
                        kernel='k_pair_force<%s> group %d' % (dominant['name'], dominant['group']),
                        avg_launch_us=avg_s*1e6, algorithmic_bytes_per_launch=bytes_per_launch,
                        list_entries=dominant['entries'], pair_kernels=pair_kernels,
                        phases_ms_per_md_step={k: round(v/8.0, 4) for k, v in phases.items()},
                        phases_ms_per_md_step_by_rank=([{k: round(v/8.0, 4) for k, v in p.items()} for p in all_phases]
                                                       if world > 1 else None),
                        pair_kernels_share_of_step=pair_ms_per_md_step/(1e3*elapsed/(args.steps*md)),
                        whole_step=(dict(bytes_per_atom_step=step_bytes, achieved=step_bytes*value/world/1e9,
                                         frac=step_bytes*value/world/1e9/peak) if step_bytes else None),
                        note='pair tiles are fp32-issue bound, not HBM bound (SURVEY 8d): algorithmic bytes = owned atoms*(24 B x '
                             '+ 16 B params + 16 B force) + 4 B per neighbour-list entry; timed in a separate eager pass '
                             'of the same step program with CUDA events on the launch stream; whole_step = SURVEY 8d '
                             'compulsory bytes per atom-step x measured rate, per GPU')

    line = dict(metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3*elapsed/args.steps, higher_is_better=True,
                scaling='strong' if dd else 'weak',
                vs_baseline=None, dtype='f32 pair forces / f64 state', data='synthetic',
                config=workload_config(args, n, dt_fs), md_steps_per_step=md,
                parallelism=(('domain decomposition over %d ranks (%s)' % (world, comm.get('exchange', 'nccl'))) if dd
                             else ('%d independent replicas' % world)),
                ns_per_day=md*args.steps*dt_fs*1e-6*86400/elapsed,
                clocks=clocks, gpu_launches=int(launches), parity=parity,
                e2e=(dict(value=e2e_value, unit='atom-steps/s', h2d_bytes_per_step=state_bytes,
                          d2h_bytes_per_step=state_bytes + 16, final_energy=energy) if e2e_value is not None else None),
                roofline=roofline, engine=dict(kernels_per_md_step=after['kernels_per_step'],
                                               list_rebuilds=after['rebuilds'], list_capacity=after['list_capacity'],
                                               list_stats=context.list_stats(), comm=comm))
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ('c2', 'c5'):
        from oracle import cport
        cores = os.cpu_count()
        port = cport.CPort(system, threads=cores, verify=False)
        g = dict(zip([integrator.getGlobalVariableName(k) for k in range(integrator.getNumGlobalVariables())],
                     integrator._global_values))
        x, v, p_eta = port.respa(pos, vel, 1, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], 0.0))
        sample = 2 if n > 1000000 else 4
        t0 = time.perf_counter()
        port.respa(x, v, sample, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], p_eta))
        cpu_elapsed = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=n*sample/cpu_elapsed, unit='atom-steps/s', cores=cores, kind='port',
                                    sample='%d outer MD steps of the same %d-atom workload (after 1 warm-up step), '
                                           'C/OpenMP float64 restatement of the reference algorithm (not OpenMM-CPU)'
                                           % (sample, n))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity['ok']:
        raise SystemExit('bench.py: first-frame parity gate FAILED: %s' % json.dumps(parity))


if __name__ == '__main__':
    main()
