#!/usr/bin/env python
"""
Benchmark of the hot path (BASELINE.json metric: atom-steps/s and ns/day; pair-kernel HBM GB/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2]

Workload (config.workload): BASELINE config 2 -- the q-SPC-FW water box replicated 4x4x4
(98 304 atoms, L = 10 nm, +-0.002 nm jitter, seed 1), RESPASystem (near force-switch 0.7/0.5 nm in
group 1, full LJ + reaction-field Coulomb rc 1.0 nm in group 2, bonded terms in group 0),
integrator TrotterSuzuki(Respa([4,2,1]), SuzukiYoshida(NoseHoover(300 K, dof, 100 fs), 3)) at 4 fs.

A bench "step" is one call integrator.step(MD_STEPS_PER_CALL) (default 100 outer MD steps): one
pass of the hot path over one batch.  `value` times K such calls with state resident in HBM;
`e2e` times the same through the public API with HOST buffers: upload of positions+velocities from
pinned host memory, the MD steps, download of positions+velocities+energies, every step.

N > 1 (default): N independent replicas of the workload, one per GPU, no data-path collective
("scaling": "weak") -- ensembles are how configs 1-4 shard.
N > 1 with --dd: ONE system integrated by all ranks with spatial domain decomposition (NCCL position
exchange before every pair-force evaluation); "scaling": "strong".  `--workload c5` selects BASELINE
config 5 (the cell replicated 14x14x14 = 4 214 784 atoms, L = 35 nm), the configuration the
decomposition is meant for:
    torchrun --nproc-per-node 8 bench.py --gpus 8 --workload c5 --dd --steps 3 --md-steps 20
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
    mass = np.array([big.getParticleMass(i).value_in_md_units() for i in range(n)])
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
    mass = np.array([big.getParticleMass(i).value_in_md_units() for i in range(n)])
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
    mass = np.array([alchemical.getParticleMass(i).value_in_md_units() for i in range(n)])
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


def run_reference(args, rank, world):
    """CPU arm: the float64 C/OpenMP restatement of the reference algorithm (oracle/c/oracle.c) on
    all host cores; each step is a bounded sample (a few outer MD steps) of the same workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import cport
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
                warmup=args.warmup, ms_per_step=1e3*elapsed/args.steps, higher_is_better=True, scaling='weak',
                vs_baseline=None, dtype='f64', data='synthetic', impl='reference',
                config=workload_config(args, n, sample),
                ns_per_day=sample*args.steps*DT_FS*1e-6*86400/elapsed,
                cpu_baseline=dict(value=value, unit='atom-steps/s', cores=cores, kind='port',
                                  sample='%d outer MD steps per step of the %d-atom workload, C/OpenMP float64 '
                                         'restatement of the reference algorithm (OpenMM itself is not installable '
                                         'offline)' % (sample, n)),
                e2e=dict(value=value, unit='atom-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(args, n, md_steps, what=None):
    dd = getattr(args, 'dd', False) and args.gpus > 1
    text = what % n if what else ('%s: q-SPC-FW water x%d^3, %d atoms, RESPASystem near 0.7/0.5 nm force-switch + '
                                  'LJ/reaction-field 1.0 nm, RESPA [4,2,1] + NoseHoover SY3, dt 4 fs'
                                  % (getattr(args, 'workload', 'c2'), args.reps, n))
    return dict(workload=text,
                atoms=n, md_steps_per_step=md_steps, dt_fs=DT_FS, loops=LOOPS,
                replicas=1 if dd else args.gpus,
                parallelism=('domain decomposition over %d ranks' % args.gpus) if dd else
                            ('%d independent replicas' % args.gpus),
                l2_policy='no flush: a bench step is %d CONSECUTIVE MD steps of one trajectory (positions, lists and '
                          'forces change every step), not a repeated identical input; at 98 304 atoms the working set '
                          '(state ~10 MB + neighbour lists ~65 MB) is L2-resident exactly as in a production run of this '
                          'size; the same engine on a working set far beyond L2 (--workload c5, 4.2 M atoms, ~3 GB of '
                          'lists) runs at a HIGHER per-atom rate (profiles/round1_c5_n1.json), so the number does not '
                          'come from cache warmth' % md_steps)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=10)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='b200')
    parser.add_argument('--reps', type=int, default=4, help='replication of the 1 536-atom cell per axis')
    parser.add_argument('--md-steps', type=int, default=MD_STEPS_PER_CALL)
    parser.add_argument('--cpu-md-steps', type=int, default=2)
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--workload', default='c2', choices=['c2', 'c3', 'c4', 'c5'])
    parser.add_argument('--dd', action='store_true', help='one system over all ranks (domain decomposition)')
    parser.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg')
    args = parser.parse_args()
    if args.workload == 'c5':
        args.reps = 14
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

    dt_fs, what = DT_FS, None
    if args.workload == 'c3':
        system, pos, vel, integrator = build_c3(args.reps)
        dt_fs, what = 2.0, ('c3: emim/B(CN)4 ionic liquid x%d^3, %%d atoms, DampedSmoothedForce + NonbondedExceptionsForce, '
                            'RESPA [4,1] + Bussi, dt 2 fs' % args.reps)
    elif args.workload == 'c4':
        system, pos, vel, integrator = build_c4()
        dt_fs, what = 4.0, 'c4: AFED, methane in water, %d atoms, soft-core lambda_vdw extended variable, outer step 4 x 1 fs'
    else:
        system, pos, vel = build_workload(args.reps)
        integrator, dof = make_integrator(system)
    n = system.getNumParticles()
    dd = args.dd and world > 1
    integrator.setRandomNumberSeed(1 if dd else 1 + rank)
    properties = {'DeviceIndex': local}
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

    # ---- end to end through the public API with host buffers --------------------------------------
    host_x = torch.from_numpy(pos.copy()).pin_memory()
    host_v = torch.from_numpy(vel.copy()).pin_memory()
    state = context.getState(getPositions=True, getVelocities=True)
    host_x.copy_(torch.from_numpy(state.getPositions(asNumpy=True).value_in_unit(unit.nanometer)))
    host_v.copy_(torch.from_numpy(state.getVelocities(asNumpy=True).value_in_unit(unit.nanometer/unit.picosecond)))
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
    state_bytes = 2*n*3*8

    # ---- roofline of the dominant kernel: CUDA events around every pair launch, eager pass --------
    context.set_profiling(True)
    integrator.step(8)
    profile = context.pair_profile()
    context.set_profiling(False)
    peak, peak_kind = measured_peak()
    used = [p for p in profile if p['launches'] > 0]
    dominant = max(used, key=lambda p: p['total_ms'])
    avg_s = dominant['total_ms']*1e-3/dominant['launches']
    bytes_per_launch = n*(24 + 16 + 16) + 4*dominant['entries']
    achieved = bytes_per_launch/avg_s/1e9
    pair_ms_per_md_step = sum(p['total_ms'] for p in used)/8.0
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as handle:
            if args.workload == 'c2' and args.reps == 4:
                traffic = json.load(handle).get(dominant['name'])
    except (OSError, ValueError):
        pass
    roofline = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved/peak, traffic=traffic,
                    peak_kind=peak_kind, kernel='k_pair_force<%s> group %d' % (dominant['name'], dominant['group']),
                    avg_launch_us=avg_s*1e6, algorithmic_bytes_per_launch=bytes_per_launch,
                    list_entries=dominant['entries'],
                    pair_kernels_share_of_step=pair_ms_per_md_step/(1e3*elapsed/(args.steps*md)),
                    note='pair tiles are fp32-issue bound, not HBM bound (SURVEY 8d): algorithmic bytes = N*(24 B x + '
                         '16 B params + 16 B force) + 4 B per neighbour-list entry; timed in a separate eager pass '
                         'of the same step program with CUDA events on the launch stream')

    line = dict(metric='atom-steps/s', value=value, unit='atom-steps/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3*elapsed/args.steps, higher_is_better=True,
                scaling='strong' if dd else 'weak',
                vs_baseline=None, dtype='f32 pair forces / f64 state', data='synthetic',
                config=workload_config(args, n, md, what), ns_per_day=md*args.steps*dt_fs*1e-6*86400/elapsed,
                clocks=clocks, gpu_launches=int(launches),
                e2e=(dict(value=e2e_value, unit='atom-steps/s', h2d_bytes_per_step=state_bytes,
                          d2h_bytes_per_step=state_bytes + 16, final_energy=energy) if e2e_value is not None else None),
                roofline=roofline, engine=dict(kernels_per_md_step=after['kernels_per_step'],
                                               list_rebuilds=after['rebuilds'], list_capacity=after['list_capacity'],
                                               list_stats=context.list_stats()))
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ('c2', 'c5'):
        from oracle import cport
        cores = os.cpu_count()
        port = cport.CPort(system, threads=cores, verify=False)
        g = dict(zip([integrator.getGlobalVariableName(k) for k in range(integrator.getNumGlobalVariables())],
                     integrator._global_values))
        x, v, p_eta = port.respa(pos, vel, 1, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], 0.0))
        sample = 4
        t0 = time.perf_counter()
        port.respa(x, v, sample, DT_FS*1e-3, LOOPS[0], LOOPS[1], (1, g['LkT'], g['Q'], p_eta))
        cpu_elapsed = time.perf_counter() - t0
        line['cpu_baseline'] = dict(value=n*sample/cpu_elapsed, unit='atom-steps/s', cores=cores, kind='port',
                                    sample='%d outer MD steps of the same %d-atom workload (after 1 warm-up step), '
                                           'C/OpenMP float64 restatement of the reference algorithm' % (sample, n))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
