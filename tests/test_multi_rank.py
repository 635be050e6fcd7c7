"""
The N > 1 path.

CPU (gloo, world_size 2): the host-side logic of the spatial decomposition -- ownership ranges from
the C library's pure-host entry point, the communicator-id broadcast, the max-over-ranks /
sum-over-ranks aggregation the benchmark uses -- and the exchange protocol of csrc/dist.cu itself
(owned-range updates, position exchange before pair forces, all-reduced sums) executed on the lowered
RESPA + Nose-Hoover program by tests/lowered_executor.py.

GPU (NCCL, world_size 2, skipped on a single-GPU box): ONE RESPA water system integrated by two
ranks with domain decomposition must reproduce the single-GPU trajectory and single-point
forces/energies.  The reference is single-process, so the single-GPU engine -- itself
checked against the oracle in test_gpu_*.py -- is the comparison.
"""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _init(rank, world, port, backend):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world)
    return dist


def _sorted_molecules(seed, nmol):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 20, size=nmol)
    return np.repeat(np.arange(nmol), sizes)


def _host_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    from atomsmm_b200 import engine
    dist = _init(rank, world, port, 'gloo')
    try:
        # the id broadcast used by Context._join_world
        payload = bytes(range(128)) if rank == 0 else None
        assert engine.broadcast_bytes(payload) == bytes(range(128))
        # every rank derives the same ownership ranges; together they tile [0, n) in whole molecules
        mol = _sorted_molecules(7, 500)
        ranges = engine.partition_ranges(mol, world)
        gathered = [None]*world
        dist.all_gather_object(gathered, ranges.tolist())
        assert all(g == gathered[0] for g in gathered)
        lo, hi = int(ranges[rank]), int(ranges[rank + 1])
        owned = torch.zeros(len(mol), dtype=torch.int64)
        owned[lo:hi] = 1
        dist.all_reduce(owned)
        assert bool((owned == 1).all())
        for cut in ranges[1:-1]:
            assert mol[cut] != mol[cut - 1]
        # aggregation used by bench.py: time = max over ranks, work = sum over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == float(world)
        out.put((rank, 'ok'))
    except Exception as error:    # pragma: no cover
        out.put((rank, repr(error)))
    finally:
        dist.destroy_process_group()


def _protocol_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist = _init(rank, world, port, 'gloo')
    try:
        import atomsmm_b200 as atomsmm
        from atomsmm_b200 import unit
        from lowered_executor import DistributedExecutor, Executor
        import test_lowered_programs as T
        system, pos = T.water_cluster(40)
        respa = atomsmm.RESPASystem(system, 7*unit.angstroms, 5*unit.angstroms)
        dof = atomsmm.countDegreesOfFreedom(respa)
        vel = T.velocities(respa, 9)

        def factory():
            nh = atomsmm.NoseHooverPropagator(300*T.K, dof, 100*T.fs)
            return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                   atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*T.fs)
        mask = 0b111 | (1 << 31)
        shared = DistributedExecutor(respa, factory(), pos, vel, dist, pair_mask=0b110 | (1 << 31), group_mask=mask)
        single = Executor(respa, factory(), pos, vel, group_mask=mask)
        shared.step(2)
        single.step(2)
        x, v = shared.full_state()
        assert 0 < shared.hi - shared.lo < len(pos)
        assert np.max(np.abs(x - single.x)) < 1e-12 and np.max(np.abs(v - single.v)) < 1e-10
        # two exchanges per RESPA[4,2,1] step: before the mid-step near force and before the end-of-step
        # near + far forces (which share one)
        assert shared.exchanges == 2*2
        names = shared.program.global_names
        assert shared.globals[names.index('p_eta')] == pytest.approx(single.globals[names.index('p_eta')], rel=1e-10)
        out.put((rank, 'ok'))
    except Exception as error:    # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc() + repr(error)))
    finally:
        dist.destroy_process_group()


def _halo_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    dist = _init(rank, world, port, 'gloo')
    try:
        import atomsmm_b200 as atomsmm
        from atomsmm_b200 import mm, unit
        from lowered_executor import Executor, HaloExecutor
        import systems
        import test_lowered_programs as T
        # the 1 536-atom water box doubled along x (5 x 2.5 x 2.5 nm): each rank owns one copy, and with
        # a list radius of 1.1 nm a slab of the other copy is NOT in its halo and goes stale
        respa, pdb = systems.respa_water()
        base = systems.positions_of(pdb)
        n = respa.getNumParticles()
        double = mm.System()
        double._masses = list(respa._masses)*2
        double.setDefaultPeriodicBoxVectors(mm.Vec3(5.0, 0, 0), mm.Vec3(0, 2.5, 0), mm.Vec3(0, 0, 2.5))
        import copy
        for force in respa.getForces():
            new = copy.deepcopy(force)
            for name, nindex in (('_particles', 0), ('_exceptions', 2), ('_exclusions', 2), ('_bonds', 2), ('_angles', 3),
                                 ('_torsions', 4)):
                rows = getattr(force, name, None)
                if rows is None or len(rows) == 0:
                    continue
                nested = isinstance(force, (mm.CustomBondForce, mm.CustomAngleForce))
                idx = mm.index_columns(rows, nindex).astype(np.int64)
                val = mm.value_columns(rows, nindex)
                idx2 = np.concatenate([idx, idx + n]) if nindex else None
                setattr(new, name, mm.PackedRows(idx2, np.concatenate([val, val]), nested=nested))
            double.addForce(new)
        rng = np.random.default_rng(3)
        pos = np.concatenate([base, base + np.array([2.5, 0, 0])]) + rng.uniform(-0.002, 0.002, size=(2*n, 3))
        mass = np.array(double._masses)
        vel = rng.standard_normal((2*n, 3))*np.sqrt(8.314472471220217e-3*300/mass)[:, None]
        dof = atomsmm.countDegreesOfFreedom(double)

        def factory():
            nh = atomsmm.NoseHooverPropagator(300*T.K, dof, 100*T.fs)
            return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([2, 2, 1]),
                                                   atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*T.fs)
        mask = 0b111
        # a tight skin so that the three steps contain both kinds of exchange (halo-only and rebuild)
        shared = HaloExecutor(double, factory(), pos, vel, dist, pair_mask=0b110, group_mask=mask, rlist=1.04, skin=0.04)
        single = Executor(double, factory(), pos, vel, group_mask=mask)
        shared.step(3)
        single.step(3)
        x, v = shared.full_state()
        assert shared.hi - shared.lo == n
        assert np.max(np.abs(x - single.x)) < 1e-12 and np.max(np.abs(v - single.v)) < 1e-10
        assert shared.exchanges == 2*3 and 1 <= shared.rebuilds < shared.exchanges, (shared.exchanges, shared.rebuilds)
        assert all(0 < h < n for h in shared.halo_sizes), shared.halo_sizes      # a strict subset of the other copy
        out.put((rank, 'ok'))
    except Exception as error:    # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc() + repr(error)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_protocol_world_size_2_gloo():
    """The peer-memory protocol (owned-only skin test, OR-ed verdict, halo-only refresh between rebuilds, stale
    positions everywhere else) on two gloo ranks reproduces the single-process trajectory."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results


def test_domain_decomposition_protocol_world_size_2_gloo():
    """The exchange / ownership / all-reduce protocol of csrc/dist.cu, executed on the CPU by two gloo
    ranks on the lowered RESPA + Nose-Hoover program, reproduces the single-process result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_protocol_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results


def test_host_logic_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_host_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results


@pytest.mark.parametrize('nranks', [1, 2, 3, 8])
def test_partition_properties(nranks):
    sys.path.insert(0, ROOT)
    from atomsmm_b200 import engine
    mol = _sorted_molecules(3, 1000)
    n = len(mol)
    ranges = engine.partition_ranges(mol, nranks)
    assert ranges[0] == 0 and ranges[-1] == n
    assert np.all(np.diff(ranges) >= 0)
    for k in range(1, nranks):
        cut = ranges[k]
        assert cut == 0 or cut == n or mol[cut] != mol[cut - 1]
        assert abs(cut - n*k//nranks) < 20      # never farther than one molecule from the even split
    # degenerate: one giant molecule cannot be split
    ranges = engine.partition_ranges(np.zeros(100, dtype=np.int32), 4)
    assert set(ranges.tolist()) <= {0, 100}


def _dd_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    torch.cuda.set_device(rank)
    dist = _init(rank, world, port, 'nccl')
    try:
        import atomsmm_b200 as atomsmm
        from atomsmm_b200 import mm, unit
        import systems
        fs, K = unit.femtoseconds, unit.kelvin
        respa, pdb = systems.respa_water()
        pos = systems.positions_of(pdb)
        big, bigpos = systems.replicate(respa, pos, np.array([2.5, 2.5, 2.5]), 2)
        n = big.getNumParticles()
        mass = np.array([big.getParticleMass(i).value_in_md_units() for i in range(n)])
        rng = np.random.default_rng(5)
        vel = rng.standard_normal((n, 3))*np.sqrt(8.314472471220217e-3*300/mass)[:, None]

        def factory():
            dof = atomsmm.countDegreesOfFreedom(big)
            nh = atomsmm.NoseHooverPropagator(300*K, dof, 100*fs)
            return atomsmm.TrotterSuzukiPropagator(atomsmm.RespaPropagator([4, 2, 1]),
                                                   atomsmm.SuzukiYoshidaPropagator(nh, 3)).integrator(4*fs)

        def run(properties, steps):
            integrator = factory()
            context = mm.Context(big, integrator, mm.Platform.getPlatformByName('B200'), properties)
            context.setPositions(bigpos)
            context.setVelocities(vel)
            first = context.getState(getForces=True, getEnergy=True)
            integrator.step(steps)
            last = context.getState(getPositions=True, getVelocities=True, getEnergy=True)
            return context, first, last

        single, f1, s1 = run({'DeviceIndex': rank}, 12)
        shared, f2, s2 = run({'DeviceIndex': rank, 'DomainDecomposition': 'true'}, 12)
        info = shared.comm_info()
        assert info['nranks'] == world and info['hi'] > info['lo'] and info['exchanges'] > 0
        fa, fb = f1._forces, f2._forces
        assert np.sqrt(np.sum((fa - fb)**2)/np.sum(fa**2)) < 1e-6
        assert abs(f1._potential - f2._potential) <= 1e-9*abs(f1._potential)
        assert np.max(np.abs(s1._positions - s2._positions)) < 1e-6
        assert np.sqrt(np.sum((s1._velocities - s2._velocities)**2)/np.sum(s1._velocities**2)) < 1e-5
        assert abs(s1._potential - s2._potential) <= 1e-6*abs(s1._potential)
        assert abs(s1._kinetic - s2._kinetic) <= 1e-6*abs(s1._kinetic)
        out.put((rank, 'ok'))
    except Exception as error:
        import traceback
        out.put((rank, traceback.format_exc() + repr(error)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_domain_decomposition_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dd_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results
