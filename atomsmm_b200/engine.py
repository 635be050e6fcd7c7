"""
``Context`` / ``State``: the execution side of the description layer, bound to the CUDA engine
through the ctypes C ABI (include/atomsmm_b200.h).

Mirrors the slice of ``openmm.Context`` / ``openmm.State`` that atomsmm and its tests use
(reference call sites: computers.py:67-88,242-246; utils.py:153-186,219-228;
tests/test_respa_forces.py:20-26).  State lives in torch CUDA tensors (double [N,3], caller's atom
order); the library works on device pointers taken from them.  There is NO CPU path: if the
shared library is missing or no CUDA device is present, construction raises.
"""

import ctypes
import math
import os

import numpy as np

from . import lowering
from . import mm
from . import unit
from .unit import md_value as _md

_LIB = None
# B2_LIBRARY selects another build of the same engine (kernel tuning experiments, scripts/build_variants.sh)
_LIB_PATH = os.environ.get('B2_LIBRARY') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libatomsmm_b200.so')

c_int_p = ctypes.POINTER(ctypes.c_int)
c_double_p = ctypes.POINTER(ctypes.c_double)
c_void = ctypes.c_void_p

_SIGNATURES = {
    'b2_create': [ctypes.c_int, ctypes.POINTER(c_void)],
    'b2_destroy': [c_void],
    'b2_set_stream': [c_void, c_void],
    'b2_synchronize': [c_void],
    'b2_set_box': [c_void, c_double_p, ctypes.c_int],
    'b2_set_particles': [c_void, ctypes.c_int, c_double_p, c_int_p],
    'b2_add_param_set': [c_void, c_double_p, c_double_p, c_double_p, c_int_p],
    'b2_set_exclusions': [c_void, ctypes.c_int, c_int_p],
    'b2_add_pair_force': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_double_p,
                          ctypes.c_int, ctypes.c_double, c_int_p],
    'b2_update_pair_force': [c_void, ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_double],
    'b2_update_pair_particles': [c_void, ctypes.c_int, c_double_p, c_double_p, c_double_p],
    'b2_bind_pair_parameter': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_int],
    'b2_add_bonded_force': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, c_double_p, ctypes.c_int,
                            ctypes.c_int, c_double_p, ctypes.c_int, c_int_p],
    'b2_add_custom_bonded_force': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, c_double_p,
                                   ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int, c_int_p, ctypes.c_int,
                                   c_double_p, ctypes.c_int, c_int_p],
    'b2_add_pme': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                   ctypes.c_double, ctypes.c_double, c_int_p],
    'b2_set_skin': [c_void, ctypes.c_double],
    'b2_set_constraints': [c_void, ctypes.c_int, c_int_p, c_double_p, ctypes.c_double],
    'b2_set_positions': [c_void, c_void],
    'b2_set_velocities': [c_void, c_void],
    'b2_get_positions': [c_void, c_void],
    'b2_get_velocities': [c_void, c_void],
    'b2_eval': [c_void, ctypes.c_uint32, ctypes.c_int, c_void, c_double_p, c_double_p],
    'b2_get_group_energies': [c_void, c_double_p, c_double_p],
    'b2_get_parameter_derivatives': [c_void, c_double_p],
    'b2_pair_set': [c_void, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_ulonglong),
                    c_void, ctypes.c_longlong],
    'b2_load_program': [c_void, c_int_p, ctypes.c_int, c_int_p, ctypes.c_int, c_double_p, ctypes.c_int,
                        c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64],
    'b2_set_globals': [c_void, ctypes.c_int, ctypes.c_int, c_double_p],
    'b2_get_globals': [c_void, ctypes.c_int, ctypes.c_int, c_double_p],
    'b2_set_perdof': [c_void, ctypes.c_int, c_void],
    'b2_get_perdof': [c_void, ctypes.c_int, c_void],
    'b2_run': [c_void, ctypes.c_int],
    'b2_get_counters': [c_void, ctypes.POINTER(ctypes.c_longlong)],
    'b2_get_list_stats': [c_void, ctypes.POINTER(ctypes.c_longlong)],
    'b2_set_profiling': [c_void, ctypes.c_int],
    'b2_get_profile': [c_void, ctypes.c_int, c_double_p, ctypes.POINTER(ctypes.c_longlong),
                       ctypes.POINTER(ctypes.c_longlong)],
    'b2_set_barostat': [c_void, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_ulonglong],
    'b2_get_box': [c_void, c_double_p],
    'b2_update_box': [c_void, c_double_p],
    'b2_get_barostat_stats': [c_void, ctypes.POINTER(ctypes.c_longlong), c_double_p],
    'b2_barostat_uniform': [ctypes.c_ulonglong, ctypes.c_ulonglong, c_double_p],
    'b2_get_phase_profile': [c_void, c_double_p],
    'b2_comm_unique_id': [ctypes.c_char_p],
    'b2_comm_init': [c_void, ctypes.c_int, ctypes.c_int, ctypes.c_char_p],
    'b2_comm_export': [c_void, ctypes.c_char_p],
    'b2_comm_import': [c_void, ctypes.c_int, ctypes.c_char_p],
    'b2_comm_mode': [c_void, c_int_p, ctypes.POINTER(ctypes.c_longlong)],
    'b2_comm_timing': [c_void, c_double_p],
    'b2_get_order': [c_void, c_int_p],
    'b2_get_jit_stats': [c_void, ctypes.POINTER(ctypes.c_longlong)],
    'b2_set_jit': [c_void, ctypes.c_int],
    'b2_jit_check': [c_int_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int],
    'b2_kinetic_energy': [c_void, c_double_p],
    'b2_partition_ranges': [ctypes.c_int, c_int_p, ctypes.c_int, c_int_p],
    'b2_hilbert_index': [c_double_p, c_double_p, ctypes.POINTER(ctypes.c_ulonglong)],
    'b2_comm_info': [c_void, c_int_p, c_int_p, c_int_p, c_int_p, ctypes.POINTER(ctypes.c_longlong)],
}


class EngineError(mm.OpenMMException):
    pass


def library():
    """Load libatomsmm_b200.so (built in-tree by ``python -m atomsmm_b200.build``)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIB_PATH):
            raise EngineError('CUDA engine library not found at %s: run `python -m atomsmm_b200.build`. '
                              'There is no CPU fallback.' % _LIB_PATH)
        lib = ctypes.CDLL(_LIB_PATH)
        for name, argtypes in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = ctypes.c_int
        lib.b2_last_error.argtypes = [c_void]
        lib.b2_last_error.restype = ctypes.c_char_p
        lib.b2_version.restype = ctypes.c_char_p
        _LIB = lib
    return _LIB


def _dptr(array):
    return np.ascontiguousarray(array, dtype=np.float64).ctypes.data_as(c_double_p)


def _iptr(array):
    return np.ascontiguousarray(array, dtype=np.int32).ctypes.data_as(c_int_p)


def partition_ranges(molecule_sorted, nranks):
    """Ownership boundaries (length nranks+1) of the spatial decomposition: cuts fall on the
    molecule boundaries nearest to k*n/nranks.  Pure host logic of the C library (no GPU)."""
    mol = np.ascontiguousarray(molecule_sorted, dtype=np.int32)
    out = np.zeros(nranks + 1, dtype=np.int32)
    code = library().b2_partition_ranges(len(mol), _iptr(mol), int(nranks), out.ctypes.data_as(c_int_p))
    if code != 0:
        raise EngineError('b2_partition_ranges failed (code %d)' % code)
    return out


def hilbert_index(position, box):
    """Index of a point along the Hilbert curve the engine orders molecules by (pure host logic)."""
    key = ctypes.c_ulonglong()
    code = library().b2_hilbert_index(_dptr(np.asarray(position, dtype=np.float64)), _dptr(np.asarray(box, dtype=np.float64)),
                                      ctypes.byref(key))
    if code != 0:
        raise EngineError('b2_hilbert_index failed (code %d)' % code)
    return key.value


def broadcast_bytes(payload, src=0, group=None):
    """Broadcast a bytes object from rank ``src`` over torch.distributed (gloo or nccl)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def _molecules(system):
    """Connected components of the bond/constraint graph (Context.getMolecules),
    numbered by their lowest atom index.  -> (molecule id per atom, list of atom-index arrays)"""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    n = system.getNumParticles()
    edges = []
    for force in system.getForces():
        if isinstance(force, (mm.HarmonicBondForce, mm.CustomBondForce)):
            if len(force._bonds):
                edges.append(mm.index_columns(force._bonds, 2))
        elif isinstance(force, (mm.HarmonicAngleForce, mm.CustomAngleForce)):
            if len(force._angles):
                a = mm.index_columns(force._angles, 3)
                edges += [a[:, :2], a[:, 1:3]]
        elif isinstance(force, mm.PeriodicTorsionForce):
            if len(force._torsions):
                t = mm.index_columns(force._torsions, 4)
                edges += [t[:, :2], t[:, 1:3], t[:, 2:4]]
    if system._constraints:
        edges.append(np.array([[c[0], c[1]] for c in system._constraints], dtype=np.int32))
    if edges:
        e = np.concatenate(edges, axis=0).astype(np.int64)
        graph = coo_matrix((np.ones(len(e), dtype=np.int8), (e[:, 0], e[:, 1])), shape=(n, n))
        _, labels = connected_components(graph, directed=False)
    else:
        labels = np.arange(n)
    # number the components in order of their first atom
    _, first = np.unique(labels, return_index=True)
    rank = np.empty(len(first), dtype=np.int64)
    rank[np.argsort(first, kind='stable')] = np.arange(len(first))
    molecule = rank[labels].astype(np.int32)
    order = np.argsort(molecule, kind='stable')
    counts = np.bincount(molecule, minlength=len(first))
    groups = np.split(order, np.cumsum(counts)[:-1]) if n else []
    return molecule, groups


def _pair_keys(pairs, n):
    """Sorted unique int64 keys min*n+max of an [m, 2] index array (exclusion sets are compared and
    merged as arrays: config 5 has 4.2 M exclusions)."""
    if len(pairs) == 0:
        return np.zeros(0, dtype=np.int64)
    p = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    return np.unique(np.minimum(p[:, 0], p[:, 1])*n + np.maximum(p[:, 0], p[:, 1]))


class State(object):
    def __init__(self, **fields):
        self.__dict__.update(fields)

    def _vectors(self, array, asNumpy, u):
        if array is None:
            raise mm.OpenMMException('Invoked a State getter for data that was not requested')
        if asNumpy:
            return unit.Quantity(array.copy(), u)
        return unit.Quantity([mm.Vec3(*row) for row in array], u)

    def getPositions(self, asNumpy=False):
        return self._vectors(self._positions, asNumpy, unit.nanometer)

    def getVelocities(self, asNumpy=False):
        return self._vectors(self._velocities, asNumpy, unit.nanometer/unit.picosecond)

    def getForces(self, asNumpy=False):
        return self._vectors(self._forces, asNumpy, unit.kilojoule_per_mole/unit.nanometer)

    def getPotentialEnergy(self):
        if self._potential is None:
            raise mm.OpenMMException('Invoked getPotentialEnergy() on a State which does not contain energies')
        return self._potential*unit.kilojoule_per_mole

    def getKineticEnergy(self):
        if self._kinetic is None:
            raise mm.OpenMMException('Invoked getKineticEnergy() on a State which does not contain energies')
        return self._kinetic*unit.kilojoule_per_mole

    def getPeriodicBoxVectors(self, asNumpy=False):
        b = self._box
        vectors = [mm.Vec3(b[0], 0, 0), mm.Vec3(0, b[1], 0), mm.Vec3(0, 0, b[2])]
        if asNumpy:
            return unit.Quantity(np.array(vectors), unit.nanometer)
        return unit.Quantity(vectors, unit.nanometer)

    def getPeriodicBoxVolume(self):
        return float(np.prod(self._box))*unit.nanometer**3

    def getParameters(self):
        return dict(self._parameters)

    def getEnergyParameterDerivatives(self):
        return dict(self._derivatives)

    def getTime(self):
        return self._time*unit.picosecond


class Context(object):
    """Execution context on one B200.  ``properties``: 'DeviceIndex' (default 0 or LOCAL_RANK),
    'Skin' (neighbour-list skin in nm, default 0.15), 'FastPaths' ('false' routes every per-DOF step
    through the generic path), 'PerDofCompiler' ('false': generic per-DOF steps run on the device-side bytecode
    interpreter instead of as NVRTC-compiled kernels), 'Precision' ('mixed', the default: State forces are the fp32-tile forces the
    integrator uses; 'double': getState evaluates and accumulates every force contribution in float64 --
    report cadence only, time stepping is always mixed precision), 'DomainDecomposition' ('true': the ranks of the initialised
    torch.distributed world integrate ONE system together, each owning a spatial range of whole
    molecules; every rank must make the same calls with the same arguments)."""

    def __init__(self, system, integrator, platform=None, properties=None):
        import torch
        self._torch = torch
        properties = dict(properties or {})
        if platform is not None and platform.getName() not in mm.Platform._NAMES:
            raise mm.OpenMMException('unknown platform %s' % platform.getName())
        self._lib = library()
        if not torch.cuda.is_available():
            raise EngineError('no CUDA device is visible: the atomsmm_b200 engine has no CPU path')
        index = int(properties.get('DeviceIndex', os.environ.get('LOCAL_RANK', 0)))
        self._device = torch.device('cuda', index)
        self._system = system
        self._integrator = integrator
        self._n = n = system.getNumParticles()
        self._time = 0.0
        self._handle = c_void()
        self._check(self._lib.b2_create(index, ctypes.byref(self._handle)), None)
        self._stream = torch.cuda.Stream(device=self._device)
        self._call('b2_set_stream', c_void(self._stream.cuda_stream))
        self._skin = float(properties.get('Skin', 0.15))
        self._fast = str(properties.get('FastPaths', 'true')).lower() != 'false'
        precision = str(properties.get('Precision', 'mixed')).lower()
        if precision not in ('mixed', 'double'):
            raise mm.OpenMMException("Precision must be 'mixed' or 'double' on this platform")
        self._double_forces = precision == 'double'
        self._call('b2_set_skin', self._skin)
        if str(properties.get('PerDofCompiler', 'true')).lower() == 'false':
            self._call('b2_set_jit', 0)          # generic per-DOF steps stay on the bytecode interpreter
        box = _md(system.getDefaultPeriodicBoxVectors())
        self._box = np.array([box[0][0], box[1][1], box[2][2]], dtype=np.float64)
        self._periodic = system.usesPeriodicBoundaryConditions()
        self._parameters = {}
        for force in system.getForces():
            if hasattr(force, 'getNumGlobalParameters'):
                for k in range(force.getNumGlobalParameters()):
                    self._parameters.setdefault(force.getGlobalParameterName(k), force.getGlobalParameterDefaultValue(k))
        self._molecule, self._molecule_groups = _molecules(system)
        self._masses = np.array(system._masses, dtype=np.float64)
        self._x = torch.zeros((n, 3), dtype=torch.float64, device=self._device)
        self._buffer = torch.zeros((n, 3), dtype=torch.float64, device=self._device)
        self._have_positions = False
        self._pair_handles = {}        # id(force) -> (handle, info)
        self._describe()
        self._rank, self._nranks = 0, 1
        if str(properties.get('DomainDecomposition', 'false')).lower() == 'true':
            self._join_world()
        self._program = None
        if integrator is not None:
            if getattr(integrator, '_context', None) is not None:
                raise mm.OpenMMException('This Integrator is already bound to a context')
            integrator._context = self
            self._load_program()

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, code, handle):
        if code != 0:
            message = self._lib.b2_last_error(handle)
            raise EngineError('%s (code %d)' % (message.decode() if message else 'engine error', code))

    def _call(self, name, *args):
        self._check(getattr(self._lib, name)(self._handle, *args), self._handle)

    def __del__(self):
        try:
            if getattr(self, '_handle', None) is not None and self._handle.value:
                self._lib.b2_destroy(self._handle)
                self._handle = c_void()
        except Exception:
            pass

    def _join_world(self):
        import torch.distributed as dist
        torch = self._torch
        if not dist.is_initialized():
            raise EngineError('DomainDecomposition needs an initialised torch.distributed process group')
        self._rank, self._nranks = dist.get_rank(), dist.get_world_size()
        if self._nranks == 1:
            return
        torch.cuda.set_device(self._device)
        ident = ctypes.create_string_buffer(128)
        if self._rank == 0:
            self._check(self._lib.b2_comm_unique_id(ident), None)
        payload = broadcast_bytes(ident.raw if self._rank == 0 else None)
        self._call('b2_comm_init', self._nranks, self._rank, ctypes.create_string_buffer(payload, 128))
        # peer-memory halo exchange: every rank maps its peers' position arrays (cudaIpc over NVLink)
        self._exchange = 'nccl all-gather'
        if os.environ.get('B2_DD_EXCHANGE', 'p2p').lower() != 'nccl':
            record = ctypes.create_string_buffer(256)
            self._call('b2_comm_export', record)
            table = [None]*self._nranks
            dist.all_gather_object(table, record.raw)
            code = self._lib.b2_comm_import(self._handle, self._nranks, ctypes.create_string_buffer(b''.join(table), 256*self._nranks))
            # all ranks must agree on the mode: one failure sends everybody to the NCCL path
            verdicts = [None]*self._nranks
            dist.all_gather_object(verdicts, int(code))
            if all(v == 0 for v in verdicts):
                self._exchange = 'peer-memory halo pull over NVLink'
            else:
                import warnings
                message = self._lib.b2_last_error(self._handle) if code != 0 else None
                self._call('b2_comm_import', -1, None)
                warnings.warn('peer-memory exchange unavailable on some rank (%s, codes %r): using the NCCL all-gather path'
                              % (message.decode() if message else 'ok here', verdicts))

    def spatial_order(self):
        """Diagnostic: caller index of the atom at every position of the engine's spatial order."""
        out = np.empty(self._n, dtype=np.int32)
        self._call('b2_get_order', out.ctypes.data_as(c_int_p))
        return out

    def comm_info(self):
        rank, nranks, lo, hi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        exchanges = ctypes.c_longlong()
        self._call('b2_comm_info', ctypes.byref(rank), ctypes.byref(nranks), ctypes.byref(lo), ctypes.byref(hi),
                   ctypes.byref(exchanges))
        p2p, halo = ctypes.c_int(), ctypes.c_longlong()
        self._call('b2_comm_mode', ctypes.byref(p2p), ctypes.byref(halo))
        timing = (ctypes.c_double*4)()
        self._call('b2_comm_timing', timing)
        return dict(rank=rank.value, nranks=nranks.value, lo=lo.value, hi=hi.value, exchanges=exchanges.value,
                    exchange=getattr(self, '_exchange', 'none') if nranks.value > 1 else 'none',
                    halo_atoms=halo.value,
                    # device-side clock of the peer-memory exchange (seconds since the context was created)
                    exchange_clock=dict(wait_for_posts_s=round(timing[0], 6), halo_copy_s=round(timing[1], 6),
                                        exchanges=int(timing[2]), wait_for_acks_s=round(timing[3], 6)))

    # -- description -> C ABI ----------------------------------------------------------------------
    def _describe(self):
        system, n = self._system, self._n
        if system.getNumConstraints() > 0:
            self._has_constraints = True
        else:
            self._has_constraints = False
        self._call('b2_set_box', _dptr(self._box), 1 if self._periodic else 0)
        self._call('b2_set_particles', n, _dptr(self._masses), _iptr(self._molecule))
        exclusions = None
        self._all_groups = 0
        volume = float(np.prod(self._box))
        for force in system.getForces():
            group = force.getForceGroup()
            if isinstance(force, mm.CMMotionRemover):
                continue
            if isinstance(force, mm.MonteCarloBarostat):
                # bar -> kJ/mol/nm^3 (OpenMM: pressure*AVOGADRO*1e-25)
                pressure = force._pressure*6.02214179e23*1e-25
                kT = 8.314472471220217e-3*force._temperature
                self._call('b2_set_barostat', pressure, kT, force._frequency, ctypes.c_ulonglong(force._seed & 0xffffffffffffffff))
                self._barostat = force
                continue
            self._all_groups |= 1 << group
            if isinstance(force, mm.CustomNonbondedForce):
                if force.getNumParticles() != n:
                    raise mm.OpenMMException('CustomNonbondedForce must have exactly as many particles as the System')
                family, cutoff, params, info = lowering.classify_pair_force(force, self._parameters)
                table = self._custom_table(force, info)
                exclusions = self._merge_exclusions(exclusions, _pair_keys(mm.index_columns(force._exclusions, 2), n))
                set_id = self._param_set(table[:, 0], table[:, 1], table[:, 2])
                econst = 0.0
                if force.getUseLongRangeCorrection():
                    econst = self._custom_lrc(force, family, params, table, cutoff, volume)
                handle = ctypes.c_int()
                p = np.array(params, dtype=np.float64)
                self._call('b2_add_pair_force', family, group, set_id, cutoff, _dptr(p), len(p), econst,
                           ctypes.byref(handle))
                self._pair_handles[id(force)] = (handle.value, info, force)
            elif isinstance(force, mm.NonbondedForce):
                if force.getNumParticles() == 0:
                    continue
                if force.getNumParticles() != n:
                    raise mm.OpenMMException('NonbondedForce must have exactly as many particles as the System')
                exclusions = self._describe_nonbonded(force, exclusions, volume)
            elif isinstance(force, mm.HarmonicBondForce):
                if force.getNumBonds():
                    atoms = mm.index_columns(force._bonds, 2)
                    params = mm.value_columns(force._bonds, 2)
                    self._add_bonded(lowering.BOND_HARMONIC, group, atoms, params, force.usesPeriodicBoundaryConditions())
            elif isinstance(force, mm.HarmonicAngleForce):
                if force.getNumAngles():
                    atoms = mm.index_columns(force._angles, 3)
                    params = mm.value_columns(force._angles, 3)
                    self._add_bonded(lowering.ANGLE_HARMONIC, group, atoms, params, force.usesPeriodicBoundaryConditions())
            elif isinstance(force, mm.PeriodicTorsionForce):
                if force.getNumTorsions():
                    atoms = mm.index_columns(force._torsions, 4)
                    params = mm.value_columns(force._torsions, 4)
                    self._add_bonded(lowering.TORSION_PERIODIC, group, atoms, params, force.usesPeriodicBoundaryConditions())
            elif isinstance(force, mm.CustomBondForce):
                if force.getNumBonds():
                    atoms = mm.index_columns(force._bonds, 2)
                    params = mm.value_columns(force._bonds, 2).reshape(len(atoms), -1)
                    family, gparams, code = lowering.classify_bond_force(force, self._parameters)
                    periodic = force.usesPeriodicBoundaryConditions()
                    if family == lowering.BOND_LJC:
                        # exclusion-type exceptions (chargeprod = epsilon = 0) contribute exactly zero
                        live = (params[:, 0] != 0.0) | (params[:, 2] != 0.0)
                        atoms, params = atoms[live], params[live]
                        if len(atoms):
                            params = np.concatenate([params, np.zeros((len(atoms), 1))], axis=1)
                            self._add_bonded(family, group, atoms, params, periodic, gparams)
                    else:
                        self._add_custom(lowering.BOND_CUSTOM, group, atoms, params, periodic, code)
            elif isinstance(force, mm.CustomAngleForce):
                if force.getNumAngles():
                    from . import expr as X
                    atoms = mm.index_columns(force._angles, 3)
                    params = mm.value_columns(force._angles, 3).reshape(len(atoms), -1)
                    names = [force.getPerAngleParameterName(k) for k in range(force.getNumPerAngleParameters())]
                    globals_ = {force.getGlobalParameterName(k): self._parameters[force.getGlobalParameterName(k)]
                                for k in range(force.getNumGlobalParameters())}
                    code = lowering.compile_custom(X.parse_inlined(force.getEnergyFunction()), 'theta', names, globals_)
                    self._add_custom(lowering.ANGLE_CUSTOM, group, atoms, params, force.usesPeriodicBoundaryConditions(), code)
            else:
                raise lowering.UnsupportedDescription('force class %s is not supported' % type(force).__name__)
        if exclusions is not None:
            pairs = np.stack([exclusions//n, exclusions % n], axis=1).astype(np.int32)
            self._call('b2_set_exclusions', len(pairs), _iptr(pairs))

    @staticmethod
    def _merge_exclusions(current, new):
        if current is not None and not np.array_equal(current, new):
            raise lowering.UnsupportedDescription('all pair forces of a System must share one exclusion list')
        return new

    def _param_set(self, q, sigma, eps):
        set_id = ctypes.c_int()
        self._call('b2_add_param_set', _dptr(q), _dptr(sigma), _dptr(eps), ctypes.byref(set_id))
        return set_id.value

    def _add_bonded(self, family, group, atoms, params, periodic, gparams=None):
        g = np.array(gparams if gparams is not None else [0.0], dtype=np.float64)
        handle = ctypes.c_int()
        self._call('b2_add_bonded_force', family, group, len(atoms), _iptr(atoms), _dptr(params), params.shape[1],
                   1 if periodic else 0, _dptr(g), len(g), ctypes.byref(handle))

    def _add_custom(self, family, group, atoms, params, periodic, code):
        handle = ctypes.c_int()
        consts = np.array(code['consts'] if code['consts'] else [0.0], dtype=np.float64)
        if params.shape[1] == 0:
            params = np.zeros((len(atoms), 1))
            stride = 0
        else:
            stride = params.shape[1]
        self._call('b2_add_custom_bonded_force', family, group, len(atoms), _iptr(atoms), _dptr(params), stride,
                   1 if periodic else 0, _iptr(np.array(code['code_e'] or [0, 0])), code['n_e'],
                   _iptr(np.array(code['code_de'] or [0, 0])), code['n_de'], _dptr(consts), len(code['consts']),
                   ctypes.byref(handle))

    def _classes(self, sigma, eps):
        table = np.stack([sigma, eps], axis=1)
        classes, counts = np.unique(table, axis=0, return_counts=True)
        return [tuple(c) for c in classes], [int(c) for c in counts]

    def _custom_lrc(self, force, family, params, table, cutoff, volume):
        if family != lowering.PAIR_LJ_VIRIAL:
            raise lowering.UnsupportedDescription('long-range correction is only supported for the LJ families')
        classes, counts = self._classes(table[:, 1], table[:, 2])
        rs = force.getSwitchingDistance().value_in_md_units() if force.getUseSwitchingFunction() else None
        return lowering.long_range_correction(classes, counts,
                                              lambda r, s, e: 24*e*(2*(s/r)**12 - (s/r)**6), cutoff, rs, volume)

    def _nonbonded_params(self, force):
        """(kernel parameter block, PME request or None) of an openmm.NonbondedForce."""
        NB = mm.NonbondedForce
        method = force.getNonbondedMethod()
        kc = 138.935456
        cutoff = force.getCutoffDistance().value_in_md_units()
        use_switch = force.getUseSwitchingFunction()
        rswitch = force.getSwitchingDistance().value_in_md_units()
        if method in (NB.NoCutoff, NB.CutoffNonPeriodic):
            raise lowering.UnsupportedDescription('NonbondedForce needs a periodic cutoff method on this engine')
        if method == NB.CutoffPeriodic:
            es = force.getReactionFieldDielectric()
            krf = (es - 1)/((2*es + 1)*cutoff**3)
            crf = 3*es/((2*es + 1)*cutoff)
            return [kc, 2.0, krf, crf, 0.0, float(use_switch), rswitch, cutoff], None
        if method != NB.PME:
            raise lowering.UnsupportedDescription('only PME is implemented for reciprocal space (not Ewald / LJPME)')
        a, nx, ny, nz = force._pme
        tol = force.getEwaldErrorTolerance()
        if a != 0.0:
            alpha, grid = a, [nx, ny, nz]
        else:
            # OpenMM's rule; no rounding to FFT-friendly sizes, like the Reference platform
            alpha = math.sqrt(-math.log(2*tol))/cutoff
            grid = [max(6, int(math.ceil(2*alpha*L/(3*tol**0.2)))) for L in self._box]
        rgroup = force.getReciprocalSpaceForceGroup()
        return [kc, 3.0, 0.0, 0.0, alpha, float(use_switch), rswitch, cutoff], \
            (force.getForceGroup() if rgroup < 0 else rgroup, alpha, grid)

    def _describe_nonbonded(self, force, exclusions, volume):
        """openmm.NonbondedForce: pair part + exception pairs + reciprocal space (PME)."""
        n = self._n
        group = force.getForceGroup()
        table = mm.value_columns(force._particles, 0).reshape(n, 3)
        kc = 138.935456
        cutoff = force.getCutoffDistance().value_in_md_units()
        use_switch = force.getUseSwitchingFunction()
        rswitch = force.getSwitchingDistance().value_in_md_units()
        params, pme = self._nonbonded_params(force)
        alpha = params[4]
        if pme is not None:
            self._pme_request = pme
        exc_atoms = mm.index_columns(force._exceptions, 2)
        exc_values = mm.value_columns(force._exceptions, 2).reshape(len(exc_atoms), 3)
        exclusions = self._merge_exclusions(exclusions, _pair_keys(exc_atoms, n))
        set_id = self._param_set(table[:, 0], table[:, 1], table[:, 2])
        if alpha > 0:
            rgroup, _, grid = self._pme_request
            self._all_groups |= 1 << rgroup
            handle = ctypes.c_int()
            self_energy = -kc*alpha/math.sqrt(math.pi)*float(np.sum(table[:, 0]**2))
            self._call('b2_add_pme', rgroup, set_id, alpha, grid[0], grid[1], grid[2], kc, self_energy,
                       ctypes.byref(handle))
        econst = 0.0
        if force.getUseDispersionCorrection():
            classes, counts = self._classes(table[:, 1], table[:, 2])
            econst = lowering.long_range_correction(
                classes, counts, lambda r, s, e: 4*e*((s/r)**12 - (s/r)**6), cutoff, rswitch if use_switch else None, volume)
        handle = ctypes.c_int()
        p = np.array(params, dtype=np.float64)
        self._call('b2_add_pair_force', lowering.PAIR_LJC, group, set_id, cutoff, _dptr(p), len(p), econst,
                   ctypes.byref(handle))
        self._pair_handles[id(force)] = (handle.value, dict(name='nonbonded'), force)
        # exceptions: own LJ + bare Coulomb, and under Ewald the erf correction with particle charges
        live = np.ones(len(exc_atoms), dtype=bool) if alpha > 0 else (exc_values[:, 0] != 0.0) | (exc_values[:, 2] != 0.0)
        if live.any():
            atoms = np.ascontiguousarray(exc_atoms[live])
            q = table[:, 0]
            params = np.concatenate([exc_values[live], (q[atoms[:, 0]]*q[atoms[:, 1]])[:, None]], axis=1)
            self._add_bonded(lowering.BOND_LJC, group, atoms, params, True, [kc, alpha])
        return exclusions

    # -- integrator ------------------------------------------------------------------------------
    def _load_program(self):
        integrator = self._integrator
        if not isinstance(integrator, mm.CustomIntegrator):
            self._program = None
            return
        self._upload_constraints()
        derivative_slots = {}
        for _, info, force in self._pair_handles.values():
            if info.get('name') == 'softcore':
                if info.get('lambda_vdw'):
                    derivative_slots[info['lambda_vdw']] = lowering.ENERGY_SLOT_DLAMBDA_VDW
                if info.get('lambda_coul'):
                    derivative_slots[info['lambda_coul']] = lowering.ENERGY_SLOT_DLAMBDA_COUL
        program = lowering.lower_program(integrator, self._all_groups, self._parameters, self._fast,
                                         constrained=self._has_constraints, derivative_slots=derivative_slots)
        self._program = program
        ops = program.packed_ops()
        code = np.array(program.bc.code if program.bc.code else [0, 0], dtype=np.int32)
        consts = np.array(program.bc.consts if program.bc.consts else [0.0], dtype=np.float64)
        values = np.array(program.global_values, dtype=np.float64)
        self._call('b2_load_program', _iptr(ops), len(ops), _iptr(code), len(program.bc.code), _dptr(consts),
                   len(program.bc.consts), _dptr(values), len(values), len(program.perdof_names),
                   ctypes.c_uint64(integrator.getRandomNumberSeed() & 0xffffffffffffffff))
        # context parameters the integrator moves itself (AFED): the pair kernels read them on the device
        self._device_parameters = set()
        assigned = set(integrator.getComputationStep(k)[1] for k in range(integrator.getNumComputations()))
        for handle, info, force in self._pair_handles.values():
            if info.get('name') != 'softcore':
                continue
            for slot, key in ((1, 'lambda_vdw'), (2, 'lambda_coul')):
                name = info.get(key)
                if name and name in assigned and name in program.global_names:
                    self._call('b2_bind_pair_parameter', handle, slot, program.gindex(name))
                    self._device_parameters.add(name)
        for k, value in enumerate(integrator._perdof_values):
            if not np.isscalar(value) and self._have_positions:
                self._set_perdof(integrator._perdof_names[k], value)
            elif np.isscalar(value) and value != 0.0 and self._have_positions:
                self._set_perdof(integrator._perdof_names[k], np.full((self._n, 3), float(value)))

    def _upload_constraints(self):
        constraints = self._system._constraints
        if not constraints:
            return
        pairs = np.array([[c[0], c[1]] for c in constraints], dtype=np.int32)
        lengths = np.array([c[2] for c in constraints], dtype=np.float64)
        tolerance = self._integrator.getConstraintTolerance() if self._integrator is not None else 1e-5
        self._call('b2_set_constraints', len(pairs), _iptr(pairs), _dptr(lengths), float(tolerance))

    def _integrator_changed(self):
        if self._program is not None:
            self._set_global('dt', self._integrator._dt)
            self._upload_constraints()

    def _get_global(self, name):
        out = ctypes.c_double()
        self._call('b2_get_globals', self._program.gindex(name), 1, ctypes.byref(out))
        return out.value

    def _set_global(self, name, value):
        v = ctypes.c_double(float(value))
        self._call('b2_set_globals', self._program.gindex(name), 1, ctypes.byref(v))

    def _get_perdof(self, name):
        torch = self._torch
        with torch.cuda.stream(self._stream):
            self._call('b2_get_perdof', self._program.perdof_names.index(name), c_void(self._buffer.data_ptr()))
            self._call('b2_synchronize')
            return [mm.Vec3(*row) for row in self._buffer.cpu().numpy()]

    def _set_perdof(self, name, array):
        torch = self._torch
        if not self._have_positions:
            return   # applied by _load_program/_after_positions once an order exists
        with torch.cuda.stream(self._stream):
            t = torch.as_tensor(np.ascontiguousarray(array, dtype=np.float64)).to(self._device)
            self._call('b2_set_perdof', self._program.perdof_names.index(name), c_void(t.data_ptr()))
            self._call('b2_synchronize')

    def _step(self, steps):
        if self._program is None:
            raise mm.OpenMMException('this integrator cannot take steps on the B200 engine')
        if not self._have_positions:
            raise mm.OpenMMException('Particle positions have not been set')
        self._call('b2_run', steps)
        self._time += steps*self._integrator._dt

    def _custom_table(self, force, info):
        """Per-particle (charge, sigma, epsilon) table of a CustomNonbondedForce as the engine stores it."""
        n = self._n
        table = mm.value_columns(force._particles, 0).reshape(n, -1)
        if table.shape[1] == 2:
            # (sigma, epsilon) only; with an interaction group the charge column carries the
            # +1/-1 set labels the soft-core kernel uses to keep only unlike pairs
            labels = np.zeros(n)
            if info.get('partition') is not None:
                labels[:] = -1.0
                labels[info['partition']] = 1.0
            table = np.concatenate([labels[:, None], table], axis=1)
        return table

    def _pair_description(self, force):
        """(params, energy constant, particle table) of a pair force from its CURRENT description."""
        volume = float(np.prod(self._current_box()))
        if isinstance(force, mm.NonbondedForce):
            params, pme = self._nonbonded_params(force)
            if pme is not None:
                raise lowering.UnsupportedDescription('updateParametersInContext of a PME NonbondedForce is not supported '
                                                      '(self energy and exception corrections depend on the charges)')
            table = mm.value_columns(force._particles, 0).reshape(self._n, 3)
            econst = 0.0
            if force.getUseDispersionCorrection():
                cutoff = force.getCutoffDistance().value_in_md_units()
                rs = force.getSwitchingDistance().value_in_md_units() if force.getUseSwitchingFunction() else None
                classes, counts = self._classes(table[:, 1], table[:, 2])
                econst = lowering.long_range_correction(
                    classes, counts, lambda r, s, e: 4*e*((s/r)**12 - (s/r)**6), cutoff, rs, volume)
            return params, econst, table
        family, cutoff, params, info = lowering.classify_pair_force(force, self._parameters)
        table = self._custom_table(force, info)
        econst = self._custom_lrc(force, family, params, table, cutoff, volume) if force.getUseLongRangeCorrection() else 0.0
        return params, econst, table

    def _parameters_changed(self, force):
        """Force.updateParametersInContext: global parameters, the long-range correction and the
        per-particle table of a pair force are re-read from the description (NonbondedForce
        exceptions and explicit-list forces cannot be updated in a live context)."""
        entry = self._pair_handles.get(id(force))
        if entry is None:
            raise lowering.UnsupportedDescription('updateParametersInContext is only supported for pair forces')
        handle = entry[0]
        params, econst, table = self._pair_description(force)
        p = np.array(params, dtype=np.float64)
        self._call('b2_update_pair_force', handle, _dptr(p), len(p), float(econst))
        self._call('b2_update_pair_particles', handle, _dptr(table[:, 0]), _dptr(table[:, 1]), _dptr(table[:, 2]))

    # -- public API --------------------------------------------------------------------------------
    def getSystem(self):
        return self._system

    def getIntegrator(self):
        return self._integrator

    def getPlatform(self):
        return mm.Platform('B200')

    def getMolecules(self):
        return [tuple(g.tolist()) for g in self._molecule_groups]

    def getParameter(self, name):
        if name in getattr(self, '_device_parameters', ()):
            self._parameters[name] = self._get_global(name)     # moved by the integrator on the device
        return self._parameters[name]

    def getParameters(self):
        for name in getattr(self, '_device_parameters', ()):
            self._parameters[name] = self._get_global(name)
        return dict(self._parameters)

    def setParameter(self, name, value):
        if name not in self._parameters:
            raise mm.OpenMMException('Called setParameter() with invalid parameter name: %s' % name)
        self._parameters[name] = float(_md(value))
        for handle, info, force in self._pair_handles.values():
            if hasattr(force, 'getNumGlobalParameters') and any(
                    force.getGlobalParameterName(k) == name for k in range(force.getNumGlobalParameters())):
                params, econst, _ = self._pair_description(force)
                p = np.array(params, dtype=np.float64)
                self._call('b2_update_pair_force', handle, _dptr(p), len(p), float(econst))
        if self._program is not None and name in self._program.global_names:
            self._set_global(name, self._parameters[name])

    def setPeriodicBoxVectors(self, a, b, c):
        """OpenMM semantics: the box changes, the atoms are not moved.  Cells, neighbour lists, PME
        influence function and long-range corrections follow; the captured step graph is released."""
        box = np.array([_md(a)[0], _md(b)[1], _md(c)[2]], dtype=np.float64)
        if not np.allclose(box, self._current_box(), rtol=0, atol=1e-14):
            with self._torch.cuda.stream(self._stream):
                self._call('b2_update_box', _dptr(box))
            self._box = box

    def _current_box(self):
        """The box as the engine holds it (a Monte Carlo barostat moves it during a run)."""
        out = (ctypes.c_double*3)()
        self._call('b2_get_box', out)
        self._box = np.array([out[0], out[1], out[2]], dtype=np.float64)
        return self._box

    def barostat_statistics(self):
        out = (ctypes.c_longlong*2)()
        scale = ctypes.c_double()
        self._call('b2_get_barostat_stats', out, ctypes.byref(scale))
        return dict(attempts=out[0], accepted=out[1], volume_scale=scale.value)

    def _upload(self, values, what):
        torch = self._torch
        if isinstance(values, torch.Tensor):
            t = values.to(device=self._device, dtype=torch.float64).reshape(self._n, 3).contiguous()
        else:
            array = np.asarray(_md(values), dtype=np.float64).reshape(-1, 3)
            if array.shape[0] != self._n:
                raise mm.OpenMMException('Called set%s() on a Context with the wrong number of %s' % (what, what.lower()))
            t = torch.from_numpy(np.ascontiguousarray(array)).to(self._device, non_blocking=False)
        return t

    def _download(self):
        """The [n][3] device buffer as a fresh numpy array, through a pinned staging buffer (DMA at full PCIe rate
        instead of a pageable copy; the array the caller receives is its own)."""
        torch = self._torch
        if getattr(self, '_staging', None) is None:
            self._staging = torch.empty(self._buffer.shape, dtype=self._buffer.dtype, pin_memory=True)
        self._staging.copy_(self._buffer)
        self._stream.synchronize()
        return self._staging.numpy().copy()

    def setPositions(self, positions):
        torch = self._torch
        with torch.cuda.stream(self._stream):
            t = self._upload(positions, 'Positions')
            self._x.copy_(t)
            self._call('b2_set_positions', c_void(self._x.data_ptr()))
            first = not self._have_positions
            self._have_positions = True
            if first and self._program is not None:
                for k, value in enumerate(self._integrator._perdof_values):
                    if not np.isscalar(value):
                        self._set_perdof(self._integrator._perdof_names[k], value)
                    elif value != 0.0:
                        self._set_perdof(self._integrator._perdof_names[k], np.full((self._n, 3), float(value)))

    def setVelocities(self, velocities):
        torch = self._torch
        if not self._have_positions:
            raise mm.OpenMMException('set positions before velocities on this engine')
        with torch.cuda.stream(self._stream):
            t = self._upload(velocities, 'Velocities')
            self._call('b2_set_velocities', c_void(t.data_ptr()))
            self._call('b2_synchronize')

    def setVelocitiesToTemperature(self, temperature, randomSeed=None):
        """Maxwell-Boltzmann velocities (numpy Philox stream; OpenMM's SFMT stream is not
        reproducible outside OpenMM)."""
        kT = 8.314472471220217e-3*float(_md(temperature))
        if self._nranks > 1:    # every rank must draw the same velocities
            import pickle
            seed = randomSeed if randomSeed is not None else int(np.random.SeedSequence().entropy % (1 << 63))
            randomSeed = pickle.loads(broadcast_bytes(pickle.dumps(seed) if self._rank == 0 else None))
        rng = np.random.Generator(np.random.Philox(randomSeed if randomSeed is not None else None))
        v = rng.standard_normal((self._n, 3))
        with np.errstate(divide='ignore'):
            scale = np.where(self._masses > 0, np.sqrt(kT/np.where(self._masses > 0, self._masses, 1.0)), 0.0)
        self.setVelocities(v*scale[:, None])

    def setState(self, state):
        self.setPositions(state.getPositions(asNumpy=True))
        if state._velocities is not None:
            self.setVelocities(state.getVelocities(asNumpy=True))

    def reinitialize(self, preserveState=False):
        raise lowering.UnsupportedDescription('reinitialize() is not supported: create a new Context')

    @staticmethod
    def _mask(groups):
        if isinstance(groups, (set, frozenset, list, tuple)):
            mask = 0
            for g in groups:
                mask |= 1 << int(g)
            return mask
        mask = int(groups)
        return 0xffffffff if mask == -1 else mask & 0xffffffff

    def getState(self, getPositions=False, getVelocities=False, getForces=False, getEnergy=False,
                 getParameters=False, getParameterDerivatives=False, enforcePeriodicBox=False, groups=-1):
        torch = self._torch
        mask = self._mask(groups)
        need_state = getPositions or getForces or getEnergy or getVelocities
        if need_state and not self._have_positions:
            raise mm.OpenMMException('Particle positions have not been set')
        fields = dict(_positions=None, _velocities=None, _forces=None, _potential=None, _kinetic=None,
                      _box=self._current_box().copy(), _parameters=self.getParameters() if getParameters else {},
                      _derivatives={}, _time=self._time)
        with torch.cuda.stream(self._stream):
            if getPositions:
                self._call('b2_get_positions', c_void(self._buffer.data_ptr()))
                self._call('b2_synchronize')
                pos = self._download()
                if enforcePeriodicBox:
                    box = fields['_box']
                    for group in self._molecule_groups:
                        centre = pos[group].mean(axis=0)
                        pos[group] -= np.floor(centre/box)*box
                fields['_positions'] = pos
            if getVelocities:
                self._call('b2_get_velocities', c_void(self._buffer.data_ptr()))
                self._call('b2_synchronize')
                fields['_velocities'] = self._download()
            if getEnergy:
                kinetic = ctypes.c_double()
                self._call('b2_kinetic_energy', ctypes.byref(kinetic))      # reduced on the device
                fields['_kinetic'] = kinetic.value
            flags = (1 if getForces else 0) | (2 if (getEnergy or getParameterDerivatives) else 0) | \
                (4 if (getForces and self._double_forces) else 0)
            if flags:
                energy, virial = ctypes.c_double(), ctypes.c_double()
                self._call('b2_eval', ctypes.c_uint32(mask), flags,
                           c_void(self._buffer.data_ptr()) if getForces else c_void(),
                           ctypes.byref(energy), ctypes.byref(virial))
                self._call('b2_synchronize')
                if getForces:
                    fields['_forces'] = self._download()
                if getEnergy:
                    if not (math.isfinite(energy.value) and math.isfinite(fields['_kinetic'])):
                        # OpenMM's behaviour: a blown-up trajectory is reported, not returned as numbers
                        raise mm.OpenMMException('Energy is NaN (potential %r, kinetic %r): the simulation has become '
                                                 'unstable' % (energy.value, fields['_kinetic']))
                    fields['_potential'] = energy.value
                    fields['_virial'] = virial.value
                if getParameterDerivatives:
                    out = (ctypes.c_double*2)()
                    self._call('b2_get_parameter_derivatives', out)
                    names = {}
                    for _, info, force in self._pair_handles.values():
                        if info.get('name') == 'softcore':
                            if info.get('lambda_vdw'):
                                names[info['lambda_vdw']] = out[0]
                            if info.get('lambda_coul'):
                                names[info['lambda_coul']] = out[1]
                    fields['_derivatives'] = names
        return State(**fields)

    # -- extras used by tests / benchmarks ---------------------------------------------------------
    def group_energies(self):
        e = (ctypes.c_double*32)()
        w = (ctypes.c_double*32)()
        self._call('b2_get_group_energies', e, w)
        return np.array(e), np.array(w)

    def pair_set(self, force, want_pairs=False):
        """(count, checksum[, pairs]) of interacting pairs of a pair force (parity tests)."""
        torch = self._torch
        handle = self._pair_handles[id(force)][0]
        count, checksum = ctypes.c_longlong(), ctypes.c_ulonglong()
        with torch.cuda.stream(self._stream):
            self._call('b2_pair_set', handle, ctypes.byref(count), ctypes.byref(checksum), c_void(), 0)
            if not want_pairs:
                return count.value, checksum.value
            pairs = torch.zeros((max(1, count.value), 2), dtype=torch.int32, device=self._device)
            self._call('b2_pair_set', handle, ctypes.byref(count), ctypes.byref(checksum), c_void(pairs.data_ptr()),
                       ctypes.c_longlong(pairs.shape[0]))
            self._call('b2_synchronize')
            return count.value, checksum.value, pairs.cpu().numpy()[:count.value]

    def counters(self):
        out = (ctypes.c_longlong*8)()
        self._call('b2_synchronize')
        self._call('b2_get_counters', out)
        return dict(launches=out[0], rebuilds=out[1], pair_launches=out[2], list_capacity=out[3],
                    list_max=out[4], graph_launches=out[5], kernels_per_step=out[6])

    def jit_stats(self):
        """Run-time compiled per-DOF / sum steps of the loaded program (csrc/jit.cu) and their launches so far."""
        out = (ctypes.c_longlong*2)()
        self._call('b2_get_jit_stats', out)
        return dict(compiled_steps=out[0], launches=out[1])

    def list_stats(self):
        out = (ctypes.c_longlong*4)()
        self._call('b2_get_list_stats', out)
        return dict(rebuilds=out[0], largest_list=out[1], fat_groups=out[2], groups=out[3])

    def set_profiling(self, on):
        self._call('b2_set_profiling', 1 if on else 0)

    def pair_profile(self):
        """[(description, group, total_ms, launches, list_entries)] for every pair force."""
        out = []
        for handle, info, force in self._pair_handles.values():
            ms, launches, entries = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_longlong()
            self._call('b2_get_profile', handle, ctypes.byref(ms), ctypes.byref(launches), ctypes.byref(entries))
            out.append(dict(name=info.get('name'), group=force.getForceGroup(), total_ms=ms.value,
                            launches=launches.value, entries=entries.value))
        return out

    def phase_profile(self):
        """Milliseconds per phase of the step accumulated in profiling mode (eager pass)."""
        out = (ctypes.c_double*6)()
        self._call('b2_get_phase_profile', out)
        names = ('other', 'skin_test_and_exchange', 'list_rebuild', 'pair_kernels', 'integrator_kernels', 'reductions')
        return dict(zip(names, [round(v, 3) for v in out]))

    def synchronize(self):
        self._call('b2_synchronize')
