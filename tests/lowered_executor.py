"""
TEST INFRASTRUCTURE: a numpy executor of *lowered* integrator programs (the op list and VM
bytecode that atomsmm_b200.lowering hands to the CUDA engine, layout in csrc/program.h).

Purpose: check the host-side lowering -- loop unrolling, force-copy folding, kick / velocity-op
fusion, thermostat-chain merging, derivative and invalidation ops, constraint ops -- on the CPU,
by executing the lowered program with the semantics the engine's kernels implement and comparing
with oracle/interp.py executing the ORIGINAL step program.  Forces, derivatives and constraints come
from the oracle, so any difference is a lowering bug.  Never used by the product.
"""

import math

import numpy as np

from atomsmm_b200 import expr as X
from atomsmm_b200 import lowering as L

OPS = {v: k for k, v in X.OPCODES.items()}


class Executor(object):
    def __init__(self, system, integrator, positions, velocities, parameters=None, constrained=None,
                 derivative_slots=None, group_mask=0xffffffff, seed=0):
        from oracle import interp, refmath
        self.refmath, self.interp = refmath, interp
        self.system = system
        self.n = system.getNumParticles()
        self.box = refmath.system_box(system)
        self.mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(self.n)])
        self.x = np.array(positions, dtype=np.float64).copy()
        self.v = np.array(velocities, dtype=np.float64).copy()
        self.parameters = dict(parameters or {})
        if constrained is None:
            constrained = system.getNumConstraints() > 0
        self.program = P = L.lower_program(integrator, group_mask, self.parameters, True, constrained=constrained,
                                           derivative_slots=derivative_slots)
        self.globals = np.array(P.global_values, dtype=np.float64)
        self.consts = np.array(P.bc.consts if P.bc.consts else [0.0])
        self.code = list(P.bc.code)
        self.perdof = []
        for k in range(integrator.getNumPerDofVariables()):
            value = integrator._perdof_values[k]
            self.perdof.append(np.full((self.n, 3), float(value)) if np.isscalar(value)
                               else np.array(value, dtype=np.float64))
        self.forces = {}            # slot -> (version, array)
        self.version = 0
        self.energies = np.zeros(96)
        self.deriv_version = -1
        self.rng = np.random.default_rng(seed)
        self.x_constrained = None
        self.param_names = {P.gindex(name): name for name in self.parameters if name in P.global_names}

    # -- VM ------------------------------------------------------------------------------------------
    def run_vm(self, start, length, per_dof=False):
        """Scalar (per_dof False) or vectorised per-DOF evaluation of `length` instructions at `start`."""
        code, stack, pc = self.code, [], 0
        while pc < length:
            op, arg = OPS[code[start + 2*pc]], code[start + 2*pc + 1]
            pc += 1
            if op == 'PUSHC':
                stack.append(self.consts[arg])
            elif op == 'PUSHG':
                stack.append(self.globals[arg])
            elif op == 'PUSHV':
                stack.append(self.x if arg == 0 else self.v if arg == 1 else self.perdof[arg - 2])
            elif op == 'PUSHM':
                stack.append(self.mass[:, None])
            elif op == 'PUSHF':
                stack.append(self.forces[arg][1])
            elif op == 'PUSHE':
                stack.append(self.energies[arg])
            elif op == 'GAUSS':
                stack.append(self.rng.standard_normal((self.n, 3)) if per_dof else float(self.rng.standard_normal()))
            elif op == 'UNIF':
                stack.append(self.rng.random((self.n, 3)) if per_dof else float(self.rng.random()))
            elif op in ('ADD', 'SUB', 'MUL', 'DIV', 'POW', 'MIN', 'MAX'):
                b, a = stack.pop(), stack.pop()
                stack.append({'ADD': lambda: a + b, 'SUB': lambda: a - b, 'MUL': lambda: a*b, 'DIV': lambda: a/b,
                              'POW': lambda: np.power(a, b), 'MIN': lambda: np.minimum(a, b),
                              'MAX': lambda: np.maximum(a, b)}[op]())
            elif op == 'NEG':
                stack.append(-stack.pop())
            elif op == 'POWI':
                stack.append(stack.pop()**arg)
            elif op in ('SQRT', 'EXP', 'LOG', 'SIN', 'COS', 'TAN', 'ABS', 'FLOOR', 'CEIL'):
                fn = {'SQRT': np.sqrt, 'EXP': np.exp, 'LOG': np.log, 'SIN': np.sin, 'COS': np.cos, 'TAN': np.tan,
                      'ABS': np.abs, 'FLOOR': np.floor, 'CEIL': np.ceil}[op]
                stack.append(fn(stack.pop()))
            elif op == 'ERF':
                stack.append(math.erf(stack.pop()))
            elif op == 'ERFC':
                stack.append(math.erfc(stack.pop()))
            elif op == 'STEP':
                stack.append(np.where(np.asarray(stack.pop()) < 0, 0.0, 1.0))
            elif op == 'DELTA':
                stack.append(np.where(np.asarray(stack.pop()) == 0, 1.0, 0.0))
            elif op == 'SELECT':
                c, b, a = stack.pop(), stack.pop(), stack.pop()
                stack.append(np.where(np.asarray(a) != 0, b, c))
            elif op == 'CMP':
                b, a = stack.pop(), stack.pop()
                r = [a == b, a < b, a > b, a != b, a <= b, a >= b][arg]
                stack.append(1.0 if r else 0.0)
            elif op == 'STOREG':
                value = float(stack.pop())
                self.globals[arg] = value
                if arg in self.param_names:
                    self.parameters[self.param_names[arg]] = value
            elif op == 'JMP':
                pc = arg
            elif op == 'JMPZ':
                if float(stack.pop()) == 0.0:
                    pc = arg
            else:
                raise NotImplementedError(op)
        return stack[-1] if stack else 0.0

    # -- forces ----------------------------------------------------------------------------------------
    def ensure(self, mask, slot):
        cached = self.forces.get(slot)
        if cached is not None and cached[0] == self.version:
            return
        mask &= 0xffffffff
        groups = None if mask == 0xffffffff else {g for g in range(32) if mask & (1 << g)}
        f = self.refmath.evaluate_system(self.system, self.x, self.box, groups, self.parameters).forces
        self.forces[slot] = (self.version, f)

    def constraints(self):
        return [(c[0], c[1], c[2]) for c in self.system._constraints]

    # -- one MD step -------------------------------------------------------------------------------------
    def step(self, count=1):
        massive = (self.mass > 0)[:, None]
        w = np.where(self.mass > 0, 1.0/np.where(self.mass > 0, self.mass, 1.0), 0.0)[:, None]
        for _ in range(count):
            self.x_constrained = self.x.copy()
            for op in self.program.ops:
                kind = op[0]
                if kind == L.OP_EVAL:
                    self.ensure(op[1], op[2])
                elif kind == L.OP_GLOBAL:
                    self.run_vm(op[2], op[3])
                elif kind == L.OP_PERDOF:
                    value = np.broadcast_to(self.run_vm(op[2], op[3], True), (self.n, 3)).astype(np.float64)
                    if op[1] == 0:
                        self.x = np.where(massive, value, self.x)
                        self.version += 1
                    elif op[1] == 1:
                        self.v = np.where(massive, value, self.v)
                    else:
                        self.perdof[op[1] - 2] = value.copy()
                elif kind == L.OP_SUM:
                    value = np.broadcast_to(self.run_vm(op[2], op[3], True), (self.n, 3))
                    self.globals[op[1]] = float(np.sum(value))
                elif kind == L.OP_KICK:
                    nterms, offset, drift, prescale, mvv, cstart, clen = op[1:8]
                    if prescale >= 0:
                        self.v = np.where(massive, self.v*self.globals[prescale], self.v)
                    if nterms > 0:
                        a = np.zeros_like(self.v)
                        for t in range(nterms):
                            slot, coef, sign = self.code[offset + 3*t: offset + 3*t + 3]
                            version, f = self.forces[slot]
                            assert version == self.version, 'kick reads a stale force slot %d' % slot
                            a += sign*self.globals[coef]*f
                        self.v = self.v + a*w
                    if drift >= 0:
                        self.x = np.where(massive, self.x + self.globals[drift]*self.v, self.x)
                        self.version += 1
                    if mvv >= 0:
                        self.globals[mvv] = float(np.sum(self.mass[:, None]*self.v*self.v))
                        if clen > 0:
                            self.run_vm(cstart, clen)
                elif kind == L.OP_DRIFT:
                    self.x = np.where(massive, self.x + self.globals[op[1]]*self.v, self.x)
                    self.version += 1
                elif kind == L.OP_SCALE:
                    self.v = self.v*self.globals[op[1]]
                elif kind == L.OP_UPDATE_STATE:
                    pass
                elif kind == L.OP_INVALIDATE:
                    self.forces = {}
                    self.deriv_version = -1
                elif kind == L.OP_ENERGY:
                    if self.deriv_version != self.version:
                        self.deriv_version = self.version
                        for name, slot in (self.program_derivative_slots or {}).items():
                            h = 1e-6
                            base = dict(self.parameters)
                            up = self.refmath.evaluate_system(self.system, self.x, self.box, None,
                                                              dict(base, **{name: base[name] + h})).energy
                            dn = self.refmath.evaluate_system(self.system, self.x, self.box, None,
                                                              dict(base, **{name: base[name] - h})).energy
                            self.energies[slot] = (up - dn)/(2*h)
                elif kind == L.OP_CONSTRAIN_X:
                    self.x = self.interp.shake(self.constraints(), self.mass, self.x, self.x_constrained)
                    self.x_constrained = self.x.copy()
                    self.version += 1
                elif kind == L.OP_MVV_FACTOR:
                    pass            # a hint for the engine's carried sum(m v.v); no effect on the state
                elif kind == L.OP_CONSTRAIN_V:
                    self.v = self.interp.rattle(self.constraints(), self.mass, self.x, self.v)
                else:
                    raise NotImplementedError('op %d' % kind)

    program_derivative_slots = None


class DistributedExecutor(Executor):
    """The domain-decomposition protocol of csrc/dist.cu on top of torch.distributed (gloo on CPU): each
    rank integrates only the atoms of its ownership range (engine.partition_ranges over whole molecules),
    positions are exchanged before a force evaluation, sums over degrees of freedom are all-reduced, and
    every rank runs the scalar programs redundantly.  Getters gather the full state."""

    def __init__(self, system, integrator, positions, velocities, dist, pair_mask=0xffffffff, **kwargs):
        super().__init__(system, integrator, positions, velocities, **kwargs)
        import torch
        from atomsmm_b200 import engine
        self.dist, self.torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        molecule, _ = engine._molecules(system)
        self.ranges = engine.partition_ranges(molecule, self.world)     # caller order = engine order here
        self.lo, self.hi = int(self.ranges[self.rank]), int(self.ranges[self.rank + 1])
        self.owned = np.zeros((self.n, 1), bool)
        self.owned[self.lo:self.hi] = True
        self.pair_mask = pair_mask      # force groups that contain pair forces: only those need foreign atoms
        self.synced = self.version
        self.exchanges = 0

    def gather(self, array):
        """All ranks end up with every rank's owned segment (the grouped broadcasts of dist.cu)."""
        t = self.torch.from_numpy(np.ascontiguousarray(array))
        for r in range(self.world):
            segment = t[int(self.ranges[r]):int(self.ranges[r + 1])]
            if segment.numel():
                self.dist.broadcast(segment, src=r)
        return t.numpy()

    def all_reduced(self, value):
        t = self.torch.tensor([float(np.sum(np.where(self.owned, value, 0.0)))], dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t)

    def ensure(self, mask, slot):
        cached = self.forces.get(slot)
        if cached is not None and cached[0] == self.version:
            return
        if (mask & self.pair_mask) and self.synced != self.version:
            self.x = self.gather(self.x)
            self.synced = self.version
            self.exchanges += 1
        super().ensure(mask, slot)

    def step(self, count=1):
        """Executor.step with every update masked to the owned range and every sum all-reduced."""
        massive = (self.mass > 0)[:, None] & self.owned
        w = np.where(self.mass > 0, 1.0/np.where(self.mass > 0, self.mass, 1.0), 0.0)[:, None]
        for _ in range(count):
            for op in self.program.ops:
                kind = op[0]
                if kind == L.OP_EVAL:
                    self.ensure(op[1], op[2])
                elif kind == L.OP_GLOBAL:
                    self.run_vm(op[2], op[3])
                elif kind == L.OP_PERDOF:
                    value = np.broadcast_to(self.run_vm(op[2], op[3], True), (self.n, 3)).astype(np.float64)
                    if op[1] == 0:
                        self.x = np.where(massive, value, self.x)
                        self.version += 1
                    elif op[1] == 1:
                        self.v = np.where(massive, value, self.v)
                    else:
                        self.perdof[op[1] - 2] = np.where(self.owned, value, self.perdof[op[1] - 2])
                elif kind == L.OP_SUM:
                    value = np.broadcast_to(self.run_vm(op[2], op[3], True), (self.n, 3))
                    self.globals[op[1]] = self.all_reduced(value)
                elif kind == L.OP_KICK:
                    nterms, offset, drift, prescale, mvv, cstart, clen = op[1:8]
                    if prescale >= 0:
                        self.v = np.where(massive, self.v*self.globals[prescale], self.v)
                    if nterms > 0:
                        a = np.zeros_like(self.v)
                        for t in range(nterms):
                            slot, coef, sign = self.code[offset + 3*t: offset + 3*t + 3]
                            version, f = self.forces[slot]
                            assert version == self.version
                            a += sign*self.globals[coef]*f
                        self.v = np.where(massive, self.v + a*w, self.v)
                    if drift >= 0:
                        self.x = np.where(massive, self.x + self.globals[drift]*self.v, self.x)
                        self.version += 1
                    if mvv >= 0:
                        self.globals[mvv] = self.all_reduced(self.mass[:, None]*self.v*self.v)
                        if clen > 0:
                            self.run_vm(cstart, clen)
                elif kind in (L.OP_UPDATE_STATE, L.OP_MVV_FACTOR):
                    pass
                else:
                    raise NotImplementedError('op %d in the distributed executor' % kind)

    def full_state(self):
        return self.gather(self.x.copy()), self.gather(self.v.copy())


class HaloExecutor(DistributedExecutor):
    """The PEER-MEMORY protocol of csrc/dist.cu + csrc/dd.cuh (halo pull): before a pair-force evaluation
    every rank tests the skin criterion on its OWN atoms, the verdicts are OR-ed over the ranks, and then
    either everybody refreshes every foreign atom and rebuilds (new reference positions, new halo = the
    foreign atoms within list radius of an owned atom), or each rank refreshes ONLY its halo; all other
    foreign atoms keep whatever stale position they had.  The trajectories must still equal the
    single-process ones: a needed atom left stale would show up as a force error."""

    def __init__(self, *args, rlist=1.1, skin=0.1, **kwargs):
        super().__init__(*args, **kwargs)
        self.rlist, self.skin = rlist, skin
        self.xref = None
        self.halo = np.zeros(self.n, bool)
        self.rebuilds = 0
        self.halo_sizes = []

    def _build(self):
        own = np.arange(self.lo, self.hi)
        self.xref = self.x[own].copy()
        d = self.x[:, None, :] - self.x[None, own, :]
        d -= self.box*np.rint(d/self.box)
        near = (np.sum(d*d, axis=2) < self.rlist**2).any(axis=1)
        near[own] = False
        self.halo = near
        self.rebuilds += 1
        self.halo_sizes.append(int(near.sum()))

    def ensure(self, mask, slot):
        cached = self.forces.get(slot)
        if cached is not None and cached[0] == self.version:
            return
        if (mask & self.pair_mask) and self.synced != self.version:
            own = slice(self.lo, self.hi)
            moved = 1.0 if self.xref is None else float(
                (np.sum((self.x[own] - self.xref)**2, axis=1) > (0.5*self.skin)**2).any())
            verdict = self.torch.tensor([moved], dtype=self.torch.float64)
            self.dist.all_reduce(verdict, op=self.dist.ReduceOp.MAX)
            fresh = self.gather(self.x.copy())            # what the owners hold; this rank may read only parts of it
            if float(verdict) > 0:
                self.x = fresh
                self._build()
            else:
                self.x[self.halo] = fresh[self.halo]
            self.synced = self.version
            self.exchanges += 1
        Executor.ensure(self, mask, slot)

    def full_state(self):
        self.synced = -1
        return super().full_state()
