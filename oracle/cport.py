"""
ORACLE (test infrastructure only -- see oracle/__init__.py).

ctypes front end of oracle/c/oracle.c, the float64 C/OpenMP restatement used (a) as the checker
for systems too large for the numpy/sympy oracle and (b) as the CPU baseline in bench.py.  A
``CPort`` is built from the same System description the engine consumes; every pair potential
it sets up is verified on construction against the generic evaluation of the force's own energy
string (oracle.refmath), so the closed forms in oracle.c cannot silently diverge from the
reference's strings.
"""

import ctypes
import math
import os
import subprocess

import numpy as np

from . import refmath

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_build', 'liboracle.so')
_lib = None
c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


def library():
    global _lib
    if _lib is None:
        # make decides whether the library is stale (a no-op when it is up to date)
        subprocess.run(['make', '-s', '-C', os.path.join(HERE, 'c')], check=True)
        lib = ctypes.CDLL(LIB_PATH)
        lib.orc_create.argtypes = [ctypes.c_int, c_double_p, c_double_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.orc_destroy.argtypes = [ctypes.c_void_p]
        lib.orc_set_threads.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.orc_set_params.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, c_double_p]
        lib.orc_set_exclusions.argtypes = [ctypes.c_void_p, ctypes.c_int, c_int_p]
        lib.orc_add_pair.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_double_p, ctypes.c_int]
        lib.orc_add_bonded.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, c_double_p,
                                       ctypes.c_int, c_double_p, ctypes.c_int]
        lib.orc_eval.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_uint, c_double_p, c_double_p, c_double_p]
        lib.orc_respa.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, c_double_p]
        lib.orc_pair_set.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong),
                                     ctypes.POINTER(ctypes.c_ulonglong)]
        lib.orc_counter.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.orc_counter.restype = ctypes.c_long
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _columns(force, attribute, nindex):
    """(int32 index columns, float64 value columns) of a description table, read in bulk (the
    description classes keep them as lists of rows or as columnar mm.PackedRows; values are in MD units)."""
    from atomsmm_b200 import mm
    rows = getattr(force, attribute)
    return mm.index_columns(rows, nindex), mm.value_columns(rows, nindex)


F_NEAR, F_DAMPED, F_LJC, F_LJ_VIRIAL = 1, 2, 3, 4
B_BOND, B_ANGLE, B_TORSION, B_LJC = 1, 2, 3, 4


def _custom_pair_params(force):
    """Family + parameter block of a CustomNonbondedForce from its globals and string features."""
    glob = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
            for k in range(force.getNumGlobalParameters())}
    text = force.getEnergyFunction().replace(' ', '')
    cutoff = force.getCutoffDistance().value_in_md_units()
    main = text.split(';')[0]
    import re

    def literal(name):
        if name in glob:
            return glob[name]
        m = re.search(r'(?:^|;)%s=([0-9.eE+-]+)' % name, text)
        return float(m.group(1)) if m else None
    if literal('rs0') is not None and literal('rc0') is not None:
        variant = 2 if 'f12' in main else (1 if '(sigma/rc0)^12' in main else 0)
        sign = -1.0 if main.startswith('-') else 1.0
        return F_NEAR, cutoff, [variant, literal('rs0'), literal('rc0'), literal('Kc'), sign, 1.0]
    if 'alpha' in glob:
        m = re.search(r'd=([0-9]+)', text)
        return F_DAMPED, cutoff, [glob['alpha'], glob['rswitch'], glob['rcut'], float(m.group(1)) if m else 1.0, glob['Kc']]
    if main.startswith('24*epsilon'):
        return F_LJ_VIRIAL, cutoff, [1.0 if force.getUseSwitchingFunction() else 0.0,
                                     force.getSwitchingDistance().value_in_md_units(), cutoff]
    raise ValueError('C port: unsupported energy function %s' % force.getEnergyFunction())


class CPort(object):
    def __init__(self, system, threads=None, verify=True):
        lib = self.lib = library()
        self.system = system
        n = self.n = system.getNumParticles()
        self.box = refmath.system_box(system)
        self.mass = np.array(system._masses, dtype=np.float64)
        self.handle = ctypes.c_void_p()
        lib.orc_create(n, _dp(self.mass), _dp(self.box), ctypes.byref(self.handle))
        if threads:
            lib.orc_set_threads(self.handle, threads)
        self.threads = threads or os.cpu_count()
        params_set = False
        self._pair_index = {}      # id(force) -> index among the pair forces of the C system
        self.econst = {}      # position-independent energies (long-range corrections) per group
        for force in system.getForces():
            kind = refmath._kind(force)
            group = force.getForceGroup()
            if kind == 'CustomNonbondedForce':
                table = _columns(force, '_particles', 0)[1].reshape(n, -1)
                self._set_params(table[:, 0], table[:, 1], table[:, 2], params_set)
                params_set = True
                excl = np.ascontiguousarray(_columns(force, '_exclusions', 2)[0])
                lib.orc_set_exclusions(self.handle, len(excl), _ip(np.ascontiguousarray(excl.reshape(-1))))
                family, cutoff, p = _custom_pair_params(force)
                p = np.array(p, dtype=np.float64)
                if verify:
                    self._verify(force, family, cutoff, p)
                self._pair_index[id(force)] = len(self._pair_index)
                lib.orc_add_pair(self.handle, family, group, cutoff, _dp(p), len(p))
            elif kind == 'NonbondedForce':
                if force.getNumParticles() == 0:
                    continue
                table = _columns(force, '_particles', 0)[1].reshape(n, 3)
                self._set_params(table[:, 0], table[:, 1], table[:, 2], params_set)
                params_set = True
                excl, exc_values = _columns(force, '_exceptions', 2)
                excl = np.ascontiguousarray(excl)
                exc_values = exc_values.reshape(len(excl), 3)
                lib.orc_set_exclusions(self.handle, len(excl), _ip(np.ascontiguousarray(excl.reshape(-1))))
                method = force.getNonbondedMethod()
                cutoff = force.getCutoffDistance().value_in_md_units()
                use_sw = 1.0 if force.getUseSwitchingFunction() else 0.0
                rs = force.getSwitchingDistance().value_in_md_units()
                alpha = 0.0
                if method == 2:
                    es = force.getReactionFieldDielectric()
                    p = [refmath.ONE_4PI_EPS0, 2.0, (es - 1)/((2*es + 1)*cutoff**3), 3*es/((2*es + 1)*cutoff), 0.0, use_sw, rs, cutoff]
                elif method >= 3:
                    alpha = refmath.pme_parameters(force, self.box)[0]
                    p = [refmath.ONE_4PI_EPS0, 3.0, 0.0, 0.0, alpha, use_sw, rs, cutoff]
                else:
                    raise ValueError('C port needs a periodic cutoff method')
                p = np.array(p, dtype=np.float64)
                self._pair_index[id(force)] = len(self._pair_index)
                lib.orc_add_pair(self.handle, F_LJC, group, cutoff, _dp(p), len(p))
                if force.getUseDispersionCorrection():
                    self.econst[group] = self.econst.get(group, 0.0) + refmath.nonbonded_lrc(force, self.box)
                keep = np.ones(len(excl), dtype=bool) if alpha > 0 else (exc_values[:, 0] != 0) | (exc_values[:, 2] != 0)
                if keep.any():
                    atoms = np.ascontiguousarray(excl[keep])
                    q = table[:, 0]
                    prm = np.concatenate([exc_values[keep], (q[atoms[:, 0]]*q[atoms[:, 1]])[:, None]], axis=1)
                    self._add_bonded(B_LJC, group, atoms, prm, [refmath.ONE_4PI_EPS0, alpha])
            elif kind == 'HarmonicBondForce':
                if force.getNumBonds():
                    self._add_bonded(B_BOND, group, *_columns(force, '_bonds', 2))
            elif kind == 'HarmonicAngleForce':
                if force.getNumAngles():
                    self._add_bonded(B_ANGLE, group, *_columns(force, '_angles', 3))
            elif kind == 'PeriodicTorsionForce':
                if force.getNumTorsions():
                    self._add_bonded(B_TORSION, group, *_columns(force, '_torsions', 4))
            elif kind == 'CustomBondForce':
                text = force.getEnergyFunction().replace(' ', '')
                names = [force.getPerBondParameterName(k) for k in range(force.getNumPerBondParameters())]
                if not (text.startswith('4*epsilon*x*(x-1)+Kc*chargeprod/r') and names == ['chargeprod', 'sigma', 'epsilon']):
                    raise ValueError('C port: unsupported CustomBondForce %s' % force.getEnergyFunction())
                import re
                m = re.search(r'Kc=([0-9.]+)', text)
                kc = float(m.group(1)) if m else force.getGlobalParameterDefaultValue(0)
                atoms, prm = _columns(force, '_bonds', 2)
                prm = np.concatenate([prm.reshape(len(atoms), 3), np.zeros((len(atoms), 1))], axis=1)
                self._add_bonded(B_LJC, group, atoms, prm, [kc, 0.0])
            elif kind in ('CMMotionRemover', 'MonteCarloBarostat'):
                continue
            else:
                raise ValueError('C port: unsupported force %s' % kind)

    def __del__(self):
        try:
            self.lib.orc_destroy(self.handle)
        except Exception:
            pass

    def _set_params(self, q, sigma, eps, already):
        q, sigma, eps = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, sigma, eps))
        if already:
            # one parameter table for all pair forces; tables imported through unit conversions may differ in the last bit
            assert all(np.allclose(a, b, rtol=1e-14, atol=0) for a, b in ((q, self._q), (sigma, self._sigma), (eps, self._eps)))
            return
        self._q, self._sigma, self._eps = q, sigma, eps
        self.lib.orc_set_params(self.handle, _dp(q), _dp(sigma), _dp(eps))

    def _add_bonded(self, family, group, atoms, params, g=(0.0,)):
        atoms = np.ascontiguousarray(atoms, dtype=np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        g = np.array(g, dtype=np.float64)
        self.lib.orc_add_bonded(self.handle, family, group, len(atoms), _ip(atoms), _dp(params), params.shape[1], _dp(g), len(g))

    def _verify(self, force, family, cutoff, p):
        """Closed form in oracle.c == the force's own energy string, on a two-particle probe."""
        import sympy
        names = ['charge', 'sigma', 'epsilon']
        glob = {force.getGlobalParameterName(k): force.getGlobalParameterDefaultValue(k)
                for k in range(force.getNumGlobalParameters())}
        syms = [sympy.Symbol('r')] + [sympy.Symbol(nm + '1') for nm in names] + [sympy.Symbol(nm + '2') for nm in names]
        _, fe, fd = refmath._pair_functions(force.getEnergyFunction(), glob, syms)
        lib = self.lib
        box = np.array([10.0, 10.0, 10.0])
        probe = ctypes.c_void_p()
        mass = np.ones(2)
        lib.orc_create(2, _dp(mass), _dp(box), ctypes.byref(probe))
        q, s, e = np.array([0.42, -0.84]), np.array([0.25, 0.3165]), np.array([0.1, 0.65])
        lib.orc_set_params(probe, _dp(q), _dp(s), _dp(e))
        lib.orc_add_pair(probe, family, 0, cutoff, _dp(p), len(p))
        for r in np.linspace(0.15, cutoff*0.999, 40):
            x = np.array([1.0, 1.0, 1.0, 1.0 + r, 1.0, 1.0])
            f = np.zeros(6)
            energy, virial = ctypes.c_double(), ctypes.c_double()
            lib.orc_eval(probe, _dp(x), 1, _dp(f), ctypes.byref(energy), ctypes.byref(virial))
            expect = float(fe(r, q[0], s[0], e[0], q[1], s[1], e[1]))
            dexpect = float(fd(r, q[0], s[0], e[0], q[1], s[1], e[1]))
            if force.getUseSwitchingFunction():
                S, dS = refmath.omm_switch(np.array(r), force.getSwitchingDistance().value_in_md_units(), cutoff)
                expect, dexpect = float(S)*expect, float(S)*dexpect + expect*float(dS)
            assert abs(energy.value - expect) <= 2e-7*max(1.0, abs(expect)), (r, energy.value, expect)
            assert abs(f[3] + dexpect) <= 2e-7*max(1.0, abs(dexpect)), (r, f[3], -dexpect)
        lib.orc_destroy(probe)

    # ---------------------------------------------------------------------------------------------
    @staticmethod
    def _mask(groups):
        if groups is None:
            return 0xffffffff
        m = 0
        for g in groups:
            m |= 1 << g
        return m

    def evaluate(self, positions, groups=None):
        x = np.ascontiguousarray(positions, dtype=np.float64)
        f = np.zeros_like(x)
        energy, virial = ctypes.c_double(), ctypes.c_double()
        self.lib.orc_eval(self.handle, _dp(x), self._mask(groups), _dp(f), ctypes.byref(energy), ctypes.byref(virial))
        constant = sum(v for g, v in self.econst.items() if groups is None or g in groups)
        return f, energy.value + constant, virial.value

    def pair_set(self, positions, force):
        """(count, checksum) of the exact interacting pair set of a pair force of the system: i < j,
        r^2 < rc^2 in float64, not excluded (the definition of the engine's Context.pair_set)."""
        which = self._pair_index[id(force)]
        x = np.ascontiguousarray(positions, dtype=np.float64)
        count, checksum = ctypes.c_longlong(), ctypes.c_ulonglong()
        code = self.lib.orc_pair_set(self.handle, _dp(x), which, ctypes.byref(count), ctypes.byref(checksum))
        assert code == 0
        return count.value, checksum.value

    def respa(self, positions, velocities, nsteps, dt, n0, n1, nose_hoover=None):
        """Advance (copies of) x, v by nsteps of RespaPropagator([n0, n1, 1]); nose_hoover =
        (nloops, LkT, Q, p_eta) wraps it in TrotterSuzuki(., SuzukiYoshida(NoseHoover, 3))."""
        x = np.array(positions, dtype=np.float64, order='C')
        v = np.array(velocities, dtype=np.float64, order='C')
        nloops, LkT, Q, p_eta = nose_hoover if nose_hoover else (1, 0.0, 1.0, 0.0)
        p = ctypes.c_double(p_eta)
        self.lib.orc_respa(self.handle, _dp(x), _dp(v), nsteps, dt, n0, n1, 1 if nose_hoover else 0, nloops, LkT, Q,
                           ctypes.byref(p))
        return x, v, p.value

    def counters(self):
        return dict(rebuilds=self.lib.orc_counter(self.handle, 0), pair_evals=self.lib.orc_counter(self.handle, 1))
