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
        if not os.path.exists(LIB_PATH):
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
        lib.orc_counter.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.orc_counter.restype = ctypes.c_long
        _lib = lib
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


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
        self.mass = np.array([system.getParticleMass(i).value_in_md_units() for i in range(n)], dtype=np.float64)
        self.handle = ctypes.c_void_p()
        lib.orc_create(n, _dp(self.mass), _dp(self.box), ctypes.byref(self.handle))
        if threads:
            lib.orc_set_threads(self.handle, threads)
        self.threads = threads or os.cpu_count()
        params_set = False
        self.econst = {}      # position-independent energies (long-range corrections) per group
        for force in system.getForces():
            kind = refmath._kind(force)
            group = force.getForceGroup()
            if kind == 'CustomNonbondedForce':
                table = np.array([force.getParticleParameters(k) for k in range(n)], dtype=np.float64)
                self._set_params(table[:, 0], table[:, 1], table[:, 2], params_set)
                params_set = True
                excl = np.array([force.getExclusionParticles(k) for k in range(force.getNumExclusions())], dtype=np.int32)
                lib.orc_set_exclusions(self.handle, len(excl), _ip(np.ascontiguousarray(excl.reshape(-1))))
                family, cutoff, p = _custom_pair_params(force)
                p = np.array(p, dtype=np.float64)
                if verify:
                    self._verify(force, family, cutoff, p)
                lib.orc_add_pair(self.handle, family, group, cutoff, _dp(p), len(p))
            elif kind == 'NonbondedForce':
                if force.getNumParticles() == 0:
                    continue
                table = np.array([[v.value_in_md_units() for v in force.getParticleParameters(k)] for k in range(n)])
                self._set_params(table[:, 0], table[:, 1], table[:, 2], params_set)
                params_set = True
                exc = [force.getExceptionParameters(k) for k in range(force.getNumExceptions())]
                excl = np.array([[e[0], e[1]] for e in exc], dtype=np.int32)
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
                lib.orc_add_pair(self.handle, F_LJC, group, cutoff, _dp(p), len(p))
                if force.getUseDispersionCorrection():
                    self.econst[group] = self.econst.get(group, 0.0) + refmath.nonbonded_lrc(force, self.box)
                keep = [e for e in exc if alpha > 0 or e[2].value_in_md_units() != 0 or e[4].value_in_md_units() != 0]
                if keep:
                    atoms = np.array([[e[0], e[1]] for e in keep], dtype=np.int32)
                    q = table[:, 0]
                    prm = np.array([[e[2].value_in_md_units(), e[3].value_in_md_units(), e[4].value_in_md_units(),
                                     q[e[0]]*q[e[1]]] for e in keep], dtype=np.float64)
                    self._add_bonded(B_LJC, group, atoms, prm, [refmath.ONE_4PI_EPS0, alpha])
            elif kind == 'HarmonicBondForce':
                b = [force.getBondParameters(k) for k in range(force.getNumBonds())]
                if b:
                    self._add_bonded(B_BOND, group, np.array([[x[0], x[1]] for x in b], dtype=np.int32),
                                     np.array([[x[2].value_in_md_units(), x[3].value_in_md_units()] for x in b]))
            elif kind == 'HarmonicAngleForce':
                a = [force.getAngleParameters(k) for k in range(force.getNumAngles())]
                if a:
                    self._add_bonded(B_ANGLE, group, np.array([x[:3] for x in a], dtype=np.int32),
                                     np.array([[x[3].value_in_md_units(), x[4].value_in_md_units()] for x in a]))
            elif kind == 'PeriodicTorsionForce':
                t = [force.getTorsionParameters(k) for k in range(force.getNumTorsions())]
                if t:
                    self._add_bonded(B_TORSION, group, np.array([x[:4] for x in t], dtype=np.int32),
                                     np.array([[float(x[4]), x[5].value_in_md_units(), x[6].value_in_md_units()] for x in t]))
            elif kind == 'CustomBondForce':
                text = force.getEnergyFunction().replace(' ', '')
                names = [force.getPerBondParameterName(k) for k in range(force.getNumPerBondParameters())]
                if not (text.startswith('4*epsilon*x*(x-1)+Kc*chargeprod/r') and names == ['chargeprod', 'sigma', 'epsilon']):
                    raise ValueError('C port: unsupported CustomBondForce %s' % force.getEnergyFunction())
                import re
                m = re.search(r'Kc=([0-9.]+)', text)
                kc = float(m.group(1)) if m else force.getGlobalParameterDefaultValue(0)
                b = [force.getBondParameters(k) for k in range(force.getNumBonds())]
                atoms = np.array([[x[0], x[1]] for x in b], dtype=np.int32)
                prm = np.array([list(x[2]) + [0.0] for x in b], dtype=np.float64)
                self._add_bonded(B_LJC, group, atoms, prm, [kc, 0.0])
            elif kind == 'CMMotionRemover':
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
            assert np.array_equal(q, self._q) and np.array_equal(sigma, self._sigma) and np.array_equal(eps, self._eps)
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
