"""
Drop-in check of the description layer: the step programs, energy strings and force settings that
atomsmm_b200's classes emit are compared with what the REFERENCE's own Python classes emit for the
same constructions (tests/golden/ref_programs.json, captured by tests/golden/capture_reference.py
from /root/reference/src).  Strings must match after whitespace removal or, failing that, be
numerically identical expressions.
"""

import json
import os
import random
import re
import sys

import pytest

import atomsmm_b200 as atomsmm
from atomsmm_b200 import app, expr, unit

import fixtures

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
from cases import build_cases  # noqa: E402

with open(os.path.join(HERE, 'golden', 'ref_programs.json')) as handle:
    REFERENCE = json.load(handle)

# the reference's literal Bussi program is reproduced with per_dof_noise=True (documented deviation)
_OURS = None


def ours():
    global _OURS
    if _OURS is None:
        real = atomsmm.propagators.VelocityRescalingPropagator

        class Literal(real):
            def __init__(self, *args, **kwargs):
                kwargs.setdefault('per_dof_noise', True)
                super().__init__(*args, **kwargs)
        atomsmm.propagators.VelocityRescalingPropagator = Literal
        try:
            _OURS = json.loads(json.dumps(build_cases(atomsmm, unit, app, fixtures.load)))
        finally:
            atomsmm.propagators.VelocityRescalingPropagator = real
    return _OURS


def squash(text):
    return re.sub(r'\s', '', text)


def same_expression(a, b):
    if squash(a) == squash(b):
        return True
    try:
        ta, tb = expr.parse_inlined(a), expr.parse_inlined(b)
    except expr.ParseError:
        return False
    names = sorted(expr.free_symbols(ta) | expr.free_symbols(tb))
    rng = random.Random(7)
    for _ in range(12):
        env = {n: rng.uniform(0.35, 0.95) for n in names}
        env['__deriv__'] = lambda e, p: 0.25 + 0.001*len(e + p)
        try:
            va, vb = expr.evaluate(ta, env), expr.evaluate(tb, env)
        except (ValueError, ZeroDivisionError, OverflowError):
            continue
        if abs(va - vb) > 1e-9*max(1.0, abs(va), abs(vb)):
            return False
    return True


@pytest.mark.parametrize('name', sorted(REFERENCE['integrators']))
def test_integrator_program(name):
    ref, new = REFERENCE['integrators'][name], ours()['integrators'][name]
    assert new['perdof'] == ref['perdof']
    assert sorted(new['globals']) == sorted(ref['globals'])
    for key in ref['globals']:
        assert new['globals'][key] == pytest.approx(ref['globals'][key], rel=1e-12, abs=1e-300), key
    assert new['dt'] == pytest.approx(ref['dt'], rel=1e-12)
    assert len(new['steps']) == len(ref['steps'])
    for k, (a, b) in enumerate(zip(new['steps'], ref['steps'])):
        assert a[0] == b[0] and a[1] == b[1], (k, a, b)
        if a[0] in (6, 7):      # block conditions
            assert squash(a[2]) == squash(b[2]), (k, a, b)
        elif b[2]:
            assert same_expression(a[2], b[2]), (k, a, b)


def compare_force(new, ref, label):
    for key in ref:
        if key == 'energy':
            assert same_expression(new[key], ref[key]), (label, new[key], ref[key])
        elif key == 'globals':
            assert sorted(new[key]) == sorted(ref[key]), label
            for g in ref[key]:
                assert new[key][g] == pytest.approx(ref[key][g], rel=1e-12), (label, g)
        elif isinstance(ref[key], float):
            assert new[key] == pytest.approx(ref[key], rel=1e-12), (label, key)
        else:
            assert new[key] == ref[key], (label, key)


@pytest.mark.parametrize('name', sorted(REFERENCE['forces']))
def test_force_description(name):
    ref, new = REFERENCE['forces'][name], ours()['forces'][name]
    if isinstance(ref, list):
        assert len(ref) == len(new)
        for k, (a, b) in enumerate(zip(new, ref)):
            compare_force(a, b, '%s[%d]' % (name, k))
    elif 'near' in ref and isinstance(ref['near'], list):
        assert len(ref['near']) == len(new['near'])
        for a, b in zip(new['near'], ref['near']):
            assert squash(a) == squash(b) or same_expression(a.split('=', 1)[-1], b.split('=', 1)[-1]), (a, b)
    else:
        compare_force(new, ref, name)


@pytest.mark.parametrize('name', sorted(REFERENCE['systems']))
def test_system_layout(name):
    ref, new = REFERENCE['systems'][name], ours()['systems'][name]
    # trailing records without a class are plain data (e.g. the decoupled solute's parameters)
    extra_ref, extra_new = [f for f in ref if 'cls' not in f], [f for f in new if 'cls' not in f]
    ref, new = [f for f in ref if 'cls' in f], [f for f in new if 'cls' in f]
    assert [f['cls'] for f in new] == [f['cls'] for f in ref]
    for k, (a, b) in enumerate(zip(new, ref)):
        compare_force(a, b, '%s[%d]' % (name, k))
    assert extra_new == extra_ref
