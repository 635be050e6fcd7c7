"""CPU checks of the boundary: the C-ABI library builds in-tree, exports every symbol declared in
include/atomsmm_b200.h, and the product refuses to run without a GPU (no CPU fallback)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'atomsmm_b200.h')).read()
    return sorted(set(re.findall(r'B2_API\s+[\w\s\*]+?\b(b2_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from atomsmm_b200 import build
    library = build.build()
    lib = ctypes.CDLL(library)
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), name


def test_engine_signatures_cover_the_header():
    from atomsmm_b200 import engine
    bound = set(engine._SIGNATURES) | {'b2_last_error', 'b2_version'}
    assert set(declared_symbols()) <= bound


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from atomsmm_b200 import engine, mm
    import systems
    system, pdb, _ = systems.water_near(None)
    with pytest.raises(mm.OpenMMException):
        mm.Platform.getPlatformByName('Reference')
    with pytest.raises(engine.EngineError):
        mm.Context(system, mm.VerletIntegrator(0.0), mm.Platform.getPlatformByName('B200'))


def test_product_never_imports_the_oracle():
    package = os.path.join(ROOT, 'atomsmm_b200')
    for folder, _, files in os.walk(package):
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(folder, name)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, re.M), name
