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


def test_hilbert_order_keeps_groups_compact():
    """The engine sorts molecules along a Hilbert curve (pure host logic of the C library): consecutive
    keys are face-adjacent cells, so the 8-atom groups of the sorted water box are compact -- which is
    what keeps neighbour lists short and ownership ranges local."""
    import numpy as np
    from atomsmm_b200 import engine
    import systems
    pdb, _ = systems.fixtures.load('q-SPC-FW')
    pos = systems.positions_of(pdb)
    box = np.array([2.5, 2.5, 2.5])
    first = pos[0::3]                                       # first atom (O) of every water
    keys = np.array([engine.hilbert_index(p, box) for p in first], dtype=np.uint64)
    assert len(set(keys.tolist())) > 0.5*len(keys)          # about one molecule per cell
    order = np.argsort(keys, kind='stable')
    atoms = (3*order[:, None] + np.arange(3)[None, :]).reshape(-1)
    sorted_pos = pos[atoms]

    def mean_extent(p):
        groups = p[:len(p)//8*8].reshape(-1, 8, 3)
        d = groups - groups[:, :1]
        d -= box*np.round(d/box)
        return float(np.mean(np.max(d, axis=1) - np.min(d, axis=1)))
    assert mean_extent(sorted_pos) < 0.75                    # nm; ~0.55 in practice
    assert mean_extent(sorted_pos) < 0.6*mean_extent(pos)    # the PDB order is far less compact
    # adjacent cells along the curve: a small step in space
    a = engine.hilbert_index([0.01, 0.01, 0.01], box)
    assert isinstance(a, int) and engine.hilbert_index([0.01 + 2.5, 0.01, 0.01 - 2.5], box) == a   # periodic


def test_barostat_random_stream_is_the_oracles():
    """b2_barostat_uniform (pure host function) and the oracle interpreter's restatement of SplitMix64 agree
    bit for bit: the basis of replaying the engine's accept / reject sequence in float64."""
    import ctypes
    from atomsmm_b200 import engine
    from oracle import interp
    lib = engine.library()
    for seed in (0, 1, 77, 2**63 + 12345):
        for counter in (0, 1, 2, 1000, 2**40):
            out = ctypes.c_double()
            assert lib.b2_barostat_uniform(seed, counter, ctypes.byref(out)) == 0
            assert out.value == interp.Interpreter.barostat_uniform(seed, counter)
            assert 0.0 <= out.value < 1.0
