"""The packed fp32x2 force functors of the pair tiles (csrc/potentials.cuh, LJCForce2: two list slots per
instruction issue on sm_100a) restate the closed forms of forces.py:448-455,541-563 and systems.py:894.  The
header also compiles for the host, so the restatement is checked here -- without a GPU -- against the float64
functors (which tests/test_gpu_forces.py and the oracle goldens pin) on random pairs, over the whole range and
inside the switching zone alone."""

import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


@pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which('nvcc')), reason='nvcc not available')
def test_packed_functors_match_float64_closed_forms(tmp_path):
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which('nvcc')
    exe = str(tmp_path/'packed_check')
    subprocess.run([nvcc, '-O2', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', exe,
                    os.path.join(HERE, 'native', 'packed_potentials_check.cu')], check=True, capture_output=True)
    lines = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split('\n')
    seen = {}
    for line in lines:
        if line.strip():
            name, rms, worst = line.split()
            seen[name] = (float(rms), float(worst))
    # the instantiations dispatched by csrc/pair.cu (12 lines) plus the switching zones of the 9 switched ones
    assert len(seen) == 21
    for name, (rms, worst) in seen.items():
        assert rms < 5e-6, (name, rms)          # errors relative to the RMS of the reference values
        assert worst < 1e-4, (name, worst)
