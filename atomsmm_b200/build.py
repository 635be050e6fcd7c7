"""
Build the CUDA engine in-tree:  python -m atomsmm_b200.build [--force] [--verbose]

Every .cu under csrc/ is compiled for sm_100a only (``-gencode arch=compute_100a,code=sm_100a
-lineinfo``) and linked into atomsmm_b200/libatomsmm_b200.so, which travels with the repository
snapshot to the GPU box.  nvcc cross-compiles without a GPU.
"""

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'csrc', '_obj')
LIBRARY = os.path.join(HERE, 'libatomsmm_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-I', OBJ,
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '-Xptxas', '-v']
# tuning experiments: B2_EXTRA_NVCC_FLAGS="-DB2_PAIR_MINB=6" python -m atomsmm_b200.build --force, B2_LIBRARY=<path> at run time
FLAGS += os.environ.get('B2_EXTRA_NVCC_FLAGS', '').split()
LIBRARY = os.environ.get('B2_BUILD_OUTPUT', LIBRARY)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers():
    found = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    found.append(os.path.join(HERE, '..', 'include', 'atomsmm_b200.h'))
    return found


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
    if _stale(obj, [src] + headers()):
        cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, proc.stdout, proc.stderr))
        with open(obj + '.log', 'w') as handle:
            handle.write(proc.stderr)
        if verbose:
            sys.stderr.write(proc.stderr)
    return obj


def _embed_rng_source():
    """csrc/_obj/vm_rng_embed.h: the Philox / RngStream section of vm.cuh as a string constant, so that the kernels
    csrc/jit.cu generates at run time draw from the very same stream as the precompiled ones."""
    with open(os.path.join(CSRC, 'vm.cuh')) as handle:
        text = handle.read()
    begin = text.index('// ---- Philox4x32-10')
    end = text.index('// ---- variable access')
    body = 'static const char* B2_RNG_SOURCE = R"B2RNG(\n' + text[begin:end] + ')B2RNG";\n'
    path = os.path.join(OBJ, 'vm_rng_embed.h')
    if not os.path.exists(path) or open(path).read() != body:
        with open(path, 'w') as handle:
            handle.write(body)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    _embed_rng_source()
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        objects = list(pool.map(lambda s: _compile(s, verbose), sources()))
    if _stale(LIBRARY, objects):
        # link under a temporary name and rename: a snapshot of the tree never sees a half-written library
        staging = LIBRARY + '.tmp%d' % os.getpid()
        cmd = [NVCC, '-shared', '-o', staging] + objects + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                            '-Xcompiler', '-fPIC', '-lcudart', '-lcufft', '-ldl']
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (proc.stdout, proc.stderr))
        os.replace(staging, LIBRARY)
    return LIBRARY


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
