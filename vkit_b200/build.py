"""Builds vkit_b200/csrc/libvkit_b200.so for sm_100a with nvcc (in-tree, no JIT cache).

    python -m vkit_b200.build [--force] [--verbose]
"""
import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB = os.path.join(CSRC, 'libvkit_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--shared',
]


def _find_nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    lib_mtime = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + [
        os.path.join(_HERE, '..', 'include', 'vkit_b200.h')]
    return any(os.path.getmtime(d) > lib_mtime for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [_find_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + sources() + ['-o', LIB]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
